// TEST STAND-INS for the reference's own headers (Lecanyu/SparsifyPoseGraph src/): only the interfaces that
// integration/spg_vertex_remover_g2o.hpp touches, written from the documented call sites — not copies of the sources.
// The per-header files next to this one (sparsity_options.h, glc_edge.h, ...) just include it.
#ifndef STUB_REFERENCE_H_
#define STUB_REFERENCE_H_
#include <list>
#include <map>
#include <g2o/core/sparse_optimizer.h>

// src/sparsity_options.h:12-29
struct SparsityOptions {
    enum SparsityTopology { Tree, Subgraph, CliqueySubgraph, Dense, CliqueyDense };
    enum LinearizationPoint { Local, Global };
    SparsityTopology topology;
    double chordRatio;
    LinearizationPoint linPoint;
    bool includeIntraClique;
    SparsityOptions() : topology(Tree), chordRatio(1), linPoint(Local), includeIntraClique(true) {}
};

// src/auto_delete_map.h: a map that owns its values
template <typename K, typename T>
class AutoDeleteMap : public std::map<K, T> {
public:
    virtual ~AutoDeleteMap() { for(auto &v : *this) delete v.second; }
    typename std::map<K, T>::size_type erase(const K &which) {
        delete (*this)[which];
        erased++;
        return std::map<K, T>::erase(which);
    }
    int erased = 0; // (stand-in only: lets the test count the calls)
};

// src/graph_wrapper.h:31-37
class GraphWrapper {
public:
    class Edge { public: virtual ~Edge() {} };
};

// src/se2_compatibility.h:20, src/se3_compatibility.h:25
class EdgeSE2ISAM : public g2o::EdgeSE2 {};
class EdgeSE3ISAM : public g2o::EdgeSE3 {};

// src/glc_reparam.h:15, glc_reparam_se2.h:30, glc_reparam_se3.h:16
class GLCReparam : public g2o::HyperGraph::HyperGraphElement { public: virtual ~GLCReparam() {} };
class GLCReparamSE2ISAM : public GLCReparam {};
class GLCReparamSE3 : public GLCReparam {};

// src/glc_edge.h:14-70
class GLCEdge : public g2o::BaseMultiEdge<-1, g2o::VectorXD> {
public:
    GLCEdge() : _reparam(NULL) {}
    virtual ~GLCEdge() { delete _reparam; }
    void setReparam(GLCReparam *reparam) { delete _reparam; _reparam = reparam; }
    const GLCReparam *reparam() const { return _reparam; }
    void setDimension(int errsize, int meassize) {
        resize(meassize);
        _dimension = errsize;
        _information.resize(errsize, errsize);
        _error.resize(errsize, 1);
        _measurement.resize(meassize, 1);
        _W.resize(errsize, meassize);
    }
    const g2o::MatrixXD &linearWeight() const { return _W; }
    void setLinearWeight(const g2o::MatrixXD &W) { _W = W; }
private:
    GLCReparam *_reparam;
    g2o::MatrixXD _W;
};

// src/multi_edge_correlated.h:13-120
template <typename EdgeType>
class MultiEdgeCorrelated : public g2o::BaseMultiEdge<-1, std::list<typename EdgeType::Measurement> > {
public:
    typedef typename EdgeType::Measurement Measurement;
    typedef g2o::BaseMultiEdge<-1, std::list<Measurement> > Base;
    typedef std::list<std::list<int> > MappingsType;
    void setDimension(int dimension_) {
        this->_vertices.resize(0);
        this->_dimension = dimension_;
        this->_information.resize(dimension_, dimension_);
    }
    void setMeasurementCount(int measurements) { setDimension(measurements * EdgeType::Dimension); }
    int measurementCount() const { return this->_dimension / EdgeType::Dimension; }
    void addMeasurement(const g2o::HyperGraph::VertexContainer &verts, const Measurement &meas) {
        std::list<int> indices;
        for(g2o::HyperGraph::Vertex *v : verts) {
            size_t pos = 0;
            while(pos < this->_vertices.size() && this->_vertices[pos] != v) pos++;
            if(pos == this->_vertices.size()) this->_vertices.push_back(v);
            indices.push_back((int) pos);
        }
        _mappings.push_back(indices);
        this->_measurement.push_back(meas);
    }
    struct const_iterator {
        const_iterator(const MultiEdgeCorrelated *e, bool end)
            : _it(end ? e->_mappings.end() : e->_mappings.begin()), _itm(e->_measurement.begin()), _edge(e) {}
        const_iterator &operator++() { ++_it; ++_itm; return *this; }
        bool operator!=(const const_iterator &c) const { return c._it != _it; }
        const const_iterator &operator*() const { return *this; }
        g2o::HyperGraph::VertexContainer vertices() const {
            g2o::HyperGraph::VertexContainer r;
            for(int i : *_it) r.push_back(_edge->_vertices[i]);
            return r;
        }
        Measurement measurement() const { return *_itm; }
        MappingsType::const_iterator _it;
        typename std::list<Measurement>::const_iterator _itm;
        const MultiEdgeCorrelated *_edge;
    };
    const_iterator begin() const { return const_iterator(this, false); }
    const_iterator end() const { return const_iterator(this, true); }
private:
    MappingsType _mappings;
};

// src/topology_provider.h:16-33, topology_provider_glc.h:15, the SE2/SE3 ISAM providers
class TopologyProvider { public: virtual ~TopologyProvider() {} };
class TopologyProviderGLC : public TopologyProvider {};
class TopologyProviderSE2ISAM : public TopologyProvider {};
class TopologyProviderSE3ISAM : public TopologyProvider {};

#endif
