// TEST STAND-IN, not g2o: the handful of g2o interfaces integration/spg_vertex_remover_g2o.hpp touches, with the member
// names and signatures of g2o (core/hyper_graph.h, core/optimizable_graph.h, types/slam2d, types/slam3d), so that the
// adapter can be compiled and run where g2o is not installed (tests/test_g2o_adapter.py). Dense types are minimal
// classes with the Eigen spelling of the operations used ((i, j), rows(), cols(), size(), setIdentity()).
#ifndef STUB_G2O_SPARSE_OPTIMIZER_H_
#define STUB_G2O_SPARSE_OPTIMIZER_H_
#include <cmath>
#include <cstddef>
#include <map>
#include <set>
#include <vector>

namespace g2o {

class MatrixXD {
public:
    MatrixXD() : _r(0), _c(0) {}
    MatrixXD(int r, int c) : _r(r), _c(c), _d((size_t) r * c, 0.0) {}
    void resize(int r, int c) { _r = r; _c = c; _d.assign((size_t) r * c, 0.0); }
    int rows() const { return _r; }
    int cols() const { return _c; }
    double &operator()(int i, int j) { return _d[i + (size_t) j * _r]; } // column-major, like Eigen's default
    double operator()(int i, int j) const { return _d[i + (size_t) j * _r]; }
    double *data() { return _d.data(); }
    const double *data() const { return _d.data(); }
    void setIdentity() { for(int j = 0; j < _c; j++) for(int i = 0; i < _r; i++) (*this)(i, j) = i == j ? 1.0 : 0.0; }
private:
    int _r, _c;
    std::vector<double> _d;
};
class VectorXD {
public:
    VectorXD() {}
    explicit VectorXD(int n) : _d(n, 0.0) {}
    void resize(int n, int = 1) { _d.assign(n, 0.0); }
    size_t size() const { return _d.size(); }
    double &operator()(int i) { return _d[i]; }
    double operator()(int i) const { return _d[i]; }
private:
    std::vector<double> _d;
};

class HyperGraph {
public:
    class Edge;
    class Vertex {
    public:
        explicit Vertex(int id = -1) : _id(id) {}
        virtual ~Vertex() {}
        int id() const { return _id; }
        void setId(int id) { _id = id; }
        const std::set<Edge *> &edges() const { return _edges; }
        std::set<Edge *> &edges() { return _edges; }
    protected:
        int _id;
        std::set<Edge *> _edges;
    };
    typedef std::vector<Vertex *> VertexContainer;
    typedef std::set<Vertex *> VertexSet;
    typedef std::set<Edge *> EdgeSet;
    typedef std::map<int, Vertex *> VertexIDMap;
    class Edge {
    public:
        virtual ~Edge() {}
        const VertexContainer &vertices() const { return _vertices; }
        VertexContainer &vertices() { return _vertices; }
        Vertex *vertex(size_t i) { return _vertices[i]; }
        const Vertex *vertex(size_t i) const { return _vertices[i]; }
        void setVertex(size_t i, Vertex *v) { _vertices[i] = v; }
        virtual void resize(size_t n) { _vertices.resize(n, NULL); }
    protected:
        VertexContainer _vertices;
    };
    class HyperGraphElement { public: virtual ~HyperGraphElement() {} };
    virtual ~HyperGraph() {
        for(Edge *e : _edges) delete e;
        for(auto &iv : _vertices) delete iv.second;
    }
    const VertexIDMap &vertices() const { return _vertices; }
    const EdgeSet &edges() const { return _edges; }
protected:
    VertexIDMap _vertices;
    EdgeSet _edges;
};

class OptimizableGraph : public HyperGraph {
public:
    class Vertex : public HyperGraph::Vertex {
    public:
        virtual bool getEstimateData(double *) const = 0;
        virtual bool setEstimateData(const double *) = 0;
        virtual int estimateDimension() const = 0;
        virtual int dimension() const = 0;
    };
    class Edge : public HyperGraph::Edge {
    public:
        virtual int dimension() const = 0;
        virtual double *informationData() = 0;
        virtual const double *informationData() const = 0;
        virtual bool getMeasurementData(double *) const { return false; }
        virtual bool setMeasurementData(const double *) { return false; }
    };
    typedef std::vector<Edge *> EdgeContainer;
    Vertex *vertex(int id) {
        auto it = _vertices.find(id);
        return it == _vertices.end() ? NULL : static_cast<Vertex *>(it->second);
    }
    bool addVertex(Vertex *v) { return _vertices.insert(std::make_pair(v->id(), v)).second; }
    bool addEdge(Edge *e) {
        for(HyperGraph::Vertex *v : e->vertices()) v->edges().insert(e);
        return _edges.insert(e).second;
    }
    bool removeEdge(HyperGraph::Edge *e) { // g2o deletes the edge object
        if(!_edges.erase(e)) return false;
        for(HyperGraph::Vertex *v : e->vertices()) v->edges().erase(e);
        delete e;
        return true;
    }
    bool removeVertex(HyperGraph::Vertex *v) { // detaches (and deletes) every edge still incident, then the vertex
        std::set<HyperGraph::Edge *> inc = v->edges();
        for(HyperGraph::Edge *e : inc) removeEdge(e);
        _vertices.erase(v->id());
        delete v;
        return true;
    }
};
class SparseOptimizer : public OptimizableGraph {};

// ---- types/slam2d, types/slam3d ------------------------------------------------------------------------------------
struct SE2 { double x, y, th; SE2() : x(0), y(0), th(0) {} };
struct Isometry3D { double t[3], q[4]; Isometry3D() : t{0, 0, 0}, q{0, 0, 0, 1} {} }; // stand-in for Eigen::Isometry3d

class VertexSE2 : public OptimizableGraph::Vertex {
public:
    bool getEstimateData(double *d) const { d[0] = _e.x; d[1] = _e.y; d[2] = _e.th; return true; }
    bool setEstimateData(const double *d) { _e.x = d[0]; _e.y = d[1]; _e.th = d[2]; return true; }
    int estimateDimension() const { return 3; }
    int dimension() const { return 3; }
private:
    SE2 _e;
};
class VertexSE3 : public OptimizableGraph::Vertex {
public:
    bool getEstimateData(double *d) const { for(int i = 0; i < 3; i++) d[i] = _e.t[i]; for(int i = 0; i < 4; i++) d[3 + i] = _e.q[i]; return true; }
    bool setEstimateData(const double *d) { for(int i = 0; i < 3; i++) _e.t[i] = d[i]; for(int i = 0; i < 4; i++) _e.q[i] = d[3 + i]; return true; }
    int estimateDimension() const { return 7; }
    int dimension() const { return 6; }
private:
    Isometry3D _e;
};

template <int D, class M>
class BaseBinaryEdgeStub : public OptimizableGraph::Edge {
public:
    typedef M Measurement;
    static const int Dimension = D;
    BaseBinaryEdgeStub() { _vertices.resize(2, NULL); for(double &x : _info) x = 0; }
    int dimension() const { return D; }
    double *informationData() { return _info; }
    const double *informationData() const { return _info; }
    const M &measurement() const { return _measurement; }
    void setMeasurement(const M &m) { _measurement = m; }
protected:
    M _measurement;
    double _info[D * D];
};
class EdgeSE2 : public BaseBinaryEdgeStub<3, SE2> {
public:
    bool getMeasurementData(double *d) const { d[0] = _measurement.x; d[1] = _measurement.y; d[2] = _measurement.th; return true; }
    bool setMeasurementData(const double *d) { _measurement.x = d[0]; _measurement.y = d[1]; _measurement.th = d[2]; return true; }
};
class EdgeSE3 : public BaseBinaryEdgeStub<6, Isometry3D> {
public:
    bool getMeasurementData(double *d) const { for(int i = 0; i < 3; i++) d[i] = _measurement.t[i]; for(int i = 0; i < 4; i++) d[3 + i] = _measurement.q[i]; return true; }
    bool setMeasurementData(const double *d) { for(int i = 0; i < 3; i++) _measurement.t[i] = d[i]; for(int i = 0; i < 4; i++) _measurement.q[i] = d[3 + i]; return true; }
};

// BaseMultiEdge<-1, Measurement>: dynamic error dimension, any number of vertices
template <int, class M>
class BaseMultiEdge : public OptimizableGraph::Edge {
public:
    typedef M Measurement;
    BaseMultiEdge() : _dimension(0) {}
    int dimension() const { return _dimension; }
    double *informationData() { return _information.data(); }
    const double *informationData() const { return _information.data(); }
    MatrixXD &information() { return _information; }
    const MatrixXD &information() const { return _information; }
    const M &measurement() const { return _measurement; }
    void setMeasurement(const M &m) { _measurement = m; }
protected:
    int _dimension;
    MatrixXD _information;
    VectorXD _error;
    M _measurement;
};

} // namespace g2o
#endif
