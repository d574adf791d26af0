// spg_vertex_remover_g2o.hpp — the reference's VertexRemover (src/vertex_remover.h:19-50) on top of libspg_b200.so.
//
// Header-only adapter for a build of the reference: it includes the reference's own headers and g2o, keeps g2o as the
// graph container, and replaces the body of VertexRemover::remove (src/vertex_remover.cpp:83-140) by
//   1. a mirror of the g2o graph in an spg_graph (vertices with their current estimates, every factor),
//   2. one spg_graph_marginalize call (wavefront rounds on the GPU, include/spg_capi.h),
//   3. the result mirrored back with the calls updateInputGraph makes (src/vertex_remover.cpp:500-546): removeEdge +
//      _edgeLookup->erase for the blanket edges, removeVertex, addEdge for the substitute factors — which are returned,
//      so that GraphWrapperG2O::marginalizeNoOptimize registers them in its edge map as before (:446-450).
// Usage inside the reference (src/graph_wrapper_g2o.cpp:431-446): replace `VertexRemover vr;` by
// `SpgVertexRemover vr(ctx);` — the calls that follow (registerTopologyProvider, setGraph, setEdgeMap,
// setSparsityOptions, remove) are the same. Link with -lspg_b200, compile with -I<repo>/include.
//
// g2o and the reference's sources are not available where this repository is built: tests/test_g2o_adapter.py
// compiles this header against minimal stand-ins of the interfaces it touches (tests/stubs/) and runs it on the CPU
// with the oracle standing in for the device; only raw-data accessors of g2o are used (getEstimateData,
// getMeasurementData, informationData, ...), so no Eigen expression crosses the boundary.
#ifndef SPG_VERTEX_REMOVER_G2O_HPP_
#define SPG_VERTEX_REMOVER_G2O_HPP_

#include <algorithm>
#include <functional>
#include <list>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <g2o/core/sparse_optimizer.h>

#include "auto_delete_map.h"          // reference src/
#include "glc_edge.h"
#include "glc_reparam_se2.h"
#include "glc_reparam_se3.h"
#include "graph_wrapper.h"
#include "multi_edge_correlated.h"
#include "se2_compatibility.h"
#include "se3_compatibility.h"
#include "sparsity_options.h"
#include "topology_provider.h"
#include "topology_provider_glc.h"

#include "spg_capi.h"

class SpgVertexRemover {
public:
    typedef g2o::OptimizableGraph::Vertex Vertex;
    typedef g2o::OptimizableGraph::Edge Edge;
    typedef AutoDeleteMap<const g2o::HyperGraph::Edge *, GraphWrapper::Edge *> EdgeMap;
    typedef MultiEdgeCorrelated<EdgeSE2ISAM> MultiEdgeSE2ISAM;
    typedef MultiEdgeCorrelated<EdgeSE3ISAM> MultiEdgeSE3ISAM;
    // what runs the removal on the mirror; the default is the GPU path (tests substitute a CPU engine)
    typedef std::function<spg_status(spg_graph *, const int32_t *, int32_t, const spg_sparsity_options *, int32_t)> Engine;

    explicit SpgVertexRemover(spg_ctx *ctx) : _ctx(ctx), _graph(NULL), _edgeLookup(NULL), _useGLC(false) {
        _engine = [this](spg_graph *g, const int32_t *which, int32_t n, const spg_sparsity_options *o, int32_t alg) {
            return spg_graph_marginalize(g, _ctx, which, n, o, alg);
        };
    }
    virtual ~SpgVertexRemover() {
        for(TopologyProvider *tp : _topologies) delete tp; // the reference's remover owns its providers
    }

    void setSparsityOptions(const SparsityOptions &opts) { _opts = opts; }
    void setGraph(g2o::OptimizableGraph *graph) { _graph = graph; }
    void setEdgeMap(EdgeMap *edgeMap) { _edgeLookup = edgeMap; }
    // the provider family decides GLC vs NFR exactly as in src/graph_wrapper_g2o.cpp:431-439
    void registerTopologyProvider(TopologyProvider *topology) {
        _topologies.push_back(topology);
        if(dynamic_cast<TopologyProviderGLC *>(topology)) _useGLC = true;
    }
    void setEngine(const Engine &engine) { _engine = engine; }
    const spg_marginalize_stats &lastStats() const { return _stats; }

    g2o::OptimizableGraph::EdgeContainer remove(Vertex *toRemove) {
        return remove(std::vector<Vertex *>(1, toRemove));
    }

    g2o::OptimizableGraph::EdgeContainer remove(const std::vector<Vertex *> &toRemove) {
        g2o::OptimizableGraph::EdgeContainer added;
        if(!_graph || toRemove.empty()) return added;
        const int dim = graphDim();
        spg_graph *g = NULL;
        check(spg_graph_create(&g, dim), "spg_graph_create");
        try {
            // ---- 1. mirror in: vertices (current estimates), then every factor; file order = the order of this loop ----
            for(const auto &iv : _graph->vertices()) {
                const Vertex *v = static_cast<const Vertex *>(iv.second);
                double pose[7] = {0, 0, 0, 0, 0, 0, 1};
                v->getEstimateData(pose); // VertexSE2: x y theta; VertexSE3: t, q (x y z w)
                check(spg_graph_add_vertex(g, v->id(), pose), "spg_graph_add_vertex");
            }
            std::vector<g2o::HyperGraph::Edge *> original; // index = uid_minor of the mirrored file edges
            std::vector<g2o::HyperGraph::Edge *> ordered(_graph->edges().begin(), _graph->edges().end());
            // a deterministic order (the EdgeSet is ordered by address): by vertex ids, like the reference's datasets
            std::stable_sort(ordered.begin(), ordered.end(), [](const g2o::HyperGraph::Edge *a, const g2o::HyperGraph::Edge *b) {
                const size_t n = std::min(a->vertices().size(), b->vertices().size());
                for(size_t i = 0; i < n; i++)
                    if(a->vertices()[i]->id() != b->vertices()[i]->id()) return a->vertices()[i]->id() < b->vertices()[i]->id();
                return a->vertices().size() < b->vertices().size();
            });
            for(g2o::HyperGraph::Edge *he : ordered) {
                mirrorEdgeIn(g, he, dim);
                original.push_back(he);
            }
            // ---- 2. the removal ----------------------------------------------------------------------------------------
            std::vector<int32_t> which;
            for(const Vertex *v : toRemove) which.push_back(v->id());
            spg_sparsity_options o;
            o.topology = (int32_t) _opts.topology; // enum values are identical (src/sparsity_options.h:12-18)
            o.lin_point = _opts.linPoint == SparsityOptions::Local ? SPG_LIN_LOCAL : SPG_LIN_GLOBAL;
            o.chord_ratio = _opts.chordRatio;
            o.include_intra_clique = _opts.includeIntraClique ? 1 : 0;
            o.flags = 0;
            const spg_status st = _engine(g, which.data(), (int32_t) which.size(), &o, _useGLC ? SPG_ALG_GLC : SPG_ALG_NFR);
            spg_graph_last_stats(g, &_stats);
            // SPG_ERR_BLANKET_FAILED: the failed blankets were left in the graph, everything else was applied — mirrored
            // back below, then reported (the reference asserts / exits at this point)
            if(st != SPG_OK && st != SPG_ERR_BLANKET_FAILED) fail("spg_graph_marginalize");
            // ---- 3. mirror out ---------------------------------------------------------------------------------------
            std::vector<char> alive(original.size(), 0);
            const int ne = spg_graph_num_edges(g);
            for(int i = 0; i < ne; i++) {
                spg_edge_desc d;
                check(spg_graph_edge_desc(g, i, &d), "spg_graph_edge_desc");
                if(d.uid_major < 0) alive[d.uid_minor] = 1;
            }
            for(size_t i = 0; i < original.size(); i++)
                if(!alive[i]) { // a blanket edge (updateInputGraph, :506-517)
                    g2o::HyperGraph::Edge *e = original[i];
                    _graph->removeEdge(e);
                    if(_edgeLookup) _edgeLookup->erase(e);
                }
            std::vector<int32_t> left(spg_graph_num_vertices(g));
            if(!left.empty()) check(spg_graph_vertex_ids(g, left.data()), "spg_graph_vertex_ids");
            std::sort(left.begin(), left.end());
            for(Vertex *v : toRemove)
                if(!std::binary_search(left.begin(), left.end(), (int32_t) v->id())) _graph->removeVertex(v); // (:520-523)
            for(int i = 0; i < ne; i++) {
                spg_edge_desc d;
                check(spg_graph_edge_desc(g, i, &d), "spg_graph_edge_desc");
                if(d.uid_major < 0) continue;
                Edge *e = mirrorEdgeOut(g, i, d, dim);
                _graph->addEdge(e); // (:526-541)
                added.push_back(e);
            }
            spg_graph_destroy(g);
            if(st == SPG_ERR_BLANKET_FAILED) _lastError = spg_last_error();
        } catch(...) {
            spg_graph_destroy(g);
            throw;
        }
        return added;
    }

    const std::string &lastError() const { return _lastError; } // non-empty: blankets failed and stayed in the graph

private:
    int graphDim() const {
        for(const auto &iv : _graph->vertices())
            return static_cast<const Vertex *>(iv.second)->estimateDimension() == 3 ? 3 : 6; // VertexSE2: 3, VertexSE3: 7
        return 3;
    }

    void mirrorEdgeIn(spg_graph *g, g2o::HyperGraph::Edge *he, int dim) {
        const int P = dim == 3 ? 3 : 7;
        std::vector<int32_t> ids;
        for(const g2o::HyperGraph::Vertex *v : he->vertices()) ids.push_back(v->id());
        if(const GLCEdge *ge = dynamic_cast<const GLCEdge *>(he)) {
            // GLC factor: measurement (d * nv) and the linear weight W, row-major (rows = error dimension)
            const g2o::MatrixXD &W = ge->linearWeight();
            const g2o::VectorXD &z = ge->measurement();
            std::vector<double> meas(z.size()), w((size_t) W.rows() * W.cols());
            for(int i = 0; i < (int) z.size(); i++) meas[i] = z(i);
            for(int r = 0; r < (int) W.rows(); r++)
                for(int c = 0; c < (int) W.cols(); c++) w[(size_t) r * W.cols() + c] = W(r, c);
            check(spg_graph_add_factor(g, SPG_EDGE_GLC, (int32_t) ids.size(), ids.data(), (int32_t) W.rows(), meas.data(), w.data(), NULL),
                  "spg_graph_add_factor(GLC)");
            return;
        }
        if(dim == 3) {
            if(const MultiEdgeSE2ISAM *me = dynamic_cast<const MultiEdgeSE2ISAM *>(he)) return mirrorMultiIn<EdgeSE2ISAM>(g, me, ids, dim);
        } else {
            if(const MultiEdgeSE3ISAM *me = dynamic_cast<const MultiEdgeSE3ISAM *>(he)) return mirrorMultiIn<EdgeSE3ISAM>(g, me, ids, dim);
        }
        // relative-pose edge (EdgeSE2ISAM / EdgeSE3ISAM and their g2o bases)
        const Edge *e = static_cast<const Edge *>(he);
        if(ids.size() != 2 || e->dimension() != dim) fail("unsupported factor type in the graph");
        double meas[7];
        e->getMeasurementData(meas); // x y theta | t, q (x y z w)
        (void) P;
        check(spg_graph_add_edge(g, ids[0], ids[1], meas, e->informationData()), "spg_graph_add_edge"); // column-major d x d
    }

    template <class EdgeType, class Multi>
    void mirrorMultiIn(spg_graph *g, const Multi *me, const std::vector<int32_t> &ids, int dim) {
        const int P = dim == 3 ? 3 : 7;
        std::vector<double> meas;
        std::vector<int32_t> pairs;
        for(typename Multi::const_iterator it = me->begin(); it != me->end(); ++it) {
            const g2o::HyperGraph::VertexContainer vs = (*it).vertices();
            for(const g2o::HyperGraph::Vertex *v : vs)
                pairs.push_back((int32_t) (std::find(ids.begin(), ids.end(), (int32_t) v->id()) - ids.begin()));
            EdgeType tmp;
            tmp.setMeasurement((*it).measurement());
            double z[7];
            tmp.getMeasurementData(z);
            meas.insert(meas.end(), z, z + P);
        }
        const int rows = me->dimension();
        check(spg_graph_add_factor(g, SPG_EDGE_MULTI, (int32_t) ids.size(), ids.data(), rows, meas.data(), me->informationData(),
                                   pairs.data()),
              "spg_graph_add_factor(MULTI)");
    }

    Edge *mirrorEdgeOut(spg_graph *g, int idx, const spg_edge_desc &d, int dim) {
        const int P = dim == 3 ? 3 : 7;
        std::vector<int32_t> ids(d.nv);
        if(d.kind == SPG_EDGE_POSE) { // the edge TopologyProviderSE2ISAM / SE3ISAM creates (topology_provider_binary.hpp:43-47)
            double meas[7], info[36];
            check(spg_graph_edge_data(g, idx, ids.data(), meas, info), "spg_graph_edge_data");
            Edge *e = dim == 3 ? static_cast<Edge *>(new EdgeSE2ISAM) : static_cast<Edge *>(new EdgeSE3ISAM);
            e->setVertex(0, _graph->vertex(ids[0]));
            e->setVertex(1, _graph->vertex(ids[1]));
            e->setMeasurementData(meas);
            std::copy(info, info + dim * dim, e->informationData()); // column-major (:531-533)
            return e;
        }
        if(d.kind == SPG_EDGE_GLC) { // TopologyProviderGLC::getEdge (src/topology_provider_glc.cpp:78-98)
            const int cols = dim * d.nv;
            std::vector<double> meas(cols), w((size_t) d.rows * cols);
            check(spg_graph_edge_data(g, idx, ids.data(), meas.data(), w.data()), "spg_graph_edge_data");
            GLCEdge *e = new GLCEdge;
            e->setReparam(dim == 3 ? static_cast<GLCReparam *>(new GLCReparamSE2ISAM) : static_cast<GLCReparam *>(new GLCReparamSE3));
            e->setDimension(d.rows, cols);
            g2o::MatrixXD W(d.rows, cols);
            for(int r = 0; r < d.rows; r++)
                for(int c = 0; c < cols; c++) W(r, c) = w[(size_t) r * cols + c];
            e->setLinearWeight(W);
            g2o::HyperGraph::VertexContainer vc; // (setDimension sized the container by columns: replaced, as the reference does)
            for(int i = 0; i < d.nv; i++) vc.push_back(_graph->vertex(ids[i]));
            e->vertices() = vc;
            g2o::VectorXD z(cols); // = reparam->reparametrize(vc) at the current estimates (computeMeasurement)
            for(int i = 0; i < cols; i++) z(i) = meas[i];
            e->setMeasurement(z);
            e->information().setIdentity();
            return e;
        }
        // MultiEdgeCorrelated (CliqueySubgraph / CliqueyDense, src/multi_edge_correlated.hpp:29-62)
        const int nm = d.rows / dim;
        std::vector<double> meas((size_t) nm * P), info((size_t) d.rows * d.rows);
        std::vector<int32_t> pairs(2 * nm);
        check(spg_graph_edge_data(g, idx, ids.data(), meas.data(), info.data()), "spg_graph_edge_data");
        check(spg_graph_edge_pairs(g, idx, pairs.data()), "spg_graph_edge_pairs");
        if(dim == 3) return multiOut<EdgeSE2ISAM, MultiEdgeSE2ISAM>(ids, pairs, meas, info, nm, P, d.rows);
        return multiOut<EdgeSE3ISAM, MultiEdgeSE3ISAM>(ids, pairs, meas, info, nm, P, d.rows);
    }

    template <class EdgeType, class Multi>
    Edge *multiOut(const std::vector<int32_t> &ids, const std::vector<int32_t> &pairs, const std::vector<double> &meas,
                   const std::vector<double> &info, int nm, int P, int rows) {
        Multi *e = new Multi;
        e->setMeasurementCount(nm);
        for(int m = 0; m < nm; m++) {
            g2o::HyperGraph::VertexContainer vs;
            vs.push_back(_graph->vertex(ids[pairs[2 * m]]));
            vs.push_back(_graph->vertex(ids[pairs[2 * m + 1]]));
            EdgeType tmp;
            tmp.setMeasurementData(meas.data() + (size_t) m * P);
            e->addMeasurement(vs, tmp.measurement());
        }
        std::copy(info.begin(), info.begin() + (size_t) rows * rows, e->informationData());
        return e;
    }

    void check(spg_status st, const char *what) const {
        if(st != SPG_OK) fail(what);
    }
    [[noreturn]] void fail(const char *what) const {
        throw std::runtime_error(std::string(what) + ": " + spg_last_error());
    }

    spg_ctx *_ctx;
    SparsityOptions _opts;
    g2o::OptimizableGraph *_graph;
    EdgeMap *_edgeLookup;
    std::list<TopologyProvider *> _topologies;
    bool _useGLC;
    Engine _engine;
    spg_marginalize_stats _stats{};
    std::string _lastError;
};

#endif /* SPG_VERTEX_REMOVER_G2O_HPP_ */
