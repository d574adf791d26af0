// ORACLE — TEST INFRASTRUCTURE ONLY (see dense.hpp header). PARITY UNPINNED.
// Graph-level restatement: a minimal multigraph standing in for g2o::OptimizableGraph, the
// sequential VertexRemover::remove loop (reference src/vertex_remover.cpp:83-251,285-392,500-557),
// decimation schedules (src/decimation.cpp:11-49), g2o text reader
// (src/graph_wrapper_g2o.cpp:107-147) and computeSubstituteEdge (src/compute_substitute_edge.cpp:13-96).
//
// Canonical order: the reference keeps blanket edges in std::set<Edge*> ordered by heap address
// (vertex_remover.h:24), i.e. its H summation order is allocation dependent. Here every edge
// carries a key (major, minor): file edges (-1, file index); edges created while removing
// which[i] get (i, provider order). Blanket edges are always visited in key order.
#pragma once
#include <deque>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include "blanket.hpp"

namespace orc {

struct GVertex;
struct GEdge {
    BEdge body;                   // body.v unused at graph level
    std::vector<GVertex *> verts;
    long keyMajor = -1, keyMinor = 0;
};
struct EdgeKeyLess {
    bool operator()(const GEdge *a, const GEdge *b) const {
        if(a->keyMajor != b->keyMajor) return a->keyMajor < b->keyMajor;
        return a->keyMinor < b->keyMinor;
    }
};
typedef std::set<GEdge *, EdgeKeyLess> EdgeSet;

struct GVertex {
    int id = 0;
    Pose est;
    EdgeSet edges;
};
struct VertexIdLess {
    bool operator()(const GVertex *a, const GVertex *b) const { return a->id < b->id; }
};
typedef std::set<GVertex *, VertexIdLess> VertexSet;

struct Graph {
    int dim = 3;
    std::map<int, GVertex *> verts;
    EdgeSet edges;
    long fileEdges = 0;

    ~Graph() {
        for(auto e : edges) delete e;
        for(auto &v : verts) delete v.second;
    }
    GVertex *vertex(int id) const {
        auto it = verts.find(id);
        return it == verts.end() ? nullptr : it->second;
    }
    GVertex *addVertex(int id, const Pose &est) {
        GVertex *v = new GVertex;
        v->id = id;
        v->est = est;
        verts[id] = v;
        return v;
    }
    void addEdge(GEdge *e) {
        edges.insert(e);
        for(GVertex *v : e->verts) v->edges.insert(e);
    }
    GEdge *addPoseEdge(int from, int to, const Pose &meas, const Mat &info) {
        GEdge *e = new GEdge;
        e->body.kind = EDGE_POSE;
        e->body.meas = meas;
        e->body.info = info;
        e->verts = {vertex(from), vertex(to)};
        e->keyMajor = -1;
        e->keyMinor = fileEdges++;
        addEdge(e);
        return e;
    }
    void removeEdge(GEdge *e) {
        for(GVertex *v : e->verts) v->edges.erase(e);
        edges.erase(e);
        delete e;
    }
    void removeVertex(GVertex *v) {
        assert(v->edges.empty());
        verts.erase(v->id);
        delete v;
    }
};

// ---- g2o text reader ---------------------------------------------------------------------
static inline Graph *loadG2o(const std::string &path) {
    std::ifstream f(path);
    if(!f) return nullptr;
    Graph *g = new Graph;
    bool dimSet = false;
    std::string line;
    struct PendingEdge { int a, b; Pose z; Mat info; };
    std::vector<PendingEdge> pend;
    while(std::getline(f, line)) {
        std::istringstream is(line);
        std::string tag;
        if(!(is >> tag)) continue;
        if(tag == "VERTEX_SE2") {
            int id; double x, y, th;
            is >> id >> x >> y >> th;
            if(!dimSet) { g->dim = 3; dimSet = true; }
            g->addVertex(id, Pose::se2(x, y, th));
        } else if(tag == "VERTEX_SE3:QUAT") {
            int id; double v[7];
            is >> id;
            for(int i = 0; i < 7; i++) is >> v[i];
            if(!dimSet) { g->dim = 6; dimSet = true; }
            g->addVertex(id, se3FromQT(v));
        } else if(tag == "EDGE_SE2") {
            PendingEdge pe;
            double x, y, th;
            is >> pe.a >> pe.b >> x >> y >> th;
            pe.z = Pose::se2(x, y, th);
            pe.info = Mat(3, 3);
            for(int i = 0; i < 3; i++)
                for(int j = i; j < 3; j++) {
                    is >> pe.info(i, j);
                    pe.info(j, i) = pe.info(i, j);
                }
            pend.push_back(pe);
        } else if(tag == "EDGE_SE3:QUAT") {
            PendingEdge pe;
            double v[7];
            is >> pe.a >> pe.b;
            for(int i = 0; i < 7; i++) is >> v[i];
            pe.z = se3FromQT(v); // g2o normalises the quaternion on read
            pe.info = Mat(6, 6);
            for(int i = 0; i < 6; i++)
                for(int j = i; j < 6; j++) {
                    is >> pe.info(i, j);
                    pe.info(j, i) = pe.info(i, j);
                }
            pend.push_back(pe);
        }
    }
    for(auto &pe : pend) g->addPoseEdge(pe.a, pe.b, pe.z, pe.info);
    return g;
}

// ---- decimation.cpp:11-49 ----------------------------------------------------------------
static inline std::vector<int> clusterDecimate(int last, int endvert, int sparsity, int clusterSize) {
    if(((last - 4) % clusterSize == 0 && last > 4) || last == endvert) {
        std::vector<int> ret;
        for(int i = int(std::ceil((last - 5) / (double) clusterSize) - 1) * clusterSize + 5; i <= last; i++)
            if(i % sparsity > 0) ret.push_back(i);
        return ret;
    }
    return std::vector<int>();
}
static inline std::vector<int> onlineDecimate(int last, int, int sparsity) {
    if(last % sparsity == 0) return std::vector<int>();
    return std::vector<int>({last});
}
static inline std::vector<int> globalDecimate(int last, int endvert, int sparsity) {
    if(last == endvert) {
        std::vector<int> which;
        for(int i = 4; i <= endvert; i++)
            if(i % sparsity != 0) which.push_back(i);
        return which;
    }
    return std::vector<int>();
}

// ---- VertexRemover -------------------------------------------------------------------------
struct RemoveLogEntry {
    int rootId;
    std::vector<int> blanketIds;     // removed first
    int nRemoved;
    int status;
    int newtonIters;
    double kld;
    std::vector<std::pair<int, int>> pattern; // new binary edges (original ids) in creation order
};

class VertexRemover {
public:
    Graph *graph = nullptr;
    SparsityOptions opts;
    int algorithm = ALG_NFR;
    std::vector<RemoveLogEntry> log;
    bool keepLog = true;
    int unsupportedLocal = 0;

    // vertex_remover.cpp:197-215
    VertexSet markovBlanketVertices(GVertex *root) const {
        VertexSet vset;
        vset.insert(root);
        for(GEdge *e : root->edges)
            for(GVertex *v : e->verts) vset.insert(v);
        return vset;
    }
    // :142-195 (the live #else branch :185-191): the id-ordered set grows while it is iterated
    VertexSet extendedMarkovBlanketVertices(GVertex *root, const std::set<GVertex *> &pickBin, VertexSet &picked) const {
        picked = VertexSet();
        VertexSet ret = markovBlanketVertices(root);
        picked.insert(root);
        for(auto v : ret) {
            if(pickBin.count(v) > 0 && picked.count(v) == 0) {
                picked.insert(v);
                VertexSet other = markovBlanketVertices(v);
                ret.insert(other.begin(), other.end());
            }
        }
        return ret;
    }
    // :225-251
    EdgeSet markovBlanketEdges(const VertexSet &mb, const VertexSet &hubs) const {
        EdgeSet edges;
        for(GVertex *vertex : mb)
            for(GEdge *edge : vertex->edges) {
                bool is_markov = true, found_hub = false;
                for(GVertex *v : edge->verts) {
                    if(mb.count(v) == 0) { is_markov = false; break; }
                    if(hubs.count(v) > 0) found_hub = true;
                }
                if(is_markov && (opts.includeIntraClique || found_hub)) edges.insert(edge);
            }
        return edges;
    }

    // :83-140
    void remove(const std::vector<int> &whichIds) {
        std::vector<GVertex *> toRemove;
        for(int id : whichIds) toRemove.push_back(graph->vertex(id));
        std::set<GVertex *> toRemoveSet(toRemove.begin(), toRemove.end()), deleted;
        for(size_t i = 0; i < toRemove.size(); i++) {
            VertexSet vmarkov, toRemoveNow;
            if(deleted.count(toRemove[i]) > 0) continue;
            if(opts.topology == SparsityOptions::Dense || opts.topology == SparsityOptions::CliqueyDense) {
                vmarkov = extendedMarkovBlanketVertices(toRemove[i], toRemoveSet, toRemoveNow);
            } else {
                vmarkov = markovBlanketVertices(toRemove[i]);
                toRemoveNow.insert(toRemove[i]);
            }
            EdgeSet emarkov = markovBlanketEdges(vmarkov, toRemoveNow);

            Blanket b;
            bool supported = buildSubgraph(toRemoveNow, vmarkov, emarkov, b);
            BlanketResult res;
            if(supported) res = processBlanket(b, opts, algorithm);
            else { res.status = ST_UNSUPPORTED; unsupportedLocal++; }

            if(keepLog) {
                RemoveLogEntry le;
                le.rootId = toRemove[i]->id;
                le.blanketIds = b.ids;
                le.nRemoved = b.nRemoved;
                le.status = res.status;
                le.newtonIters = res.nfr.newtonIters;
                le.kld = res.nfr.kld;
                for(const BEdge &e : res.edges)
                    if(e.v.size() == 2) le.pattern.push_back({b.ids[b.nRemoved + e.v[0]], b.ids[b.nRemoved + e.v[1]]});
                log.push_back(le);
            }
            updateInputGraph(toRemoveNow, emarkov, b, res, (long) i);
            deleted.insert(toRemoveNow.begin(), toRemoveNow.end());
        }
    }

private:
    // :285-392. The Local non-star case runs the restated Levenberg-Marquardt of blanket.hpp (g2o is third party).
    bool buildSubgraph(const VertexSet &toRemove, const VertexSet &blanketVertices, const EdgeSet &blanketEdges, Blanket &b) {
        b.dim = graph->dim;
        b.nRemoved = (int) toRemove.size();
        std::map<GVertex *, int> forward;
        int k = 0;
        for(GVertex *v : toRemove) { forward[v] = k++; b.ids.push_back(v->id); b.poses.push_back(v->est); }
        for(GVertex *v : blanketVertices)
            if(toRemove.count(v) == 0) { forward[v] = k++; b.ids.push_back(v->id); b.poses.push_back(v->est); }
        for(GEdge *e : blanketEdges) {
            BEdge be = e->body;
            be.v.clear();
            for(GVertex *v : e->verts) be.v.push_back(forward[v]);
            b.edges.push_back(be);
        }
        bool closedFormEstimate = false;
        GVertex *first = *toRemove.begin();
        if(opts.linPoint == SparsityOptions::Local) {
            std::map<GVertex *, int> nconnections;
            closedFormEstimate = true;
            for(GVertex *v : blanketVertices)
                if(v != first) nconnections[v] = 0;
            for(GEdge *e : blanketEdges)
                for(GVertex *v : e->verts)
                    if(v != first) {
                        nconnections[v]++;
                        // initialEstimatePossible: g2o pose edges return 1, GLCEdge -1
                        // (glc_edge.cpp:57-62), MultiEdgeCorrelated sums its pose edges
                        closedFormEstimate = closedFormEstimate && (e->body.kind != EDGE_GLC);
                    }
            for(auto &nc : nconnections)
                if(nc.second > 1) { closedFormEstimate = false; break; }
        }
        if(closedFormEstimate) {
            // :363-381 — removed vertex at the origin, neighbours from the measurements
            // (g2o's binary initialEstimate ignores its second argument)
            b.poses[0] = Pose::identity(b.dim);
            for(const BEdge &be : b.edges) {
                auto apply = [&](int vi, int vj, const Pose &z) {
                    if(vi == 0) b.poses[vj] = compose(b.poses[vi], z);
                    else b.poses[vi] = compose(b.poses[vj], inverse(z));
                };
                if(be.kind == EDGE_POSE) {
                    apply(be.v[0], be.v[1], be.meas);
                    apply(be.v[0], be.v[1], be.meas);
                } else if(be.kind == EDGE_MULTI) {
                    // MultiEdgeCorrelated::initialEstimate (multi_edge_correlated.hpp:142-157),
                    // called once per vertex of the edge: every measurement containing it
                    for(size_t vi = 0; vi < be.v.size(); vi++)
                        for(size_t m = 0; m < be.pairs.size(); m++)
                            if(be.pairs[m][0] == (int) vi || be.pairs[m][1] == (int) vi)
                                apply(be.v[be.pairs[m][0]], be.v[be.pairs[m][1]], be.mmeas[m]);
                }
            }
            return true;
        } else if(opts.linPoint != SparsityOptions::Global) {
            // :382-391 — Local without a closed form: 10 LM iterations on the subgraph, removed vertex fixed
            localOptimize(b, 10);
        }
        return true;
    }

    // :500-546
    void updateInputGraph(const VertexSet &toRemove, const EdgeSet &blanketEdges, const Blanket &b,
                          const BlanketResult &res, long major) {
        std::vector<GEdge *> be(blanketEdges.begin(), blanketEdges.end());
        for(GEdge *e : be) graph->removeEdge(e);
        std::vector<GVertex *> tr(toRemove.begin(), toRemove.end());
        for(GVertex *v : tr) graph->removeVertex(v);
        long minor = 0;
        for(const BEdge &ne : res.edges) {
            GEdge *e = new GEdge;
            e->body = ne;
            for(int vi : ne.v) e->verts.push_back(graph->vertex(b.ids[b.nRemoved + vi]));
            e->body.v.clear();
            e->keyMajor = major;
            e->keyMinor = minor++;
            graph->addEdge(e);
        }
    }
};

// ---- computeSubstituteEdge, compute_substitute_edge.cpp:13-96 ------------------------------
// Only binary POSE edges carry measurement()/information() in the façade.
static inline Mat invertSmall(const Mat &A) { return luInverse(A); } // Eigen MatrixXd::inverse() = PartialPivLU
static inline void computeSubstituteEdge(const Graph *gw, const std::set<int> &marginalized, int maxid,
                                         int &from, int &to, Pose &edgemeas, Mat &edgeinfo) {
    std::set<int> visited;
    std::deque<std::set<int>> frontiers;
    std::set<int> newFrontier;
    int minid = std::numeric_limits<int>::max();
    int toConnect = std::max(from, to);
    int toReplace = std::min(from, to);
    newFrontier.insert(toReplace);
    visited.insert(toConnect);
    visited.insert(toReplace);
    do {
        frontiers.push_back(newFrontier);
        newFrontier.clear();
        for(int r : frontiers.back()) {
            if(marginalized.count(r) == 0 && r != from && r != to) {
                minid = std::min(minid, r);
            } else {
                visited.insert(r);
                for(GEdge *e : gw->vertex(r)->edges) {
                    if(e->verts.size() == 2) {
                        int idother = e->verts[0]->id == r ? e->verts[1]->id : e->verts[0]->id;
                        if(visited.count(idother) == 0 && idother <= maxid && idother != 0) newFrontier.insert(idother);
                    }
                }
            }
        }
    } while(minid == std::numeric_limits<int>::max());
    visited.clear();
    visited.insert(toConnect);
    frontiers.push_front(visited);
    frontiers.pop_back();

    int dim = gw->dim;
    Mat covsum(dim, dim);
    Pose meas = Pose::identity(dim);
    int reach = minid;
    while(!frontiers.empty()) {
        std::set<int> lastFrontier = frontiers.back();
        frontiers.pop_back();
        for(GEdge *e : gw->vertex(reach)->edges) {
            if(e->verts.size() != 2) continue;
            int id0 = e->verts[0]->id, id1 = e->verts[1]->id;
            if(lastFrontier.count(id0) || lastFrontier.count(id1)) {
                covsum = covsum + invertSmall(e->body.info);
                if(from == toConnect) {
                    if(id1 == reach) meas = compose(e->body.meas, meas);
                    else meas = compose(inverse(e->body.meas), meas);
                } else {
                    if(id1 == reach) meas = compose(meas, inverse(e->body.meas));
                    else meas = compose(meas, e->body.meas);
                }
                reach = id1 == reach ? id0 : id1;
                break;
            }
        }
    }
    edgeinfo = invertSmall(covsum);
    edgeinfo = 0.5 * (edgeinfo + edgeinfo.transpose());
    edgemeas = meas;
    if(from == toConnect) to = minid;
    else from = minid;
}

} // namespace orc
