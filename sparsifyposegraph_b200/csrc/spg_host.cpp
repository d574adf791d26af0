// spg_host.cpp — host side of the node-removal path: pose-graph container, g2o text reader,
// decimation schedules, wavefront round scheduler + packer (VertexRemover) and
// computeSubstituteEdge. See spg_host.h for the reference files each piece mirrors.
// All numerical work of a removal happens on the GPU through spg_remove_round(); this file only
// moves integers, copies doubles and composes a few poses (Local linearisation point, R2).
#include "spg_host.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <condition_variable>
#include <functional>
#include <limits>
#include <mutex>
#include <sstream>

#include <cuda_runtime.h>

#include "../../include/spg_record.h"
#include "spg_ctx.h"

namespace spg {

// ---- small pose helpers -----------------------------------------------------------------------
static double normalizeTheta(double theta) {
    if(theta >= -M_PI && theta < M_PI) return theta;
    double multiplier = std::floor(theta / (2 * M_PI));
    theta = theta - multiplier * 2 * M_PI;
    if(theta >= M_PI) theta -= 2 * M_PI;
    if(theta < -M_PI) theta += 2 * M_PI;
    return theta;
}
static void qmul(const double *a, const double *b, double *o) { // x y z w
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
static void qrot(const double *q, const double *v, double *o) {
    // o = R(q) v
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * (y * v[2] - z * v[1]), ty = 2 * (z * v[0] - x * v[2]), tz = 2 * (x * v[1] - y * v[0]);
    o[0] = v[0] + w * tx + (y * tz - z * ty);
    o[1] = v[1] + w * ty + (z * tx - x * tz);
    o[2] = v[2] + w * tz + (x * ty - y * tx);
}
static void qnorm(double *q) {
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double s = (q[3] < 0 ? -1.0 : 1.0) / n;
    for(int i = 0; i < 4; i++) q[i] *= s;
}
void poseCompose(int dim, const double *a, const double *b, double *out) {
    if(dim == 3) {
        const double c = std::cos(a[2]), s = std::sin(a[2]);
        double x = a[0] + c * b[0] - s * b[1], y = a[1] + s * b[0] + c * b[1];
        out[0] = x; out[1] = y; out[2] = normalizeTheta(a[2] + b[2]);
    } else {
        double qa[4] = {a[3], a[4], a[5], a[6]}, qb[4] = {b[3], b[4], b[5], b[6]};
        qnorm(qa); qnorm(qb);
        double t[3], q[4];
        qrot(qa, b, t);
        qmul(qa, qb, q);
        qnorm(q);
        out[0] = a[0] + t[0]; out[1] = a[1] + t[1]; out[2] = a[2] + t[2];
        out[3] = q[0]; out[4] = q[1]; out[5] = q[2]; out[6] = q[3];
    }
}
void poseInverse(int dim, const double *a, double *out) {
    if(dim == 3) {
        const double th = normalizeTheta(-a[2]);
        const double c = std::cos(th), s = std::sin(th);
        double x = c * (-a[0]) - s * (-a[1]), y = s * (-a[0]) + c * (-a[1]);
        out[0] = x; out[1] = y; out[2] = th;
    } else {
        double q[4] = {-a[3], -a[4], -a[5], a[6]};
        qnorm(q);
        double mt[3] = {-a[0], -a[1], -a[2]}, t[3];
        qrot(q, mt, t);
        out[0] = t[0]; out[1] = t[1]; out[2] = t[2];
        out[3] = q[0]; out[4] = q[1]; out[5] = q[2]; out[6] = q[3];
    }
}
static void poseIdentity(int dim, double *out) {
    for(int i = 0; i < 7; i++) out[i] = 0;
    if(dim == 6) out[6] = 1;
}

// ---- decimation (src/decimation.cpp:11-49) ---------------------------------------------------------
std::vector<int> clusterDecimate(int last, int endvert, const DecimateOptions &opts) {
    if(((last - 4) % opts.clusterSize == 0 && last > 4) || last == endvert) {
        std::vector<int> ret;
        for(int i = int(std::ceil((last - 5) / (double) opts.clusterSize) - 1) * opts.clusterSize + 5; i <= last; i++)
            if(i % opts.sparsity > 0) ret.push_back(i);
        return ret;
    }
    return std::vector<int>();
}
std::vector<int> onlineDecimate(int last, int, const DecimateOptions &opts) {
    if(last % opts.sparsity == 0) return std::vector<int>();
    return std::vector<int>({last});
}
std::vector<int> globalDecimate(int last, int endvert, const DecimateOptions &opts) {
    if(last == endvert) {
        std::vector<int> which;
        for(int i = 4; i <= endvert; i++)
            if(i % opts.sparsity != 0) which.push_back(i);
        return which;
    }
    return std::vector<int>();
}

// ---- Graph -------------------------------------------------------------------------------------------
bool Graph::hasVertex(int id) const {
    const int i = indexOf(id);
    return i >= 0 && verts[i].alive;
}
GraphVertex *Graph::vertex(int id) {
    const int i = indexOf(id);
    return (i < 0 || !verts[i].alive) ? nullptr : &verts[i];
}
const GraphVertex *Graph::vertex(int id) const {
    const int i = indexOf(id);
    return (i < 0 || !verts[i].alive) ? nullptr : &verts[i];
}
bool Graph::addVertex(int id, const double *pose) {
    if(hasVertex(id)) return false;
    GraphVertex v;
    v.id = id;
    std::memcpy(v.pose, pose, sizeof(double) * poseWords());
    if(dim == 3) v.pose[2] = normalizeTheta(v.pose[2]);
    else qnorm(v.pose + 3);
    if((unsigned) id < (1u << 26)) {
        if((size_t) id >= dense.size()) dense.resize(std::max<size_t>((size_t) id + 1, dense.size() * 2), -1);
        dense[id] = (int) verts.size();
    } else {
        sparse[id] = (int) verts.size();
    }
    verts.push_back(v);
    aliveVertices++;
    return true;
}
int Graph::addEdge(GraphEdge &&e) {
    const int ei = (int) edges.size();
    const int n = (int) e.v.size();
    e.alive = true;
    e.vx0 = n > 0 ? indexOf(e.v[0]) : -1;
    e.vx1 = n > 1 ? indexOf(e.v[1]) : -1;
    e.vxn.clear();
    if(n > 2)
        for(int id : e.v) e.vxn.push_back(indexOf(id));
    for(int q = 0; q < n; q++) {
        GraphVertex &gv = verts[e.vx(q)];
        gv.edges.push_back(ei);
        gv.peer.push_back(n == 2 ? ((e.vx(1 - q) << 1) | (q == 0)) : -1);
    }
    edges.push_back(std::move(e));
    aliveEdges++;
    version++;
    return ei;
}
int Graph::addPoseEdge(int from, int to, const double *meas, const double *info) {
    if(!hasVertex(from) || !hasVertex(to)) return -1;
    GraphEdge e;
    e.kind = SPG_EDGE_POSE;
    e.v = {from, to};
    e.rows = dim;
    e.setPayload(meas, poseWords(), info, dim * dim);
    if(dim == 3) e.meas()[2] = normalizeTheta(e.meas()[2]);
    else qnorm(e.meas() + 3);
    e.uidMajor = -1;
    e.uidMinor = fileEdges++;
    return addEdge(std::move(e));
}
void Graph::removeEdge(int ei) {
    GraphEdge &e = edges[ei];
    if(!e.alive) return;
    e.alive = false;
    for(int q = 0, n = e.nv(); q < n; q++) {
        GraphVertex &gv = verts[e.vx(q)];
        std::vector<int> &ve = gv.edges;
        auto it = std::find(ve.begin(), ve.end(), ei);
        if(it != ve.end()) {
            gv.peer[it - ve.begin()] = gv.peer.back();
            gv.peer.pop_back();
            *it = ve.back();
            ve.pop_back();
        }
    }
    std::vector<double>().swap(e.payload);
    aliveEdges--;
    version++;
}
void Graph::removeVertex(int id) {
    GraphVertex *v = vertex(id);
    if(!v) return;
    // g2o's HyperGraph::removeVertex detaches every edge still incident to the vertex
    while(!v->edges.empty()) removeEdge(v->edges.back());
    v->alive = false;
    aliveVertices--;
}
int Graph::maxVertexId() const {
    int m = -1;
    for(const GraphVertex &v : verts)
        if(v.alive) m = std::max(m, v.id);
    return m;
}
std::vector<int> Graph::vertexIds() const {
    std::vector<int> ids;
    for(const GraphVertex &v : verts)
        if(v.alive) ids.push_back(v.id);
    std::sort(ids.begin(), ids.end());
    return ids;
}
std::vector<int> Graph::edgeOrder() const {
    std::vector<int> o;
    for(size_t i = 0; i < edges.size(); i++)
        if(edges[i].alive) o.push_back((int) i);
    std::sort(o.begin(), o.end(), [&](int a, int b) {
        if(edges[a].uidMajor != edges[b].uidMajor) return edges[a].uidMajor < edges[b].uidMajor;
        return edges[a].uidMinor < edges[b].uidMinor;
    });
    return o;
}

// g2o text I/O. Reader: the four tags of the reference's datasets (graph_wrapper_g2o.cpp:107-147) plus the factor types a
// sparsified graph carries when the reference saves it (GraphWrapperG2O::write -> g2o save, tags registered in
// src/edge_types.cpp:68-76):
//   EDGE_SE2_ISAM a b  x y th  <upper info>                                  (EdgeSE2ISAM, se2_compatibility.h)
//   GLC_EDGE v0 v1 .. || GLC_REPARAM_{SE2_ISAM,SE2,SE3} rows cols  meas[cols]  W[rows x cols, row-major]  <upper I_rows>
//                                                                            (GLCEdge::write, src/glc_edge.cpp:94-118)
//   MULTI_EDGE_{SE2,SE2_ISAM,SE3,SE3_ISAM} v0 v1 .. || nmeas nrelevant  (meas tokens) x nmeas  <upper info, d*nmeas>
//                                                                            (MultiEdgeCorrelated::write, multi_edge_correlated.hpp:226-267)
// The reference's multi-edge writer does not store which vertices each measurement connects (_mappings), so its own
// reader cannot restore them either; the writer here puts them on a comment line in front of the edge
// ("#SPG_MULTI_PAIRS a0 b0 a1 b1 ..", indices into the edge's vertex list), which stock g2o skips and this reader uses.
static bool readUpper(std::istream &is, int n, double *full) {
    for(int i = 0; i < n; i++)
        for(int j = i; j < n; j++) {
            double v;
            if(!(is >> v)) return false;
            full[i + j * n] = v;
            full[j + i * n] = v;
        }
    return true;
}

Graph *Graph::loadG2o(const std::string &path, std::string *err) {
    std::ifstream f(path);
    if(!f) {
        if(err) *err = "cannot open " + path;
        return nullptr;
    }
    Graph *g = nullptr;
    std::vector<GraphEdge> pend; // edges may precede their vertices in a file
    std::vector<int> pendingPairs;
    std::string line;
    int lineNo = 0;
    auto fail = [&](const std::string &what) -> Graph * {
        if(err) *err = path + ":" + std::to_string(lineNo) + ": " + what;
        delete g;
        return nullptr;
    };
    while(std::getline(f, line)) {
        lineNo++;
        std::istringstream is(line);
        std::string tag;
        if(!(is >> tag)) continue;
        if(tag == "#SPG_MULTI_PAIRS") {
            pendingPairs.clear();
            int x;
            while(is >> x) pendingPairs.push_back(x);
            continue;
        }
        if(tag[0] == '#') continue;
        if(tag == "VERTEX_SE2" || tag == "VERTEX_SE3:QUAT") {
            int d = tag == "VERTEX_SE2" ? 3 : 6;
            if(!g) g = new Graph(d);
            if(g->dim != d) return fail("vertex dimension differs from the graph's");
            int id;
            double p[7] = {0, 0, 0, 0, 0, 0, 1};
            is >> id;
            for(int i = 0; i < g->poseWords(); i++) is >> p[i];
            if(!is) return fail("short vertex line");
            g->addVertex(id, p);
        } else if(tag == "EDGE_SE2" || tag == "EDGE_SE2_ISAM" || tag == "EDGE_SE3:QUAT") {
            int d = tag == "EDGE_SE3:QUAT" ? 6 : 3, P = d == 3 ? 3 : 7;
            GraphEdge e;
            int a, b;
            double z[7], info[36];
            is >> a >> b;
            for(int i = 0; i < P; i++) is >> z[i];
            if(!is || !readUpper(is, d, info)) return fail("short edge line");
            e.kind = SPG_EDGE_POSE;
            e.v = {a, b};
            e.rows = d;
            e.setPayload(z, P, info, d * d);
            pend.push_back(std::move(e));
        } else if(tag == "GLC_EDGE" || tag.rfind("MULTI_EDGE_", 0) == 0) {
            GraphEdge e;
            std::string tok;
            while(is >> tok && tok != "||") e.v.push_back(atoi(tok.c_str()));
            if(tok != "||" || e.v.empty()) return fail("multi-vertex edge without the '||' separator");
            if(tag == "GLC_EDGE") {
                std::string reparam;
                int rows = 0, cols = 0;
                is >> reparam >> rows >> cols;
                const int d = reparam == "GLC_REPARAM_SE3" ? 6 : 3;
                if(!is || reparam.rfind("GLC_REPARAM_", 0) != 0 || rows < 0 || cols != d * (int) e.v.size())
                    return fail("bad GLC_EDGE header");
                e.kind = SPG_EDGE_GLC;
                e.rows = rows;
                e.allocPayload(cols, rows * cols);
                for(int i = 0; i < cols; i++) is >> e.meas()[i];
                for(int i = 0; i < rows * cols; i++) is >> e.info()[i];
                std::vector<double> I((size_t) rows * rows);
                if(!is || !readUpper(is, rows, I.data())) return fail("short GLC_EDGE line");
                e.pairs.assign(1, d); // remembers the dimension until the graph exists
            } else {
                const int d = tag.find("SE3") != std::string::npos ? 6 : 3, P = d == 3 ? 3 : 7;
                int nmeas = 0, nrel = 0;
                is >> nmeas >> nrel;
                if(!is || nmeas < 1 || nrel != P) return fail("bad MULTI_EDGE header");
                if((int) pendingPairs.size() != 2 * nmeas)
                    return fail("MULTI_EDGE without its #SPG_MULTI_PAIRS line: the reference's format does not store which "
                                "vertices a measurement connects");
                e.kind = SPG_EDGE_MULTI;
                e.rows = d * nmeas;
                e.allocPayload(nmeas * P, e.rows * e.rows);
                for(int i = 0; i < nmeas * P; i++) is >> e.meas()[i];
                if(!is || !readUpper(is, e.rows, e.info())) return fail("short MULTI_EDGE line");
                e.pairs = pendingPairs;
                for(int x : e.pairs)
                    if(x < 0 || x >= (int) e.v.size()) return fail("#SPG_MULTI_PAIRS index out of range");
                e.pairs.push_back(-d); // dimension marker, removed below
                pendingPairs.clear();
            }
            pend.push_back(std::move(e));
        }
        // FIX, PARAMS_* and unknown tags are skipped like g2o's loader skips unregistered ones
    }
    if(!g) {
        if(err) *err = "no vertices in " + path;
        return nullptr;
    }
    lineNo = 0;
    for(GraphEdge &e : pend) {
        int d = g->dim;
        if(e.kind == SPG_EDGE_POSE) d = e.rows;
        else if(e.kind == SPG_EDGE_GLC) { d = e.pairs[0]; e.pairs.clear(); }
        else { d = -e.pairs.back(); e.pairs.pop_back(); }
        if(d != g->dim) return fail("edge dimension differs from the graph's");
        for(int id : e.v)
            if(!g->hasVertex(id)) return fail("edge refers to unknown vertex " + std::to_string(id));
        if(e.kind == SPG_EDGE_POSE) {
            if(g->dim == 3) e.meas()[2] = normalizeTheta(e.meas()[2]);
            else qnorm(e.meas() + 3);
        }
        e.uidMajor = -1;
        e.uidMinor = g->fileEdges++;
        g->addEdge(std::move(e));
    }
    return g;
}

// Writer: what GraphWrapperG2O::write (-> g2o's OptimizableGraph::save) emits for a graph of this wrapper — vertices in
// ascending id, then edges in the canonical order, numbers with 17 significant digits (g2o's default stream precision
// of 6 would not survive a round trip at the 1e-9 parity bar).
bool Graph::saveG2o(const std::string &path, std::string *err) const {
    std::ofstream f(path);
    if(!f) {
        if(err) *err = "cannot open " + path + " for writing";
        return false;
    }
    f.precision(17);
    const int P = poseWords();
    for(int id : vertexIds()) {
        const GraphVertex *v = vertex(id);
        f << (dim == 3 ? "VERTEX_SE2 " : "VERTEX_SE3:QUAT ") << id;
        for(int i = 0; i < P; i++) f << ' ' << v->pose[i];
        f << '\n';
    }
    auto upper = [&](const double *M, int n) {
        for(int i = 0; i < n; i++)
            for(int j = i; j < n; j++) f << ' ' << M[i + j * n];
    };
    for(int ei : edgeOrder()) {
        const GraphEdge &e = edges[ei];
        if(e.kind == SPG_EDGE_POSE) {
            f << (dim == 3 ? "EDGE_SE2_ISAM " : "EDGE_SE3:QUAT ") << e.v[0] << ' ' << e.v[1];
            for(int i = 0; i < P; i++) f << ' ' << e.meas()[i];
            upper(e.info(), dim);
        } else if(e.kind == SPG_EDGE_GLC) {
            f << "GLC_EDGE";
            for(int id : e.v) f << ' ' << id;
            const int cols = dim * e.nv();
            f << " || " << (dim == 3 ? "GLC_REPARAM_SE2_ISAM " : "GLC_REPARAM_SE3 ") << e.rows << ' ' << cols;
            for(int i = 0; i < cols; i++) f << ' ' << e.meas()[i];
            for(int i = 0; i < e.rows * cols; i++) f << ' ' << e.info()[i];
            for(int i = 0; i < e.rows; i++)
                for(int j = i; j < e.rows; j++) f << ' ' << (i == j ? 1 : 0);
        } else {
            const int nmeas = e.rows / dim;
            f << "#SPG_MULTI_PAIRS";
            for(int x : e.pairs) f << ' ' << x;
            f << '\n' << (dim == 3 ? "MULTI_EDGE_SE2_ISAM" : "MULTI_EDGE_SE3_ISAM");
            for(int id : e.v) f << ' ' << id;
            f << " || " << nmeas << ' ' << P;
            for(int i = 0; i < nmeas * P; i++) f << ' ' << e.meas()[i];
            upper(e.info(), e.rows);
        }
        f << '\n';
    }
    f.flush();
    if(!f) {
        if(err) *err = "write error on " + path;
        return false;
    }
    return true;
}

// ---- VertexRemover --------------------------------------------------------------------------------
VertexRemover::VertexRemover() {}
VertexRemover::~VertexRemover() {
    for(TopologyProvider *t : _topologies) delete t;
}
void VertexRemover::setSparsityOptions(const SparsityOptions &opts) {
    _opts = opts;
    for(TopologyProvider *t : _topologies) t->setSparsityOptions(opts);
}

// src/vertex_remover.cpp:197-215
std::vector<int> VertexRemover::markovBlanketVertices(int root) const {
    std::set<int> vset;
    vset.insert(root);
    const GraphVertex *v = _graph->vertex(root);
    if(v)
        for(int ei : v->edges)
            for(int id : _graph->edges[ei].v) vset.insert(id);
    return std::vector<int>(vset.begin(), vset.end());
}
// :142-195 (live branch :185-191): the id-ordered set grows while it is being iterated, so the closure
// is transitive only towards larger ids
std::vector<int> VertexRemover::extendedMarkovBlanketVertices(int root, const std::set<int> &pickBin,
                                                               std::vector<int> &picked) const {
    std::set<int> pk, ret;
    for(int id : markovBlanketVertices(root)) ret.insert(id);
    pk.insert(root);
    for(auto it = ret.begin(); it != ret.end(); ++it) {
        int v = *it;
        if(pickBin.count(v) > 0 && pk.count(v) == 0) {
            pk.insert(v);
            for(int id : markovBlanketVertices(v)) ret.insert(id);
        }
    }
    picked.assign(pk.begin(), pk.end());
    return std::vector<int>(ret.begin(), ret.end());
}
// :225-251
std::vector<int> VertexRemover::markovBlanketEdges(const std::vector<int> &mb, const std::vector<int> &hubs) const {
    std::set<int> inb(mb.begin(), mb.end()), hub(hubs.begin(), hubs.end());
    std::set<int> es;
    for(int id : mb) {
        const GraphVertex *v = _graph->vertex(id);
        if(!v) continue;
        for(int ei : v->edges) {
            const GraphEdge &e = _graph->edges[ei];
            bool is_markov = true, found_hub = false;
            for(int x : e.v) {
                if(inb.count(x) == 0) { is_markov = false; break; }
                if(hub.count(x) > 0) found_hub = true;
            }
            if(is_markov && (_opts.includeIntraClique || found_hub)) es.insert(ei);
        }
    }
    std::vector<int> out(es.begin(), es.end());
    std::sort(out.begin(), out.end(), [&](int a, int b) {
        const GraphEdge &x = _graph->edges[a], &y = _graph->edges[b];
        if(x.uidMajor != y.uidMajor) return x.uidMajor < y.uidMajor;
        return x.uidMinor < y.uidMinor;
    });
    return out;
}

bool VertexRemover::buildUnit(int root, int listIndex, const std::set<int> &toRemoveSet, RemovalUnit &u) const {
    u.removed.clear(); u.kept.clear(); u.edges.clear(); u.ridx.clear(); u.kidx.clear(); // keep the capacity
    u.listIndex = listIndex;
    std::vector<int> vmarkov;
    if(_opts.topology == SparsityOptions::Dense || _opts.topology == SparsityOptions::CliqueyDense) {
        vmarkov = extendedMarkovBlanketVertices(root, toRemoveSet, u.removed);
    } else {
        // single removal (Tree / Subgraph / CliqueySubgraph): the same sets as markovBlanketVertices / -Edges, walked
        // by vertex index with a per-thread stamp array for membership — this runs once per pending vertex and round
        static thread_local std::vector<unsigned> mark;
        static thread_local unsigned token = 0;
        static thread_local std::vector<std::pair<int, int>> mb;                    // (id, vertex index) of the blanket
        static thread_local std::vector<std::pair<unsigned long long, int>> keyed;  // (canonical order key, edge)
        const std::vector<GraphVertex> &V = _graph->verts;
        const ChunkedVector<GraphEdge> &E = _graph->edges;
        if(mark.size() < V.size()) mark.resize(V.size() + V.size() / 4 + 64, 0);
        if(++token == 0) { std::fill(mark.begin(), mark.end(), 0u); token = 1; }
        const int ri = _graph->indexOf(root);
        mb.clear();
        keyed.clear();
        if(ri >= 0 && V[ri].alive) {
            mark[ri] = token;
            mb.emplace_back(root, ri);
            const GraphVertex &rv = V[ri];
            for(size_t k = 0; k < rv.edges.size(); k++) {
                const int p = rv.peer[k];
                if(p >= 0) {
                    const int xi = p >> 1;
                    if(mark[xi] != token) { mark[xi] = token; mb.emplace_back(V[xi].id, xi); }
                    continue;
                }
                const GraphEdge &e = E[rv.edges[k]];
                for(int q = 0, n = e.nv(); q < n; q++) {
                    const int xi = e.vx(q);
                    if(mark[xi] != token) { mark[xi] = token; mb.emplace_back(V[xi].id, xi); }
                }
            }
        } else {
            mb.emplace_back(root, -1);
        }
        // the second walk reads the adjacency arrays of every blanket vertex (heap blocks): start fetching them now
        for(const auto &pr : mb)
            if(pr.second >= 0) {
                __builtin_prefetch(V[pr.second].edges.data());
                __builtin_prefetch(V[pr.second].peer.data());
            }
        std::sort(mb.begin(), mb.end());
        u.removed.push_back(root);
        for(const auto &pr : mb)
            if(pr.first != root) { u.kept.push_back(pr.first); u.kidx.push_back(pr.second); }
        if(ri >= 0) u.ridx.push_back(ri);
        for(const auto &pr : mb) {
            const int xi = pr.second;
            if(xi < 0 || !V[xi].alive) continue;
            const GraphVertex &gv = V[xi];
            for(size_t k = 0; k < gv.edges.size(); k++) {
                const int ei = gv.edges[k], p = gv.peer[k];
                if(p >= 0) { // two-vertex edge: decided from the adjacency entry alone
                    if(!(p & 1)) continue; // met once per endpoint: taken at its first one
                    const int yi = p >> 1;
                    if(mark[yi] != token) continue;
                    if(_opts.includeIntraClique || xi == ri || yi == ri) {
                        __builtin_prefetch(&E[ei]);
                        keyed.emplace_back(~0ULL, ei); // key filled in below, once the edge has arrived
                    }
                    continue;
                }
                const GraphEdge &e = E[ei];
                const int n = e.nv();
                if(e.vx(0) != xi) continue; // every blanket edge is met once per vertex: take it at its first one
                bool is_markov = true, found_hub = false;
                for(int q = 0; q < n; q++) {
                    const int yi = e.vx(q);
                    if(mark[yi] != token) { is_markov = false; break; }
                    if(yi == ri) found_hub = true;
                }
                if(is_markov && (_opts.includeIntraClique || found_hub)) keyed.emplace_back(e.uidKey(), ei);
            }
        }
        for(auto &ke : keyed)
            if(ke.first == ~0ULL) ke.first = E[ke.second].uidKey();
        std::sort(keyed.begin(), keyed.end());
        keyed.erase(std::unique(keyed.begin(), keyed.end()), keyed.end());
        for(const auto &ke : keyed) u.edges.push_back(ke.second);
        return !u.edges.empty();
    }
    std::set<int> rem(u.removed.begin(), u.removed.end());
    for(int id : vmarkov)
        if(!rem.count(id)) u.kept.push_back(id);
    u.edges = markovBlanketEdges(vmarkov, u.removed);
    for(int id : u.removed) u.ridx.push_back(_graph->indexOf(id));
    for(int id : u.kept) u.kidx.push_back(_graph->indexOf(id));
    return !u.edges.empty();
}

// src/vertex_remover.cpp:452-463: first registered provider that is applicable
TopologyProvider *VertexRemover::chooseTopologyProvider(const RemovalUnit &u) const {
    std::set<int> kinds;
    for(int ei : u.edges) kinds.insert(_graph->edges[ei].kind);
    for(TopologyProvider *tp : _topologies)
        if(tp->applicable(_graph->dim, kinds)) return tp;
    return nullptr;
}

// per-thread map vertex index -> local index inside the blanket being packed (entries of other blankets are stale
// but never read: every endpoint of a blanket edge is a blanket vertex)
static thread_local std::vector<int> t_local;

// Local linearisation point: does the blanket lack the closed-form estimate (:304-342)? True when a kept vertex
// appears in more than one blanket edge or a GLC factor is involved (GLCEdge::initialEstimatePossible == -1).
bool VertexRemover::needsSubgraphOptimisation(const RemovalUnit &u) const {
    if(_opts.linPoint != SparsityOptions::Local) return false;
    const int root = u.ridx.empty() ? -1 : u.ridx[0];
    std::vector<std::pair<int, int>> nconn; // (vertex index, count): blankets are small
    for(int ei : u.edges) {
        const GraphEdge &e = _graph->edges[ei];
        for(int q = 0, n = e.nv(); q < n; q++) {
            const int xi = e.vx(q);
            if(xi == root) continue;
            if(e.kind == SPG_EDGE_GLC) return true;
            auto it = std::find_if(nconn.begin(), nconn.end(), [&](const std::pair<int, int> &pr) { return pr.first == xi; });
            if(it == nconn.end()) nconn.emplace_back(xi, 1);
            else if(++it->second > 1) return true;
        }
    }
    return false;
}

// vertex_remover.cpp:382-391: the blanket subgraph optimised for 10 Levenberg-Marquardt iterations with the (first)
// removed vertex fixed — spg_graph_optimize on a copy of the blanket, i.e. on the GPU like every other numerical step.
// poses: nv x P in record order (removed first, kept ascending).
spg_status VertexRemover::localLinearise(const RemovalUnit &u, std::vector<double> &poses) const {
    const int P = _graph->poseWords();
    spg_graph sub;
    sub.g = new Graph(_graph->dim);
    std::vector<int> order;
    for(size_t i = 0; i < u.ridx.size(); i++) order.push_back(u.ridx[i]);
    for(size_t i = 0; i < u.kidx.size(); i++) order.push_back(u.kidx[i]);
    for(int xi : order) sub.g->addVertex(_graph->verts[xi].id, _graph->verts[xi].pose);
    for(int ei : u.edges) {
        GraphEdge e = _graph->edges[ei]; // copy (ids, payload, pairs, canonical key)
        sub.g->addEdge(std::move(e));
    }
    const int32_t fixed = u.removed[0];
    spg_status st = spg_graph_optimize(_ctx, &sub, &fixed, 1, 10, nullptr);
    if(st == SPG_OK) {
        poses.resize(order.size() * (size_t) P);
        for(size_t i = 0; i < order.size(); i++)
            std::memcpy(&poses[i * P], sub.g->vertex(_graph->verts[order[i]].id)->pose, sizeof(double) * P);
    }
    delete sub.g;
    sub.g = nullptr;
    return st;
}

// record size of a unit, 0 when the blanket needs the Local non-star linearisation point
int64_t VertexRemover::unitWords(const RemovalUnit &u) const {
    const int dim = _graph->dim;
    const int nv = (int) (u.removed.size() + u.kept.size()), ne = (int) u.edges.size();
    int64_t words = spgr_record_fixed_words(dim, nv, ne);
    for(int ei : u.edges) {
        const GraphEdge &e = _graph->edges[ei];
        words += spgr_edge_words(dim, e.kind, e.nv(), e.rows);
    }
    return (words + 1) & ~(int64_t) 1;
}

// buildSubgraph (src/vertex_remover.cpp:285-392) + record packing into rec[0, words). Returns false when the
// blanket needs the Local non-star linearisation point (g2o LM on the subgraph; not part of this path).
bool VertexRemover::packUnit(const RemovalUnit &u, uint64_t *rec, int64_t words, const double *linPoses) const {
    const int dim = _graph->dim, P = _graph->poseWords();
    const int nrem = (int) u.removed.size(), nv = nrem + (int) u.kept.size(), ne = (int) u.edges.size();
    const std::vector<GraphVertex> &V = _graph->verts;
    if(t_local.size() < V.size()) t_local.resize(V.size() + V.size() / 4 + 64, -1);
    // every word of the record is written exactly once below (no memset of the whole record first: the records of a
    // round are gigabytes); the padding slots of the int32 tables and the tail are zeroed explicitly
    // (tests: SPG_POISON_RECORDS=1 fills the record with 0xFF first — the result must not change)
    static const bool poison = getenv("SPG_POISON_RECORDS") != nullptr;
    if(poison) std::memset(rec, 0xFF, (size_t) words * 8);
    int32_t *h = reinterpret_cast<int32_t *>(rec);
    h[0] = nv; h[1] = nrem; h[2] = ne; h[3] = dim; h[4] = (int32_t) words; h[5] = 0; h[6] = u.listIndex; h[7] = 0;
    int32_t *rid = reinterpret_cast<int32_t *>(rec + spgr_ids_off());
    if(nv & 1) rid[nv] = 0;
    double *poses = reinterpret_cast<double *>(rec + spgr_poses_off(nv));
    for(int i = 0; i < nv; i++) {
        const int xi = i < nrem ? u.ridx[i] : u.kidx[i - nrem];
        t_local[xi] = i;
        rid[i] = V[xi].id;
        std::memcpy(poses + (size_t) i * P, V[xi].pose, sizeof(double) * P);
    }

    if(linPoses) {
        // Local linearisation point of a non-star blanket: the subgraph was optimised in planRound (:382-391)
        std::memcpy(poses, linPoses, sizeof(double) * nv * P);
    } else if(_opts.linPoint == SparsityOptions::Local) {
        // closed-form estimate only for star-shaped blankets (:304-342)
        bool closedForm = true;
        std::vector<int> nconn(nv, 0);
        for(int ei : u.edges) {
            const GraphEdge &e = _graph->edges[ei];
            for(int q = 0, n = e.nv(); q < n; q++) {
                const int li = t_local[e.vx(q)];
                if(li != 0) {
                    nconn[li]++;
                    closedForm = closedForm && (e.kind != SPG_EDGE_GLC); // GLCEdge::initialEstimatePossible == -1
                }
            }
        }
        for(int i = 1; i < nv; i++)
            if(nconn[i] > 1) { closedForm = false; break; }
        if(!closedForm) return false;
        // :363-381 — removed vertex at the origin, each neighbour from its measurement
        poseIdentity(dim, &poses[0]);
        double tmp[7], inv[7];
        auto apply = [&](int vi, int vj, const double *z) {
            if(vi == 0) {
                poseCompose(dim, &poses[(size_t) vi * P], z, tmp);
                std::memcpy(&poses[(size_t) vj * P], tmp, sizeof(double) * P);
            } else {
                poseInverse(dim, z, inv);
                poseCompose(dim, &poses[(size_t) vj * P], inv, tmp);
                std::memcpy(&poses[(size_t) vi * P], tmp, sizeof(double) * P);
            }
        };
        for(int ei : u.edges) {
            const GraphEdge &e = _graph->edges[ei];
            if(e.kind == SPG_EDGE_POSE) {
                apply(t_local[e.vx(0)], t_local[e.vx(1)], e.meas());
            } else if(e.kind == SPG_EDGE_MULTI) {
                for(int vi = 0; vi < e.nv(); vi++)
                    for(size_t m = 0; m < e.pairs.size() / 2; m++)
                        if(e.pairs[2 * m] == vi || e.pairs[2 * m + 1] == vi)
                            apply(t_local[e.vx(e.pairs[2 * m])], t_local[e.vx(e.pairs[2 * m + 1])], e.meas() + m * P);
            }
        }
    }

    int32_t *etab = reinterpret_cast<int32_t *>(rec + spgr_edgetab_off(dim, nv));
    if(ne & 1) etab[ne] = 0;
    int64_t eoff = spgr_record_fixed_words(dim, nv, ne);
    // an edge costs two dependent cache misses (the edge, then its payload block on the heap): fetch the edges four
    // ahead and the payloads two ahead of the copy
    for(int i = 0; i < ne && i < 4; i++) __builtin_prefetch(&_graph->edges[u.edges[i]]);
    for(int i = 0; i < ne && i < 2; i++) {
        const char *pp = reinterpret_cast<const char *>(_graph->edges[u.edges[i]].meas());
        for(int l = 0; l < 6; l++) __builtin_prefetch(pp + 64 * l);
    }
    for(int i = 0; i < ne; i++) {
        if(i + 4 < ne) __builtin_prefetch(&_graph->edges[u.edges[i + 4]]);
        if(i + 2 < ne) {
            const char *pp = reinterpret_cast<const char *>(_graph->edges[u.edges[i + 2]].meas());
            for(int l = 0; l < 6; l++) __builtin_prefetch(pp + 64 * l);
        }
        const GraphEdge &e = _graph->edges[u.edges[i]];
        etab[i] = (int32_t) eoff;
        uint64_t *ew = rec + eoff;
        int32_t *eh = reinterpret_cast<int32_t *>(ew);
        const int nve = e.nv();
        eh[0] = e.kind; eh[1] = nve; eh[2] = e.rows; eh[3] = 0;
        int32_t *vi = reinterpret_cast<int32_t *>(ew + 2);
        for(int q = 0; q < nve; q++) vi[q] = t_local[e.vx(q)];
        if(nve & 1) vi[nve] = 0;
        double *pl = reinterpret_cast<double *>(ew + 2 + spgr_pad2(nve));
        if(e.kind == SPG_EDGE_POSE) {
            std::memcpy(pl, e.meas(), sizeof(double) * P);
            std::memcpy(pl + P, e.info(), sizeof(double) * dim * dim);
        } else if(e.kind == SPG_EDGE_GLC) {
            std::memcpy(pl, e.meas(), sizeof(double) * dim * nve);
            std::memcpy(pl + dim * nve, e.info(), sizeof(double) * e.rows * dim * nve);
        } else {
            const int nm = e.rows / dim;
            int32_t *pr = reinterpret_cast<int32_t *>(ew + 2 + spgr_pad2(nve));
            for(int q = 0; q < 2 * nm; q++) pr[q] = e.pairs[q];
            double *pm = reinterpret_cast<double *>(ew + 2 + spgr_pad2(nve) + spgr_pad2(2 * nm));
            std::memcpy(pm, e.meas(), sizeof(double) * nm * P);
            std::memcpy(pm + (size_t) nm * P, e.info(), sizeof(double) * e.rows * e.rows);
        }
        eoff += spgr_edge_words(dim, e.kind, nve, e.rows);
    }
    for(int64_t w = eoff; w < words; w++) rec[w] = 0; // (the record is rounded up to an even number of words)
    return true;
}

// ---- host buffers of a round: page-locked when a CUDA device is there (the copies of spg_remove_round then overlap
// the kernels), plain memory otherwise (CPU tests drive the planner without a GPU) -------------------------------------
uint64_t *HostBuf::reserve(size_t words) {
    if(words <= cap) return p;
    release();
    const size_t want = words + words / 4 + 1024;
    void *q = nullptr;
    static const bool have_gpu = [] { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess && n > 0; }();
    if(have_gpu && cudaHostAlloc(&q, want * 8, cudaHostAllocDefault) == cudaSuccess) pinned = true;
    else {
        (void) cudaGetLastError();
        q = std::malloc(want * 8);
        pinned = false;
    }
    p = static_cast<uint64_t *>(q);
    cap = p ? want : 0;
    return p;
}
void HostBuf::release() {
    if(p) {
        if(pinned) cudaFreeHost(p);
        else std::free(p);
    }
    p = nullptr;
    cap = 0;
}

// Persistent worker pool of the host scheduler: the parallel passes of a removal (staleness test, blanket extraction,
// packing, splicing) run dozens of times per round, so threads are created once — thread creation and the first
// touch of the per-thread scratch arrays would otherwise cost more than the passes themselves.
namespace {
class WorkerPool {
public:
    explicit WorkerPool(unsigned n) {
        for(unsigned t = 0; t < n; t++) _threads.emplace_back([this] { loop(); });
    }
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(_mx);
            _stop = true;
        }
        _cv.notify_all();
        for(auto &th : _threads) th.join();
    }
    unsigned size() const { return (unsigned) _threads.size(); }
    // runs job(k) for k in [0, parts) on the workers and the calling thread; returns when all are done
    void run(unsigned parts, const std::function<void(unsigned)> &job) {
        std::unique_lock<std::mutex> lk(_mx);
        _job = &job;
        _next = 0;
        _parts = parts;
        _pending = parts;
        _gen++;
        lk.unlock();
        _cv.notify_all();
        for(;;) { // the caller works too
            lk.lock();
            if(_next >= _parts) break;
            const unsigned k = _next++;
            lk.unlock();
            job(k);
            lk.lock();
            _pending--;
            lk.unlock();
        }
        _done.wait(lk, [this] { return _pending == 0; });
        _job = nullptr;
    }
private:
    void loop() {
        std::unique_lock<std::mutex> lk(_mx);
        unsigned long long seen = 0;
        for(;;) {
            _cv.wait(lk, [&] { return _stop || (_gen != seen && _job && _next < _parts); });
            if(_stop) return;
            seen = _gen;
            while(_job && _next < _parts) {
                const unsigned k = _next++;
                const std::function<void(unsigned)> *job = _job;
                lk.unlock();
                (*job)(k);
                lk.lock();
                if(--_pending == 0) _done.notify_all();
            }
        }
    }
    std::vector<std::thread> _threads;
    std::mutex _mx;
    std::condition_variable _cv, _done;
    const std::function<void(unsigned)> *_job = nullptr;
    unsigned _next = 0, _parts = 0, _pending = 0;
    unsigned long long _gen = 0;
    bool _stop = false;
};
} // namespace

// run fn(begin, end) over [0, n) on up to 16 host threads (one call on this thread when n is small)
template <class F>
static void parallelFor(size_t n, size_t min_per_thread, F fn) {
    static const unsigned max_thr = [] {
        const char *e = getenv("SPG_HOST_THREADS");
        return e ? (unsigned) std::max(1, atoi(e)) : std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    }();
    const unsigned nthr = (unsigned) std::min<size_t>(max_thr, std::max<size_t>(1, n / std::max<size_t>(1, min_per_thread)));
    if(nthr <= 1) {
        fn((size_t) 0, n);
        return;
    }
    static WorkerPool pool(max_thr - 1); // + the calling thread
    static std::mutex one_at_a_time;     // removals on different graphs may run from several caller threads
    std::lock_guard<std::mutex> guard(one_at_a_time);
    const size_t chunk = (n + nthr - 1) / nthr;
    pool.run(nthr, [&](unsigned t) {
        const size_t b0 = std::min(n, (size_t) t * chunk), b1 = std::min(n, b0 + chunk);
        if(b0 < b1) fn(b0, b1);
    });
}

// plain threads, not the worker pool: a remover may be destroyed when the pool is gone already (process exit)
void UnitCache::reset(size_t n) {
    auto spread = [](size_t count, auto fn) {
        const unsigned nthr = count < 65536 ? 1u : std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
        if(nthr == 1) { fn((size_t) 0, count); return; }
        std::vector<std::thread> th;
        const size_t per = (count + nthr - 1) / nthr;
        for(unsigned t = 0; t < nthr; t++) {
            const size_t b0 = std::min(count, (size_t) t * per), b1 = std::min(count, b0 + per);
            if(b0 < b1) th.emplace_back(fn, b0, b1);
        }
        for(auto &x : th) x.join();
    };
    if(_p) {
        RemovalUnit *p = _p;
        spread(_n, [p](size_t b0, size_t b1) { for(size_t i = b0; i < b1; i++) p[i].~RemovalUnit(); });
        ::operator delete(static_cast<void *>(_p));
        _p = nullptr;
        _n = 0;
    }
    if(n) {
        RemovalUnit *p = static_cast<RemovalUnit *>(::operator new(n * sizeof(RemovalUnit)));
        spread(n, [p](size_t b0, size_t b1) { for(size_t i = b0; i < b1; i++) new(p + i) RemovalUnit(); });
        _p = p;
        _n = n;
    }
}

namespace {
struct HostProf {
    double stale = 0, extract = 0, select = 0, pack = 0, apply = 0;
    long long n_extract = 0, n_visit = 0, n_sel = 0;
    double a1 = 0, a2 = 0, a3 = 0;
    bool on = getenv("SPG_HOST_PROF") != nullptr;
    ~HostProf() {
        if(on) fprintf(stderr, "[spg host] stale-check %.3f s, extract %.3f s, select %.3f s, pack %.3f s, splice %.3f s; %lld extractions, %lld window visits, %lld selected; splice passes %.3f %.3f %.3f\n", stale, extract, select, pack, apply, n_extract, n_visit, n_sel, a1, a2, a3);
    }
} g_prof;
inline double nowS() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
} // namespace

std::vector<int> VertexRemover::remove(int toRemove, spg_status *status) {
    return remove(std::vector<int>(1, toRemove), status);
}

// src/vertex_remover.cpp:83-140, regrouped into wavefront rounds.
//
// Two removal units U (earlier in the list) and V commute — removing them in either order, hence
// also both from one snapshot, gives the same graph — iff V's blanket is untouched by U:
//   removed(V) ∩ blanket(U) = ∅,  kept(V) ∩ removed(U) = ∅  and  |kept(U) ∩ kept(V)| <= 1
// (U deletes / creates edges only among vertices of blanket(U); with includeIntraClique an edge
// inside kept(U) ∩ kept(V) would belong to both blankets).
// A round is built by scanning the pending list in order. A unit is selected when it commutes with
// every unit already selected AND with every earlier unit that had to be deferred, using for the
// deferred ones a conservative region: their blanket plus the blankets of all units they were
// deferred behind (their neighbourhood can only grow through those). Selected units are then
// processed from the same snapshot, and spliced in list order.
std::vector<int> VertexRemover::remove(const std::vector<int> &toRemove, spg_status *status) {
    using clk = std::chrono::steady_clock;
    if(status) *status = SPG_OK;
    spg_status st = beginRemoval(toRemove);
    while(st == SPG_OK) {
        auto t0 = clk::now();
        const bool sharded = spg_comm_nranks(_ctx) > 1;
        st = planRound(sharded); // one GPU: the records are packed chunk by chunk inside the streamed call below
        if(st != SPG_OK || _round.sel.empty()) break;
        auto t1 = clk::now();
        uint64_t *out = spg_ctx_pinned(_ctx, 1, (size_t) _round.outOff.back()); // every word is written by the kernels
        if(!out) out = _outBuf.reserve((size_t) _round.outOff.back());
        if(!out) {
            error = "out of host memory";
            st = SPG_ERR_INVALID;
            break;
        }
        spg_round_in rin = roundDescriptor();
        spg_round_out rout{};
        rout.out = out;
        if(sharded) {
            // a context with a communicator (spg_comm_init) shards every round over its ranks; all ranks hold the same
            // graph, get the complete output (all-gather) and splice the same edges
            st = spg_remove_round_sharded(_ctx, &rin, &rout, -1, nullptr);
        } else {
            // the substitute edges of the round need fresh blocks of the edge store (page faults, constructors): a helper
            // thread prepares them while the GPU works; the records of every pipeline chunk are spliced as they arrive
            // (the host is idle otherwise, waiting for the device at the end of the round)
            struct Pipe {
                VertexRemover *self;
                const uint64_t *out;
                std::thread grow;
                double spliceS = 0;
                static int32_t fill(void *u, int32_t b0, int32_t b1) {
                    VertexRemover *vr = static_cast<Pipe *>(u)->self;
                    const double t = nowS();
                    const bool ok = vr->packRange((size_t) b0, (size_t) b1);
                    g_prof.pack += nowS() - t;
                    if(!ok) vr->_packFailed = true;
                    return ok ? 0 : 1;
                }
                static int32_t drain(void *u, int32_t b0, int32_t b1) {
                    Pipe *p = static_cast<Pipe *>(u);
                    const double t = nowS();
                    if(p->grow.joinable()) p->grow.join();
                    p->self->applyRange(p->out, (size_t) b0, (size_t) b1);
                    p->spliceS += nowS() - t;
                    return 0;
                }
            } pipe{this, out, {}, 0};
            _packFailed = false;
            applyBegin();
            static const bool no_drain = getenv("SPG_NO_CHUNKED_SPLICE") != nullptr;
            pipe.grow = std::thread([this] {
                _graph->edges.reserve((size_t) _apply.e0 + (size_t) _apply.base.back(), [](GraphEdge &e) { e.alive = false; });
            });
            st = spg_remove_round_pipelined(_ctx, &rin, &rout, &Pipe::fill, no_drain ? nullptr : &Pipe::drain, &pipe);
            if(pipe.grow.joinable()) pipe.grow.join();
            if(_packFailed) { // `error` names the reason (packRange); chunks spliced before it stay spliced
                if(_apply.spliced) applyEnd();
                st = SPG_ERR_UNSUPPORTED;
                break;
            }
            if(st != SPG_OK) {
                error = spg_last_error();
                if(_apply.spliced) applyEnd();
                break;
            }
            auto t2 = clk::now();
            if(_apply.spliced < _round.sel.size()) applyRange(out, _apply.spliced, _round.sel.size());
            applyEnd();
            auto t3 = clk::now();
            stats.pack_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
            stats.gpu_ms += std::chrono::duration<double, std::milli>(t2 - t1).count() - 1e3 * pipe.spliceS;
            stats.splice_ms += std::chrono::duration<double, std::milli>(t3 - t2).count() + 1e3 * pipe.spliceS;
            continue;
        }
        if(st != SPG_OK) {
            error = spg_last_error();
            break;
        }
        auto t2 = clk::now();
        applyRound(out);
        auto t3 = clk::now();
        stats.pack_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
        stats.gpu_ms += std::chrono::duration<double, std::milli>(t2 - t1).count();
        stats.splice_ms += std::chrono::duration<double, std::milli>(t3 - t2).count();
    }
    if(st == SPG_OK) st = failureStatus();
    if(status) *status = st;
    return _added;
}

// SPG_ERR_BLANKET_FAILED once any blanket of the call came back with a status != OK (it stays in the graph)
spg_status VertexRemover::failureStatus() {
    if(stats.n_failed == 0) return SPG_OK;
    error = "blanket of list entry " + std::to_string(stats.first_failed_index) + " (vertex " +
            std::to_string(_pending[stats.first_failed_index]) + ") failed with blanket status " +
            std::to_string(stats.first_failed_status) + "; " + std::to_string(stats.n_failed) +
            " failed blanket(s) were left in the graph untouched";
    return SPG_ERR_BLANKET_FAILED;
}

spg_status VertexRemover::beginRemoval(const std::vector<int> &toRemove) {
    stats = spg_marginalize_stats{};
    stats.first_failed_index = -1;
    _added.clear();
    // option combinations the path cannot serve are refused before the graph is touched
    for(TopologyProvider *tp : _topologies)
        if(tp->algorithm() == SPG_ALG_GLC &&
           (!(_opts.topology == SparsityOptions::Dense || _opts.topology == SparsityOptions::Tree) ||
            _opts.linPoint != SparsityOptions::Global)) {
            // asserts of TopologyProviderGLC::topology (src/topology_provider_glc.cpp:107-111)
            error = "GLC needs Dense|Tree topology and the Global linearisation point";
            return SPG_ERR_UNSUPPORTED;
        }
    _pending = toRemove;
    _toRemoveSet.clear(); // only the extended (Dense) blankets consult it
    if(_opts.topology == SparsityOptions::Dense || _opts.topology == SparsityOptions::CliqueyDense)
        _toRemoveSet = std::set<int>(toRemove.begin(), toRemove.end());
    _cursor = 0;
    _leftover.clear();
    _window = 0;
    if(const char *e = getenv("SPG_PLAN_WINDOW")) _window = (size_t) std::max(1LL, atoll(e));
    _done.assign(toRemove.size(), 0);
    _remaining = toRemove.size();
    _round.sel.clear();
    _round.recOff.assign(1, 0);
    _round.outOff.assign(1, 0);
    _unitCache.reset(toRemove.size());
    _unitBuilt.assign(toRemove.size(), 0);
    _rootIdx.assign(toRemove.size(), -1);
    _extended = _opts.topology == SparsityOptions::Dense || _opts.topology == SparsityOptions::CliqueyDense;
    _dupInList.assign(toRemove.size(), 0);
    {
        std::vector<unsigned char> seen(_graph->verts.size(), 0);
        for(size_t i = 0; i < toRemove.size(); i++) {
            _rootIdx[i] = _graph->indexOf(toRemove[i]);
            if(_rootIdx[i] >= 0) {
                if(seen[_rootIdx[i]]) _dupInList[i] = 1;
                seen[_rootIdx[i]] = 1;
            }
        }
    }
    _stamp.assign(_graph->verts.size(), 0);
    _vtouch.assign(_graph->verts.size(), VTouch());
    _planNo = 0;
    return SPG_OK;
}

spg_round_in VertexRemover::roundDescriptor() const {
    spg_round_in rin{};
    rin.dim = _graph->dim;
    rin.algorithm = _round.algorithm;
    rin.opts.topology = _opts.topology;
    rin.opts.lin_point = _opts.linPoint;
    rin.opts.chord_ratio = _opts.chordRatio;
    rin.opts.include_intra_clique = _opts.includeIntraClique;
    rin.opts.flags = _opts.flags | (_round.poseOnly ? SPG_OPT_POSE_EDGES_ONLY : 0);
    rin.n_blankets = (int32_t) _round.sel.size();
    rin.rec_off = _round.recOff.data();
    rin.records = _round.rec;
    rin.out_off = _round.outOff.data();
    return rin;
}


// Select and pack the next wavefront round (empty round: nothing left).
//
// A unit U may run in this round iff it commutes with every earlier pending unit V (selected or deferred):
// removed(U) misses V's blanket, kept(U) misses removed(V) and U shares at most one kept vertex with V. A deferred
// unit waits for the units it hit; its blanket after they ran lies inside the union of their blankets and its own, so
// the two are merged into one component (union-find over region ids) and later units are tested against whole
// components. Only earlier list entries matter for a unit, so the scan may stop anywhere: a round is drawn from a
// WINDOW — the units deferred so far plus the next entries of the list — which keeps the per-round work (staleness
// test, blanket re-extraction, selection) proportional to the window instead of to everything still pending. Any such
// schedule ends in the same graph (the units of a round commute with each other and with everything before them).
spg_status VertexRemover::planRound(bool packNow) {
    _round.sel.clear();
    _round.recOff.assign(1, 0);
    _round.outOff.assign(1, 0);
    if(_remaining == 0) return SPG_OK;
    double tp0 = nowS();
    const int dim = _graph->dim;
    const std::vector<int> &toRemove = _pending;
    const int Vn = (int) _graph->verts.size();
    _planNo++;
    if((int) _stamp.size() < Vn) _stamp.resize(Vn, 0);
    if((int) _vtouch.size() < Vn) _vtouch.resize(Vn);

    // ---- window: the deferred units (list order) and then fresh list entries ---------------------------------------
    const size_t W = _window ? _window : std::max<size_t>(32768, _remaining / 12);
    std::vector<int> win;
    win.reserve(_leftover.size() + W);
    auto admit = [&](int i) {
        if(_done[i]) return;
        // gone already: merged into an earlier extended blanket (:91), or listed twice. (Single-removal topologies can
        // only lose a listed vertex through its own entry, i.e. a duplicate: those are flagged once in beginRemoval,
        // which spares this sequential loop a cache miss per unit on the vertex table.)
        if(_rootIdx[i] < 0 || ((_extended || _dupInList[i]) && !_graph->verts[_rootIdx[i]].alive)) {
            _done[i] = 1;
            _remaining--;
            return;
        }
        win.push_back(i);
    };
    for(int i : _leftover) admit(i);
    const size_t fresh_goal = std::max(W / 4, W > win.size() ? W - win.size() : 0);
    for(size_t fresh0 = win.size(); _cursor < toRemove.size() && win.size() - fresh0 < fresh_goal; _cursor++) admit((int) _cursor);
    if(win.empty()) return SPG_OK;

    // ---- re-extract the stale cached blankets of the window (reads the graph only: spread over the host threads) ---
    {
        std::vector<int> todo;
        std::vector<char> staleFlag(win.size(), 0);
        parallelFor(win.size(), 4096, [&](size_t b0, size_t b1) {
            for(size_t q = b0; q < b1; q++) {
                const int i = win[q];
                const RemovalUnit &u = _unitCache[i];
                bool stale = !_unitBuilt[i];
                if(!stale) {
                    for(int xi : u.ridx) if(_stamp[xi] >= _unitBuilt[i]) { stale = true; break; }
                    if(!stale) for(int xi : u.kidx) if(_stamp[xi] >= _unitBuilt[i]) { stale = true; break; }
                }
                staleFlag[q] = stale;
            }
        });
        for(size_t q = 0; q < win.size(); q++)
            if(staleFlag[q]) todo.push_back(win[q]);
        g_prof.stale += nowS() - tp0; tp0 = nowS();
        g_prof.n_extract += (long long) todo.size();
        g_prof.n_visit += (long long) win.size();
        std::atomic<int> bad(-1);
        parallelFor(todo.size(), 512, [&](size_t b0, size_t b1) {
            for(size_t q = b0; q < b1; q++) {
                const int i = todo[q];
                if(q + 2 < b1 && _rootIdx[todo[q + 2]] >= 0) __builtin_prefetch(&_graph->verts[_rootIdx[todo[q + 2]]]);
                if(q + 1 < b1 && _rootIdx[todo[q + 1]] >= 0) {
                    const GraphVertex &nr = _graph->verts[_rootIdx[todo[q + 1]]];
                    __builtin_prefetch(nr.edges.data());
                    __builtin_prefetch(nr.peer.data());
                }
                if(!buildUnit(toRemove[i], i, _toRemoveSet, _unitCache[i])) {
                    int expect = -1;
                    bad.compare_exchange_strong(expect, i);
                }
                _unitBuilt[i] = _planNo;
            }
        });
        if(bad.load() >= 0) {
            // isolated vertex: the reference asserts blanketEdges.size() > 0
            error = "vertex " + std::to_string(toRemove[bad.load()]) + " has no edges";
            return SPG_ERR_INVALID;
        }
        g_prof.extract += nowS() - tp0; tp0 = nowS();
    }

    // ---- select (sequential: the decision for a unit depends on every earlier unit of the window) -----------------
    {
        std::vector<int> &sel = _round.sel;
        _touchNodes.clear();
        std::vector<int> touched;   // vertex indices whose per-round lists must be reset
        std::vector<int> parent;    // union-find over region ids
        parent.reserve(win.size());
        auto find = [&](int x) { while(parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
        std::vector<int> hit, comps;
        std::vector<std::pair<int, int>> shared;
        _leftover.clear();
        for(size_t wq = 0; wq < win.size(); wq++) {
            const int i = win[wq];
            // the loop is a chain of dependent cache misses (unit -> per-vertex touch heads): fetch the unit eight entries
            // ahead and the touch heads three ahead
            if(wq + 8 < win.size()) { // (the index lists are stored inside the unit: four cache lines)
                const char *pu = reinterpret_cast<const char *>(&_unitCache[win[wq + 8]]);
                for(size_t l = 0; l < sizeof(RemovalUnit); l += 64) __builtin_prefetch(pu + l);
            }
            if(wq + 3 < win.size()) {
                const RemovalUnit &n1 = _unitCache[win[wq + 3]];
                for(int xi : n1.ridx) __builtin_prefetch(&_vtouch[xi]);
                for(int xi : n1.kidx) __builtin_prefetch(&_vtouch[xi]);
            }
            const RemovalUnit &u = _unitCache[i];
            hit.clear();
            shared.clear();
            auto note = [&](int c) { if(std::find(hit.begin(), hit.end(), c) == hit.end()) hit.push_back(c); };
            for(int xi : u.ridx) {
                for(int t = _vtouch[xi].head; t >= 0; t = _touchNodes[t].next) note(find(_touchNodes[t].region));
            }
            for(int xi : u.kidx) {
                if(_vtouch[xi].removedBy >= 0) note(find(_vtouch[xi].removedBy));
                comps.clear();
                for(int t = _vtouch[xi].head; t >= 0; t = _touchNodes[t].next) {
                    const int c = find(_touchNodes[t].region);
                    if(std::find(comps.begin(), comps.end(), c) == comps.end()) comps.push_back(c);
                }
                for(int c : comps) {
                    auto it = std::find_if(shared.begin(), shared.end(), [&](const std::pair<int, int> &pr) { return pr.first == c; });
                    if(it == shared.end()) shared.emplace_back(c, 1);
                    else if(++it->second >= 2) note(c);
                }
            }
            const int rid = (int) parent.size();
            parent.push_back(rid);
            const bool select = hit.empty();
            for(int c : hit) parent[find(c)] = rid;
            auto reg = [&](int xi, bool removed) {
                VTouch &vt = _vtouch[xi];
                if(vt.head < 0 && vt.removedBy < 0) touched.push_back(xi);
                _touchNodes.push_back(TouchNode{vt.head, rid});
                vt.head = (int) _touchNodes.size() - 1;
                if(removed) vt.removedBy = rid;
            };
            for(int xi : u.ridx) reg(xi, true);
            for(int xi : u.kidx) reg(xi, false);
            if(select) sel.push_back(i);
            else _leftover.push_back(i);
        }
        for(int xi : touched) _vtouch[xi] = VTouch();
        g_prof.select += nowS() - tp0; tp0 = nowS();
        if(sel.empty()) return SPG_OK;
    }

    // ---- Local linearisation point of non-star blankets: optimise the blanket subgraph (GPU), one blanket at a time ---
    const std::vector<int> &sel = _round.sel;
    const size_t ns = sel.size();
    _round.linPoses.clear();
    if(_opts.linPoint == SparsityOptions::Local) {
        _round.linPoses.resize(ns);
        for(size_t q = 0; q < ns; q++) {
            const RemovalUnit &u = _unitCache[sel[q]];
            if(!needsSubgraphOptimisation(u)) continue;
            if(!_ctx) {
                error = "Local linearisation point on a non-star blanket needs the subgraph optimiser: give the remover a context";
                return SPG_ERR_UNSUPPORTED;
            }
            const spg_status ls = localLinearise(u, _round.linPoses[q]);
            if(ls != SPG_OK) {
                error = std::string("subgraph optimisation of the blanket of vertex ") + std::to_string(u.removed[0]) + " failed: " + spg_last_error();
                return ls;
            }
            stats.n_local_optimised++;
        }
    }
    // ---- pack: sizes and providers per unit, prefix sums, then every record written in place by the host threads ---
    TopologyProvider *byMask[8] = {};
    bool haveMask[8] = {};
    for(int m = 1; m < 8; m++) {
        std::set<int> kinds;
        for(int k = 0; k < 3; k++) if(m & (1 << k)) kinds.insert(k);
        for(TopologyProvider *tp : _topologies)
            if(tp->applicable(dim, kinds)) { byMask[m] = tp; break; }
        haveMask[m] = true;
    }
    std::vector<int64_t> &recOff = _round.recOff, &outOff = _round.outOff;
    recOff.assign(ns + 1, 0);
    outOff.assign(ns + 1, 0);
    std::vector<unsigned char> mask(ns, 0);
    parallelFor(ns, 2048, [&](size_t b0, size_t b1) {
        for(size_t q = b0; q < b1; q++) {
            const RemovalUnit &u = _unitCache[sel[q]];
            unsigned char m = 0;
            for(int ei : u.edges) m |= (unsigned char) (1u << _graph->edges[ei].kind);
            mask[q] = m;
            recOff[q + 1] = unitWords(u);
        }
    });
    unsigned char allMask = 0;
    int64_t maxNew = 0; // substitute edges of the round at most
    int maxv = stats.max_blanket_vertices;
    TopologyProvider *tp = nullptr;
    for(size_t q = 0; q < ns; q++) {
        allMask |= mask[q];
        TopologyProvider *t = byMask[mask[q] & 7];
        if(!t) {
            error = "No valid topology provider for Markov blanket";
            return SPG_ERR_UNSUPPORTED;
        }
        if(tp && t->algorithm() != tp->algorithm()) {
            error = "mixed providers within one round";
            return SPG_ERR_UNSUPPORTED;
        }
        tp = t;
        const RemovalUnit &u = _unitCache[sel[q]];
        maxv = std::max<int>(maxv, (int) (u.removed.size() + u.kept.size()));
        outOff[q + 1] = outOff[q] + spgr_out_record_words(dim, t->algorithm(), _opts.topology, _opts.chordRatio, (int) u.kept.size());
        maxNew += spgr_out_edge_count(t->algorithm(), _opts.topology, _opts.chordRatio, (int) u.kept.size());
        recOff[q + 1] += recOff[q];
    }
    stats.max_blanket_vertices = maxv;
    _round.maxNewEdges = maxNew;
    uint64_t *rec = _ctx ? spg_ctx_pinned(_ctx, 0, (size_t) recOff[ns]) : nullptr; // the context keeps its staging buffers
    if(!rec) rec = _round.records.reserve((size_t) recOff[ns]);
    _round.rec = rec;
    if(!rec) {
        error = "out of host memory";
        return SPG_ERR_INVALID;
    }
    if(packNow && !packRange(0, ns)) return SPG_ERR_UNSUPPORTED;
    g_prof.pack += nowS() - tp0;
    _round.algorithm = tp->algorithm();
    _round.poseOnly = (allMask == (1u << SPG_EDGE_POSE));
    if(_round.algorithm == SPG_ALG_GLC) {
        // asserts of TopologyProviderGLC::topology (src/topology_provider_glc.cpp:107-111)
        if(!(_opts.topology == SparsityOptions::Dense || _opts.topology == SparsityOptions::Tree) ||
           _opts.linPoint != SparsityOptions::Global) {
            error = "GLC needs Dense|Tree topology and the Global linearisation point";
            return SPG_ERR_UNSUPPORTED;
        }
    }
    return SPG_OK;
}

// records of the selected units [q0, q1) of the planned round, written in place by the host threads
bool VertexRemover::packRange(size_t q0, size_t q1) {
    const std::vector<int> &sel = _round.sel;
    const std::vector<int64_t> &recOff = _round.recOff;
    uint64_t *rec = _round.rec;
    std::atomic<int> badPack(-1);
    parallelFor(q1 - q0, 1024, [&](size_t b0, size_t b1) {
        for(size_t q = q0 + b0; q < q0 + b1; q++) {
            if(q + 1 < q0 + b1) { // the next unit's vertices and first edges, while this one is copied
                const RemovalUnit &nx = _unitCache[sel[q + 1]];
                for(int xi : nx.ridx) __builtin_prefetch(&_graph->verts[xi]);
                for(int xi : nx.kidx) __builtin_prefetch(&_graph->verts[xi]);
                for(size_t i = 0; i < nx.edges.size() && i < 4; i++) __builtin_prefetch(&_graph->edges[nx.edges[i]]);
            }
            if(!packUnit(_unitCache[sel[q]], rec + recOff[q], recOff[q + 1] - recOff[q],
                         (q < _round.linPoses.size() && !_round.linPoses[q].empty()) ? _round.linPoses[q].data() : nullptr)) {
                int expect = -1;
                badPack.compare_exchange_strong(expect, (int) q);
            }
        }
    });
    if(badPack.load() >= 0) {
        error = "Local linearisation point on a non-star blanket needs the subgraph optimiser (not on this path)";
        return false;
    }
    return true;
}

// Splice the output records of the planned round, in list order (updateInputGraph,
// src/vertex_remover.cpp:500-546). Three steps, so that the units of a round can be spliced range by range while the
// GPU still works on the rest of the round (remove() drains every pipeline chunk as its records arrive):
//   applyBegin  — edge-store slots of every unit's substitutes from the providers' upper bounds (prefix sum);
//   applyRange  — units [q0, q1): status and staleness bookkeeping, blanket edges and removed vertices die, substitutes
//                 are built in their slots, adjacency of the kept vertices;
//   applyEnd    — counters, the list of added edges.
// Slots a unit does not use (failed blanket, GLC factors of rank 0) stay dead edges.
void VertexRemover::applyRound(const uint64_t *out) {
    applyBegin();
    const size_t ns = _round.sel.size();
    // test hook: SPG_SPLICE_RANGES=k splices the round in k ranges, like the pipelined path does per chunk
    size_t parts = 1;
    if(const char *e = getenv("SPG_SPLICE_RANGES")) parts = (size_t) std::max(1, atoi(e));
    for(size_t pi = 0; pi < parts; pi++) applyRange(out, ns * pi / parts, ns * (pi + 1) / parts);
    applyEnd();
}

void VertexRemover::applyBegin() {
    const double ta0 = nowS();
    const std::vector<int> &sel = _round.sel;
    const size_t ns = sel.size();
    ApplyState &A = _apply;
    A.base.assign(ns + 1, 0);
    A.used.assign(ns, 0);
    A.okUnit.assign(ns, 0);
    A.removedEdges = 0;
    A.removedVerts = 0;
    A.sized = false;
    A.spliced = 0;
    parallelFor(ns, 4096, [&](size_t q0, size_t q1) {
        for(size_t ui = q0; ui < q1; ui++)
            A.base[ui + 1] = spgr_out_edge_count(_round.algorithm, _opts.topology, _opts.chordRatio, (int) _unitCache[sel[ui]].kept.size());
    });
    for(size_t ui = 0; ui < ns; ui++) A.base[ui + 1] += A.base[ui];
    A.e0 = (int) _graph->edges.size();
    g_prof.a1 += nowS() - ta0;
    g_prof.apply += nowS() - ta0;
}

// the slots of the round's substitutes: blocks prepared ahead (remove(): helper thread) or here; every slot starts dead
void VertexRemover::applySizeEdgeStore() {
    ApplyState &A = _apply;
    if(A.sized) return;
    ChunkedVector<GraphEdge> &E = _graph->edges;
    const size_t end = (size_t) A.e0 + (size_t) A.base.back();
    E.reserve(end, [](GraphEdge &e) { e.alive = false; });
    E.resize(end);
    // slots in the block that was in use already were constructed alive
    for(size_t i = (size_t) A.e0; i < end && i < E.blockEnd((size_t) A.e0); i++) E[i].alive = false;
    A.sized = true;
}

void VertexRemover::applyRange(const uint64_t *out, size_t r0, size_t r1) {
    if(r1 <= r0) return;
    const int dim = _graph->dim, P = _graph->poseWords(), algorithm = _round.algorithm;
    const bool cliquey = _opts.topology == SparsityOptions::CliqueySubgraph || _opts.topology == SparsityOptions::CliqueyDense;
    const double ta0 = nowS();
    const std::vector<int> &sel = _round.sel;
    std::vector<GraphVertex> &V = _graph->verts;
    ChunkedVector<GraphEdge> &E = _graph->edges;
    ApplyState &A = _apply;
    applySizeEdgeStore();
    const std::vector<int> &base = A.base;
    std::vector<char> &okUnit = A.okUnit;
    const int e0 = A.e0;
    const size_t nr = r1 - r0;

    // ---- pass 1 (headers only, host threads): status, staleness stamps ---------------------------------------------
    std::atomic<int> failedA(0), firstBad(-1);
    parallelFor(nr, 2048, [&](size_t q0, size_t q1) {
        for(size_t ui = r0 + q0; ui < r0 + q1; ui++) {
            if(ui + 2 < r0 + q1) {
                __builtin_prefetch(out + _round.outOff[ui + 2]);
                __builtin_prefetch(&_unitCache[sel[ui + 2]]);
            }
            const RemovalUnit &u = _unitCache[sel[ui]];
            const int32_t *oh = reinterpret_cast<const int32_t *>(out + _round.outOff[ui]);
            _done[u.listIndex] = 1; // (one byte per list entry, every entry belongs to one unit)
            if(oh[0] != SPG_BLANKET_OK) {
                failedA++;
                int cur = firstBad.load();
                while((cur < 0 || (int) ui < cur) && !firstBad.compare_exchange_weak(cur, (int) ui)) {}
                continue;
            }
            okUnit[ui] = 1;
            // cached blankets containing these vertices are stale; two units of a round may share a kept vertex and then
            // store the same value
            for(int xi : u.ridx) __atomic_store_n(&_stamp[xi], _planNo, __ATOMIC_RELAXED);
            for(int xi : u.kidx) __atomic_store_n(&_stamp[xi], _planNo, __ATOMIC_RELAXED);
        }
    });
    _remaining -= nr;
    if(failedA.load()) {
        // The reference asserts / exits here. A failed blanket is left in the graph untouched (vertex, edges, no
        // substitutes); the units of a round commute, so the others are unaffected. The call reports
        // SPG_ERR_BLANKET_FAILED with the first failing list index.
        const int fb = firstBad.load();
        if(stats.n_failed == 0) { // (ranges are spliced in list order: the first one seen is the first of the call)
            stats.first_failed_index = _unitCache[sel[fb]].listIndex;
            stats.first_failed_status = reinterpret_cast<const int32_t *>(out + _round.outOff[fb])[0];
        }
        stats.n_failed += failedA.load();
        if(!_toRemoveSet.empty())
            for(size_t ui = r0; ui < r1; ui++)
                if(!okUnit[ui])
                    for(int id : _unitCache[sel[ui]].removed) _toRemoveSet.erase(id);
    }
    double tq = nowS();
    g_prof.a1 += tq - ta0;

    // ---- pass 2 (parallel over units; blankets of a round share no edge and no removed vertex): blanket edges and
    // removed vertices die, the substitute edges are built in place --------------------------------------------------
    std::atomic<int> removedEdgesA(0), removedVertsA(0), droppedA(0), blanketsA(0);
    std::mutex garbageMx;
    std::vector<std::vector<double>> garbage; // payload blocks nobody took over: freed off the critical path
    parallelFor(nr, 256, [&](size_t q0, size_t q1) {
        int dead = 0, deadV = 0, dropped = 0, nblk = 0;
        std::vector<int> spare;
        std::vector<std::vector<double>> mine;
        for(size_t ui = r0 + q0; ui < r0 + q1; ui++) {
            // dependent cache misses again (unit -> edge list -> edges -> payload blocks): three, two and one units ahead
            if(ui + 3 < r0 + q1) {
                const char *pu = reinterpret_cast<const char *>(&_unitCache[sel[ui + 3]]);
                for(size_t l = 0; l < sizeof(RemovalUnit); l += 64) __builtin_prefetch(pu + l);
            }
            if(ui + 2 < r0 + q1) __builtin_prefetch(_unitCache[sel[ui + 2]].edges.data());
            if(ui + 1 < r0 + q1) {
                const RemovalUnit &nx = _unitCache[sel[ui + 1]];
                for(int ei : nx.edges) __builtin_prefetch(&E[ei]);
                for(int xi : nx.ridx) __builtin_prefetch(&V[xi]);
                __builtin_prefetch(out + _round.outOff[ui + 1]);
            }
            if(!okUnit[ui]) continue;
            const RemovalUnit &u = _unitCache[sel[ui]];
            nblk++;
            deadV += (int) u.removed.size();
            spare.clear();
            auto kill = [&](int ei) {
                GraphEdge &e = E[ei];
                if(!e.alive) return;
                e.alive = false;
                if(e.payload.capacity()) spare.push_back(ei); // its payload block is handed to a substitute edge below
                dead++;
            };
            for(int ei : u.edges) kill(ei);
            for(int xi : u.ridx) {
                // g2o's HyperGraph::removeVertex detaches every edge still incident to the vertex; its other ends are
                // blanket vertices, whose adjacency is cleaned below
                for(int ei : V[xi].edges) kill(ei);
                V[xi].alive = false;
                std::vector<int>().swap(V[xi].edges);
                std::vector<int>().swap(V[xi].peer);
            }
            const uint64_t *o = out + _round.outOff[ui];
            const int nnew = reinterpret_cast<const int32_t *>(o)[1];
            const int nk = (int) u.kept.size();
            const int64_t slot = spgr_out_slot_words(dim, algorithm, _opts.topology, nk);
            int minor = 0;
            int64_t centry = SPG_OUT_HEADER_WORDS; // correlated topologies: word offset of the next entry
            for(int e = 0; e < nnew; e++) {
                const uint64_t *sl = o + SPG_OUT_HEADER_WORDS + (int64_t) e * slot;
                const int32_t *si = reinterpret_cast<const int32_t *>(sl);
                if(algorithm != SPG_ALG_NFR && si[1] == 0) continue;
                GraphEdge &ge = E[(size_t) e0 + base[ui] + minor];
                if(!spare.empty()) {
                    ge.payload.swap(E[spare.back()].payload);
                    spare.pop_back();
                }
                ge.uidMajor = u.listIndex;
                ge.uidMinor = minor++;
                ge.alive = true;
                if(algorithm == SPG_ALG_NFR && cliquey) {
                    // correlated topologies: variable-size entries (spg_record.h); nmeas == 1 is a plain pose edge, more a
                    // MultiEdgeCorrelated whose vertex list grows in order of appearance (addMeasurement,
                    // multi_edge_correlated.hpp:29-62)
                    sl = o + centry;
                    si = reinterpret_cast<const int32_t *>(sl);
                    const int nm = si[0], rows = si[1];
                    const int32_t *ab = si + 2;
                    const double *pm = reinterpret_cast<const double *>(sl + 1 + spgr_pad2(2 * nm));
                    centry += spgr_out_entry_words(dim, nm);
                    if(nm == 1) {
                        ge.kind = SPG_EDGE_POSE;
                        ge.v = {u.kept[ab[0]], u.kept[ab[1]]};
                        ge.vx0 = u.kidx[ab[0]];
                        ge.vx1 = u.kidx[ab[1]];
                        ge.rows = dim;
                        ge.setPayload(pm, P, pm + P, dim * dim);
                    } else {
                        ge.kind = SPG_EDGE_MULTI;
                        ge.rows = rows;
                        std::vector<int> order; // kept-list indices in order of appearance
                        for(int q = 0; q < 2 * nm; q++) {
                            int pos = (int) (std::find(order.begin(), order.end(), ab[q]) - order.begin());
                            if(pos == (int) order.size()) order.push_back(ab[q]);
                            ge.pairs.push_back(pos);
                        }
                        for(int ki : order) ge.v.push_back(u.kept[ki]);
                        const int nve = (int) order.size();
                        ge.vx0 = nve > 0 ? u.kidx[order[0]] : -1;
                        ge.vx1 = nve > 1 ? u.kidx[order[1]] : -1;
                        if(nve > 2)
                            for(int ki : order) ge.vxn.push_back(u.kidx[ki]);
                        ge.setPayload(pm, nm * P, pm + (size_t) nm * P, rows * rows);
                    }
                } else if(algorithm == SPG_ALG_NFR) {
                    ge.kind = SPG_EDGE_POSE;
                    ge.v = {u.kept[si[0]], u.kept[si[1]]};
                    ge.vx0 = u.kidx[si[0]];
                    ge.vx1 = u.kidx[si[1]];
                    ge.rows = dim;
                    const double *pm = reinterpret_cast<const double *>(sl + 1);
                    ge.setPayload(pm, P, pm + P, dim * dim);
                } else {
                    const int nvcap = (_opts.topology == SparsityOptions::Dense || nk == 1) ? nk : 2;
                    const int c = dim * nvcap, nve = si[0], rank = si[1];
                    ge.kind = SPG_EDGE_GLC;
                    const int32_t *vi = reinterpret_cast<const int32_t *>(sl + 1);
                    for(int q = 0; q < nve; q++) ge.v.push_back(u.kept[vi[q]]);
                    ge.vx0 = nve > 0 ? u.kidx[vi[0]] : -1;
                    ge.vx1 = nve > 1 ? u.kidx[vi[1]] : -1;
                    if(nve > 2)
                        for(int q = 0; q < nve; q++) ge.vxn.push_back(u.kidx[vi[q]]);
                    ge.rows = rank;
                    const double *pm = reinterpret_cast<const double *>(sl + 1 + spgr_pad2(nvcap));
                    const double *W = pm + c;
                    ge.allocPayload(dim * nve, rank * dim * nve);
                    std::memcpy(ge.meas(), pm, sizeof(double) * dim * nve);
                    double *gi = ge.info();
                    for(int r = 0; r < rank; r++)
                        for(int q = 0; q < dim * nve; q++) gi[(size_t) r * dim * nve + q] = W[(size_t) r * c + q];
                }
            }
            A.used[ui] = minor;
            dropped += nnew - minor; // GLC: getEdge returned NULL for rank-0 factors (src/topology_provider_glc.cpp:85-89)
            for(int ei : spare) mine.emplace_back(std::move(E[ei].payload));
        }
        removedEdgesA += dead;
        removedVertsA += deadV;
        droppedA += dropped;
        blanketsA += nblk;
        std::lock_guard<std::mutex> lk(garbageMx);
        for(auto &g : mine) garbage.emplace_back(std::move(g));
    });
    if(garbage.size() > 4096) std::thread([g = std::move(garbage)]() mutable { g.clear(); }).detach();
    A.removedEdges += removedEdgesA.load();
    A.removedVerts += removedVertsA.load();
    stats.n_dropped_edges += droppedA.load();
    stats.n_blankets += blanketsA.load();
    stats.n_applied += removedVertsA.load();
    g_prof.a2 += nowS() - tq; tq = nowS();

    // ---- pass 3 (parallel over units, one byte lock per vertex: two blankets of a round may share one kept vertex):
    // adjacency of the kept vertices — dead edges out, substitutes in. The order inside an adjacency list carries no
    // meaning (every consumer sorts by the canonical edge key).
    if(_vlock.size() < V.size()) _vlock = std::vector<std::atomic<unsigned char>>(V.size() + V.size() / 4 + 64);
    parallelFor(nr, 256, [&](size_t q0, size_t q1) {
        for(size_t ui = r0 + q0; ui < r0 + q1; ui++) {
            if(ui + 3 < r0 + q1) {
                const char *pu = reinterpret_cast<const char *>(&_unitCache[sel[ui + 3]]);
                for(size_t l = 0; l < sizeof(RemovalUnit); l += 64) __builtin_prefetch(pu + l);
            }
            if(ui + 2 < r0 + q1)
                for(int xi : _unitCache[sel[ui + 2]].kidx) __builtin_prefetch(&V[xi]);
            if(ui + 1 < r0 + q1)
                for(int xi : _unitCache[sel[ui + 1]].kidx) { // (a racing writer may be resizing these: only a hint)
                    __builtin_prefetch(V[xi].edges.data());
                    __builtin_prefetch(V[xi].peer.data());
                }
            if(!okUnit[ui]) continue;
            const RemovalUnit &u = _unitCache[sel[ui]];
            const int nb0 = e0 + base[ui], nb1 = nb0 + A.used[ui];
            for(int xi : u.kidx) {
                std::atomic<unsigned char> &lk = _vlock[xi];
                while(lk.exchange(1, std::memory_order_acquire)) { /* spin: held for a few dozen instructions */ }
                GraphVertex &gv = V[xi];
                // the liveness flags read below are one edge (cache miss) per adjacency entry: all of them in flight first
                for(int ei : gv.edges) __builtin_prefetch(&E[ei].alive);
                size_t w = 0;
                for(size_t k = 0; k < gv.edges.size(); k++)
                    if(E[gv.edges[k]].alive) { gv.edges[w] = gv.edges[k]; gv.peer[w] = gv.peer[k]; w++; }
                gv.edges.resize(w);
                gv.peer.resize(w);
                for(int ei = nb0; ei < nb1; ei++) {
                    const GraphEdge &ge = E[ei];
                    for(int q = 0, n = ge.nv(); q < n; q++)
                        if(ge.vx(q) == xi) {
                            gv.edges.push_back(ei);
                            gv.peer.push_back(n == 2 ? ((ge.vx(1 - q) << 1) | (q == 0)) : -1);
                        }
                }
                lk.store(0, std::memory_order_release);
            }
        }
    });
    g_prof.a3 += nowS() - tq;
    A.spliced += nr;
    g_prof.apply += nowS() - ta0;
}

void VertexRemover::applyEnd() {
    const double ta0 = nowS();
    ApplyState &A = _apply;
    const size_t ns = _round.sel.size();
    int added = 0;
    for(size_t ui = 0; ui < ns; ui++) {
        for(int k = 0; k < A.used[ui]; k++) _added.push_back(A.e0 + A.base[ui] + k);
        added += A.used[ui];
    }
    _graph->version++;
    _graph->aliveEdges += added - A.removedEdges;
    _graph->aliveVertices -= A.removedVerts;
    stats.n_rounds++;
    stats.max_round_width = std::max<int>(stats.max_round_width, (int) ns);
    _round.sel.clear();
    g_prof.apply += nowS() - ta0;
}

// ---- computeSubstituteEdge (src/compute_substitute_edge.cpp:13-96) ---------------------------------
static void invertSmall(int n, const double *A, double *X) { // Eigen MatrixXd::inverse(): partial-pivot LU
    std::vector<double> LU(A, A + n * n);
    std::vector<int> piv(n);
    for(int k = 0; k < n; k++) {
        int p = k;
        double big = std::fabs(LU[k + k * n]);
        for(int i = k + 1; i < n; i++)
            if(std::fabs(LU[i + k * n]) > big) { big = std::fabs(LU[i + k * n]); p = i; }
        piv[k] = p;
        if(p != k)
            for(int j = 0; j < n; j++) std::swap(LU[k + j * n], LU[p + j * n]);
        for(int i = k + 1; i < n; i++) {
            LU[i + k * n] /= LU[k + k * n];
            for(int j = k + 1; j < n; j++) LU[i + j * n] -= LU[i + k * n] * LU[k + j * n];
        }
    }
    for(int c = 0; c < n; c++) {
        std::vector<double> x(n, 0.0);
        x[c] = 1;
        for(int k = 0; k < n; k++)
            if(piv[k] != k) std::swap(x[k], x[piv[k]]);
        for(int i = 0; i < n; i++)
            for(int k = 0; k < i; k++) x[i] -= LU[i + k * n] * x[k];
        for(int i = n - 1; i >= 0; i--) {
            for(int k = i + 1; k < n; k++) x[i] -= LU[i + k * n] * x[k];
            x[i] /= LU[i + i * n];
        }
        for(int i = 0; i < n; i++) X[i + c * n] = x[i];
    }
}

void computeSubstituteEdge(const Graph *gw, const std::set<int> &marginalized, int maxid, int &from, int &to,
                           double *edgemeas, double *edgeinfo) {
    std::set<int> visited;
    std::deque<std::set<int>> frontiers;
    std::set<int> newFrontier;
    int minid = std::numeric_limits<int>::max();
    const int toConnect = std::max(from, to), toReplace = std::min(from, to);
    newFrontier.insert(toReplace);
    visited.insert(toConnect);
    visited.insert(toReplace);
    auto sortedEdges = [&](int id) {
        std::vector<int> es = gw->vertex(id)->edges;
        std::sort(es.begin(), es.end(), [&](int a, int b) {
            const GraphEdge &x = gw->edges[a], &y = gw->edges[b];
            if(x.uidMajor != y.uidMajor) return x.uidMajor < y.uidMajor;
            return x.uidMinor < y.uidMinor;
        });
        return es;
    };
    do {
        frontiers.push_back(newFrontier);
        newFrontier.clear();
        for(int r : frontiers.back()) {
            if(marginalized.count(r) == 0 && r != from && r != to) {
                minid = std::min(minid, r);
            } else {
                visited.insert(r);
                for(int ei : gw->vertex(r)->edges) {
                    const GraphEdge &e = gw->edges[ei];
                    if(e.v.size() == 2) {
                        int idother = e.v[0] == r ? e.v[1] : e.v[0];
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

    const int dim = gw->dim, P = gw->poseWords();
    std::vector<double> covsum(dim * dim, 0.0), inv(dim * dim);
    double meas[7], tmp[7], zi[7];
    poseIdentity(dim, meas);
    int reach = minid;
    while(!frontiers.empty()) {
        std::set<int> lastFrontier = frontiers.back();
        frontiers.pop_back();
        for(int ei : sortedEdges(reach)) {
            const GraphEdge &e = gw->edges[ei];
            if(e.v.size() != 2 || e.kind != SPG_EDGE_POSE) continue;
            if(lastFrontier.count(e.v[0]) || lastFrontier.count(e.v[1])) {
                invertSmall(dim, e.info(), inv.data());
                for(int q = 0; q < dim * dim; q++) covsum[q] += inv[q];
                poseInverse(dim, e.meas(), zi);
                if(from == toConnect) {
                    if(e.v[1] == reach) poseCompose(dim, e.meas(), meas, tmp);
                    else poseCompose(dim, zi, meas, tmp);
                } else {
                    if(e.v[1] == reach) poseCompose(dim, meas, zi, tmp);
                    else poseCompose(dim, meas, e.meas(), tmp);
                }
                std::memcpy(meas, tmp, sizeof(double) * P);
                reach = e.v[1] == reach ? e.v[0] : e.v[1];
                break;
            }
        }
    }
    invertSmall(dim, covsum.data(), inv.data());
    for(int i = 0; i < dim; i++)
        for(int j = 0; j < dim; j++) edgeinfo[i + j * dim] = 0.5 * (inv[i + j * dim] + inv[j + i * dim]);
    std::memcpy(edgemeas, meas, sizeof(double) * P);
    if(from == toConnect) to = minid;
    else from = minid;
}

} // namespace spg

// ---- C ABI, graph level ------------------------------------------------------------------------------
extern "C" {

spg_status spg_graph_create(spg_graph **g, int32_t dim) {
    if(!g || (dim != 3 && dim != 6)) return SPG_ERR_INVALID;
    *g = new spg_graph;
    (*g)->g = new spg::Graph(dim);
    return SPG_OK;
}
void spg_graph_destroy(spg_graph *g) {
    if(!g) return;
    delete g->session;
    delete g->g;
    delete g;
}
spg_status spg_graph_load_g2o(spg_graph **g, const char *path) {
    if(!g || !path) return SPG_ERR_INVALID;
    std::string err;
    spg::Graph *gr = spg::Graph::loadG2o(path, &err);
    if(!gr) return SPG_ERR_IO;
    *g = new spg_graph;
    (*g)->g = gr;
    return SPG_OK;
}
spg_status spg_graph_save_g2o(const spg_graph *g, const char *path) {
    if(!g || !path) return SPG_ERR_INVALID;
    std::string err;
    if(!g->g->saveG2o(path, &err)) {
        spg_set_err(err);
        return SPG_ERR_IO;
    }
    return SPG_OK;
}
spg_status spg_graph_add_factor(spg_graph *g, int32_t kind, int32_t nv, const int32_t *vert_ids, int32_t rows, const double *meas,
                                const double *info_or_w, const int32_t *pairs) {
    if(!g || !vert_ids || !meas || !info_or_w || nv < 1 || rows < 0) return SPG_ERR_INVALID;
    const int dim = g->g->dim, P = g->g->poseWords();
    for(int i = 0; i < nv; i++)
        if(!g->g->hasVertex(vert_ids[i])) return SPG_ERR_INVALID;
    spg::GraphEdge e;
    e.kind = kind;
    for(int i = 0; i < nv; i++) e.v.push_back(vert_ids[i]);
    e.rows = rows;
    if(kind == SPG_EDGE_POSE) {
        if(nv != 2 || rows != dim) return SPG_ERR_INVALID;
        return g->g->addPoseEdge(vert_ids[0], vert_ids[1], meas, info_or_w) >= 0 ? SPG_OK : SPG_ERR_INVALID;
    } else if(kind == SPG_EDGE_GLC) {
        if(rows > dim * nv) return SPG_ERR_INVALID;
        e.setPayload(meas, dim * nv, info_or_w, rows * dim * nv);
    } else if(kind == SPG_EDGE_MULTI) {
        if(rows % dim != 0 || rows == 0 || !pairs) return SPG_ERR_INVALID;
        const int nm = rows / dim;
        for(int q = 0; q < 2 * nm; q++) {
            if(pairs[q] < 0 || pairs[q] >= nv) return SPG_ERR_INVALID;
            e.pairs.push_back(pairs[q]);
        }
        e.setPayload(meas, nm * P, info_or_w, rows * rows);
    } else {
        return SPG_ERR_INVALID;
    }
    e.uidMajor = -1;
    e.uidMinor = g->g->fileEdges++;
    g->g->addEdge(std::move(e));
    return SPG_OK;
}
spg_status spg_graph_add_vertex(spg_graph *g, int32_t id, const double *pose) {
    if(!g || !pose) return SPG_ERR_INVALID;
    return g->g->addVertex(id, pose) ? SPG_OK : SPG_ERR_INVALID;
}
spg_status spg_graph_add_edge(spg_graph *g, int32_t from, int32_t to, const double *meas, const double *info) {
    if(!g || !meas || !info) return SPG_ERR_INVALID;
    return g->g->addPoseEdge(from, to, meas, info) >= 0 ? SPG_OK : SPG_ERR_INVALID;
}
int32_t spg_graph_dim(const spg_graph *g) { return g ? g->g->dim : 0; }
int32_t spg_graph_num_vertices(const spg_graph *g) { return g ? g->g->aliveVertices : 0; }
int32_t spg_graph_num_edges(const spg_graph *g) { return g ? g->g->aliveEdges : 0; }
int32_t spg_graph_max_vertex_id(const spg_graph *g) { return g ? g->g->maxVertexId() : -1; }

static int32_t copyOut(const std::vector<int> &r, int32_t *out, int32_t cap) {
    for(size_t i = 0; i < r.size() && (int32_t) i < cap; i++) out[i] = r[i];
    return (int32_t) r.size();
}
int32_t spg_decimate_global(int32_t last, int32_t endvert, int32_t sparsity, int32_t *out, int32_t cap) {
    if(sparsity <= 0) return -1;
    return copyOut(spg::globalDecimate(last, endvert, spg::DecimateOptions{sparsity, 0}), out, cap);
}
int32_t spg_decimate_online(int32_t last, int32_t endvert, int32_t sparsity, int32_t *out, int32_t cap) {
    if(sparsity <= 0) return -1;
    return copyOut(spg::onlineDecimate(last, endvert, spg::DecimateOptions{sparsity, 0}), out, cap);
}
int32_t spg_decimate_cluster(int32_t last, int32_t endvert, int32_t sparsity, int32_t cluster_size, int32_t *out,
                             int32_t cap) {
    if(sparsity <= 0 || cluster_size <= 0) return -1;
    return copyOut(spg::clusterDecimate(last, endvert, spg::DecimateOptions{sparsity, cluster_size}), out, cap);
}

// GraphWrapperG2O::marginalizeNoOptimize (src/graph_wrapper_g2o.cpp:398-453): provider registration
// order as :431-439, then VertexRemover::remove.
spg_status spg_graph_marginalize(spg_graph *g, spg_ctx *ctx, const int32_t *which, int32_t n_which,
                                 const spg_sparsity_options *opts, int32_t algorithm) {
    if(!g || !ctx || !which || !opts || n_which < 0) return SPG_ERR_INVALID;
    for(int i = 0; i < n_which; i++)
        if(!g->g->hasVertex(which[i])) return SPG_ERR_INVALID; // "vertex needs to exist in order to be marginalized" (:406-407)
    spg::VertexRemover vr;
    if(algorithm == SPG_ALG_GLC) {
        vr.registerTopologyProvider(new spg::TopologyProviderGLC);
    } else {
        vr.registerTopologyProvider(new spg::TopologyProviderSE2);
        vr.registerTopologyProvider(new spg::TopologyProviderSE2ISAM);
        vr.registerTopologyProvider(new spg::TopologyProviderSE3);
        vr.registerTopologyProvider(new spg::TopologyProviderSE3ISAM);
    }
    vr.setGraph(g->g);
    vr.setContext(ctx);
    spg::SparsityOptions o;
    o.topology = (spg::SparsityOptions::SparsityTopology) opts->topology;
    o.chordRatio = opts->chord_ratio;
    o.linPoint = (spg::SparsityOptions::LinearizationPoint) opts->lin_point;
    o.includeIntraClique = opts->include_intra_clique != 0;
    o.flags = opts->flags;
    vr.setSparsityOptions(o);
    spg_status st = SPG_OK;
    vr.remove(std::vector<int>(which, which + n_which), &st);
    g->stats = vr.stats;
    return st;
}
// Round-by-round variant of spg_graph_marginalize, for callers that run the blankets of a round
// themselves (e.g. sharded over several GPUs / ranks): begin -> { next -> compute -> apply }*.
static spg::VertexRemover *makeRemover(spg_graph *g, const spg_sparsity_options *opts, int32_t algorithm) {
    spg::VertexRemover *vr = new spg::VertexRemover;
    if(algorithm == SPG_ALG_GLC) {
        vr->registerTopologyProvider(new spg::TopologyProviderGLC);
    } else {
        vr->registerTopologyProvider(new spg::TopologyProviderSE2);
        vr->registerTopologyProvider(new spg::TopologyProviderSE2ISAM);
        vr->registerTopologyProvider(new spg::TopologyProviderSE3);
        vr->registerTopologyProvider(new spg::TopologyProviderSE3ISAM);
    }
    vr->setGraph(g->g);
    spg::SparsityOptions o;
    o.topology = (spg::SparsityOptions::SparsityTopology) opts->topology;
    o.chordRatio = opts->chord_ratio;
    o.linPoint = (spg::SparsityOptions::LinearizationPoint) opts->lin_point;
    o.includeIntraClique = opts->include_intra_clique != 0;
    o.flags = opts->flags;
    vr->setSparsityOptions(o);
    return vr;
}
spg_status spg_graph_rounds_begin(spg_graph *g, const int32_t *which, int32_t n_which, const spg_sparsity_options *opts,
                                  int32_t algorithm) {
    if(!g || !which || !opts || n_which < 0) return SPG_ERR_INVALID;
    for(int i = 0; i < n_which; i++)
        if(!g->g->hasVertex(which[i])) return SPG_ERR_INVALID;
    delete g->session;
    g->session = makeRemover(g, opts, algorithm);
    return g->session->beginRemoval(std::vector<int>(which, which + n_which));
}
spg_status spg_graph_round_next(spg_graph *g, spg_round_in *round) {
    if(!g || !g->session || !round) return SPG_ERR_INVALID;
    spg_status st = g->session->planRound();
    *round = g->session->roundDescriptor();
    if(st != SPG_OK || round->n_blankets == 0) g->stats = g->session->stats;
    return st;
}
spg_status spg_graph_round_apply(spg_graph *g, const uint64_t *out) {
    if(!g || !g->session || !out) return SPG_ERR_INVALID;
    g->session->applyRound(out);
    g->stats = g->session->stats;
    return g->session->failureStatus();
}

spg_status spg_graph_last_stats(const spg_graph *g, spg_marginalize_stats *stats) {
    if(!g || !stats) return SPG_ERR_INVALID;
    *stats = g->stats;
    return SPG_OK;
}

// edge read-back in canonical order (idx is a position in that order; O(E log E) per call batch: the
// order is cached per query run by callers that iterate 0..E-1)
static thread_local std::vector<int> t_order;
static thread_local unsigned long long t_order_serial = 0, t_order_version = ~0ULL;
static const spg::GraphEdge *edgeAt(const spg_graph *g, int idx) {
    // cached canonical order, keyed by the graph's serial number and mutation count (not by its address: a new graph
    // may be allocated where an old one lived)
    if(t_order_serial != g->g->serial || t_order_version != g->g->version) {
        t_order = g->g->edgeOrder();
        t_order_serial = g->g->serial;
        t_order_version = g->g->version;
    }
    if(idx < 0 || idx >= (int) t_order.size()) return nullptr;
    return &g->g->edges[t_order[idx]];
}
spg_status spg_graph_edge_desc(const spg_graph *g, int32_t idx, spg_edge_desc *d) {
    if(!g || !d) return SPG_ERR_INVALID;
    const spg::GraphEdge *e = edgeAt(g, idx);
    if(!e) return SPG_ERR_INVALID;
    d->kind = e->kind;
    d->nv = (int32_t) e->v.size();
    d->rows = e->rows;
    d->uid_major = e->uidMajor;
    d->uid_minor = e->uidMinor;
    return SPG_OK;
}
spg_status spg_graph_edge_data(const spg_graph *g, int32_t idx, int32_t *vert_ids, double *meas, double *info_or_w) {
    if(!g) return SPG_ERR_INVALID;
    const spg::GraphEdge *e = edgeAt(g, idx);
    if(!e) return SPG_ERR_INVALID;
    if(vert_ids) for(size_t i = 0; i < e->v.size(); i++) vert_ids[i] = e->v[i];
    if(meas) std::memcpy(meas, e->meas(), sizeof(double) * e->nMeas);
    if(info_or_w) std::memcpy(info_or_w, e->info(), sizeof(double) * e->nInfo());
    return SPG_OK;
}
spg_status spg_graph_edge_pairs(const spg_graph *g, int32_t idx, int32_t *pairs) {
    if(!g || !pairs) return SPG_ERR_INVALID;
    const spg::GraphEdge *e = edgeAt(g, idx);
    if(!e) return SPG_ERR_INVALID;
    for(size_t i = 0; i < e->pairs.size(); i++) pairs[i] = e->pairs[i];
    return SPG_OK;
}
spg_status spg_graph_vertex_ids(const spg_graph *g, int32_t *ids) {
    if(!g || !ids) return SPG_ERR_INVALID;
    std::vector<int> v = g->g->vertexIds();
    for(size_t i = 0; i < v.size(); i++) ids[i] = v[i];
    return SPG_OK;
}
spg_status spg_graph_vertex_pose(const spg_graph *g, int32_t id, double *pose) {
    if(!g || !pose) return SPG_ERR_INVALID;
    const spg::GraphVertex *v = g->g->vertex(id);
    if(!v) return SPG_ERR_INVALID;
    std::memcpy(pose, v->pose, sizeof(double) * g->g->poseWords());
    return SPG_OK;
}
spg_status spg_graph_set_vertex_pose(spg_graph *g, int32_t id, const double *pose) {
    if(!g || !pose) return SPG_ERR_INVALID;
    spg::GraphVertex *v = g->g->vertex(id);
    if(!v) return SPG_ERR_INVALID;
    std::memcpy(v->pose, pose, sizeof(double) * g->g->poseWords());
    g->g->version++;
    return SPG_OK;
}
spg_status spg_compute_substitute_edge(const spg_graph *g, const int32_t *marginalized, int32_t n_marginalized,
                                       int32_t maxid, int32_t *from, int32_t *to, double *meas, double *info) {
    if(!g || !from || !to || !meas || !info) return SPG_ERR_INVALID;
    std::set<int> m(marginalized, marginalized + n_marginalized);
    int f = *from, t = *to;
    if(!g->g->hasVertex(f) || !g->g->hasVertex(t)) return SPG_ERR_INVALID;
    spg::computeSubstituteEdge(g->g, m, maxid, f, t, meas, info);
    *from = f;
    *to = t;
    return SPG_OK;
}

} // extern "C"
