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
#include <limits>
#include <sstream>

#include "../../include/spg_record.h"

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
    auto it = index.find(id);
    return it != index.end() && verts[it->second].alive;
}
GraphVertex *Graph::vertex(int id) {
    auto it = index.find(id);
    if(it == index.end() || !verts[it->second].alive) return nullptr;
    return &verts[it->second];
}
const GraphVertex *Graph::vertex(int id) const {
    auto it = index.find(id);
    if(it == index.end() || !verts[it->second].alive) return nullptr;
    return &verts[it->second];
}
bool Graph::addVertex(int id, const double *pose) {
    if(hasVertex(id)) return false;
    GraphVertex v;
    v.id = id;
    std::memcpy(v.pose, pose, sizeof(double) * poseWords());
    if(dim == 3) v.pose[2] = normalizeTheta(v.pose[2]);
    else qnorm(v.pose + 3);
    index[id] = (int) verts.size();
    verts.push_back(v);
    aliveVertices++;
    return true;
}
int Graph::addEdge(const GraphEdge &e) {
    int ei = (int) edges.size();
    edges.push_back(e);
    edges.back().alive = true;
    for(int id : e.v) verts[index[id]].edges.push_back(ei);
    aliveEdges++;
    return ei;
}
int Graph::addPoseEdge(int from, int to, const double *meas, const double *info) {
    if(!hasVertex(from) || !hasVertex(to)) return -1;
    GraphEdge e;
    e.kind = SPG_EDGE_POSE;
    e.v = {from, to};
    e.rows = dim;
    e.meas.assign(meas, meas + poseWords());
    if(dim == 3) e.meas[2] = normalizeTheta(e.meas[2]);
    else qnorm(e.meas.data() + 3);
    e.info.assign(info, info + dim * dim);
    e.uidMajor = -1;
    e.uidMinor = fileEdges++;
    return addEdge(e);
}
void Graph::removeEdge(int ei) {
    GraphEdge &e = edges[ei];
    if(!e.alive) return;
    e.alive = false;
    for(int id : e.v) {
        std::vector<int> &ve = verts[index[id]].edges;
        auto it = std::find(ve.begin(), ve.end(), ei);
        if(it != ve.end()) { *it = ve.back(); ve.pop_back(); }
    }
    std::vector<double>().swap(e.info);
    aliveEdges--;
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

// g2o text reader for the four tags the reference's datasets use (graph_wrapper_g2o.cpp:107-147).
Graph *Graph::loadG2o(const std::string &path, std::string *err) {
    std::ifstream f(path);
    if(!f) {
        if(err) *err = "cannot open " + path;
        return nullptr;
    }
    Graph *g = nullptr;
    struct Pend { int a, b; double z[7]; double info[36]; };
    std::vector<Pend> pend;
    std::string line;
    while(std::getline(f, line)) {
        std::istringstream is(line);
        std::string tag;
        if(!(is >> tag)) continue;
        if(tag == "VERTEX_SE2" || tag == "VERTEX_SE3:QUAT") {
            int d = tag == "VERTEX_SE2" ? 3 : 6;
            if(!g) g = new Graph(d);
            int id;
            double p[7] = {0, 0, 0, 0, 0, 0, 1};
            is >> id;
            for(int i = 0; i < g->poseWords(); i++) is >> p[i];
            g->addVertex(id, p);
        } else if(tag == "EDGE_SE2" || tag == "EDGE_SE3:QUAT") {
            int d = tag == "EDGE_SE2" ? 3 : 6, P = d == 3 ? 3 : 7;
            Pend pe;
            is >> pe.a >> pe.b;
            for(int i = 0; i < P; i++) is >> pe.z[i];
            for(int i = 0; i < d; i++)
                for(int j = i; j < d; j++) {
                    double v;
                    is >> v;
                    pe.info[i + j * d] = v;
                    pe.info[j + i * d] = v;
                }
            pend.push_back(pe);
        }
    }
    if(!g) {
        if(err) *err = "no vertices in " + path;
        return nullptr;
    }
    for(const Pend &pe : pend) g->addPoseEdge(pe.a, pe.b, pe.z, pe.info);
    return g;
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
        // single removal (Tree / Subgraph / CliqueySubgraph): the same sets as markovBlanketVertices / -Edges with
        // sorted vectors instead of node-based sets — this runs once per pending vertex and round
        const GraphVertex *rv = _graph->vertex(root);
        std::vector<int> &mb = vmarkov;
        mb.push_back(root);
        if(rv)
            for(int ei : rv->edges)
                for(int id : _graph->edges[ei].v) mb.push_back(id);
        std::sort(mb.begin(), mb.end());
        mb.erase(std::unique(mb.begin(), mb.end()), mb.end());
        u.removed.push_back(root);
        for(int id : mb)
            if(id != root) u.kept.push_back(id);
        std::vector<int> &es = u.edges;
        for(int id : mb) {
            const GraphVertex *v = _graph->vertex(id);
            if(!v) continue;
            for(int ei : v->edges) {
                const GraphEdge &e = _graph->edges[ei];
                bool is_markov = true, found_hub = false;
                for(int x : e.v) {
                    if(!std::binary_search(mb.begin(), mb.end(), x)) { is_markov = false; break; }
                    if(x == root) found_hub = true;
                }
                if(is_markov && (_opts.includeIntraClique || found_hub)) es.push_back(ei);
            }
        }
        std::sort(es.begin(), es.end(), [&](int a, int b) {
            const GraphEdge &x = _graph->edges[a], &y = _graph->edges[b];
            if(x.uidMajor != y.uidMajor) return x.uidMajor < y.uidMajor;
            if(x.uidMinor != y.uidMinor) return x.uidMinor < y.uidMinor;
            return a < b;
        });
        es.erase(std::unique(es.begin(), es.end()), es.end());
        for(int id : u.removed) u.ridx.push_back(_graph->index.at(id));
        for(int id : u.kept) u.kidx.push_back(_graph->index.at(id));
        return !es.empty();
    }
    std::set<int> rem(u.removed.begin(), u.removed.end());
    for(int id : vmarkov)
        if(!rem.count(id)) u.kept.push_back(id);
    u.edges = markovBlanketEdges(vmarkov, u.removed);
    for(int id : u.removed) u.ridx.push_back(_graph->index.at(id));
    for(int id : u.kept) u.kidx.push_back(_graph->index.at(id));
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

// buildSubgraph (src/vertex_remover.cpp:285-392) + record packing. Returns false when the blanket
// needs the Local non-star linearisation point (g2o LM on the subgraph; not part of this path).
bool VertexRemover::packUnit(const RemovalUnit &u, std::vector<uint64_t> &rec) const {
    const int dim = _graph->dim, P = _graph->poseWords();
    const int nrem = (int) u.removed.size(), nv = nrem + (int) u.kept.size(), ne = (int) u.edges.size();
    std::unordered_map<int, int> local;
    std::vector<int> ids;
    for(int id : u.removed) { local[id] = (int) ids.size(); ids.push_back(id); }
    for(int id : u.kept) { local[id] = (int) ids.size(); ids.push_back(id); }
    std::vector<double> poses((size_t) nv * P);
    for(int i = 0; i < nv; i++) std::memcpy(&poses[(size_t) i * P], _graph->vertex(ids[i])->pose, sizeof(double) * P);

    if(_opts.linPoint == SparsityOptions::Local) {
        // closed-form estimate only for star-shaped blankets (:304-342)
        bool closedForm = true;
        std::vector<int> nconn(nv, 0);
        for(int ei : u.edges) {
            const GraphEdge &e = _graph->edges[ei];
            for(int id : e.v) {
                int li = local[id];
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
                apply(local[e.v[0]], local[e.v[1]], e.meas.data());
            } else if(e.kind == SPG_EDGE_MULTI) {
                for(size_t vi = 0; vi < e.v.size(); vi++)
                    for(size_t m = 0; m < e.pairs.size() / 2; m++)
                        if(e.pairs[2 * m] == (int) vi || e.pairs[2 * m + 1] == (int) vi)
                            apply(local[e.v[e.pairs[2 * m]]], local[e.v[e.pairs[2 * m + 1]]], &e.meas[m * P]);
            }
        }
    }

    int64_t words = spgr_record_fixed_words(dim, nv, ne);
    std::vector<int64_t> eoff(ne);
    for(int i = 0; i < ne; i++) {
        const GraphEdge &e = _graph->edges[u.edges[i]];
        eoff[i] = words;
        words += spgr_edge_words(dim, e.kind, (int) e.v.size(), e.rows);
    }
    words = (words + 1) & ~(int64_t) 1;
    rec.assign((size_t) words, 0);
    int32_t *h = reinterpret_cast<int32_t *>(rec.data());
    h[0] = nv; h[1] = nrem; h[2] = ne; h[3] = dim; h[4] = (int32_t) words; h[5] = 0; h[6] = u.listIndex; h[7] = 0;
    int32_t *rid = reinterpret_cast<int32_t *>(rec.data() + spgr_ids_off());
    for(int i = 0; i < nv; i++) rid[i] = ids[i];
    std::memcpy(rec.data() + spgr_poses_off(nv), poses.data(), sizeof(double) * nv * P);
    int32_t *etab = reinterpret_cast<int32_t *>(rec.data() + spgr_edgetab_off(dim, nv));
    for(int i = 0; i < ne; i++) {
        const GraphEdge &e = _graph->edges[u.edges[i]];
        etab[i] = (int32_t) eoff[i];
        uint64_t *ew = rec.data() + eoff[i];
        int32_t *eh = reinterpret_cast<int32_t *>(ew);
        const int nve = (int) e.v.size();
        eh[0] = e.kind; eh[1] = nve; eh[2] = e.rows; eh[3] = 0;
        int32_t *vi = reinterpret_cast<int32_t *>(ew + 2);
        for(int q = 0; q < nve; q++) vi[q] = local[e.v[q]];
        double *pl = reinterpret_cast<double *>(ew + 2 + spgr_pad2(nve));
        if(e.kind == SPG_EDGE_POSE) {
            std::memcpy(pl, e.meas.data(), sizeof(double) * P);
            std::memcpy(pl + P, e.info.data(), sizeof(double) * dim * dim);
        } else if(e.kind == SPG_EDGE_GLC) {
            std::memcpy(pl, e.meas.data(), sizeof(double) * dim * nve);
            std::memcpy(pl + dim * nve, e.info.data(), sizeof(double) * e.rows * dim * nve);
        } else {
            const int nm = e.rows / dim;
            int32_t *pr = reinterpret_cast<int32_t *>(ew + 2 + spgr_pad2(nve));
            for(int q = 0; q < 2 * nm; q++) pr[q] = e.pairs[q];
            double *pm = reinterpret_cast<double *>(ew + 2 + spgr_pad2(nve) + spgr_pad2(2 * nm));
            std::memcpy(pm, e.meas.data(), sizeof(double) * nm * P);
            std::memcpy(pm + (size_t) nm * P, e.info.data(), sizeof(double) * e.rows * e.rows);
        }
    }
    return true;
}

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
        st = planRound();
        if(st != SPG_OK || _round.units.empty()) break;
        auto t1 = clk::now();
        std::vector<uint64_t> out((size_t) _round.outOff.back(), 0);
        spg_round_in rin = roundDescriptor();
        spg_round_out rout{};
        rout.out = out.data();
        // a context with a communicator (spg_comm_init) shards every round over its ranks; all ranks hold the same
        // graph, get the complete output (all-gather) and splice the same edges
        st = spg_comm_nranks(_ctx) > 1 ? spg_remove_round_sharded(_ctx, &rin, &rout, -1, nullptr) : spg_remove_round(_ctx, &rin, &rout);
        if(st != SPG_OK) {
            error = spg_last_error();
            break;
        }
        auto t2 = clk::now();
        applyRound(out.data());
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
    if(_opts.topology == SparsityOptions::CliqueySubgraph || _opts.topology == SparsityOptions::CliqueyDense) {
        error = "CliqueySubgraph / CliqueyDense (MultiEdgeCorrelated substitutes) are not produced on the device";
        return SPG_ERR_UNSUPPORTED;
    }
    for(TopologyProvider *tp : _topologies)
        if(tp->algorithm() == SPG_ALG_GLC &&
           (!(_opts.topology == SparsityOptions::Dense || _opts.topology == SparsityOptions::Tree) ||
            _opts.linPoint != SparsityOptions::Global)) {
            // asserts of TopologyProviderGLC::topology (src/topology_provider_glc.cpp:107-111)
            error = "GLC needs Dense|Tree topology and the Global linearisation point";
            return SPG_ERR_UNSUPPORTED;
        }
    _pending = toRemove;
    _toRemoveSet = std::set<int>(toRemove.begin(), toRemove.end());
    _done.assign(toRemove.size(), 0);
    _remaining = toRemove.size();
    _round = Round();
    _unitCache.assign(toRemove.size(), RemovalUnit());
    _unitBuilt.assign(toRemove.size(), 0);
    _rootIdx.assign(toRemove.size(), -1);
    for(size_t i = 0; i < toRemove.size(); i++) {
        auto it = _graph->index.find(toRemove[i]);
        if(it != _graph->index.end()) _rootIdx[i] = it->second;
    }
    _stamp.assign(_graph->verts.size(), 0);
    _touchHead.assign(_graph->verts.size(), -1);
    _removedBy.assign(_graph->verts.size(), -1);
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
    rin.n_blankets = (int32_t) _round.units.size();
    rin.rec_off = _round.recOff.data();
    rin.records = _round.records.data();
    rin.out_off = _round.outOff.data();
    return rin;
}

namespace {
struct HostProf {
    double stale = 0, extract = 0, select = 0, pack = 0, apply = 0;
    bool on = getenv("SPG_HOST_PROF") != nullptr;
    ~HostProf() {
        if(on) fprintf(stderr, "[spg host] stale-check %.3f s, extract %.3f s, select %.3f s, pack %.3f s, splice %.3f s\n", stale, extract, select, pack, apply);
    }
} g_prof;
inline double nowS() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
} // namespace

// Select and pack the next wavefront round (empty round: nothing left).
spg_status VertexRemover::planRound() {
    _round = Round();
    if(_remaining == 0) return SPG_OK;
    double tp0 = nowS();
    const int dim = _graph->dim;
    const std::vector<int> &toRemove = _pending;
    {
        // ---- select a round -------------------------------------------------------------------
        // A unit U may run in this round iff it commutes with every earlier pending unit V (selected or
        // deferred): removed(U) misses V's blanket, kept(U) misses removed(V) and U shares at most one kept
        // vertex with V. A deferred unit waits for the units it hit; its blanket after they ran lies inside
        // the union of their blankets and its own, so the two are merged into one component (union-find over
        // region ids) and later units are tested against whole components. Linear in the blanket sizes;
        // blankets are re-extracted only when a vertex of theirs was touched by an applied round.
        std::vector<RemovalUnit> &units = _round.units;
        const int Vn = (int) _graph->verts.size();
        _planNo++;
        if((int) _stamp.size() < Vn) _stamp.resize(Vn, 0);
        if((int) _touchHead.size() < Vn) { _touchHead.resize(Vn, -1); _removedBy.resize(Vn, -1); }
        _touchNext.clear();
        _touchRegion.clear();
        std::vector<int> touched;   // vertex indices whose per-round lists must be reset
        std::vector<int> parent;    // union-find over region ids
        auto find = [&](int x) { while(parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
        std::vector<int> hit, comps;
        std::vector<std::pair<int, int>> shared;
        // pass 1: re-extract the stale cached blankets. Extraction only reads the graph, so it is spread over the
        // host threads (it is ~70 % of the planning time on large graphs); the selection below stays sequential.
        {
            std::vector<int> todo;
            for(size_t i = 0; i < toRemove.size(); i++) {
                if(_done[i] || _rootIdx[i] < 0 || !_graph->verts[_rootIdx[i]].alive) continue;
                const RemovalUnit &u = _unitCache[i];
                bool stale = !_unitBuilt[i];
                if(!stale) {
                    for(int xi : u.ridx) if(_stamp[xi] >= _unitBuilt[i]) { stale = true; break; }
                    if(!stale) for(int xi : u.kidx) if(_stamp[xi] >= _unitBuilt[i]) { stale = true; break; }
                }
                if(stale) todo.push_back((int) i);
            }
            g_prof.stale += nowS() - tp0; tp0 = nowS();
            std::atomic<int> bad(-1);
            auto work = [&](size_t b0, size_t b1) {
                for(size_t q = b0; q < b1; q++) {
                    const int i = todo[q];
                    if(!buildUnit(toRemove[i], i, _toRemoveSet, _unitCache[i])) {
                        int expect = -1;
                        bad.compare_exchange_strong(expect, i);
                    }
                    _unitBuilt[i] = _planNo;
                }
            };
            unsigned nthr = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
            if(todo.size() < 4096) nthr = 1;
            if(nthr <= 1) work(0, todo.size());
            else {
                std::vector<std::thread> pool;
                const size_t chunk = (todo.size() + nthr - 1) / nthr;
                for(unsigned t = 0; t < nthr; t++) {
                    const size_t b0 = std::min(todo.size(), t * chunk), b1 = std::min(todo.size(), b0 + chunk);
                    if(b0 < b1) pool.emplace_back(work, b0, b1);
                }
                for(auto &th : pool) th.join();
            }
            if(bad.load() >= 0) {
                // isolated vertex: the reference asserts blanketEdges.size() > 0
                error = "vertex " + std::to_string(toRemove[bad.load()]) + " has no edges";
                return SPG_ERR_INVALID;
            }
            g_prof.extract += nowS() - tp0; tp0 = nowS();
        }
        for(size_t i = 0; i < toRemove.size(); i++) {
            if(_done[i]) continue;
            if(_rootIdx[i] < 0 || !_graph->verts[_rootIdx[i]].alive) { // merged into an earlier extended blanket (:91)
                _done[i] = 1;
                _remaining--;
                continue;
            }
            RemovalUnit &u = _unitCache[i];
            hit.clear();
            shared.clear();
            auto note = [&](int c) { if(std::find(hit.begin(), hit.end(), c) == hit.end()) hit.push_back(c); };
            for(int xi : u.ridx) {
                for(int t = _touchHead[xi]; t >= 0; t = _touchNext[t]) note(find(_touchRegion[t]));
            }
            for(int xi : u.kidx) {
                if(_removedBy[xi] >= 0) note(find(_removedBy[xi]));
                comps.clear();
                for(int t = _touchHead[xi]; t >= 0; t = _touchNext[t]) {
                    const int c = find(_touchRegion[t]);
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
                if(_touchHead[xi] < 0 && _removedBy[xi] < 0) touched.push_back(xi);
                _touchNext.push_back(_touchHead[xi]);
                _touchRegion.push_back(rid);
                _touchHead[xi] = (int) _touchNext.size() - 1;
                if(removed) _removedBy[xi] = rid;
            };
            for(int xi : u.ridx) reg(xi, true);
            for(int xi : u.kidx) reg(xi, false);
            if(select) units.push_back(u);
        }
        for(int xi : touched) { _touchHead[xi] = -1; _removedBy[xi] = -1; }
        g_prof.select += nowS() - tp0; tp0 = nowS();
        if(units.empty()) return SPG_OK;
    }
    // ---- pack ----------------------------------------------------------------------------
    TopologyProvider *tp = nullptr;
    _round.recOff.assign(1, 0);
    _round.outOff.assign(1, 0);
    std::vector<uint64_t> rec;
    bool poseOnly = true;
    for(const RemovalUnit &u : _round.units) {
        TopologyProvider *t = chooseTopologyProvider(u);
        if(!t) {
            error = "No valid topology provider for Markov blanket";
            return SPG_ERR_UNSUPPORTED;
        }
        if(tp && t->algorithm() != tp->algorithm()) {
            error = "mixed providers within one round";
            return SPG_ERR_UNSUPPORTED;
        }
        tp = t;
        for(int ei : u.edges) poseOnly = poseOnly && (_graph->edges[ei].kind == SPG_EDGE_POSE);
        if(!packUnit(u, rec)) {
            error = "Local linearisation point on a non-star blanket needs the subgraph optimiser (not on this path)";
            return SPG_ERR_UNSUPPORTED;
        }
        _round.records.insert(_round.records.end(), rec.begin(), rec.end());
        _round.recOff.push_back((int64_t) _round.records.size());
        _round.outOff.push_back(_round.outOff.back() + spgr_out_record_words(dim, tp->algorithm(), _opts.topology, _opts.chordRatio,
                                                                             (int) u.kept.size()));
        stats.max_blanket_vertices = std::max<int>(stats.max_blanket_vertices, (int) (u.removed.size() + u.kept.size()));
    }
    g_prof.pack += nowS() - tp0;
    _round.algorithm = tp->algorithm();
    _round.poseOnly = poseOnly;
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

// Splice the output records of the planned round, in list order (updateInputGraph,
// src/vertex_remover.cpp:500-546).
void VertexRemover::applyRound(const uint64_t *out) {
    const int dim = _graph->dim, P = _graph->poseWords(), algorithm = _round.algorithm;
    const double ta0 = nowS();
    for(size_t ui = 0; ui < _round.units.size(); ui++) {
        const RemovalUnit &u = _round.units[ui];
        const uint64_t *o = out + _round.outOff[ui];
        const int32_t *oh = reinterpret_cast<const int32_t *>(o);
        const int bstatus = oh[0], nnew = oh[1];
        if(bstatus != SPG_BLANKET_OK) {
            // The reference asserts / exits here. The blanket is left in the graph untouched (vertex, edges, no
            // substitutes); the units of a round commute, so the others are unaffected. The call reports
            // SPG_ERR_BLANKET_FAILED with the first failing list index.
            if(stats.n_failed++ == 0) {
                stats.first_failed_index = u.listIndex;
                stats.first_failed_status = bstatus;
            }
            for(int id : u.removed) _toRemoveSet.erase(id);
            _done[u.listIndex] = 1;
            _remaining--;
            continue;
        }
        for(int xi : u.ridx) _stamp[xi] = _planNo; // cached blankets containing these are stale
        for(int xi : u.kidx) _stamp[xi] = _planNo;
        for(int ei : u.edges) _graph->removeEdge(ei);
        for(int id : u.removed) _graph->removeVertex(id);
        const int nk = (int) u.kept.size();
        const int64_t slot = spgr_out_slot_words(dim, algorithm, _opts.topology, nk);
        int minor = 0;
        for(int e = 0; e < nnew; e++) {
            const uint64_t *sl = o + SPG_OUT_HEADER_WORDS + (int64_t) e * slot;
            const int32_t *si = reinterpret_cast<const int32_t *>(sl);
            GraphEdge ge;
            ge.uidMajor = u.listIndex;
            if(algorithm == SPG_ALG_NFR) {
                ge.kind = SPG_EDGE_POSE;
                ge.v = {u.kept[si[0]], u.kept[si[1]]};
                ge.rows = dim;
                const double *pm = reinterpret_cast<const double *>(sl + 1);
                ge.meas.assign(pm, pm + P);
                ge.info.assign(pm + P, pm + P + dim * dim);
            } else {
                const int nvcap = (_opts.topology == SparsityOptions::Dense || nk == 1) ? nk : 2;
                const int c = dim * nvcap, nve = si[0], rank = si[1];
                if(rank == 0) { // getEdge returned NULL (src/topology_provider_glc.cpp:85-89)
                    stats.n_dropped_edges++;
                    continue;
                }
                ge.kind = SPG_EDGE_GLC;
                const int32_t *vi = reinterpret_cast<const int32_t *>(sl + 1);
                for(int q = 0; q < nve; q++) ge.v.push_back(u.kept[vi[q]]);
                ge.rows = rank;
                const double *pm = reinterpret_cast<const double *>(sl + 1 + spgr_pad2(nvcap));
                ge.meas.assign(pm, pm + dim * nve);
                const double *W = pm + c;
                ge.info.resize((size_t) rank * dim * nve);
                for(int r = 0; r < rank; r++)
                    for(int q = 0; q < dim * nve; q++) ge.info[(size_t) r * dim * nve + q] = W[(size_t) r * c + q];
            }
            ge.uidMinor = minor++;
            _added.push_back(_graph->addEdge(ge));
        }
        _done[u.listIndex] = 1;
        _remaining--;
        stats.n_blankets++;
        stats.n_applied += (int) u.removed.size();
    }
    stats.n_rounds++;
    stats.max_round_width = std::max<int>(stats.max_round_width, (int) _round.units.size());
    _round.units.clear();
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
                invertSmall(dim, e.info.data(), inv.data());
                for(int q = 0; q < dim * dim; q++) covsum[q] += inv[q];
                poseInverse(dim, e.meas.data(), zi);
                if(from == toConnect) {
                    if(e.v[1] == reach) poseCompose(dim, e.meas.data(), meas, tmp);
                    else poseCompose(dim, zi, meas, tmp);
                } else {
                    if(e.v[1] == reach) poseCompose(dim, meas, zi, tmp);
                    else poseCompose(dim, meas, e.meas.data(), tmp);
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
static thread_local const spg::Graph *t_order_of = nullptr;
static thread_local size_t t_order_edges = 0;
static const spg::GraphEdge *edgeAt(const spg_graph *g, int idx) {
    if(t_order_of != g->g || t_order_edges != g->g->edges.size() || (int) t_order.size() != g->g->aliveEdges) {
        t_order = g->g->edgeOrder();
        t_order_of = g->g;
        t_order_edges = g->g->edges.size();
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
    if(meas) std::memcpy(meas, e->meas.data(), sizeof(double) * e->meas.size());
    if(info_or_w) std::memcpy(info_or_w, e->info.data(), sizeof(double) * e->info.size());
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
