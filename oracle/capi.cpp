// ORACLE — TEST INFRASTRUCTURE ONLY. Built into oracle/_build/libspg_oracle.so and loaded ONLY by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// PARITY UNPINNED: CPU restatement of the reference's node-removal path (the reference itself
// needs Eigen/g2o/iSAM/CHOLMOD and cannot be built here); pinned by finite differences, analytic
// invariants and a numpy/LAPACK twin (tests/), not by reference golden vectors (none exist).
//
// C entry points (prefix orc_) over the same packed records as include/spg_capi.h so the parity
// tests feed identical bytes to the oracle and to the CUDA path.
#include <atomic>
#include <chrono>
#include <thread>
#include "../include/spg_record.h"
#include "graph.hpp"

using namespace orc;

namespace {

inline const int32_t *i32(const uint64_t *w) { return reinterpret_cast<const int32_t *>(w); }
inline int32_t *i32(uint64_t *w) { return reinterpret_cast<int32_t *>(w); }
inline const double *f64(const uint64_t *w) { return reinterpret_cast<const double *>(w); }
inline double *f64(uint64_t *w) { return reinterpret_cast<double *>(w); }

SparsityOptions toOpts(const spg_sparsity_options *o) {
    SparsityOptions s;
    s.topology = (SparsityOptions::SparsityTopology) o->topology;
    s.chordRatio = o->chord_ratio;
    s.linPoint = (SparsityOptions::LinearizationPoint) o->lin_point;
    s.includeIntraClique = o->include_intra_clique != 0;
    return s;
}

bool parseRecord(const uint64_t *rec, Blanket &b) {
    const int32_t *h = i32(rec);
    int nvert = h[0], nrem = h[1], nedges = h[2], dim = h[3];
    if(dim != 3 && dim != 6) return false;
    b.dim = dim;
    b.nRemoved = nrem;
    int P = (int) spgr_pose_words(dim);
    const int32_t *ids = i32(rec + spgr_ids_off());
    const double *poses = f64(rec + spgr_poses_off(nvert));
    const int32_t *etab = i32(rec + spgr_edgetab_off(dim, nvert));
    for(int i = 0; i < nvert; i++) {
        b.ids.push_back(ids[i]);
        b.poses.push_back(poseFromFlat(dim, poses + (size_t) P * i));
    }
    for(int e = 0; e < nedges; e++) {
        const uint64_t *ew = rec + etab[e];
        const int32_t *eh = i32(ew);
        int kind = eh[0], nv = eh[1], rows = eh[2];
        BEdge be;
        be.kind = kind;
        const int32_t *vidx = i32(ew + 2);
        for(int i = 0; i < nv; i++) be.v.push_back(vidx[i]);
        const double *pl = f64(ew + 2 + spgr_pad2(nv));
        if(kind == SPG_EDGE_POSE) {
            be.meas = poseFromFlat(dim, pl);
            be.info = Mat(dim, dim);
            std::memcpy(be.info.a.data(), pl + P, sizeof(double) * dim * dim);
        } else if(kind == SPG_EDGE_GLC) {
            int c = dim * nv;
            be.gmeas.assign(pl, pl + c);
            be.W = Mat(rows, c);
            for(int r = 0; r < rows; r++)
                for(int j = 0; j < c; j++) be.W(r, j) = pl[c + (size_t) r * c + j];
        } else {
            int nmeas = rows / dim;
            const int32_t *pairs = i32(ew + 2 + spgr_pad2(nv));
            const double *pm = f64(ew + 2 + spgr_pad2(nv) + spgr_pad2(2 * nmeas));
            for(int m = 0; m < nmeas; m++) {
                be.pairs.push_back({pairs[2 * m], pairs[2 * m + 1]});
                be.mmeas.push_back(poseFromFlat(dim, pm + (size_t) P * m));
            }
            be.info = Mat(rows, rows);
            std::memcpy(be.info.a.data(), pm + (size_t) P * nmeas, sizeof(double) * rows * rows);
        }
        b.edges.push_back(be);
    }
    return true;
}

void writeOut(const Blanket &b, const BlanketResult &res, int algorithm, const SparsityOptions &opts,
              uint64_t *out, int64_t outWords) {
    std::memset(out, 0, sizeof(uint64_t) * outWords);
    int dim = b.dim, P = (int) spgr_pose_words(dim);
    int nk = (int) b.ids.size() - b.nRemoved;
    int32_t *h = i32(out);
    h[0] = res.status;
    h[1] = (int32_t) res.edges.size();
    h[2] = res.nfr.newtonIters;
    h[3] = (res.nfr.lineSearchFailed ? 1 : 0);
    f64(out)[2] = res.nfr.kld;
    int64_t slot = spgr_out_slot_words(dim, algorithm, opts.topology, nk);
    uint64_t *w = out + SPG_OUT_HEADER_WORDS;
    bool cliquey = opts.topology == SparsityOptions::CliqueyDense || opts.topology == SparsityOptions::CliqueySubgraph;
    for(size_t ei = 0; ei < res.edges.size(); ei++, w += slot) {
        const BEdge &e = res.edges[ei];
        if((w - out) + slot > outWords) { h[0] = ST_UNSUPPORTED; return; }
        if(algorithm == ALG_NFR && !cliquey && e.kind == EDGE_POSE) {
            i32(w)[0] = e.v[0];
            i32(w)[1] = e.v[1];
            poseToFlat(e.meas, f64(w + 1));
            std::memcpy(f64(w + 1 + P), e.info.a.data(), sizeof(double) * dim * dim);
        } else if(algorithm == ALG_GLC) {
            int nvcap = (opts.topology == SparsityOptions::Dense || nk == 1) ? nk : 2;
            int c = dim * nvcap, nv = (int) e.v.size();
            i32(w)[0] = nv;
            i32(w)[1] = e.W.rows();
            for(int i = 0; i < nv; i++) i32(w + 1)[i] = e.v[i];
            double *m = f64(w + 1 + spgr_pad2(nvcap));
            for(int i = 0; i < dim * nv; i++) m[i] = e.gmeas[i];
            double *W = m + c;
            for(int r = 0; r < e.W.rows(); r++)
                for(int j = 0; j < dim * nv; j++) W[(size_t) r * c + j] = e.W(r, j);
        } else {
            h[0] = ST_UNSUPPORTED;
            return;
        }
    }
}

// correlated topologies: the sequence of variable-size entries of spg_record.h (spgr_out_entry_words)
void writeOutCliquey(const Blanket &b, const BlanketResult &res, uint64_t *out, int64_t outWords) {
    std::memset(out, 0, sizeof(uint64_t) * outWords);
    int dim = b.dim, P = (int) spgr_pose_words(dim);
    int32_t *h = i32(out);
    h[0] = res.status;
    h[1] = (int32_t) res.edges.size();
    h[2] = res.nfr.newtonIters;
    h[3] = (res.nfr.lineSearchFailed ? 1 : 0);
    f64(out)[2] = res.nfr.kld;
    uint64_t *w = out + SPG_OUT_HEADER_WORDS;
    for(const BEdge &e : res.edges) {
        const int nm = e.kind == EDGE_POSE ? 1 : (int) e.pairs.size(), rows = dim * nm;
        if((w - out) + spgr_out_entry_words(dim, nm) > outWords) { h[0] = ST_UNSUPPORTED; return; }
        i32(w)[0] = nm;
        i32(w)[1] = rows;
        int32_t *ab = i32(w + 1);
        double *m = f64(w + 1 + spgr_pad2(2 * nm));
        if(e.kind == EDGE_POSE) {
            ab[0] = e.v[0];
            ab[1] = e.v[1];
            poseToFlat(e.meas, m);
        } else {
            for(int q = 0; q < nm; q++) {
                ab[2 * q] = e.v[e.pairs[q][0]];
                ab[2 * q + 1] = e.v[e.pairs[q][1]];
                poseToFlat(e.mmeas[q], m + (size_t) q * P);
            }
        }
        std::memcpy(m + (size_t) nm * P, e.info.a.data(), sizeof(double) * rows * rows);
        w += spgr_out_entry_words(dim, nm);
    }
}

struct OGraphHandle {
    Graph *g = nullptr;
    std::vector<RemoveLogEntry> log;
};

} // namespace

extern "C" {

const char *orc_version() { return "spg-oracle 0.1 (CPU restatement; parity unpinned)"; }

// Same contract as spg_remove_round on host buffers. n_threads <= 0: hardware concurrency.
// Returns wall seconds spent in the blanket loop.
double orc_remove_round(const spg_round_in *in, spg_round_out *out, int n_threads) {
    SparsityOptions opts = toOpts(&in->opts);
    int nb = in->n_blankets;
    if(n_threads <= 0) n_threads = (int) std::thread::hardware_concurrency();
    if(n_threads < 1) n_threads = 1;
    if(n_threads > nb) n_threads = nb > 0 ? nb : 1;
    std::atomic<int> next(0);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&]() {
        for(;;) {
            int i = next.fetch_add(1);
            if(i >= nb) break;
            Blanket b;
            uint64_t *o = out->out + in->out_off[i];
            int64_t ow = in->out_off[i + 1] - in->out_off[i];
            if(!parseRecord(in->records + in->rec_off[i], b)) {
                std::memset(o, 0, sizeof(uint64_t) * ow);
                i32(o)[0] = ST_UNSUPPORTED;
                continue;
            }
            // SPG_OPT_DBG_WEIGHTS_IN: the Chow-Liu weights are an INPUT (exact-tie tests)
            g_overrideWeights = ((in->opts.flags & SPG_OPT_DBG_WEIGHTS_IN) && out->dbg_weights && out->dbg_weights_off)
                                        ? out->dbg_weights + out->dbg_weights_off[i] : nullptr;
            BlanketResult res = processBlanket(b, opts, in->algorithm);
            g_overrideWeights = nullptr;
            if(in->algorithm == ALG_NFR && (opts.topology == SparsityOptions::CliqueyDense || opts.topology == SparsityOptions::CliqueySubgraph))
                writeOutCliquey(b, res, o, ow);
            else
                writeOut(b, res, in->algorithm, opts, o, ow);
            if(out->dbg_target && out->dbg_target_off) {
                double *t = out->dbg_target + out->dbg_target_off[i];
                int64_t cap = out->dbg_target_off[i + 1] - out->dbg_target_off[i];
                if((int64_t) res.target.a.size() <= cap) std::memcpy(t, res.target.a.data(), sizeof(double) * res.target.a.size());
            }
            if(out->dbg_weights && out->dbg_weights_off) {
                double *t = out->dbg_weights + out->dbg_weights_off[i];
                int64_t cap = out->dbg_weights_off[i + 1] - out->dbg_weights_off[i];
                for(int64_t q = 0; q < cap; q++) t[q] = q < (int64_t) res.weights.size() ? res.weights[q] : 0.0;
            }
        }
    };
    if(n_threads == 1) {
        work();
    } else {
        std::vector<std::thread> th;
        for(int t = 0; t < n_threads; t++) th.emplace_back(work);
        for(auto &t : th) t.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- graph level --------------------------------------------------------------------------
void *orc_graph_load_g2o(const char *path) {
    Graph *g = loadG2o(path);
    if(!g) return nullptr;
    OGraphHandle *h = new OGraphHandle;
    h->g = g;
    return h;
}
void *orc_graph_create(int dim) {
    OGraphHandle *h = new OGraphHandle;
    h->g = new Graph;
    h->g->dim = dim;
    return h;
}
void orc_graph_destroy(void *hp) {
    OGraphHandle *h = (OGraphHandle *) hp;
    if(!h) return;
    delete h->g;
    delete h;
}
void orc_graph_add_vertex(void *hp, int id, const double *pose) {
    Graph *g = ((OGraphHandle *) hp)->g;
    g->addVertex(id, poseFromFlat(g->dim, pose));
}
void orc_graph_add_edge(void *hp, int from, int to, const double *meas, const double *info) {
    Graph *g = ((OGraphHandle *) hp)->g;
    Mat I(g->dim, g->dim);
    std::memcpy(I.a.data(), info, sizeof(double) * g->dim * g->dim);
    g->addPoseEdge(from, to, poseFromFlat(g->dim, meas), I);
}
int orc_graph_dim(void *hp) { return ((OGraphHandle *) hp)->g->dim; }
int orc_graph_num_vertices(void *hp) { return (int) ((OGraphHandle *) hp)->g->verts.size(); }
int orc_graph_num_edges(void *hp) { return (int) ((OGraphHandle *) hp)->g->edges.size(); }
int orc_graph_max_vertex_id(void *hp) {
    Graph *g = ((OGraphHandle *) hp)->g;
    return g->verts.empty() ? -1 : g->verts.rbegin()->first;
}
void orc_graph_vertex_ids(void *hp, int *ids) {
    int i = 0;
    for(auto &v : ((OGraphHandle *) hp)->g->verts) ids[i++] = v.first;
}
int orc_graph_vertex_pose(void *hp, int id, double *pose) {
    GVertex *v = ((OGraphHandle *) hp)->g->vertex(id);
    if(!v) return 1;
    poseToFlat(v->est, pose);
    return 0;
}

// Sequential VertexRemover::remove(which). Returns the number of blankets whose status != OK.
int orc_graph_marginalize(void *hp, const int *which, int n, const spg_sparsity_options *o, int algorithm) {
    OGraphHandle *h = (OGraphHandle *) hp;
    VertexRemover vr;
    vr.graph = h->g;
    vr.opts = toOpts(o);
    vr.algorithm = algorithm;
    vr.remove(std::vector<int>(which, which + n));
    h->log = vr.log;
    int bad = 0;
    for(auto &le : h->log) bad += le.status != ST_OK;
    return bad;
}
int orc_graph_log_size(void *hp) { return (int) ((OGraphHandle *) hp)->log.size(); }
// entry: root id, n blanket vertices, n removed, status, newton iters, n pattern pairs
void orc_graph_log_entry(void *hp, int i, int *meta, double *kld) {
    const RemoveLogEntry &le = ((OGraphHandle *) hp)->log[i];
    meta[0] = le.rootId;
    meta[1] = (int) le.blanketIds.size();
    meta[2] = le.nRemoved;
    meta[3] = le.status;
    meta[4] = le.newtonIters;
    meta[5] = (int) le.pattern.size();
    *kld = le.kld;
}
void orc_graph_log_detail(void *hp, int i, int *blanket_ids, int *pattern_pairs) {
    const RemoveLogEntry &le = ((OGraphHandle *) hp)->log[i];
    for(size_t k = 0; k < le.blanketIds.size(); k++) blanket_ids[k] = le.blanketIds[k];
    for(size_t k = 0; k < le.pattern.size(); k++) {
        pattern_pairs[2 * k] = le.pattern[k].first;
        pattern_pairs[2 * k + 1] = le.pattern[k].second;
    }
}

static GEdge *edgeAt(Graph *g, int idx) {
    auto it = g->edges.begin();
    std::advance(it, idx);
    return *it;
}
void orc_graph_edge_desc(void *hp, int idx, spg_edge_desc *d) {
    GEdge *e = edgeAt(((OGraphHandle *) hp)->g, idx);
    d->kind = e->body.kind;
    d->nv = (int) e->verts.size();
    d->rows = e->body.rows();
    d->uid_major = (int) e->keyMajor;
    d->uid_minor = (int) e->keyMinor;
}
void orc_graph_edge_data(void *hp, int idx, int *vert_ids, double *meas, double *info_or_w) {
    Graph *g = ((OGraphHandle *) hp)->g;
    GEdge *e = edgeAt(g, idx);
    for(size_t i = 0; i < e->verts.size(); i++) vert_ids[i] = e->verts[i]->id;
    const BEdge &b = e->body;
    if(b.kind == EDGE_POSE) {
        poseToFlat(b.meas, meas);
        std::memcpy(info_or_w, b.info.a.data(), sizeof(double) * b.info.a.size());
    } else if(b.kind == EDGE_GLC) {
        for(size_t i = 0; i < b.gmeas.size(); i++) meas[i] = b.gmeas[i];
        int c = b.W.cols();
        for(int r = 0; r < b.W.rows(); r++)
            for(int j = 0; j < c; j++) info_or_w[(size_t) r * c + j] = b.W(r, j);
    } else {
        int P = poseWords(g->dim);
        for(size_t m = 0; m < b.mmeas.size(); m++) poseToFlat(b.mmeas[m], meas + P * m);
        std::memcpy(info_or_w, b.info.a.data(), sizeof(double) * b.info.a.size());
    }
}

int orc_decimate_global(int last, int endvert, int sparsity, int *out, int cap) {
    std::vector<int> r = globalDecimate(last, endvert, sparsity);
    for(size_t i = 0; i < r.size() && (int) i < cap; i++) out[i] = r[i];
    return (int) r.size();
}
int orc_decimate_online(int last, int endvert, int sparsity, int *out, int cap) {
    std::vector<int> r = onlineDecimate(last, endvert, sparsity);
    for(size_t i = 0; i < r.size() && (int) i < cap; i++) out[i] = r[i];
    return (int) r.size();
}
int orc_decimate_cluster(int last, int endvert, int sparsity, int cluster, int *out, int cap) {
    std::vector<int> r = clusterDecimate(last, endvert, sparsity, cluster);
    for(size_t i = 0; i < r.size() && (int) i < cap; i++) out[i] = r[i];
    return (int) r.size();
}
void orc_compute_substitute_edge(void *hp, const int *marg, int nmarg, int maxid, int *from, int *to,
                                 double *meas, double *info) {
    Graph *g = ((OGraphHandle *) hp)->g;
    std::set<int> m(marg, marg + nmarg);
    Pose z;
    Mat I;
    computeSubstituteEdge(g, m, maxid, *from, *to, z, I);
    poseToFlat(z, meas);
    std::memcpy(info, I.a.data(), sizeof(double) * I.a.size());
}

// ---- low-level hooks for the finite-difference / known-answer tests ---------------------------
void orc_edge_error(int dim, const double *z, const double *xi, const double *xj, double *err) {
    edgeError(poseFromFlat(dim, z), poseFromFlat(dim, xi), poseFromFlat(dim, xj), err);
}
// Ji, Jj: d x d column-major
void orc_edge_jacobians(int dim, const double *z, const double *xi, const double *xj, double *Ji, double *Jj) {
    Mat A, B;
    edgeJacobians(poseFromFlat(dim, z), poseFromFlat(dim, xi), poseFromFlat(dim, xj), A, B);
    std::memcpy(Ji, A.a.data(), sizeof(double) * dim * dim);
    std::memcpy(Jj, B.a.data(), sizeof(double) * dim * dim);
}
void orc_oplus(int dim, const double *x, const double *delta, double *out) {
    poseToFlat(oplus(poseFromFlat(dim, x), delta), out);
}
void orc_compose(int dim, const double *a, const double *b, double *out) {
    poseToFlat(compose(poseFromFlat(dim, a), poseFromFlat(dim, b)), out);
}
void orc_inverse(int dim, const double *a, double *out) { poseToFlat(inverse(poseFromFlat(dim, a)), out); }
// symmetric eigen-decomposition (column-major A n x n) -> w ascending, V columns
int orc_sym_eig(int n, const double *A, double *w, double *V) {
    Mat M(n, n);
    std::memcpy(M.a.data(), A, sizeof(double) * n * n);
    SymEig e(M);
    std::memcpy(w, e.w.data(), sizeof(double) * n);
    std::memcpy(V, e.V.a.data(), sizeof(double) * n * n);
    return e.ok ? 0 : 1;
}
double orc_ldlt_sumlogd(int n, const double *A, int *positive) {
    Mat M(n, n);
    std::memcpy(M.a.data(), A, sizeof(double) * n * n);
    LDLT l(M);
    *positive = l.positive();
    return l.sumLogD();
}
// GLC reparametrisation of n poses (flat), meas d*n -> r (d*n) and J (dn x dn column-major)
void orc_glc_reparam(int dim, int n, const double *poses, const double *meas, double *r, double *J) {
    int P = poseWords(dim);
    std::vector<Pose> vs;
    for(int i = 0; i < n; i++) vs.push_back(poseFromFlat(dim, poses + P * i));
    std::vector<double> m(meas, meas + dim * n);
    std::vector<double> rr = glcReparametrize(dim, vs, m);
    Mat JJ = glcJacobian(dim, vs, m);
    std::memcpy(r, rr.data(), sizeof(double) * dim * n);
    std::memcpy(J, JJ.a.data(), sizeof(double) * JJ.a.size());
}

// LogdetFunctionWithConstraints on a mapping of `nmeas` measurements, each with `nblk[m]` Jacobian
// blocks (rows[m] x cols[m][b], column-major, concatenated in Jdata) at offsets off[m][b]
// (flattened in `offs`). Evaluates value / gradient / hessian at x (test_logdet.cpp shape).
// g: xsize, H: xsize*xsize column-major (may be NULL).
double orc_logdet_eval(int k, const double *target, int nmeas, const int *rows, const int *nblk,
                       const int *cols, const int *offs, const double *Jdata, double rho, const double *x,
                       double *g, double *H, int *closed_form) {
    Mat T(k, k);
    std::memcpy(T.a.data(), target, sizeof(double) * k * k);
    JacobianMapping mapping;
    int bi = 0;
    const double *jp = Jdata;
    for(int m = 0; m < nmeas; m++) {
        mapping.push_back(MeasurementJacobian());
        for(int b = 0; b < nblk[m]; b++, bi++) {
            Mat J(rows[m], cols[bi]);
            std::memcpy(J.a.data(), jp, sizeof(double) * rows[m] * cols[bi]);
            jp += rows[m] * cols[bi];
            mapping.back().push_back(std::make_pair(J, offs[bi]));
        }
    }
    LogdetFunctionWithConstraints fun(mapping, T);
    fun.setRho(rho);
    if(closed_form) *closed_form = fun.hasClosedFormSolution();
    int n = fun.xsize();
    Vec xv(x, x + n), gv;
    double f = fun.value(xv);
    fun.gradient(xv, gv);
    if(g) std::memcpy(g, gv.data(), sizeof(double) * n);
    if(H) {
        Mat HH(n, n);
        fun.hessian(xv, HH);
        std::memcpy(H, HH.a.data(), sizeof(double) * n * n);
    }
    return f;
}

} // extern "C"
