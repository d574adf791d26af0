// ORACLE — TEST INFRASTRUCTURE ONLY (see dense.hpp header). PARITY UNPINNED.
// One Markov blanket through the reference's per-vertex pipeline:
//   computeTargetInformation  (src/vertex_remover.cpp:394-450)
//   TopologyProviderBinary    (src/topology_provider_binary.hpp:22-70)   NFR skeleton edges
//   TopologyProviderGLC       (src/topology_provider_glc.cpp:42-185)      GLC edges
//   buildJacobianMapping      (src/vertex_remover.cpp:466-498)
//   optimizeInformation       (src/optimizer.cpp:16-81)
#pragma once
#include <array>
#include "chow_liu.hpp"
#include "nfr.hpp"
#include "poses.hpp"

namespace orc {

enum { EDGE_POSE = 0, EDGE_GLC = 1, EDGE_MULTI = 2 };
enum { ALG_NFR = 0, ALG_GLC = 1 };
enum {
    ST_OK = 0, ST_NOT_PD_MARGINAL = 1, ST_NOT_PD_CHOWLIU = 2, ST_EIG_NOCONV = 3, ST_NOT_PD_CLOSED = 4,
    ST_TOO_LARGE = 5, ST_LINESEARCH_FAIL = 6, ST_KLD_INF = 7, ST_UNSUPPORTED = 8, ST_NOT_PD_JOINT = 9
};

// An edge in blanket-local (or kept-local) vertex indices.
struct BEdge {
    int kind = EDGE_POSE;
    std::vector<int> v;            // vertex indices
    // POSE
    Pose meas;
    Mat info;                      // d x d (POSE) or rows x rows (MULTI)
    // GLC (glc_edge.h): error = W * reparam(vertices, gmeas), information = I
    std::vector<double> gmeas;     // d * nv
    Mat W;                         // rows x d*nv
    // MULTI (multi_edge_correlated.h): list of binary measurements over v
    std::vector<std::array<int, 2>> pairs; // indices INTO v
    std::vector<Pose> mmeas;
    int rows() const {
        if(kind == EDGE_POSE) return meas.dim;
        if(kind == EDGE_GLC) return W.rows();
        return (int) mmeas.size() * mmeas.front().dim;
    }
};

struct Blanket {
    int dim = 6;
    int nRemoved = 1;
    std::vector<int> ids;      // removed first, then kept ascending (vertex_remover.cpp:349-356)
    std::vector<Pose> poses;   // linearisation point
    std::vector<BEdge> edges;  // canonical summation order
};

struct BlanketResult {
    int status = ST_OK;
    Mat target;                     // Lambda_t, k x k
    std::vector<double> weights;    // Chow-Liu MI (lexicographic pairs), empty if not computed
    std::vector<BEdge> edges;       // new edges, vertex indices into the KEPT list
    std::vector<int> glcRank;       // per GLC edge
    int droppedEdges = 0;           // GLC rank-0 edges (getEdge returned NULL)
    NfrStats nfr;
};

// ---- GLC reparametrisation, reference src/glc_reparam_binary.hpp:34-120 -------------------
// errorToMeasurement: SE2 -> SE2(err) (glc_reparam_se2.h:38-40), SE3 -> fromVectorMQT (glc_reparam_se3.h:25-27)
static inline Pose glcErrorToMeasurement(int dim, const double *e) {
    if(dim == 3) return Pose::se2(e[0], e[1], e[2]);
    return se3FromMQT(e);
}
static inline std::vector<double> glcReparametrize(int dim, const std::vector<Pose> &vs, const std::vector<double> &meas) {
    int d = dim;
    std::vector<double> ret(d * vs.size());
    Pose zero = Pose::identity(dim);
    edgeError(glcErrorToMeasurement(dim, &meas[0]), zero, vs[0], &ret[0]);
    for(size_t i = 1; i < vs.size(); i++)
        edgeError(glcErrorToMeasurement(dim, &meas[d * i]), vs[0], vs[i], &ret[d * i]);
    return ret;
}
static inline Mat glcJacobian(int dim, const std::vector<Pose> &vs, const std::vector<double> &meas) {
    int d = dim, n = (int) vs.size();
    Mat J(d * n, d * n);
    Pose zero = Pose::identity(dim);
    Mat Ji, Jj;
    edgeJacobians(glcErrorToMeasurement(dim, &meas[0]), zero, vs[0], Ji, Jj);
    J.setBlock(0, 0, Jj);
    for(int i = 1; i < n; i++) {
        edgeJacobians(glcErrorToMeasurement(dim, &meas[d * i]), vs[0], vs[i], Ji, Jj);
        J.setBlock(d * i, 0, Ji);
        J.setBlock(d * i, d * i, Jj);
    }
    return J;
}

// Per-vertex Jacobian blocks (rows x d each) of an edge at the given poses: what
// e->linearizeOplus(jw) leaves in the workspace (vertex_remover.cpp:486-494 and g2o buildSystem).
static inline std::vector<Mat> edgeJacobianBlocks(int dim, const BEdge &e, const std::vector<Pose> &poses) {
    std::vector<Mat> out;
    if(e.kind == EDGE_POSE) {
        Mat Ji, Jj;
        edgeJacobians(e.meas, poses[e.v[0]], poses[e.v[1]], Ji, Jj);
        out.push_back(Ji);
        out.push_back(Jj);
    } else if(e.kind == EDGE_GLC) {
        // GLCEdge::linearizeOplus, glc_edge.cpp:40-49
        std::vector<Pose> vs;
        for(int vi : e.v) vs.push_back(poses[vi]);
        Mat J = glcJacobian(dim, vs, e.gmeas);
        for(size_t i = 0; i < e.v.size(); i++) out.push_back(e.W * J.block(0, dim * (int) i, J.rows(), dim));
    } else {
        // MultiEdgeCorrelated::linearizeOplus, multi_edge_correlated.hpp:96-140
        int rows = e.rows();
        for(size_t i = 0; i < e.v.size(); i++) out.push_back(Mat(rows, dim));
        for(size_t m = 0; m < e.pairs.size(); m++) {
            Mat Ji, Jj;
            int a = e.pairs[m][0], b = e.pairs[m][1];
            edgeJacobians(e.mmeas[m], poses[e.v[a]], poses[e.v[b]], Ji, Jj);
            out[a].setBlock(dim * (int) m, 0, Ji);
            out[b].setBlock(dim * (int) m, 0, Jj);
        }
    }
    return out;
}
static inline Mat edgeInformation(const BEdge &e) {
    if(e.kind == EDGE_GLC) return Mat::identity(e.W.rows());
    return e.info;
}

// H = sum_e J^T Omega J over the blanket, g2o BlockSolver::buildSystem + the symmetric mirror of
// utils.cpp:99-123 (triplets). Row/col order = blanket vertex order (removed first).
static inline Mat assembleH(const Blanket &b) {
    int d = b.dim, N = d * (int) b.ids.size();
    Mat H(N, N);
    for(const BEdge &e : b.edges) {
        std::vector<Mat> J = edgeJacobianBlocks(d, e, b.poses);
        Mat Om = edgeInformation(e);
        for(size_t i = 0; i < e.v.size(); i++) {
            Mat JtO = J[i].transpose() * Om;
            for(size_t j = 0; j < e.v.size(); j++) {
                // g2o stores only upper blocks and the reference mirrors them; summing both
                // (i,j) and (j,i) products is the same matrix.
                H.addBlock(d * e.v[i], d * e.v[j], JtO * J[j]);
            }
        }
    }
    return H;
}

// Edge error vector at the given poses (rows of the edge): g2o computeError of the three factor kinds.
static inline std::vector<double> edgeErrorVector(int dim, const BEdge &e, const std::vector<Pose> &poses) {
    std::vector<double> err;
    if(e.kind == EDGE_POSE) {
        err.resize(dim);
        edgeError(e.meas, poses[e.v[0]], poses[e.v[1]], err.data());
    } else if(e.kind == EDGE_GLC) { // GLCEdge::computeError, glc_edge.cpp:28-32
        std::vector<Pose> vs;
        for(int vi : e.v) vs.push_back(poses[vi]);
        std::vector<double> r = glcReparametrize(dim, vs, e.gmeas);
        err.assign(e.W.rows(), 0.0);
        for(int i = 0; i < e.W.rows(); i++)
            for(int j = 0; j < e.W.cols(); j++) err[i] += e.W(i, j) * r[j];
    } else { // MultiEdgeCorrelated::computeError, multi_edge_correlated.hpp:64-77
        err.resize(e.rows());
        for(size_t m = 0; m < e.pairs.size(); m++)
            edgeError(e.mmeas[m], poses[e.v[e.pairs[m][0]]], poses[e.v[e.pairs[m][1]]], &err[dim * m]);
    }
    return err;
}

// Local linearisation point without a closed form (vertex_remover.cpp:382-391): the blanket subgraph optimised for
// 10 iterations with the first removed vertex fixed. g2o's OptimizationAlgorithmLevenberg is third party (absent here);
// restated from its published algorithm with the default parameters: lambda_0 = 1e-5 * max diag(H), a step is accepted
// when rho = (chi - chi') / (sum x (lambda x + b) + 1e-3) > 0, lambda *= max(1/3, min(1 - (2 rho - 1)^3, 2/3)) on
// success, lambda *= nu, nu *= 2 on failure (at most 10 trials), Terminate when no trial succeeded or rho == 0.
static inline void localOptimize(Blanket &b, int iterations = 10) {
    const int d = b.dim, nv = (int) b.ids.size(), n = d * (nv - 1);
    if(n <= 0) return;
    auto chi2 = [&](const std::vector<Pose> &poses) {
        double c = 0;
        for(const BEdge &e : b.edges) {
            std::vector<double> err = edgeErrorVector(d, e, poses);
            Mat Om = edgeInformation(e);
            for(int i = 0; i < Om.rows(); i++)
                for(int j = 0; j < Om.cols(); j++) c += err[i] * Om(i, j) * err[j];
        }
        return c;
    };
    double lambda = 0, ni = 2, chi = chi2(b.poses);
    for(int it = 0; it < iterations; it++) {
        Mat H(n, n), g(n, 1);
        for(const BEdge &e : b.edges) {
            std::vector<Mat> J = edgeJacobianBlocks(d, e, b.poses);
            std::vector<double> err = edgeErrorVector(d, e, b.poses);
            Mat Om = edgeInformation(e), ev((int) err.size(), 1);
            for(size_t i = 0; i < err.size(); i++) ev((int) i, 0) = err[i];
            for(size_t i = 0; i < e.v.size(); i++) {
                if(e.v[i] == 0) continue; // the fixed vertex
                Mat JtO = J[i].transpose() * Om, gi = JtO * ev;
                for(int q = 0; q < d; q++) g(d * (e.v[i] - 1) + q, 0) += gi(q, 0);
                for(size_t j = 0; j < e.v.size(); j++)
                    if(e.v[j] != 0) H.addBlock(d * (e.v[i] - 1), d * (e.v[j] - 1), JtO * J[j]);
            }
        }
        if(it == 0) {
            double md = 0;
            for(int i = 0; i < n; i++) md = std::max(md, H(i, i));
            lambda = 1e-5 * md;
            ni = 2;
        }
        double rho = 0;
        int qmax = 0;
        do {
            Mat A = H;
            for(int i = 0; i < n; i++) A(i, i) += lambda;
            LLT llt(A);
            Mat rhs(n, 1);
            for(int i = 0; i < n; i++) rhs(i, 0) = -g(i, 0);
            Mat x = llt.solve(rhs);
            std::vector<Pose> trial = b.poses;
            if(llt.ok)
                for(int v = 1; v < nv; v++) {
                    double dx[6];
                    for(int q = 0; q < d; q++) dx[q] = x(d * (v - 1) + q, 0);
                    trial[v] = oplus(b.poses[v], dx);
                }
            double tempChi = llt.ok ? chi2(trial) : std::numeric_limits<double>::max();
            rho = chi - tempChi;
            double scale = 0;
            for(int i = 0; i < n; i++) scale += x(i, 0) * (lambda * x(i, 0) - g(i, 0));
            rho /= scale + 1e-3;
            if(rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1.0 - std::pow(2 * rho - 1, 3);
                alpha = std::min(alpha, 2.0 / 3.0);
                lambda *= std::max(1.0 / 3.0, alpha);
                ni = 2;
                chi = tempChi;
                b.poses = trial;
            } else {
                lambda *= ni;
                ni *= 2;
                if(!std::isfinite(lambda)) break;
            }
            qmax++;
        } while(rho < 0 && qmax < 10);
        if(qmax == 10 || rho == 0 || !std::isfinite(lambda)) break;
    }
}

// Schur complement, vertex_remover.cpp:443-449
static inline Mat schurTarget(const Mat &H, int m, bool *ok) {
    int N = H.rows(), k = N - m;
    std::vector<int> im, ik;
    for(int i = 0; i < m; i++) im.push_back(i);
    for(int i = m; i < N; i++) ik.push_back(i);
    LLT chol(selectVariables(H, im));
    *ok = chol.ok;
    Mat mixed = selectVariables(H, im, ik);
    Mat info = selectVariables(H, ik) - mixed.transpose() * chol.solve(mixed);
    mirrorUpperToLower(info);
    (void) k;
    return info;
}

// ---- GLC, reference src/topology_provider_glc.cpp ------------------------------------------
static const double glc_eps = 1e-8;

// posdef_pinv, :42-56
static inline Mat posdef_pinv(const Mat &a) {
    SymEig eig(a);
    int n = a.rows();
    double maxabs = 0;
    for(double w : eig.w) maxabs = std::max(maxabs, std::fabs(w));
    double tolerance = std::numeric_limits<double>::epsilon() * std::max(a.cols(), a.rows()) * maxabs;
    Mat VD(n, n);
    for(int j = 0; j < n; j++) {
        double inv = (eig.w[j] > tolerance) ? 1.0 / eig.w[j] : 0.0;
        for(int i = 0; i < n; i++) VD(i, j) = eig.V(i, j) * inv;
    }
    return VD * eig.V.transpose();
}
// glc_chol, :59-71 — returns V_keep * sqrt(D_keep), (dn x r)
static inline Mat glc_chol(const Mat &J, const Mat &m) {
    int i = 0;
    Mat invJ = luInverse(J);
    Mat m2 = invJ.transpose() * m * invJ;
    SymEig eig(m2);
    int n = m.cols();
    while(i < n && eig.w[i] < glc_eps) i++;
    Mat out(n, n - i);
    for(int c = 0; c < n - i; c++) {
        double s = std::sqrt(eig.w[i + c]);
        for(int r = 0; r < n; r++) out(r, c) = eig.V(r, i + c) * s;
    }
    return out;
}
// getEdge, :73-98. `verts` index the kept list; returns false when rank 0 (NULL edge).
static inline bool glcGetEdge(int dim, const Mat &targetInfo, const std::vector<int> &verts,
                              const std::vector<Pose> &keptPoses, BEdge &edge, int &rank) {
    std::vector<Pose> vs;
    for(int v : verts) vs.push_back(keptPoses[v]);
    std::vector<double> zero(dim * verts.size(), 0.0);
    std::vector<double> meas = glcReparametrize(dim, vs, zero);
    Mat J = glcJacobian(dim, vs, meas);
    Mat W = glc_chol(J, targetInfo).transpose();
    rank = W.rows();
    if(W.rows() == 0) return false;
    edge = BEdge();
    edge.kind = EDGE_GLC;
    edge.v = verts;
    edge.W = W;
    edge.gmeas = glcReparametrize(dim, vs, zero); // computeMeasurement(), glc_edge.cpp:23-26
    return true;
}

static inline void glcTopology(const SparsityOptions &opts, int dim, const Mat &information,
                               const std::vector<Pose> &keptPoses, BlanketResult &res) {
    int n = (int) keptPoses.size();
    std::vector<int> all;
    for(int i = 0; i < n; i++) all.push_back(i);
    BEdge edge;
    int rank = 0;
    if(n == 1 || opts.topology == SparsityOptions::Dense) {
        if(glcGetEdge(dim, information, all, keptPoses, edge, rank)) {
            res.edges.push_back(edge);
            res.glcRank.push_back(rank);
        } else {
            res.droppedEdges++;
        }
        return;
    }
    PseudoChowLiu cl(opts, information, n, dim);
    cl.computeSparsityPattern();
    if(!cl.ok) res.status = ST_NOT_PD_CHOWLIU;
    res.weights = cl.weights;
    const PseudoChowLiu::SparsityPattern &sp = cl.getSparsityPattern();
    int root = sp.front().front().first;
    bool ok = true;
    Mat rootInfo = cl.marginal(root, &ok);
    if(!ok && res.status == ST_OK) res.status = ST_NOT_PD_JOINT;
    if(glcGetEdge(dim, rootInfo, std::vector<int>(1, root), keptPoses, edge, rank)) {
        res.edges.push_back(edge);
        res.glcRank.push_back(rank);
    } else {
        res.droppedEdges++;
    }
    for(const PseudoChowLiu::CorrelatedSkeletonTree &tree : sp) {
        int a = tree.front().first, b = tree.front().second;
        int d1 = dim, d2 = dim;
        Mat jointInfo = cl.jointMarginal(a, b, &ok);
        if(!ok && res.status == ST_OK) res.status = ST_NOT_PD_JOINT;
        Mat b1 = jointInfo.block(d1, 0, d2, d1);
        Mat b2 = posdef_pinv(jointInfo.block(0, 0, d1, d1));
        Mat b3 = jointInfo.block(0, d1, d1, d2);
        Mat m4 = b1 * b2 * b3;
        Mat target(d1 + d2, d1 + d2);
        target.setBlock(0, 0, jointInfo.block(0, 0, d1, d1));
        target.setBlock(0, d1, jointInfo.block(0, d1, d1, d2));
        target.setBlock(d1, 0, jointInfo.block(d1, 0, d2, d1));
        target.setBlock(d1, d1, m4);
        std::vector<int> vc = {a, b};
        if(glcGetEdge(dim, selfadjointUpper(target), vc, keptPoses, edge, rank)) {
            res.edges.push_back(edge);
            res.glcRank.push_back(rank);
        } else {
            res.droppedEdges++;
        }
    }
}

// ---- NFR, reference src/topology_provider_binary.hpp:22-70 + vertex_remover.cpp:123-126 -----
static inline void nfrTopology(const SparsityOptions &opts, int dim, const Mat &information,
                               const std::vector<Pose> &keptPoses, BlanketResult &res) {
    int n = (int) keptPoses.size();
    if(n < 2) return;
    PseudoChowLiu cl(opts, information, n, dim);
    cl.computeSparsityPattern();
    if(!cl.ok) res.status = ST_NOT_PD_CHOWLIU;
    res.weights = cl.weights;
    for(const auto &tree : cl.getSparsityPattern()) {
        BEdge e;
        if(tree.size() == 1) {
            int a = tree.front().first, b = tree.front().second;
            e.kind = EDGE_POSE;
            e.v = {a, b};
            e.meas = compose(inverse(keptPoses[a]), keptPoses[b]); // setMeasurementFromState()
        } else {
            e.kind = EDGE_MULTI;
            for(const auto &pr : tree) {
                int idx[2];
                int vv[2] = {pr.first, pr.second};
                for(int s = 0; s < 2; s++) { // addMeasurement, multi_edge_correlated.hpp:26-62
                    auto where = std::find(e.v.begin(), e.v.end(), vv[s]);
                    if(where == e.v.end()) {
                        idx[s] = (int) e.v.size();
                        e.v.push_back(vv[s]);
                    } else {
                        idx[s] = int(where - e.v.begin());
                    }
                }
                e.pairs.push_back({idx[0], idx[1]});
                e.mmeas.push_back(compose(inverse(keptPoses[pr.first]), keptPoses[pr.second]));
            }
        }
        res.edges.push_back(e);
    }
    if(res.edges.empty()) return;
    // buildJacobianMapping, vertex_remover.cpp:466-498
    JacobianMapping mapping;
    for(const BEdge &e : res.edges) {
        std::vector<Mat> J = edgeJacobianBlocks(dim, e, keptPoses);
        mapping.push_back(MeasurementJacobian());
        for(size_t i = 0; i < e.v.size(); i++) mapping.back().push_back(std::make_pair(J[i], dim * e.v[i]));
    }
    std::list<Mat> infos = optimizeInformation(mapping, information, res.nfr);
    auto it = infos.begin();
    for(BEdge &e : res.edges) e.info = *it++;
    if(res.status == ST_OK) {
        if(res.nfr.kldInf) res.status = ST_KLD_INF;
        else if(res.nfr.notPd && res.nfr.closedForm) res.status = ST_NOT_PD_CLOSED;
    }
}

// Body of the loop of VertexRemover::remove for one blanket whose linearisation point is fixed.
static inline BlanketResult processBlanket(const Blanket &b, const SparsityOptions &opts, int algorithm) {
    BlanketResult res;
    Mat H = assembleH(b);
    bool ok = true;
    res.target = schurTarget(H, b.dim * b.nRemoved, &ok);
    if(!ok) res.status = ST_NOT_PD_MARGINAL;
    std::vector<Pose> kept(b.poses.begin() + b.nRemoved, b.poses.end());
    if(algorithm == ALG_GLC) {
        glcTopology(opts, b.dim, res.target, kept, res);
    } else {
        nfrTopology(opts, b.dim, res.target, kept, res);
    }
    return res;
}

} // namespace orc
