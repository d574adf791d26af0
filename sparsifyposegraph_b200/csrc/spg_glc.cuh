// spg_glc.cuh — GLC substitute factors on device (R6): TopologyProviderGLC::topology / getEdge /
// glc_chol / posdef_pinv (reference src/topology_provider_glc.cpp:42-185) and
// PseudoChowLiu::marginal / jointMarginal (src/pseudo_chow_liu.cpp:130-167), CTA-cooperative on
// shared-memory matrices.
#pragma once
#include "../../include/spg_record.h"
#include "spg_device.cuh"

namespace spg {

// error -> measurement of the GLC reparametrisation: SE2(err) / fromVectorMQT(err)
// (glc_reparam_se2.h:38-40, glc_reparam_se3.h:25-27)
template <int D>
__device__ __forceinline__ void glc_error_to_pose(const double *e, double *p) {
    if constexpr(D == 6) se3_from_mqt(e, p);
    else se2_from_flat(e, p);
}
template <int D>
__device__ __forceinline__ void pose_identity(double *p) {
    if constexpr(D == 6) {
        p[0] = 1; p[1] = 0; p[2] = 0; p[3] = 0; p[4] = 1; p[5] = 0; p[6] = 0; p[7] = 0; p[8] = 1; p[9] = 0; p[10] = 0; p[11] = 0;
    } else {
        p[0] = 0; p[1] = 0; p[2] = 0; p[3] = 1; p[4] = 0; p[5] = 0;
    }
}

// Blocks of the GLC reparametrisation Jacobian (GLCReparamBinary::jacobian, glc_reparam_binary.hpp:76-120)
// for vertex i of the edge: i == 0 -> A = J(0,0); i >= 1 -> A = J(i,0), B = J(i,i). D x D column-major.
template <int D>
__device__ __forceinline__ void glc_reparam_blocks(int i, const double *meas, const double *x0, const double *xi, double *A, double *B) {
    constexpr int PS = PoseStride<D>::value;
    double Z[PS], J[D * 2 * D];
    glc_error_to_pose<D>(meas + i * D, Z);
    if(i == 0) {
        double I0[PS];
        pose_identity<D>(I0);
        edge_jacobians<D>(Z, I0, x0, J);
        for(int q = 0; q < D * D; q++) A[q] = J[D * D + q];
    } else {
        edge_jacobians<D>(Z, x0, xi, J);
        for(int q = 0; q < D * D; q++) { A[q] = J[q]; B[q] = J[D * D + q]; }
    }
}

// Eigen::PartialPivLU inverse of a D x D matrix (column-major), registers/local memory, one thread.
template <int D>
__device__ void lu_inverse_small(const double *A, double *X) {
    double LU[D * D];
    int piv[D];
    for(int q = 0; q < D * D; q++) LU[q] = A[q];
    for(int k = 0; k < D; k++) {
        int p = k;
        double big = fabs(LU[k + k * D]);
        for(int i = k + 1; i < D; i++)
            if(fabs(LU[i + k * D]) > big) { big = fabs(LU[i + k * D]); p = i; }
        piv[k] = p;
        if(p != k)
            for(int j = 0; j < D; j++) { double t = LU[k + j * D]; LU[k + j * D] = LU[p + j * D]; LU[p + j * D] = t; }
        const double d = LU[k + k * D];
        for(int i = k + 1; i < D; i++) {
            LU[i + k * D] /= d;
            const double l = LU[i + k * D];
            for(int j = k + 1; j < D; j++) LU[i + j * D] -= l * LU[k + j * D];
        }
    }
    for(int c = 0; c < D; c++) {
        double x[D];
        for(int i = 0; i < D; i++) x[i] = (i == c) ? 1.0 : 0.0;
        for(int k = 0; k < D; k++)
            if(piv[k] != k) { double t = x[k]; x[k] = x[piv[k]]; x[piv[k]] = t; }
        for(int i = 0; i < D; i++)
            for(int k = 0; k < i; k++) x[i] -= LU[i + k * D] * x[k];
        for(int i = D - 1; i >= 0; i--) {
            for(int k = i + 1; k < D; k++) x[i] -= LU[i + k * D] * x[k];
            x[i] /= LU[i + i * D];
        }
        for(int i = 0; i < D; i++) X[i + c * D] = x[i];
    }
}

// GLC measurement r(x^) of the kept vertices `verts` (GLCReparamBinary::reparametrize with zero
// measurement, glc_reparam_binary.hpp:34-74): block 0 = toVector(x0), block i = toVector(x0^-1 xi).
template <int D>
__device__ void glc_measurement(int i, const double *x0, const double *xi, double *out) {
    constexpr int PS = PoseStride<D>::value;
    if constexpr(D == 6) {
        if(i == 0) se3_to_mqt(x0, out);
        else {
            double T[PS], B[PS];
            se3_inverse(x0, T);
            se3_compose(T, xi, B);
            se3_to_mqt(B, out);
        }
    } else {
        if(i == 0) { out[0] = x0[0]; out[1] = x0[1]; out[2] = normalize_theta(x0[2]); }
        else {
            double T[PS], B[PS];
            se2_inverse(x0, T);
            se2_compose(T, xi, B);
            out[0] = B[0]; out[1] = B[1]; out[2] = normalize_theta(B[2]);
        }
    }
}

// PseudoChowLiu::marginal(keep) for keep = the dimensions of kept vertices va (and vb if nb == 2):
// S = T_kk - T_kr chol(T_rr)^-1 T_rk, then selfadjointView<Upper>. T: k x k (ldT). Out: S (c x c, ld c).
// work: (k-c) x odd(k-c) Cholesky buffer; xs: (k-c) x c solve buffer. Returns false if T_rr is not PD.
template <int D, int NT>
__device__ bool glc_marginal(const double *T, int k, int ldT, int va, int vb, int nb, double *S, double *work, double *xs) {
    const int tid = threadIdx.x;
    const int c = D * nb, nr = k - c, ldw = odd_ld(nr > 0 ? nr : 1);
    auto keepIdx = [&](int i) { return (i < D ? va * D + i : vb * D + i - D); };
    auto restIdx = [&](int i) { // i-th index of the sorted complement
        if(nb == 1) return i < va * D ? i : i + D;
        int x = i;
        if(x >= va * D) x += D;
        if(x >= vb * D) x += D;
        return x;
    };
    if(nr == 0) {
        for(int t = tid; t < c * c; t += NT) {
            const int i = t % c, j = t / c;
            const int a = i <= j ? i : j, b = i <= j ? j : i; // upper triangle mirrored
            S[i + j * c] = T[keepIdx(a) + (size_t) keepIdx(b) * ldT];
        }
        gsync<NT>();
        return true;
    }
    // Register-tiled partial sweep (spg_device.cuh): T permuted to [rest, keep], the nr rest pivots swept, and the
    // trailing block is S — the same Schur complement as the LLT route below at a fraction of its barriers (the LLT
    // route factors T_rr cooperatively but solves with one thread per right-hand side: c of NT threads busy).
    if(sweep_fits<NT>(k)) {
        const int ldp = odd_ld(k);
        auto permIdx = [&](int i) { return i < nr ? restIdx(i) : keepIdx(i - nr); };
        for(int t = tid; t < k * k; t += NT) {
            const int i = t % k, j = t / k;
            work[i + j * ldp] = T[permIdx(i) + (size_t) permIdx(j) * ldT];
        }
        gsync<NT>();
        const int r = sweep_spd_auto<D, NT>(work, ldp, work, ldp, k, nr, xs);
        if(r == 0) return false;
        if(r == 1) {
            for(int t = tid; t < c * c; t += NT) {
                const int i = t % c, j = t / c;
                const int a = i <= j ? i : j, b = i <= j ? j : i; // upper triangle mirrored
                S[i + j * c] = work[(nr + a) + (size_t) (nr + b) * ldp];
            }
            gsync<NT>();
            return true;
        }
    }
    for(int t = tid; t < nr * nr; t += NT) {
        const int i = t % nr, j = t / nr;
        work[i + j * ldw] = T[restIdx(i) + (size_t) restIdx(j) * ldT];
    }
    for(int t = tid; t < nr * c; t += NT) {
        const int i = t % nr, j = t / nr;
        xs[i + j * nr] = T[restIdx(i) + (size_t) keepIdx(j) * ldT]; // mixed^T = T_rk
    }
    gsync<NT>();
    if(!chol_lower<NT>(work, nr, ldw)) return false;
    for(int col = tid; col < c; col += NT) { // chol.solve(mixed^T), thread per right-hand side
        double *x = xs + (size_t) col * nr;
        for(int i = 0; i < nr; i++) {
            double s = x[i];
            for(int p = 0; p < i; p++) s -= work[i + p * ldw] * x[p];
            x[i] = s / work[i + i * ldw];
        }
        for(int i = nr - 1; i >= 0; i--) {
            double s = x[i];
            for(int p = i + 1; p < nr; p++) s -= work[p + i * ldw] * x[p];
            x[i] = s / work[i + i * ldw];
        }
    }
    gsync<NT>();
    for(int t = tid; t < c * c; t += NT) {
        const int i = t % c, j = t / c;
        if(i <= j) {
            double s = 0;
            for(int p = 0; p < nr; p++) s += T[keepIdx(i) + (size_t) restIdx(p) * ldT] * xs[p + (size_t) j * nr];
            const double v = T[keepIdx(i) + (size_t) keepIdx(j) * ldT] - s;
            S[i + j * c] = v;
            S[j + i * c] = v;
        }
    }
    gsync<NT>();
    return true;
}

// TopologyProviderGLC::getEdge for nv <= 2 vertices (tree edges and the root): target (c x c, ld c,
// full symmetric). Writes the slot if rank > 0. Scratch `sc` needs 8*c*c + 4*c doubles.
// Returns the rank (0: edge dropped, reference returns NULL), or -1 if the eigen-solver failed.
template <int D, int NT>
__device__ int glc_get_edge_small(const double *target, int nv, const int *kv, const double *s_pose_kept, double *sc,
                                  uint64_t *slot, int cslot, int nvcap) {
    constexpr int PS = PoseStride<D>::value;
    const int tid = (NT <= 32) ? (int) (threadIdx.x & 31) : (int) threadIdx.x; // NT = 32: any one warp of the CTA
    const int c = D * nv, ldm = odd_ld(c);
    double *meas = sc;               // c
    double *AB = meas + c;           // 2 * nv * D*D
    double *invJ = AB + 2 * nv * D * D; // c*c (ld c)
    double *T1 = invJ + c * c;       // c*c
    double *M2 = T1 + c * c;         // c*ldm
    double *V = M2 + c * ldm;        // c*ldm
    double *cs = V + c * ldm;        // 2c + 4
    double *red = cs + 2 * c + 4;    // 4
    int *order = reinterpret_cast<int *>(red + 4); // c ints
    const double *x0 = s_pose_kept + PS * kv[0];
    for(int i = tid; i < nv; i += NT) glc_measurement<D>(i, x0, s_pose_kept + PS * kv[i], meas + i * D);
    gsync<NT>();
    for(int i = tid; i < nv; i += NT)
        glc_reparam_blocks<D>(i, meas, x0, s_pose_kept + PS * kv[i], AB + i * D * D, AB + (nv + i) * D * D);
    for(int t = tid; t < c * c; t += NT) invJ[t] = 0.0;
    gsync<NT>();
    if(tid == 0) { // J^-1 by blocks (J is block lower-triangular: [[A0, 0], [A1, B1]])
        double X0[D * D];
        lu_inverse_small<D>(AB, X0);
        for(int j = 0; j < D; j++)
            for(int i = 0; i < D; i++) invJ[i + j * c] = X0[i + j * D];
        if(nv == 2) {
            double X1[D * D], Tm[D * D];
            lu_inverse_small<D>(AB + (nv + 1) * D * D, X1);
            const double *A1 = AB + D * D;
            for(int j = 0; j < D; j++)
                for(int i = 0; i < D; i++) {
                    double s = 0;
                    for(int q = 0; q < D; q++) s += A1[i + q * D] * X0[q + j * D];
                    Tm[i + j * D] = s;
                }
            for(int j = 0; j < D; j++)
                for(int i = 0; i < D; i++) {
                    double s = 0;
                    for(int q = 0; q < D; q++) s += X1[i + q * D] * Tm[q + j * D];
                    invJ[(D + i) + j * c] = -s;
                    invJ[(D + i) + (D + j) * c] = X1[i + j * D];
                }
        }
    }
    gsync<NT>();
    for(int t = tid; t < c * c; t += NT) { // T1 = target * invJ
        const int i = t % c, j = t / c;
        double s = 0;
        for(int q = 0; q < c; q++) s += target[i + q * c] * invJ[q + j * c];
        T1[i + j * c] = s;
    }
    gsync<NT>();
    for(int t = tid; t < c * c; t += NT) { // M2 = invJ^T * T1
        const int i = t % c, j = t / c;
        double s = 0;
        for(int q = 0; q < c; q++) s += invJ[q + i * c] * T1[q + j * c];
        M2[i + j * ldm] = s;
    }
    gsync<NT>();
    // SelfAdjointEigenSolver reads the lower triangle: symmetrise from it
    for(int t = tid; t < c * c; t += NT) {
        const int i = t % c, j = t / c;
        if(i < j) M2[i + j * ldm] = M2[j + i * ldm];
    }
    gsync<NT>();
    const int sweeps = jacobi_eig<NT>(M2, c, ldm, V, ldm, cs, red);
    int rank = 0;
    if(tid == 0) {
        for(int i = 0; i < c; i++) { // ascending order
            int rk = 0;
            const double wi = M2[i + i * ldm];
            for(int j = 0; j < c; j++) {
                const double wj = M2[j + j * ldm];
                rk += (wj < wi) || (wj == wi && j < i);
            }
            order[rk] = i;
        }
        int i0 = 0;
        while(i0 < c && M2[order[i0] + order[i0] * ldm] < 1e-8) i0++; // glc_eps (absolute), :18,:67
        red[1] = (double) (c - i0);
        red[2] = (double) i0;
    }
    gsync<NT>();
    rank = (int) red[1];
    const int i0 = (int) red[2];
    if(sweeps < 0) return -1;
    if(rank > 0) {
        int32_t *si = reinterpret_cast<int32_t *>(slot);
        double *sm = reinterpret_cast<double *>(slot + 1 + spgr_pad2(nvcap));
        double *sw = sm + cslot;
        if(tid == 0) {
            si[0] = nv;
            si[1] = rank;
            for(int i = 0; i < nv; i++) si[2 + i] = kv[i];
        }
        for(int t = tid; t < c; t += NT) sm[t] = meas[t];
        for(int t = tid; t < rank * c; t += NT) { // W = (V_keep sqrt(D_keep))^T, row-major rows of cslot
            const int l = t / c, q = t % c;
            const int col = order[i0 + l];
            sw[(size_t) l * cslot + q] = V[q + col * ldm] * sqrt(M2[col + col * ldm]);
        }
    }
    gsync<NT>();
    return rank;
}

// posdef_pinv (topology_provider_glc.cpp:42-56) of a D x D block A (ld lda): eig, tolerance
// eps * D * max|lambda|, V diag(1/lambda | 0) V^T -> out (D x D, ld D). sc: 2*D*odd(D) + 2D + 16 doubles.
template <int D, int NT>
__device__ void glc_posdef_pinv(const double *A, int lda, double *out, double *sc) {
    const int tid = (NT <= 32) ? (int) (threadIdx.x & 31) : (int) threadIdx.x; // NT = 32: any one warp of the CTA
    constexpr int LD = D | 1;
    double *M = sc, *V = sc + D * LD, *cs = V + D * LD, *red = cs + 2 * D + 4;
    for(int t = tid; t < D * D; t += NT) {
        const int i = t % D, j = t / D;
        M[i + j * LD] = (i >= j) ? A[i + j * lda] : A[j + i * lda]; // eig reads the lower triangle
    }
    gsync<NT>();
    jacobi_eig<NT>(M, D, LD, V, LD, cs, red);
    if(tid == 0) {
        double mx = 0;
        for(int i = 0; i < D; i++) mx = fmax(mx, fabs(M[i + i * LD]));
        red[1] = 2.220446049250313e-16 * D * mx;
    }
    gsync<NT>();
    const double tol = red[1];
    for(int t = tid; t < D * D; t += NT) {
        const int i = t % D, j = t / D;
        double s = 0;
        for(int l = 0; l < D; l++) {
            const double w = M[l + l * LD];
            if(w > tol) s += V[i + l * LD] * V[j + l * LD] / w;
        }
        out[i + j * D] = s;
    }
    gsync<NT>();
}

// getEdge for the Dense topology (one n-ary edge over all kept vertices) or a single kept vertex.
// T (k x k, ldT) is overwritten by J^-T Lambda_t J^-1 and diagonalised in place; buf1/buf2: k x ldk each.
// sc: k + 2*nk*D*D + (k + k/2 + 8) + 8 + k/2 + 1 doubles. Returns rank, or -1 (eig failure).
template <int D, int NT>
__device__ int glc_get_edge_dense(double *T, int k, int ldT, int nk, const double *s_pose_kept, double *buf1, double *buf2,
                                  int ldk, double *sc, uint64_t *slot) {
    constexpr int PS = PoseStride<D>::value;
    const int tid = threadIdx.x;
    double *meas = sc, *AB = meas + k, *cs = AB + 2 * nk * D * D, *red = cs + (k + k / 2 + 8);
    int *order = reinterpret_cast<int *>(red + 8);
    double *invJ = buf1, *tmp = buf2;
    const double *x0 = s_pose_kept;
    for(int i = tid; i < nk; i += NT) glc_measurement<D>(i, x0, s_pose_kept + PS * i, meas + i * D);
    for(int t = tid; t < k * ldk; t += NT) invJ[t] = 0.0;
    gsync<NT>();
    for(int i = tid; i < nk; i += NT)
        glc_reparam_blocks<D>(i, meas, x0, s_pose_kept + PS * i, AB + i * D * D, AB + (nk + i) * D * D);
    gsync<NT>();
    if(tid == 0) {
        double X0[D * D];
        lu_inverse_small<D>(AB, X0);
        for(int j = 0; j < D; j++)
            for(int i = 0; i < D; i++) invJ[i + j * ldk] = X0[i + j * D];
    }
    gsync<NT>();
    for(int v = 1 + tid; v < nk; v += NT) {
        double Xi[D * D], Tm[D * D];
        lu_inverse_small<D>(AB + (nk + v) * D * D, Xi);
        const double *Ai = AB + v * D * D;
        for(int j = 0; j < D; j++)
            for(int i = 0; i < D; i++) {
                double s = 0;
                for(int q = 0; q < D; q++) s += Ai[i + q * D] * invJ[q + j * ldk];
                Tm[i + j * D] = s;
            }
        for(int j = 0; j < D; j++)
            for(int i = 0; i < D; i++) {
                double s = 0;
                for(int q = 0; q < D; q++) s += Xi[i + q * D] * Tm[q + j * D];
                invJ[(v * D + i) + j * ldk] = -s;
                invJ[(v * D + i) + (v * D + j) * ldk] = Xi[i + j * D];
            }
    }
    gsync<NT>();
    for(int t = tid; t < k * k; t += NT) { // tmp = Lambda_t * invJ
        const int i = t % k, j = t / k;
        double s = 0;
        for(int q = 0; q < k; q++) s += T[i + (size_t) q * ldT] * invJ[q + j * ldk];
        tmp[i + j * ldk] = s;
    }
    gsync<NT>();
    for(int t = tid; t < k * k; t += NT) { // M2 = invJ^T * tmp, lower triangle mirrored (eig reads lower)
        const int i = t % k, j = t / k;
        if(i >= j) {
            double s = 0;
            for(int q = 0; q < k; q++) s += invJ[q + i * ldk] * tmp[q + j * ldk];
            T[i + (size_t) j * ldT] = s;
            T[j + (size_t) i * ldT] = s;
        }
    }
    gsync<NT>();
    double *V = buf1;
    const int sweeps = jacobi_eig<NT>(T, k, ldT, V, ldk, cs, red);
    for(int i = tid; i < k; i += NT) {
        const double wi = T[i + (size_t) i * ldT];
        int rk = 0;
        for(int j = 0; j < k; j++) {
            const double wj = T[j + (size_t) j * ldT];
            rk += (wj < wi) || (wj == wi && j < i);
        }
        order[rk] = i;
    }
    gsync<NT>();
    if(tid == 0) {
        int i0 = 0;
        while(i0 < k && T[order[i0] + (size_t) order[i0] * ldT] < 1e-8) i0++;
        red[1] = (double) (k - i0);
        red[2] = (double) i0;
    }
    gsync<NT>();
    const int rank = (int) red[1], i0 = (int) red[2];
    if(sweeps < 0) return -1;
    if(rank > 0) {
        int32_t *si = reinterpret_cast<int32_t *>(slot);
        double *sm = reinterpret_cast<double *>(slot + 1 + spgr_pad2(nk));
        double *sw = sm + k;
        if(tid == 0) { si[0] = nk; si[1] = rank; }
        for(int i = tid; i < nk; i += NT) si[2 + i] = i;
        for(int t = tid; t < k; t += NT) sm[t] = meas[t];
        for(int t = tid; t < rank * k; t += NT) {
            const int l = t / k, q = t % k;
            const int col = order[i0 + l];
            sw[(size_t) l * k + q] = V[q + col * ldk] * sqrt(T[col + (size_t) col * ldT]);
        }
    }
    gsync<NT>();
    return rank;
}

} // namespace spg
