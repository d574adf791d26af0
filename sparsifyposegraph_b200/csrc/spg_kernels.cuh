// spg_kernels.cuh — the fused blanket kernel (sm_100a): one CTA per Markov blanket runs the whole
// per-vertex body of VertexRemover::remove (reference src/vertex_remover.cpp:89-139) on chip:
//   R3  assembly  H = sum_e J^T Omega J                (g2o buildSystem, vertex_remover.cpp:397-409)
//   R4  Schur     Lambda_t = H_kk - H_mk^T H_mm^-1 H_mk (vertex_remover.cpp:443-449)
//   R5  Chow-Liu  C = (Lambda_t + I)^-1, MI weights, max-heap Kruskal (pseudo_chow_liu.cpp)
//   R7  Jacobians of the new edges at zero error       (vertex_remover.cpp:466-498)
//   R8  NFR closed form: eig(Lambda_t), Sigma blocks, X_e = (J Sigma J^T)^-1 (logdet_function.cpp:14-64,236-279)
// H, Lambda_t, C and the eigenvectors never leave shared memory; HBM sees the packed record in and
// the substitute-edge record out (SURVEY.md §8d "compulsory traffic").
#pragma once
#include "../../include/spg_capi.h"
#include "../../include/spg_record.h"
#include "spg_device.cuh"
#include "spg_plan.h"
#include "spg_glc.cuh"

namespace spg {

// std::priority_queue<WeightedEdge> emulation (libstdc++ __push_heap / __adjust_heap), compared on
// weight only (pseudo_chow_liu.h:49-53) so that equal weights pop in the reference's order.
struct HeapView {
    double *w;
    int *ab; // packed a | b << 16
    int len;
    __device__ void push(double vw, int vab) {
        int hole = len++;
        int parent = (hole - 1) / 2;
        while(hole > 0 && w[parent] < vw) {
            w[hole] = w[parent]; ab[hole] = ab[parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        w[hole] = vw; ab[hole] = vab;
    }
    __device__ void pop(double &tw, int &tab) {
        tw = w[0]; tab = ab[0];
        if(len > 1) {
            const int n = len - 1;
            const double vw = w[n];
            const int vab = ab[n];
            w[n] = w[0]; ab[n] = ab[0];
            int hole = 0, second = 0;
            while(second < (n - 1) / 2) {
                second = 2 * (second + 1);
                if(w[second] < w[second - 1]) second--;
                w[hole] = w[second]; ab[hole] = ab[second];
                hole = second;
            }
            if((n & 1) == 0 && second == (n - 2) / 2) {
                second = 2 * (second + 1);
                w[hole] = w[second - 1]; ab[hole] = ab[second - 1];
                hole = second - 1;
            }
            int parent = (hole - 1) / 2;
            while(hole > 0 && w[parent] < vw) {
                w[hole] = w[parent]; ab[hole] = ab[parent];
                hole = parent;
                parent = (hole - 1) / 2;
            }
            w[hole] = vw; ab[hole] = vab;
        }
        len--;
    }
};

// log-determinant pieces of the Chow-Liu weight (pseudo_chow_liu.cpp:169-183), all in registers.
// Lj: Cholesky factor of C_jj is not needed; we use
//   logdet C_{ij,ij} = logdet C_ii + logdet (C_jj - C_ji C_ii^-1 C_ij)
// with Li = chol(C_ii) from shared memory (column-major D x D, lower).
template <int D>
__device__ __forceinline__ double schur_logdet(const double *C, int ld, int i, int j, const double *Li) {
    double Y[D][D]; // Y = Li^-1 * C_ji^T   (C_ji = rows of j, cols of i: the lower-triangle block)
#pragma unroll
    for(int c = 0; c < D; c++) {
#pragma unroll
        for(int r = 0; r < D; r++) {
            double s = C[(j * D + c) + (i * D + r) * ld];
#pragma unroll
            for(int p = 0; p < r; p++) s -= Li[r + p * D] * Y[p][c];
            Y[r][c] = s / Li[r + r * D];
        }
    }
    double S[D][D];
#pragma unroll
    for(int c = 0; c < D; c++)
#pragma unroll
        for(int r = c; r < D; r++) {
            double s = C[(j * D + r) + (j * D + c) * ld];
#pragma unroll
            for(int p = 0; p < D; p++) s -= Y[p][r] * Y[p][c];
            S[r][c] = s;
        }
    // LDL^T without pivoting; logdet = sum log d
    double ld_sum = 0;
    double dinv[D];
#pragma unroll
    for(int c = 0; c < D; c++) {
        double d = S[c][c];
#pragma unroll
        for(int p = 0; p < c; p++) d -= S[c][p] * S[c][p] * dinv[p]; // l_cp^2 d_p, with S[c][p] = l_cp * d_p
        ld_sum += log(d);
        dinv[c] = 1.0 / d;
#pragma unroll
        for(int r = c + 1; r < D; r++) {
            double s = S[r][c];
#pragma unroll
            for(int p = 0; p < c; p++) s -= S[r][p] * S[c][p] * dinv[p];
            S[r][c] = s; // = l_rc * d_c
        }
    }
    return ld_sum;
}

// D x D Cholesky in registers -> Lout (column-major), returns logdet. ok=false if not PD.
template <int D>
__device__ __forceinline__ double chol_small(const double *A, int ld, double *Lout, bool &ok) {
    double L[D][D];
    double lds = 0;
    ok = true;
#pragma unroll
    for(int c = 0; c < D; c++) {
        double d = A[c + c * ld];
#pragma unroll
        for(int p = 0; p < c; p++) d -= L[c][p] * L[c][p];
        if(!(d > 0)) { ok = false; d = 1.0; }
        const double l = sqrt(d);
        L[c][c] = l;
        lds += log(d);
#pragma unroll
        for(int r = c + 1; r < D; r++) {
            double s = A[r + c * ld];
#pragma unroll
            for(int p = 0; p < c; p++) s -= L[r][p] * L[c][p];
            L[r][c] = s / l;
        }
    }
#pragma unroll
    for(int c = 0; c < D; c++)
#pragma unroll
        for(int r = 0; r < D; r++) Lout[r + c * D] = (r >= c) ? L[r][c] : 0.0;
    return lds;
}

// X = A^-1 for a symmetric PD D x D matrix (reads the lower triangle, column-major ld=D), via LLT
// like Eigen's chol.solve(Identity) (logdet_function.cpp:273-274). Xout column-major.
template <int D>
__device__ __forceinline__ bool spd_inverse_small(const double *A, double *Xout) {
    double L[D * D];
    bool ok;
    chol_small<D>(A, D, L, ok);
    if(!ok) return false;
#pragma unroll
    for(int c = 0; c < D; c++) {
        double x[D];
#pragma unroll
        for(int i = 0; i < D; i++) {
            double s = (i == c) ? 1.0 : 0.0;
#pragma unroll
            for(int p = 0; p < i; p++) s -= L[i + p * D] * x[p];
            x[i] = s / L[i + i * D];
        }
#pragma unroll
        for(int i = D - 1; i >= 0; i--) {
            double s = x[i];
#pragma unroll
            for(int p = i + 1; p < D; p++) s -= L[p + i * D] * x[p];
            x[i] = s / L[i + i * D];
        }
#pragma unroll
        for(int i = 0; i < D; i++) Xout[i + c * D] = x[i];
    }
    return true;
}


} // namespace spg
#include "spg_nfr.cuh"
namespace spg {

// pair index (a | b << 16) helpers
__device__ __forceinline__ int pk(int a, int b) { return a | (b << 16); }
__device__ __forceinline__ int pk_a(int v) { return v & 0xffff; }
__device__ __forceinline__ int pk_b(int v) { return v >> 16; }

// H += J^T J for one GLC edge, J = W * J_reparam (Omega = I). Returns false (uniformly) when the
// scratch does not fit.
template <int D, int NT>
__device__ bool assemble_glc_edge(const uint64_t *ew, const double *s_pose, double *H, int ldH, double *scratch, int cap) {
    constexpr int PS = PoseStride<D>::value;
    const int tid = threadIdx.x;
    const int32_t *eh = reinterpret_cast<const int32_t *>(ew);
    const int nv = eh[1], r = eh[2], c = D * nv;
    const int32_t *vi = reinterpret_cast<const int32_t *>(ew + 2);
    const double *meas = reinterpret_cast<const double *>(ew + 2 + spgr_pad2(nv));
    const double *W = meas + c; // r x c row-major
    double *JA = scratch, *JB = scratch + nv * D * D, *Jf = scratch + 2 * nv * D * D; // Jf: r x c column-major (ld r)
    if(2 * nv * D * D + r * c > cap) return false;
    for(int i = tid; i < nv; i += NT)
        glc_reparam_blocks<D>(i, meas, s_pose + PS * vi[0], s_pose + PS * vi[i], JA + i * D * D, JB + i * D * D);
    gsync<NT>();
    for(int t = tid; t < r * c; t += NT) {
        const int row = t % r, col = t / r, bi = col / D, j = col % D;
        double s = 0;
        if(bi == 0) {
            for(int i = 0; i < nv; i++)
                for(int q = 0; q < D; q++) s += W[(size_t) row * c + i * D + q] * JA[i * D * D + q + j * D];
        } else {
            for(int q = 0; q < D; q++) s += W[(size_t) row * c + bi * D + q] * JB[bi * D * D + q + j * D];
        }
        Jf[row + (size_t) col * r] = s;
    }
    gsync<NT>();
    for(int t = tid; t < c * c; t += NT) {
        const int a = t % c, b = t / c;
        double s = 0;
        for(int row = 0; row < r; row++) s += Jf[row + (size_t) a * r] * Jf[row + (size_t) b * r];
        H[(vi[a / D] * D + a % D) + (size_t) (vi[b / D] * D + b % D) * ldH] += s;
    }
    gsync<NT>();
    return true;
}

// H += J^T Omega J for one MultiEdgeCorrelated (multi_edge_correlated.hpp:96-140): nmeas pose measurements over the
// nv vertices of the edge, J stacked (rows = D * nmeas), Omega rows x rows. Returns false (uniformly) when the scratch
// (J and Omega J, rows x D nv each) does not fit.
template <int D, int NT>
__device__ bool assemble_multi_edge(const uint64_t *ew, const double *s_pose, double *H, int ldH, double *scratch, int cap) {
    constexpr int PS = PoseStride<D>::value, PW = (D == 6) ? 7 : 3;
    const int tid = threadIdx.x;
    const int32_t *eh = reinterpret_cast<const int32_t *>(ew);
    const int nv = eh[1], rows = eh[2], nm = rows / D, c = D * nv;
    const int32_t *vi = reinterpret_cast<const int32_t *>(ew + 2);
    const int32_t *pr = reinterpret_cast<const int32_t *>(ew + 2 + spgr_pad2(nv));
    const double *meas = reinterpret_cast<const double *>(ew + 2 + spgr_pad2(nv) + spgr_pad2(2 * nm));
    const double *Om = meas + (size_t) nm * PW; // rows x rows, column-major
    double *Jf = scratch, *M = scratch + (size_t) rows * c;
    if(2 * (int64_t) rows * c > cap) return false;
    for(int t = tid; t < rows * c; t += NT) Jf[t] = 0.0;
    gsync<NT>();
    for(int p = tid; p < nm; p += NT) {
        double Z[PS], J[2 * D * D];
        if constexpr(D == 6) se3_from_flat(meas + (size_t) p * PW, Z);
        else se2_from_flat(meas + (size_t) p * PW, Z);
        const int a = pr[2 * p], b = pr[2 * p + 1];
        edge_jacobians<D>(Z, s_pose + PS * vi[a], s_pose + PS * vi[b], J);
        for(int q = 0; q < D; q++)
            for(int r = 0; r < D; r++) {
                Jf[(p * D + r) + (size_t) (a * D + q) * rows] = J[r + q * D];
                Jf[(p * D + r) + (size_t) (b * D + q) * rows] = J[D * D + r + q * D];
            }
    }
    gsync<NT>();
    for(int t = tid; t < rows * c; t += NT) {
        const int row = t % rows, col = t / rows;
        double s = 0;
        for(int q = 0; q < rows; q++) s += Om[row + (size_t) q * rows] * Jf[q + (size_t) col * rows];
        M[t] = s;
    }
    gsync<NT>();
    for(int t = tid; t < c * c; t += NT) {
        const int a = t % c, b = t / c;
        double s = 0;
        for(int row = 0; row < rows; row++) s += Jf[row + (size_t) a * rows] * M[row + (size_t) b * rows];
        H[(vi[a / D] * D + a % D) + (size_t) (vi[b / D] * D + b % D) * ldH] += s;
    }
    gsync<NT>();
    return true;
}

#define SPG_T(i)                                                                                     \
    do {                                                                                             \
        if(P.prof && tid == 0) {                                                                     \
            const long long t_now = clock64();                                                       \
            atomicAdd(&P.prof[i], (unsigned long long) (t_now - t_last));                            \
            t_last = t_now;                                                                          \
        }                                                                                            \
    } while(0)

// SPILL = true: the variant for blankets whose working set exceeds shared memory. Same code; the record is
// read in place and every buffer of the shared-memory plan lives in a per-CTA slice of a global workspace
// (served by L1/L2). A separate instantiation so that the on-chip variant keeps LDS/STS addressing.
// LEAN = true: 256 threads at 128 registers with a small buf2, two CTAs per SM, for NFR rounds made of POSE edges
// only (N <= 80): two independent pivot chains per SM.
template <int D, int NT, bool SPILL = false, bool LEAN = false>
__global__ void __launch_bounds__(NT, LEAN ? 2 : ((NT >= 256) ? 1 : (NT == 128 ? 3 : 512 / NT))) blanket_kernel(const KernelParams P) {
    extern __shared__ double smem_dyn[];
    double *smem = smem_dyn;
    if constexpr(SPILL) smem = P.gws + (size_t) blockIdx.x * (size_t) P.gws_stride;
    constexpr int PS = PoseStride<D>::value;
    constexpr int PW = (D == 6) ? 7 : 3; // pose words in records
    constexpr int JW = D * 2 * D;        // doubles of one edge Jacobian [Ji Jj]
    constexpr int SW = 4 * D * D;        // doubles of one 2D x 2D block
    const int tid = threadIdx.x;

    uint64_t *s_rec_buf = reinterpret_cast<uint64_t *>(smem);
    double *s_pose = smem + P.off_pose;
    double *buf0 = smem + P.off_buf0;
    double *buf1 = smem + P.off_buf1;
    double *buf2 = smem + P.off_buf2;
    double *s_small = smem + P.off_small;
    const int kmax = D * (P.max_nv - 1);
    const int nkmax = P.max_nv - 1;
    const int pairs_max = nkmax * (nkmax - 1) / 2;
    // carve the small region (sizes must match plan_smem)
    double *s_w = s_small;
    int *s_order = reinterpret_cast<int *>(s_w + kmax);
    double *s_cs = s_w + kmax + (kmax + 1) / 2;
    double *s_red = s_cs + (kmax + kmax / 2 + 4);
    double *s_wt = s_red + 34;
    double *s_heapw = s_wt + pairs_max;
    int *s_heapab = reinterpret_cast<int *>(s_heapw + pairs_max);
    int *s_tree = reinterpret_cast<int *>(s_heapw + 2 * pairs_max);
    int *s_uf = reinterpret_cast<int *>(s_heapw + 2 * pairs_max + (pairs_max > 0 ? pairs_max : 1));
    double *s_Lfac = s_heapw + 2 * pairs_max + (pairs_max > 0 ? pairs_max : 1) + (nkmax + 1) / 2 + 1;
    double *s_logd = s_Lfac + nkmax * D * D;
    int *s_misc = reinterpret_cast<int *>(s_logd + nkmax);

    const int n_list = P.n_list_dev ? *P.n_list_dev : P.n_list; // retry launches: the count is on the device
    for(int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int b = P.list ? P.list[li] : li;
        const uint64_t *grec = P.records + P.rec_off[b];
        uint64_t *gout = P.out + P.out_off[b];
        const int out_words = (int) (P.out_off[b + 1] - P.out_off[b]);
        const int32_t *gh = reinterpret_cast<const int32_t *>(grec);
        const int nv = gh[0], nrem = gh[1], ne = gh[2], rdim = gh[3], rec_words = gh[4];
        const int nk = nv - nrem;
        const int N = D * nv, m = D * nrem, k = D * nk;
        const int ldH = odd_ld(N);
        const int ldk = odd_ld(k > 0 ? k : 1);

        int status = SPG_BLANKET_OK;
        int newton_iters = 0, out_flags = 0;
        double out_kld = 0;
        if(rdim != D || nv > P.max_nv || ne > P.max_e || rec_words > P.max_rec_words || nrem < 1 || nk < 0)
            status = SPG_BLANKET_TOO_LARGE;

        // zero the output record (unused slots must read as zeros)
        for(int t = tid; t < out_words; t += NT) gout[t] = 0;
        gsync<NT>(); // also protects the shared buffers of the previous blanket

        if(status != SPG_BLANKET_OK) {
            if(tid == 0) reinterpret_cast<int32_t *>(gout)[0] = status;
            continue;
        }

        long long t_last = clock64();
        // ---- S0: stage the record in shared memory, expand poses ---------------------------------
        const uint64_t *s_rec = SPILL ? grec : s_rec_buf;
        if constexpr(!SPILL)
            for(int t = tid; t < rec_words; t += NT) s_rec_buf[t] = grec[t];
        if(tid == 0) { s_misc[0] = SPG_BLANKET_OK; s_misc[1] = 0; }
        gsync<NT>();
        const double *r_pose = reinterpret_cast<const double *>(s_rec + spgr_poses_off(nv));
        const int32_t *r_etab = reinterpret_cast<const int32_t *>(s_rec + spgr_edgetab_off(D, nv));
        for(int v = tid; v < nv; v += NT) {
            if constexpr(D == 6) se3_from_flat(r_pose + PW * v, s_pose + PS * v);
            else se2_from_flat(r_pose + PW * v, s_pose + PS * v);
        }
        for(int t = tid; t < N * ldH; t += NT) buf0[t] = 0.0;
        gsync<NT>();

        SPG_T(0);
        // ---- S1: assembly H = sum_e J^T Omega J (edge order = record order) -------------------------
        double *H = buf0;
        const int asm_chunk = max(1, min(32, P.buf1_doubles / (2 * JW))); // edges linearised per pass
        int e0 = 0;
        while(e0 < ne) {
            const int kind0 = reinterpret_cast<const int32_t *>(s_rec + r_etab[e0])[0];
            if(kind0 != SPG_EDGE_POSE) {
                // GLCEdge (J = W * J_reparam, Omega = I; glc_edge.cpp:40-49) — one edge at a time, whole CTA.
                // Scratch: buf1 and buf2 are contiguous and both free during assembly.
                bool okg = false;
                if(kind0 == SPG_EDGE_GLC)
                    okg = assemble_glc_edge<D, NT>(s_rec + r_etab[e0], s_pose, H, ldH, buf1, P.buf1_doubles + P.buf2_doubles);
                else if(kind0 == SPG_EDGE_MULTI)
                    okg = assemble_multi_edge<D, NT>(s_rec + r_etab[e0], s_pose, H, ldH, buf1, P.buf1_doubles + P.buf2_doubles);
                if(!okg) status = SPG_BLANKET_UNSUPPORTED; // uniform
                e0++;
                continue;
            }
            int ce = 1;
            while(ce < asm_chunk && e0 + ce < ne &&
                  reinterpret_cast<const int32_t *>(s_rec + r_etab[e0 + ce])[0] == SPG_EDGE_POSE) ce++;
            for(int e = tid; e < ce; e += NT) { // linearise: one thread per edge
                const uint64_t *ew = s_rec + r_etab[e0 + e];
                const int32_t *eh = reinterpret_cast<const int32_t *>(ew);
                double *J = buf1 + (size_t) e * 2 * JW;
                if(eh[0] == SPG_EDGE_POSE && eh[1] == 2) {
                    const int32_t *vi = reinterpret_cast<const int32_t *>(ew + 2);
                    const double *pl = reinterpret_cast<const double *>(ew + 3);
                    double Z[PS];
                    if constexpr(D == 6) se3_from_flat(pl, Z);
                    else se2_from_flat(pl, Z);
                    edge_jacobians<D>(Z, s_pose + PS * vi[0], s_pose + PS * vi[1], J);
                } else {
                    s_misc[0] = SPG_BLANKET_UNSUPPORTED;
                    for(int q = 0; q < JW; q++) J[q] = 0;
                }
            }
            gsync<NT>();
            for(int t = tid; t < ce * JW; t += NT) { // M = Omega * J  (D x 2D per edge)
                const int e = t / JW, q = t % JW, r = q % D, c = q / D;
                const uint64_t *ew = s_rec + r_etab[e0 + e];
                const double *Om = reinterpret_cast<const double *>(ew + 3) + PW; // column-major D x D
                const double *J = buf1 + (size_t) e * 2 * JW;
                double s = 0;
#pragma unroll
                for(int p = 0; p < D; p++) s += Om[r + p * D] * J[p + c * D];
                buf1[(size_t) e * 2 * JW + JW + q] = s;
            }
            gsync<NT>();
            for(int e = 0; e < ce; e++) { // sequential over edges: fixed summation order
                const uint64_t *ew = s_rec + r_etab[e0 + e];
                const int32_t *vi = reinterpret_cast<const int32_t *>(ew + 2);
                const int va = vi[0], vb = vi[1];
                const double *J = buf1 + (size_t) e * 2 * JW;
                const double *M = J + JW;
                for(int t = tid; t < SW; t += NT) {
                    const int r = t % (2 * D), c = t / (2 * D);
                    double s = 0;
#pragma unroll
                    for(int p = 0; p < D; p++) s += J[p + r * D] * M[p + c * D];
                    const int gr = (r < D ? va * D + r : vb * D + r - D);
                    const int gc = (c < D ? va * D + c : vb * D + c - D);
                    H[gr + gc * ldH] += s;
                }
                gsync<NT>();
            }
            e0 += ce;
        }
        if(s_misc[0] != SPG_BLANKET_OK) status = s_misc[0];

        SPG_T(1);
        // ---- S2: Schur complement onto the kept variables ----------------------------------------
        double *T = H + m + (size_t) m * ldH; // Lambda_t in place, leading dimension ldH
        if(status == SPG_BLANKET_OK) {
            // register-tiled partial sweep of the m removed pivots: H_kk <- H_kk - H_km H_mm^-1 H_mk
            // (the store mirrors the upper triangle onto the lower one, vertex_remover.cpp:447-449)
            const int sw = sweep_spd_auto<D, NT, LEAN>(H, ldH, H, ldH, N, m, buf1, 0.0, true);
            if(sw == 0) status = SPG_BLANKET_NOT_PD_MARGINAL;
            if(sw < 0) { // blanket larger than the register tiles: LLT route
                if(!chol_lower<NT>(H, m, ldH)) status = SPG_BLANKET_NOT_PD_MARGINAL;
                if(status == SPG_BLANKET_OK && k > 0) {
                    for(int c = m + tid; c < N; c += NT) { // Y = L^-1 H_mk : thread per kept column
                        double *y = H + (size_t) c * ldH;
                        for(int i = 0; i < m; i++) {
                            double s = y[i];
                            for(int p = 0; p < i; p++) s -= H[i + p * ldH] * y[p];
                            y[i] = s / H[i + i * ldH];
                        }
                    }
                    gsync<NT>();
                    for(int t = tid; t < k * k; t += NT) {
                        const int i = t % k, j = t / k;
                        if(i <= j) {
                            const double *yi = H + (size_t) (m + i) * ldH, *yj = H + (size_t) (m + j) * ldH;
                            double s = 0;
                            for(int p = 0; p < m; p++) s += yi[p] * yj[p];
                            const double v = T[i + (size_t) j * ldH] - s;
                            T[i + (size_t) j * ldH] = v;
                            if(i != j) T[j + (size_t) i * ldH] = v;
                        }
                    }
                    gsync<NT>();
                }
            }
        }
        if(status == SPG_BLANKET_OK && k > 0) {
            if(P.dbg_target) {
                double *g = P.dbg_target + P.dbg_target_off[b];
                if(P.dbg_target_off[b + 1] - P.dbg_target_off[b] >= (int64_t) k * k)
                    for(int t = tid; t < k * k; t += NT) g[t] = T[(t % k) + (size_t) (t / k) * ldH];
            }
        }

        SPG_T(2);
        // ---- S3: sparsity pattern (PseudoChowLiu::computeSparsityPattern, pseudo_chow_liu.cpp:33-87) ---
        int n_out = 0;
        double glog_part = 0; // this thread's share of logdet(Lambda_rr) (pivots of the anchored sweep): projected KLD of the closed form
        bool g_ready = false; // Lambda_rr^-1 already sits in buf2 (computed next to the Chow-Liu inverse)
        const bool glc_tree = (P.algorithm == SPG_ALG_GLC && P.topology == SPG_TOPO_TREE);
        if(status == SPG_BLANKET_OK && (P.algorithm == SPG_ALG_NFR || glc_tree) && nk >= 2) {
            const int mch = (int) ((1 + P.chord_ratio) * (nk - 1));
            const int all = nk * (nk - 1) / 2;
            const bool full = mch >= all;
            const bool cliquey = (P.topology == SPG_TOPO_CLIQUEY_DENSE || P.topology == SPG_TOPO_CLIQUEY_SUBGRAPH);
            if(nk == 2) {
                n_out = 1;
                if(tid == 0) { s_tree[0] = pk(0, 1); s_heapab[0] = 0; s_misc[14] = 1; }
            } else if(P.topology == SPG_TOPO_DENSE || (P.topology == SPG_TOPO_SUBGRAPH && full)) {
                n_out = all;
                for(int t = tid; t < all; t += NT) {
                    int i = 0, rem = t;
                    while(rem >= nk - 1 - i) { rem -= nk - 1 - i; i++; }
                    s_tree[t] = pk(i, i + 1 + rem);
                }
            } else {
                n_out = (P.topology == SPG_TOPO_TREE || cliquey) ? nk - 1 : mch; // correlated topologies group the spanning tree
                // C = (Lambda_t + 1 I)^-1  (fillEdges, pseudo_chow_liu.cpp:185-190)
                double *Lc = buf1, *C = buf1; // buf2 is scratch
                // CTAs twice as wide as the sweep's thread grid invert the anchored block Lambda_rr of the NFR
                // gauge shortcut (S4) at the same time, on their second group of threads: two independent
                // pivot chains per SM instead of one. The sweeps read Lambda_t (+ I) straight from T.
                if constexpr(SweepDual<NT>::value && !SPILL) {
                    const int kk = k - D;
                    if(P.algorithm == SPG_ALG_NFR && n_out * D == kk && !(P.flags & 1) && k <= SweepDual<NT>::NMAX) g_ready = true; // uniform
                }
                if(g_ready) {
                    if constexpr(SweepDual<NT>::value && !SPILL) {
                        constexpr int GS = SweepDual<NT>::GS;
                        const int kk = k - D, ldg = odd_ld(kk > 0 ? kk : 1);
                        double *cb = reinterpret_cast<double *>(s_rec_buf); // the record copy is dead after the assembly
                        if(tid < GS) {
                            const int sw = sweep_spd_group<D, NT, 1>(T, ldH, Lc, ldk, k, k, cb, tid, 1.0);
                            if(tid == 0) s_misc[8] = sw;
                        } else if(tid < 2 * GS) {
                            const int sw = sweep_spd_group<D, NT, 2>(T, ldH, buf2, ldg, kk, kk, cb + SweepDual<NT>::CB, tid - GS, 0.0, &glog_part);
                            if(tid == GS) s_misc[9] = sw;
                        }
                        gsync<NT>();
                        if(s_misc[8] == 0) status = SPG_BLANKET_NOT_PD_CHOWLIU;
                    }
                } else if(sweep_fits<NT, LEAN>(k)) {
                    const int sw = sweep_spd_auto<D, NT, LEAN>(T, ldH, Lc, ldk, k, k, buf2, 1.0, false);
                    if(sw == 0) status = SPG_BLANKET_NOT_PD_CHOWLIU;
                } else { // larger than the register tiles: blocked Cholesky + triangular inverse
                    for(int t = tid; t < k * k; t += NT) {
                        const int i = t % k, j = t / k;
                        Lc[i + j * ldk] = T[i + (size_t) j * ldH] + (i == j ? 1.0 : 0.0);
                    }
                    gsync<NT>();
                    if(!chol_lower<NT>(Lc, k, ldk)) status = SPG_BLANKET_NOT_PD_CHOWLIU;
                    else chol_inverse_inplace<NT>(Lc, k, ldk, buf2);
                }
                if(status == SPG_BLANKET_OK) {
                    SPG_T(3);
                    SPG_T(4);
                    // per-vertex Cholesky of the diagonal blocks + their log-determinants
                    for(int v = tid; v < nk; v += NT) {
                        bool ok;
                        s_logd[v] = chol_small<D>(C + (size_t) v * D + (size_t) v * D * ldk, ldk, s_Lfac + v * D * D, ok);
                        if(!ok) s_misc[0] = SPG_BLANKET_NOT_PD_CHOWLIU;
                    }
                    gsync<NT>();
                    SPG_T(5);
                    // weight(i,j) = logdet C_ii + logdet C_jj - logdet C_{ij,ij}   (:169-183)
                    //             = logdet C_jj - logdet (C_jj - C_ji C_ii^-1 C_ij)
                    for(int t = tid; t < all; t += NT) {
                        int i = 0, rem = t;
                        while(rem >= nk - 1 - i) { rem -= nk - 1 - i; i++; }
                        const int j = i + 1 + rem;
                        s_wt[t] = s_logd[j] - schur_logdet<D>(C, ldk, i, j, s_Lfac + i * D * D);
                    }
                    gsync<NT>();
                    if(s_misc[0] != SPG_BLANKET_OK) status = s_misc[0];
                    if(P.dbg_weights) {
                        double *g = P.dbg_weights + P.dbg_weights_off[b];
                        const int cap = (int) (P.dbg_weights_off[b + 1] - P.dbg_weights_off[b]);
                        if(P.flags & SPG_OPT_DBG_WEIGHTS_IN) { // test hook: weights are an input (exact ties)
                            for(int t = tid; t < all && t < cap; t += NT) s_wt[t] = g[t];
                            gsync<NT>();
                        } else {
                            for(int t = tid; t < cap; t += NT) g[t] = t < all ? s_wt[t] : 0.0;
                        }
                    }
                    SPG_T(6);
                    // doKruskal (:253-289): pops of a max-heap keyed on the weight only, accepted edges first
                    // then rejected. Without exactly equal (or NaN) weights the pop order is simply the
                    // descending order, which all threads establish by ranking; otherwise one thread
                    // replays libstdc++'s heap so ties break as in the reference.
                    int *s_sorted = s_tree + all; // second half of the tree scratch
                    if(tid == 0) s_misc[3] = 0;
                    gsync<NT>();
                    for(int t = tid; t < all; t += NT) {
                        const double wt = s_wt[t];
                        int rank = 0, tie = (wt != wt);
                        for(int u = 0; u < all; u++) {
                            const double wu = s_wt[u];
                            rank += (wu > wt);
                            tie |= (u != t) && (wu == wt);
                        }
                        if(tie) s_misc[3] = 1;
                        else {
                            int i = 0, rem = t;
                            while(rem >= nk - 1 - i) { rem -= nk - 1 - i; i++; }
                            s_sorted[rank] = pk(i, i + 1 + rem);
                        }
                    }
                    gsync<NT>();
                    if(s_misc[3]) out_flags |= 128; // diagnostic: exactly equal weights, heap replayed
                    if(tid == 0) {
                        for(int v = 0; v < nk; v++) s_uf[v] = v;
                        int nacc = 0, nrej = 0;
                        if(!s_misc[3]) {
                            const bool tree_only = (n_out == nk - 1);
                            for(int q = 0; q < all; q++) {
                                const int ab = s_sorted[q];
                                int ra = pk_a(ab), rb = pk_b(ab);
                                while(s_uf[ra] != ra) ra = s_uf[ra];
                                while(s_uf[rb] != rb) rb = s_uf[rb];
                                if(ra != rb) {
                                    s_uf[rb] = ra;
                                    s_tree[nacc++] = ab;
                                    if(tree_only && nacc == nk - 1) break; // only the spanning tree is used
                                } else {
                                    s_heapab[nrej++] = ab;
                                }
                            }
                            if(!tree_only)
                                for(int q = 0; q < nrej && nacc + q < n_out; q++) s_tree[nacc + q] = s_heapab[q];
                        } else {
                            HeapView hp{s_heapw, s_heapab, 0};
                            int t = 0;
                            for(int i = 0; i < nk - 1; i++)
                                for(int j = i + 1; j < nk; j++, t++) hp.push(s_wt[t], pk(i, j));
                            // rejected edges are staged at the tail of the heap storage, then appended
                            while(hp.len > 0) {
                                double w; int ab;
                                hp.pop(w, ab);
                                int ra = pk_a(ab), rb = pk_b(ab);
                                while(s_uf[ra] != ra) ra = s_uf[ra];
                                while(s_uf[rb] != rb) rb = s_uf[rb];
                                if(ra != rb) {
                                    s_uf[rb] = ra;
                                    s_tree[nacc++] = ab;
                                } else {
                                    s_heapab[all - 1 - nrej] = ab; // behind the live heap: free slots
                                    nrej++;
                                }
                            }
                            for(int q = 0; q < nrej && nacc + q < all; q++) s_tree[nacc + q] = s_heapab[all - 1 - q];
                        }
                    }
                    gsync<NT>();
                    if(cliquey) {
                        // Correlated skeleton trees (pseudo_chow_liu.cpp:60-68, fillCliques :198-251): the n-1 tree edges, in
                        // acceptance order, are grouped into cliques; s_heapab[e] = clique of tree edge e, s_misc[14] =
                        // number of cliques (-1: not representable here). Integer logic, one thread.
                        if(tid == 0) {
                            int *clq = s_heapab;
                            const int ne_t = nk - 1;
                            if(P.topology == SPG_TOPO_CLIQUEY_DENSE || full) {
                                for(int e = 0; e < ne_t; e++) clq[e] = 0;
                                s_misc[14] = 1;
                            } else if(nk > 64) {
                                s_misc[14] = -1;
                            } else {
                                unsigned long long cm[64];
                                int ncl = ne_t;
                                for(int e = 0; e < ne_t; e++) cm[e] = (1ull << pk_a(s_tree[e])) | (1ull << pk_b(s_tree[e]));
                                bool joined = true;
                                for(int nedges = nk - 1, maxfill = 1; nedges < mch && joined; maxfill++) {
                                    joined = false;
                                    int minfill = 0x7fffffff;
                                    for(int i = 0; i < ncl; i++)
                                        for(int j = i + 1; j < ncl; j++)
                                            if(cm[i] & cm[j]) {
                                                const int thisfill = (__popcll(cm[i]) - 1) * (__popcll(cm[j]) - 1);
                                                minfill = min(thisfill, minfill);
                                                if(thisfill <= maxfill && nedges + thisfill <= mch) {
                                                    cm[i] |= cm[j];
                                                    nedges += thisfill;
                                                    for(int q = j; q + 1 < ncl; q++) cm[q] = cm[q + 1];
                                                    ncl--;
                                                    joined = true;
                                                    j--;
                                                }
                                            }
                                    if(!joined && minfill > maxfill) {
                                        joined = true;
                                        maxfill = minfill - 1;
                                    }
                                }
                                int ok = 1;
                                for(int e = 0; e < ne_t; e++) {
                                    const unsigned long long both = (1ull << pk_a(s_tree[e])) | (1ull << pk_b(s_tree[e]));
                                    int owner = -1, cnt = 0;
                                    for(int j = 0; j < ncl; j++)
                                        if((cm[j] & both) == both) { if(owner < 0) owner = j; cnt++; }
                                    clq[e] = owner;
                                    if(cnt != 1) ok = 0; // an edge inside two cliques: more rows than the closed form has (iterative fit)
                                }
                                s_misc[14] = ok ? ncl : -1;
                            }
                        }
                        gsync<NT>();
                        if(s_misc[14] < 0) status = SPG_BLANKET_UNSUPPORTED;
                    }
                }
            }
        }

        SPG_T(7);
        // ---- S4: NFR information fit (optimizeInformation, optimizer.cpp:16-81) ------------------------
        if(status == SPG_BLANKET_OK && P.algorithm == SPG_ALG_NFR && n_out > 0) {
            gsync<NT>(); // s_tree visible
            const int r = k - D;
            const bool closed = (n_out * D == r); // hasClosedFormSolution, logdet_function.cpp:83-86
            double *V = buf1;
            int smalleigs = 0;
            // -- gauge shortcut ---------------------------------------------------------------------------
            // A blanket made only of relative-pose edges is gauge free: Lambda_t has a d-dimensional null
            // space N and every new-edge Jacobian satisfies J N = 0. The reference's
            // Sigma = U S U^T (d smallest eigen-directions dropped, logdet_function.cpp:33-41) is then the
            // pseudo-inverse, and J Sigma J^T = J G J^T for ANY generalised inverse G of Lambda_t. We take
            // G = [[Lambda_rr^-1, 0], [0, 0]] (last kept vertex anchored): one Cholesky instead of an
            // eigen-decomposition. The shortcut is taken only when the reference would be in the same
            // branch: (i) ||G||_F <= 1e5 => largest eigenvalue of G <= 1e5 => smallest eigenvalue of Lambda_rr >= 1e-5 => by interlacing
            // lambda_{d+1} >= cutoff, i.e. smalleigs <= d; (ii) max diag < 1e8 so the null eigenvalues
            // (~ k eps ||Lambda||) stay below the cutoff, i.e. smalleigs >= d. Otherwise: general path.
            bool fast = false;
            const int kk = k - D, ldg = odd_ld(kk > 0 ? kk : 1);
            double *G = g_ready ? buf2 : buf1; // the inverse overwrites the Cholesky factor; buf2 is its scratch
            if(closed && !(P.flags & 1)) {
                // guard (ii): every diagonal entry of Lambda_t below 1e8 (checked by all threads, OR-reduced)
                int bigdiag = 0;
                for(int i = tid; i < k; i += NT) bigdiag |= !(fabs(T[i + (size_t) i * ldH]) < 1e8);
                bigdiag = gsync_or<NT>(bigdiag);
                int swg;
                if(g_ready) swg = s_misc[9];
                else if(sweep_fits<NT, LEAN>(kk)) swg = sweep_spd_auto<D, NT, LEAN>(T, ldH, buf1, ldg, kk, kk, buf2, 0.0, false, &glog_part);
                else { // larger than the register tiles
                    for(int t = tid; t < kk * kk; t += NT) {
                        const int i = t % kk, j = t / kk;
                        buf1[i + j * ldg] = T[i + (size_t) j * ldH];
                    }
                    gsync<NT>();
                    swg = chol_lower<NT>(buf1, kk, ldg) ? 1 : 0;
                    if(swg) {
                        for(int i = tid; i < kk; i += NT) glog_part += 2.0 * log(buf1[i + (size_t) i * ldg]);
                        chol_inverse_inplace<NT>(buf1, kk, ldg, buf2);
                    }
                }
                if(swg <= 0) out_flags |= 32; // diagnostic: anchored block not positive definite
                if(bigdiag) out_flags |= 64;
                if(swg > 0) {
                    SPG_T(8);
                    // guard (i): ||G||_F <= 1e5  (every thread gets the same sum: fixed reduction order). The trace
                    // bound refuses the hubs of a decimated grid (trace(G) ~ 3e4 at 28 vertices, > 1e5 beyond 35,
                    // while the largest eigenvalue stays below 5e3); the Frobenius norm is as cheap and tight enough.
                    double fp = 0;
                    for(int t = tid; t < kk * kk; t += NT) {
                        const double v = G[(t % kk) + (t / kk) * ldg];
                        fp += v * v;
                    }
                    const double frob2 = block_sum<NT>(fp, s_red);
                    fast = (frob2 <= 1e10) && !bigdiag; // false for NaN too
                    if(!fast) out_flags |= 16; // diagnostic: gauge shortcut refused, general eigen path taken
                }
            }
            if(!fast) {
                // eig(Lambda_t): A = T in place (destroyed), V in buf1
                const int sweeps = jacobi_eig<NT>(T, k, ldH, V, ldk, s_cs, s_red);
                if(sweeps < 0) status = SPG_BLANKET_EIG_NOCONV;
                for(int i = tid; i < k; i += NT) s_w[i] = T[i + (size_t) i * ldH];
                gsync<NT>();
                // ascending order (SelfAdjointEigenSolver sorts increasingly); ties by index
                int mysmall = 0;
                for(int i = tid; i < k; i += NT) {
                    const double wi = s_w[i];
                    int rank = 0;
                    for(int j = 0; j < k; j++) {
                        const double wj = s_w[j];
                        rank += (wj < wi) || (wj == wi && j < i);
                    }
                    s_order[rank] = i;
                    if(wi < 1e-5) mysmall++;
                }
                if(mysmall) atomicAdd(&s_misc[1], mysmall);
                gsync<NT>();
                smalleigs = s_misc[1];
            }
            SPG_T(9);
            // new-edge Jacobians at the linearisation point with measurement == state (:466-498),
            // measurements written straight to the output record
            double *Jn = buf0;                       // n_out * JW   (buf0 is free: eigenvalues are saved)
            double *Sg = buf0 + (size_t) n_out * JW; // n_out * SW
            double *Bk = Sg + (size_t) n_out * SW;   // n_out * D*D
            const int slot = 1 + PW + D * D;
            const bool cliquey4 = (P.topology == SPG_TOPO_CLIQUEY_DENSE || P.topology == SPG_TOPO_CLIQUEY_SUBGRAPH);
            for(int e = tid; e < n_out; e += NT) {
                const int a = pk_a(s_tree[e]), bb = pk_b(s_tree[e]);
                const double *Xa = s_pose + PS * (nrem + a), *Xb = s_pose + PS * (nrem + bb);
                double Z[PS], Ti[PS];
                if constexpr(D == 6) { se3_inverse(Xa, Ti); se3_compose(Ti, Xb, Z); }
                else { se2_inverse(Xa, Ti); se2_compose(Ti, Xb, Z); }
                edge_jacobians_zero_error<D>(Z, Xa, Xb, Jn + (size_t) e * JW);
                if(cliquey4) continue; // correlated topologies write their entries after the fit
                uint64_t *sl = gout + SPG_OUT_HEADER_WORDS + (size_t) e * slot;
                int32_t *si = reinterpret_cast<int32_t *>(sl);
                si[0] = a; si[1] = bb;
                double *sm = reinterpret_cast<double *>(sl + 1);
                if constexpr(D == 6) se3_to_flat(Z, sm);
                else { sm[0] = Z[0]; sm[1] = Z[1]; sm[2] = Z[2]; }
            }
            // S (inverse eigenvalues) and the kept eigen-directions, logdet_function.cpp:33-61
            double *s_S = s_cs; // r values (cs is free after the eigen-solver)
            int ooff = 0; // kept direction l is eigen-column s_order[ooff + l]
            if(fast) {
                gsync<NT>();
            } else if(smalleigs <= D) {
                ooff = D;
                for(int l = tid; l < r; l += NT) s_S[l] = 1.0 / s_w[s_order[D + l]];
                gsync<NT>();
            } else {
                // chooseDimensions (:66-81): among the `smalleigs` smallest directions drop the D with
                // the smallest || J u ||  (sparseJacobian drops |J| < eps entries, :335)
                gsync<NT>(); // Jn complete
                double *cn = Bk; // || J u_c || per candidate (Bk is not in use yet)
                for(int c = tid; c < smalleigs; c += NT) {
                    const double *u = V + (size_t) s_order[c] * ldk;
                    double acc = 0;
                    for(int e = 0; e < n_out; e++) {
                        const int a = pk_a(s_tree[e]), bb = pk_b(s_tree[e]);
                        const double *J = Jn + (size_t) e * JW;
                        for(int row = 0; row < D; row++) {
                            double s = 0;
                            for(int q = 0; q < D; q++) {
                                const double ja = J[row + q * D], jb = J[row + (D + q) * D];
                                if(fabs(ja) >= 2.220446049250313e-16) s += ja * u[a * D + q];
                                if(fabs(jb) >= 2.220446049250313e-16) s += jb * u[bb * D + q];
                            }
                            acc += s * s;
                        }
                    }
                    cn[c] = sqrt(acc);
                }
                gsync<NT>();
                if(tid == 0) {
                    // std::sort of (norm, index) pairs ascending; the first D are dropped
                    const double wmax = s_w[s_order[k - 1]];
                    int *drop = s_misc + 4;
                    int nd = 0;
                    for(int d0 = 0; d0 < D; d0++) {
                        int best = -1;
                        for(int c = 0; c < smalleigs; c++) {
                            bool used = false;
                            for(int q = 0; q < nd; q++) used |= (drop[q] == c);
                            if(used) continue;
                            if(best < 0 || cn[c] < cn[best]) best = c;
                        }
                        drop[nd++] = best;
                    }
                    int jj = 0;
                    for(int i = 0; i < k; i++) { // in-place compaction (jj <= i)
                        bool dropped = false;
                        for(int q = 0; q < nd; q++) dropped |= (drop[q] == i);
                        if(!dropped) {
                            const int col = s_order[i];
                            s_S[jj] = fmin(fabs(1.0 / s_w[col]), 1e6 / wmax);
                            s_order[jj] = col;
                            jj++;
                        }
                    }
                }
                gsync<NT>();
            }
            if(status == SPG_BLANKET_OK && !closed) {
                // ---- R9: interior-point / Newton loop (Subgraph, Dense with >= 3 kept vertices) --------
                gsync<NT>(); // Jn, S, order complete
                if(P.nfr_ws == nullptr || nfr_work_doubles(n_out, D, r) > P.nfr_ws_stride) {
                    status = SPG_BLANKET_TOO_LARGE;
                } else {
                    double *ws = P.nfr_ws + (size_t) blockIdx.x * P.nfr_ws_stride;
                    int fl = 0;
                    double kld = 0;
                    nfr_iterative<D, NT>(ws, n_out, k, r, V, ldk, s_order + ooff, s_S, Jn, s_tree, s_red, s_misc + 2, newton_iters, fl, kld);
                    out_flags |= fl;
                    out_kld = kld;
                    if(fl & 2) status = SPG_BLANKET_KLD_INF;
                    const NfrWork w = nfr_carve(ws, n_out, D, r);
                    for(int t = tid; t < n_out * D * D; t += NT) { // decondense: selfadjointView<Lower> (:88-99)
                        const int e = t / (D * D), qq = t % (D * D), i = qq % D, j = qq / D;
                        const double *xe = w.x + (size_t) e * D * D;
                        double *sx = reinterpret_cast<double *>(gout + SPG_OUT_HEADER_WORDS + (size_t) e * slot + 1 + PW);
                        sx[qq] = (i >= j) ? xe[i + j * D] : xe[j + i * D];
                    }
                    gsync<NT>();
                }
            } else if(status == SPG_BLANKET_OK && cliquey4) {
                // ---- correlated topologies (CliqueySubgraph / CliqueyDense): one information block per clique of tree
                // edges, X_c = (J_c Sigma J_c^T)^-1 with J_c the stacked Jacobians of the clique's measurements
                // (closedFormSolution, logdet_function.cpp:236-279; single clique: the sparse-Jacobian branch :243-247).
                // The measurements of a spanning tree always add up to k - d rows, so the closed form always exists.
                gsync<NT>(); // Jn, S, order complete
                const int *clq = s_heapab;
                const int nc = s_misc[14];
                int *map = s_uf; // tree edges of the current clique
                double *Blk = buf0 + (size_t) n_out * JW;          // R x R
                double *Tm = fast ? ((G == buf2) ? buf1 : buf2) : buf2; // fast: J G~ (R x kk); eigen path: J U (R x r); then scratch
                int64_t woff = SPG_OUT_HEADER_WORDS;
                int emitted = 0;
                double lp = fast ? glog_part : 0.0;
                for(int c = 0; c < nc && status == SPG_BLANKET_OK; c++) {
                    if(tid == 0) {
                        int m0 = 0;
                        for(int e = 0; e < n_out; e++)
                            if(clq[e] == c) map[m0++] = e;
                        s_misc[15] = m0;
                    }
                    gsync<NT>();
                    const int m = s_misc[15], R = D * m, ldR = odd_ld(R);
                    if(fast) {
                        for(int t = tid; t < R * kk; t += NT) { // Tm = J_c G~, G~ = [[Lambda_rr^-1, 0], [0, 0]] (lower triangle of G)
                            const int row = t % R, col = t / R, e = map[row / D], i = row % D;
                            const int a = pk_a(s_tree[e]), bb = pk_b(s_tree[e]);
                            const double *J = Jn + (size_t) e * JW;
                            double acc = 0;
#pragma unroll
                            for(int j = 0; j < D; j++) {
                                const int ra = a * D + j, rb = bb * D + j;
                                if(ra < kk) acc += J[i + j * D] * (ra >= col ? G[ra + (size_t) col * ldg] : G[col + (size_t) ra * ldg]);
                                if(rb < kk) acc += J[i + (D + j) * D] * (rb >= col ? G[rb + (size_t) col * ldg] : G[col + (size_t) rb * ldg]);
                            }
                            Tm[row + (size_t) col * R] = acc;
                        }
                        gsync<NT>();
                        for(int t = tid; t < R * R; t += NT) { // B = Tm J_c^T
                            const int row = t % R, col = t / R, e = map[col / D], j = col % D;
                            const int a = pk_a(s_tree[e]), bb = pk_b(s_tree[e]);
                            const double *J = Jn + (size_t) e * JW;
                            double acc = 0;
#pragma unroll
                            for(int l = 0; l < D; l++) {
                                const int ca = a * D + l, cb2 = bb * D + l;
                                if(ca < kk) acc += Tm[row + (size_t) ca * R] * J[j + l * D];
                                if(cb2 < kk) acc += Tm[row + (size_t) cb2 * R] * J[j + (D + l) * D];
                            }
                            Blk[row + (size_t) col * ldR] = acc;
                        }
                    } else {
                        for(int t = tid; t < R * r; t += NT) { // JU = J_c U (kept eigen-directions)
                            const int row = t % R, l = t / R, e = map[row / D], i = row % D;
                            const int a = pk_a(s_tree[e]), bb = pk_b(s_tree[e]);
                            const double *J = Jn + (size_t) e * JW;
                            const double *u = V + (size_t) s_order[ooff + l] * ldk;
                            double acc = 0;
#pragma unroll
                            for(int j = 0; j < D; j++) acc += J[i + j * D] * u[a * D + j] + J[i + (D + j) * D] * u[bb * D + j];
                            Tm[row + (size_t) l * R] = acc;
                        }
                        gsync<NT>();
                        for(int t = tid; t < R * R; t += NT) { // B = JU S JU^T
                            const int row = t % R, col = t / R;
                            double acc = 0;
                            for(int l = 0; l < r; l++) acc += Tm[row + (size_t) l * R] * s_S[l] * Tm[col + (size_t) l * R];
                            Blk[row + (size_t) col * ldR] = acc;
                        }
                    }
                    gsync<NT>();
                    for(int t = tid; t < R * R; t += NT) { // symmetrise (:249-270)
                        const int i = t % R, j = t / R;
                        if(i > j) {
                            const double v = 0.5 * (Blk[i + (size_t) j * ldR] + Blk[j + (size_t) i * ldR]);
                            Blk[i + (size_t) j * ldR] = v;
                            Blk[j + (size_t) i * ldR] = v;
                        }
                    }
                    gsync<NT>();
                    // X_c = B^-1: register-tiled sweep (pivot <= 0 is LLT's failure), blocked Cholesky beyond the tiles
                    double mylog = 0;
                    int sw = -1;
                    if(sweep_fits<NT, LEAN>(R)) sw = sweep_spd_auto<D, NT, LEAN>(Blk, ldR, Blk, ldR, R, R, Tm, 0.0, false, &mylog);
                    if(sw < 0) {
                        sw = chol_lower<NT>(Blk, R, ldR) ? 1 : 0;
                        if(sw) {
                            for(int i = tid; i < R; i += NT) mylog += 2.0 * log(Blk[i + (size_t) i * ldR]);
                            chol_inverse_inplace<NT>(Blk, R, ldR, Tm);
                        }
                    }
                    if(sw == 0) { status = SPG_BLANKET_NOT_PD_CLOSED; break; } // uniform
                    lp += mylog;
                    uint64_t *w = gout + woff;
                    if(tid == 0) {
                        int32_t *wi = reinterpret_cast<int32_t *>(w);
                        wi[0] = m; wi[1] = R;
                        for(int p = 0; p < m; p++) { wi[2 + 2 * p] = pk_a(s_tree[map[p]]); wi[3 + 2 * p] = pk_b(s_tree[map[p]]); }
                    }
                    double *wm = reinterpret_cast<double *>(w + 1 + spgr_pad2(2 * m));
                    for(int p = tid; p < m; p += NT) {
                        const int e = map[p];
                        const double *Xa = s_pose + PS * (nrem + pk_a(s_tree[e])), *Xb = s_pose + PS * (nrem + pk_b(s_tree[e]));
                        double Z[PS], Ti[PS];
                        if constexpr(D == 6) { se3_inverse(Xa, Ti); se3_compose(Ti, Xb, Z); se3_to_flat(Z, wm + (size_t) p * PW); }
                        else { se2_inverse(Xa, Ti); se2_compose(Ti, Xb, Z); wm[p * PW] = Z[0]; wm[p * PW + 1] = Z[1]; wm[p * PW + 2] = Z[2]; }
                    }
                    double *wx = wm + (size_t) m * PW;
                    for(int t = tid; t < R * R; t += NT) {
                        const int i = t % R, j = t / R;
                        wx[t] = (i >= j) ? Blk[i + (size_t) j * ldR] : Blk[j + (size_t) i * ldR]; // selfadjointView<Lower>
                    }
                    woff += spgr_out_entry_words(D, m);
                    emitted++;
                    gsync<NT>();
                }
                n_out = emitted;
                if(status == SPG_BLANKET_OK) {
                    // projected KLD (logdet_function.cpp:119-133): the trace term is exactly r, J^T X J = J_all^T blockdiag(X_c)
                    // J_all with the square, unit-determinant tree Jacobian in anchored coordinates, so
                    // KLD = 1/2 [logdet Lambda_rr + sum_c logdet B_c]. Evaluated on the shortcut path only.
                    if(fast) out_kld = 0.5 * block_sum<NT>(lp, s_red);
                    else out_flags |= 8; // diagnostic: KLD not evaluated (eigen path of a correlated topology)
                }
            } else if(status == SPG_BLANKET_OK) {
                SPG_T(10);
                // Sigma blocks: Sg_e = U[ab,:] S U[ab,:]^T, lower triangle mirrored up (:239-240)
                for(int t = tid; t < n_out * SW; t += NT) {
                    const int e = t / SW, q = t % SW, i = q % (2 * D), j = q / (2 * D);
                    if(i >= j) {
                        const int a = pk_a(s_tree[e]), bb = pk_b(s_tree[e]);
                        const int ri = (i < D ? a * D + i : bb * D + i - D);
                        const int rj = (j < D ? a * D + j : bb * D + j - D);
                        double s = 0;
                        if(fast) {
                            if(ri < kk) s = G[ri + rj * ldg]; // rj <= ri: lower triangle of G
                        } else {
                            for(int l = 0; l < r; l++) {
                                const double *u = V + (size_t) s_order[ooff + l] * ldk;
                                s += u[ri] * s_S[l] * u[rj];
                            }
                        }
                        Sg[(size_t) e * SW + i + j * 2 * D] = s;
                        Sg[(size_t) e * SW + j + i * 2 * D] = s;
                    }
                }
                gsync<NT>();
                // block_e = J Sg_e J^T, symmetrised (:249-270), in two stages: Tm = J Sg_e (D x 2D), then Tm J^T
                double *Tm = buf2; // n_out * JW doubles; buf2 is free here
                for(int t = tid; t < n_out * JW; t += NT) {
                    const int e = t / JW, q = t % JW, rr = q % D, j = q / D;
                    const double *J = Jn + (size_t) e * JW;
                    const double *S2 = Sg + (size_t) e * SW;
                    double acc = 0;
#pragma unroll
                    for(int i = 0; i < 2 * D; i++) acc += J[rr + i * D] * S2[i + j * 2 * D];
                    Tm[t] = acc;
                }
                gsync<NT>();
                for(int t = tid; t < n_out * D * D; t += NT) {
                    const int e = t / (D * D), q = t % (D * D), rr = q % D, cc = q / D;
                    if(rr >= cc) {
                        const double *J = Jn + (size_t) e * JW;
                        const double *Te = Tm + (size_t) e * JW;
                        double s1 = 0, s2 = 0;
#pragma unroll
                        for(int j = 0; j < 2 * D; j++) {
                            s1 += Te[rr + j * D] * J[cc + j * D];
                            s2 += Te[cc + j * D] * J[rr + j * D];
                        }
                        const double v = 0.5 * (s1 + s2);
                        Bk[(size_t) e * D * D + rr + cc * D] = v;
                        Bk[(size_t) e * D * D + cc + rr * D] = v;
                    }
                }
                gsync<NT>();
                SPG_T(11);
                // X_e = block_e^-1 (:273-274; the reference goes through LLT, here all D x D blocks are inverted
                // together by D symmetric Gauss-Jordan sweeps, one thread per entry; a pivot <= 0 is the same
                // "not positive definite" condition), straight to the output record
                double *src = Bk, *dst = Sg; // Sg is free again: ping-pong between the two
                for(int s0 = 0; s0 < D; s0++) {
                    for(int t = tid; t < n_out * D * D; t += NT) {
                        const int e = t / (D * D), q = t % (D * D), i = q % D, j = q / D;
                        const double *B = src + (size_t) e * D * D;
                        const double d = B[s0 + s0 * D];
                        if(!(d > 0)) s_misc[0] = SPG_BLANKET_NOT_PD_CLOSED;
                        if(fast && i == s0 && j == s0) s_cs[e * D + s0] = d; // pivots: logdet(X_e^-1); s_cs is free on the shortcut path
                        const double inv = 1.0 / d, bis = B[i + s0 * D], bsj = B[s0 + j * D];
                        double v = B[i + j * D] - bis * bsj * inv;
                        if(j == s0) v = bis * inv;
                        if(i == s0) v = bsj * inv;
                        if(i == s0 && j == s0) v = -inv;
                        dst[t] = v;
                    }
                    gsync<NT>();
                    double *tmp = src; src = dst; dst = tmp;
                }
                for(int t = tid; t < n_out * D * D; t += NT) {
                    const int e = t / (D * D), q = t % (D * D);
                    double *sx = reinterpret_cast<double *>(gout + SPG_OUT_HEADER_WORDS + (size_t) e * slot + 1 + PW);
                    sx[q] = (s_misc[0] == SPG_BLANKET_OK) ? -src[t] : 0.0;
                }
                gsync<NT>();
                if(s_misc[0] != SPG_BLANKET_OK) status = s_misc[0];
                // ---- projected KLD at the closed form: LogdetFunction::value (logdet_function.cpp:119-133) as
                // optimizeInformation's diagnostics evaluate it (optimizer.cpp:22-24) ------------------------------
                if(status == SPG_BLANKET_OK) {
                    if(fast) {
                        // X_e = (J Sigma J^T)^-1 makes tr(S U^T J^T X J U) = r exactly, and Lambda_t and J^T X J share
                        // the gauge null space, so the ratio of their pseudo-determinants is the ratio of the
                        // determinants of their anchored blocks. J^T X J is tree structured with |det J_child| = 1:
                        //   KLD = 1/2 [ logdet Lambda_rr - sum_e logdet X_e ]
                        // = half the sum of the logs of the pivots of the anchored sweep and of the D x D inversions.
                        double lp = glog_part;
                        for(int t = tid; t < n_out * D; t += NT) lp += log(s_cs[t]);
                        out_kld = 0.5 * block_sum<NT>(lp, s_red);
                    } else {
                        // literal: A = U^T J^T X J U (r x r), value = 1/2 [tr(A S) - logdet A - logdet S - r]
                        double *JU = buf0 + (size_t) n_out * JW; // R x r column-major, R = n_out * D = r; beyond Jn
                        double *Xs = buf2;                        // n_out * D*D
                        double *A = buf1;                         // V is dead once JU is formed
                        const int R = n_out * D;
                        for(int t = tid; t < n_out * D * D; t += NT) Xs[t] = -src[t];
                        gsync<NT>(); // src (Sg / Bk) lies where JU goes
                        for(int t = tid; t < R * r; t += NT) {
                            const int row = t % R, col = t / R, e = row / D, i = row % D;
                            const int a = pk_a(s_tree[e]), bb = pk_b(s_tree[e]);
                            const double *u = V + (size_t) s_order[ooff + col] * ldk;
                            const double *J = Jn + (size_t) e * JW;
                            double acc = 0;
#pragma unroll
                            for(int j = 0; j < D; j++) {
                                const double ja = J[i + j * D], jb = J[i + (D + j) * D];
                                if(fabs(ja) >= 2.220446049250313e-16) acc += ja * u[a * D + j]; // sparseJacobian(), :335
                                if(fabs(jb) >= 2.220446049250313e-16) acc += jb * u[bb * D + j];
                            }
                            JU[t] = acc;
                        }
                        gsync<NT>();
                        for(int t = tid; t < r * r; t += NT) {
                            const int i = t % r, j = t / r;
                            if(i <= j) { // upper triangle, mirrored (:121-122)
                                double acc = 0;
                                for(int e = 0; e < n_out; e++) {
                                    const double *xe = Xs + (size_t) e * D * D;
                                    const double *ui = JU + (size_t) e * D + (size_t) i * R, *uj = JU + (size_t) e * D + (size_t) j * R;
#pragma unroll
                                    for(int bq = 0; bq < D; bq++) {
                                        double w = 0;
#pragma unroll
                                        for(int aq = 0; aq < D; aq++) w += ui[aq] * xe[aq + bq * D];
                                        acc += w * uj[bq];
                                    }
                                }
                                A[i + (size_t) j * ldk] = acc;
                                A[j + (size_t) i * ldk] = acc;
                            }
                        }
                        gsync<NT>();
                        double tp = 0;
                        for(int i = tid; i < r; i += NT) tp += A[i + (size_t) i * ldk] * s_S[i] - log(s_S[i]);
                        const double tr_minus_logS = block_sum<NT>(tp, s_red);
                        if(!chol_lower<NT>(A, r, ldk)) out_kld = INFINITY;
                        else {
                            double dp = 0;
                            for(int i = tid; i < r; i += NT) dp += log(A[i + (size_t) i * ldk]);
                            out_kld = 0.5 * (tr_minus_logS - 2.0 * block_sum<NT>(dp, s_red) - (double) r);
                        }
                        gsync<NT>();
                    }
                }
            }
        }

        SPG_T(12);
        // ---- S5: GLC substitute factors (TopologyProviderGLC::topology, topology_provider_glc.cpp:100-185) ----
        if(status == SPG_BLANKET_OK && P.algorithm == SPG_ALG_GLC && nk >= 1) {
            gsync<NT>();
            double *gsc = smem + P.off_glc;
            const double *kept_pose = s_pose + PS * nrem;
            const bool dense = (P.topology == SPG_TOPO_DENSE) || nk == 1;
            const int nvcap = dense ? nk : 2, cslot = D * nvcap;
            const int64_t slotw = 1 + spgr_pad2(nvcap) + cslot + (int64_t) cslot * cslot;
            int emitted = 0;
            if(dense) {
                const int rank = glc_get_edge_dense<D, NT>(T, k, ldH, nk, kept_pose, buf1, buf2, ldk, gsc, gout + SPG_OUT_HEADER_WORDS);
                if(rank < 0) status = SPG_BLANKET_EIG_NOCONV;
                else if(rank > 0) emitted = 1;
            } else {
                // Tree: the root's unary factor and one binary factor per tree edge (:134-176). Phase A (whole CTA): the
                // marginal of every factor's vertices, parked in the information area of the factor's own output slot
                // (global memory, L2-resident; the slot is rewritten in phase B). Phase B: the factors are independent, so
                // warp w finishes factors w, w + W, ... on its own scratch with warp-level barriers only (pseudo-inverse,
                // conditional target, J^-1, 12 x 12 eigen-decomposition, W) — up to W factors side by side instead of
                // one factor at a time behind CTA-wide barriers.
                const int nf = 1 + n_out; // factor 0: root; factor f: tree edge f - 1
                const int c2 = 2 * D;
                int *s_rank = s_heapab; // rank of every factor (the Chow-Liu scratch is dead)
                for(int f = 0; f < nf && status == SPG_BLANKET_OK; f++) {
                    double *park = reinterpret_cast<double *>(gout + SPG_OUT_HEADER_WORDS + (int64_t) f * slotw + 1 + spgr_pad2(nvcap) + cslot);
                    const int a = f ? pk_a(s_tree[f - 1]) : pk_a(s_tree[0]), bb = f ? pk_b(s_tree[f - 1]) : 0;
                    if(!glc_marginal<D, NT>(T, k, ldH, a, bb, f ? 2 : 1, park, buf1, buf2)) status = SPG_BLANKET_NOT_PD_JOINT;
                }
                SPG_T(14);
                __threadfence_block();
                gsync<NT>();
                if(status == SPG_BLANKET_OK) {
                    const int W = min(P.glc_warps > 0 ? P.glc_warps : 1, (NT + 31) / 32);
                    const int warp = tid >> 5, lane = tid & 31;
                    if(warp < W) {
                        double *ws = gsc + (size_t) warp * P.glc_warp_doubles;
                        double *Sj = ws, *tgt = Sj + 4 * D * D, *pin = tgt + 4 * D * D, *psc = pin + D * D;
                        double *esc = psc + (2 * D * (D | 1) + 2 * D + 16);
                        int *kv = reinterpret_cast<int *>(esc + (8 * c2 * c2 + 8 * c2 + 64));
                        for(int f = warp; f < nf; f += W) {
                            uint64_t *slotp = gout + SPG_OUT_HEADER_WORDS + (int64_t) f * slotw;
                            const double *park = reinterpret_cast<const double *>(slotp + 1 + spgr_pad2(nvcap) + cslot);
                            const int cf = f ? c2 : D;
                            for(int t = lane; t < cf * cf; t += 32) Sj[t] = park[t];
                            __syncwarp();
                            const double *target = Sj;
                            if(f) {
                                // target = [Jaa Jab; Jba Jba pinv(Jaa) Jab], then selfadjointView<Upper> (:146-176)
                                glc_posdef_pinv<D, 32>(Sj, 2 * D, pin, psc);
                                for(int t = lane; t < 4 * D * D; t += 32) {
                                    const int i = t % (2 * D), j = t / (2 * D);
                                    if(i <= j) {
                                        double v;
                                        if(i >= D) { // lower-right block: Jba * pinv * Jab
                                            v = 0;
                                            for(int p = 0; p < D; p++) {
                                                double u = 0;
                                                for(int q = 0; q < D; q++) u += Sj[i + q * 2 * D] * pin[q + p * D];
                                                v += u * Sj[p + j * 2 * D];
                                            }
                                        } else {
                                            v = Sj[i + j * 2 * D];
                                        }
                                        tgt[i + j * 2 * D] = v;
                                        tgt[j + i * 2 * D] = v;
                                    }
                                }
                                target = tgt;
                            }
                            if(lane == 0) {
                                kv[0] = f ? pk_a(s_tree[f - 1]) : pk_a(s_tree[0]);
                                kv[1] = f ? pk_b(s_tree[f - 1]) : 0;
                            }
                            __syncwarp();
                            // the parked marginal is consumed: clear the slot before the factor is written into it
                            for(int t = lane; t < (int) slotw; t += 32) slotp[t] = 0;
                            __syncwarp();
                            const int rank = glc_get_edge_small<D, 32>(target, f ? 2 : 1, kv, kept_pose, esc, slotp, cslot, nvcap);
                            if(lane == 0) s_rank[f] = rank;
                            __syncwarp();
                        }
                    }
                    __threadfence_block();
                    gsync<NT>();
                    SPG_T(15);
                    // compaction: factors of rank 0 are dropped (getEdge returned NULL, :85-89), the others keep their order
                    for(int f = 0; f < nf; f++) {
                        const int rank = s_rank[f];
                        if(rank < 0) status = SPG_BLANKET_EIG_NOCONV;
                        else if(rank > 0) {
                            if(emitted != f) {
                                uint64_t *dst = gout + SPG_OUT_HEADER_WORDS + (int64_t) emitted * slotw;
                                const uint64_t *src = gout + SPG_OUT_HEADER_WORDS + (int64_t) f * slotw;
                                for(int t = tid; t < (int) slotw; t += NT) dst[t] = src[t];
                                gsync<NT>();
                            }
                            emitted++;
                        }
                    }
                    for(int f = emitted; f < nf; f++) { // slots behind the last emitted factor read as empty
                        uint64_t *dst = gout + SPG_OUT_HEADER_WORDS + (int64_t) f * slotw;
                        for(int t = tid; t < (int) slotw; t += NT) dst[t] = 0;
                    }
                }
            }
            n_out = emitted;
        }

        SPG_T(13);
        gsync<NT>();
        if(tid == 0) {
            int32_t *oh = reinterpret_cast<int32_t *>(gout);
            oh[0] = status;
            oh[1] = (status == SPG_BLANKET_OK) ? n_out : 0;
            oh[2] = newton_iters;
            oh[3] = out_flags;
            reinterpret_cast<double *>(gout)[2] = out_kld;
        }
    }
}

} // namespace spg
