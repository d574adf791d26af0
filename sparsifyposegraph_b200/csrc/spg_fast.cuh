// spg_fast.cuh — the NFR-tree kernel (sm_100a): the path BASELINE.json's metric is quoted on and the one the
// shipped datasets take (Chow-Liu tree, closed-form information fit), for blankets of relative-pose edges with one
// removed vertex. Same stages as blanket_kernel (spg_kernels.cuh) — assembly, Schur complement, Chow-Liu, gauge
// shortcut, closed form (reference src/vertex_remover.cpp:394-450, src/pseudo_chow_liu.cpp, src/logdet_function.cpp:
// 236-279) — but built around ONE data structure: the symmetric k x k matrix lives in REGISTERS, as TR x D tiles of
// its lower triangle, one tile per thread, from the assembly to the last sweep.
//   * H_kk is accumulated straight into the tiles (no N x N matrix in shared memory), the rank-d Schur update is
//     applied in registers, and both SPD inverses (C = (Lambda_t + I)^-1 and the anchored G = Lambda_rr^-1) are
//     symmetric Gauss-Jordan sweeps over the lower-triangular tiles: half the FMAs and half the threads of the
//     full-matrix sweep, 18 FMAs + 6 shared-memory loads per thread and pivot.
//   * Lambda_t, C and G are stored PACKED (lower triangle): 33 KB instead of 65 KB per matrix at k = 90, and G
//     overwrites C. A 16-vertex SE3 blanket needs ~92 KB and <= 128 registers: two CTAs of 256 threads per SM
//     (16 warps, two independent pivot chains) instead of one at 229.7 KB; a 6-vertex blanket is ONE warp and 15 KB:
//     ~14 blankets in flight per SM with __syncwarp() as the only barrier.
//   * The kernel carries nothing but this path (no GLC, no Newton loop, no Jacobi eigen-solver). Whatever it does not
//     take — GLC / MULTI input edges, several removed vertices, a refused gauge guard, any non-positive pivot — is
//     appended to a device-side retry list and re-run by blanket_kernel, which reports the blanket's status.
#pragma once
#include "spg_kernels.cuh"

namespace spg {

template <bool ONEWARP>
__device__ __forceinline__ void fsync() {
    if constexpr(ONEWARP) __syncwarp();
    else __syncthreads();
}
template <bool ONEWARP>
__device__ __forceinline__ int fsync_or(int pred) {
    if constexpr(ONEWARP) return __any_sync(0xffffffffu, pred);
    else return __syncthreads_or(pred);
}
// sum over the CTA in a fixed order; every thread gets the result. red: >= 16 doubles.
template <bool ONEWARP>
__device__ __forceinline__ double fsum(double v, double *red) {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if constexpr(ONEWARP) return v;
    else {
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
        __syncthreads();
        if(l == 0) red[w] = v;
        __syncthreads();
        double s = 0;
        for(int i = 0; i < nw; i++) s += red[i];
        return s;
    }
}

// packed lower triangle, column-major: element (i, l), i >= l, of a k x k symmetric matrix
__device__ __forceinline__ int pidx(int k, int i, int l) { return l * k - ((l * (l - 1)) >> 1) + (i - l); }
__device__ __forceinline__ double psym(const double *A, int k, int i, int l) { return i >= l ? A[pidx(k, i, l)] : A[pidx(k, l, i)]; }

// One pivot step of the symmetric sweep on TR x D register tiles of the lower triangle. Pivot j = D J + C0.
// `cur` holds column j of the current matrix (all rows, by symmetry), the pivot and its reciprocal; the threads
// holding column / row j+1 update those elements first and publish them in `nxt` (look-ahead), then everybody
// updates the rest of the tile: one barrier per pivot.
template <int D, int TR, int C0>
__device__ __forceinline__ void fast_step(double (&a)[TR][D], const double *cur, double *nxt, int NP, int J, int rb, int cb, int nsweep,
                                          bool &bad, double &mypiv) {
    constexpr int RPB = D / TR;
    constexpr int R0 = C0 % TR;
    const int j = D * J + C0;
    const double d = cur[NP], inv = cur[NP + 1];
    bad |= !(d > 0);
    if((int) threadIdx.x == j) mypiv = d;
    double ci[TR], cl[D];
#pragma unroll
    for(int r = 0; r < TR; r++) ci[r] = cur[TR * rb + r];
#pragma unroll
    for(int c = 0; c < D; c++) cl[c] = cur[D * cb + c] * inv;
    const bool colj = (cb == J);
    const bool rowj = (rb == RPB * J + C0 / TR);
    constexpr bool wrap = (C0 + 1 == D);
    constexpr int C1 = wrap ? 0 : C0 + 1;
    constexpr int R1 = C1 % TR;
    const int Jn = wrap ? J + 1 : J;
    const int rbn = RPB * Jn + C1 / TR;
    const bool more = (j + 1 < nsweep);
    const bool pubcol = more && (cb == Jn);              // this tile holds rows of column j+1 (at / below its diagonal block)
    const bool pubrow = more && (rb == rbn) && (cb < Jn); // this tile holds row j+1 left of the diagonal block
    if(pubcol) {
#pragma unroll
        for(int r = 0; r < TR; r++) {
            double v = a[r][C1] - ci[r] * cl[C1];
            if(rowj && r == R0) v = cl[C1];
            a[r][C1] = v;
            nxt[TR * rb + r] = v;
        }
        if(rb == rbn) {
            nxt[NP] = a[R1][C1];
            nxt[NP + 1] = fast_rcp(a[R1][C1]);
        }
    }
    if(pubrow) {
#pragma unroll
        for(int c = 0; c < D; c++) {
            double v = a[R1][c] - ci[R1] * cl[c];
            if(colj && c == C0) v = ci[R1] * inv;
            a[R1][c] = v;
            nxt[D * cb + c] = v;
        }
    }
#pragma unroll
    for(int r = 0; r < TR; r++)
#pragma unroll
        for(int c = 0; c < D; c++) {
            const bool ahead = (pubcol && c == C1) || (pubrow && r == R1);
            if(!ahead) a[r][c] -= ci[r] * cl[c];
        }
    if(rowj) {
#pragma unroll
        for(int c = 0; c < D; c++) a[R0][c] = cl[c];
    }
    if(colj) {
#pragma unroll
        for(int r = 0; r < TR; r++) a[r][C0] = ci[r] * inv;
        if(rowj) a[R0][C0] = -inv;
    }
}

template <int D, int TR, bool ONEWARP, int C0>
__device__ __forceinline__ void fast_steps(double (&a)[TR][D], double *colbuf, int NP, int J, int rb, int cb, int nsweep, bool active,
                                           bool &bad, double &mypiv) {
    if constexpr(C0 < D) {
        const int j = D * J + C0;
        const int cs = NP + 2;
        if(active) fast_step<D, TR, C0>(a, colbuf + (j & 1) * cs, colbuf + ((j + 1) & 1) * cs, NP, J, rb, cb, nsweep, bad, mypiv);
        fsync<ONEWARP>();
        fast_steps<D, TR, ONEWARP, C0 + 1>(a, colbuf, NP, J, rb, cb, nsweep, active, bad, mypiv);
    }
}

// Sweeps the first nsweep (a multiple of D) pivots of the symmetric matrix held in the tiles: a <- -(A^-1) on the
// leading nsweep x nsweep block. Returns false (uniformly) on a non-positive pivot. logpiv: log of pivot
// threadIdx.x (0 beyond nsweep): summed over the CTA it is the log-determinant.
template <int D, int TR, bool ONEWARP>
__device__ __forceinline__ bool fast_sweep(double (&a)[TR][D], double *colbuf, int NP, int rb, int cb, int nsweep, bool active, double *logpiv) {
    if(active && cb == 0) { // publish column 0
#pragma unroll
        for(int r = 0; r < TR; r++) colbuf[TR * rb + r] = a[r][0];
        if(rb == 0) {
            colbuf[NP] = a[0][0];
            colbuf[NP + 1] = fast_rcp(a[0][0]);
        }
    }
    fsync<ONEWARP>();
    bool bad = false;
    double mypiv = 1.0;
    const int nJ = nsweep / D;
#pragma unroll 1
    for(int J = 0; J < nJ; J++) fast_steps<D, TR, ONEWARP, 0>(a, colbuf, NP, J, rb, cb, nsweep, active, bad, mypiv);
    if(fsync_or<ONEWARP>(bad)) return false;
    if(logpiv) *logpiv = ((int) threadIdx.x < nsweep) ? log(mypiv) : 0.0;
    return true;
}

// D x D Cholesky of the diagonal block v of the packed matrix -> Lout (column-major), returns logdet
template <int D>
__device__ __forceinline__ double chol_block_packed(const double *Cp, int k, int v, double *Lout, bool &ok) {
    double A[D * D];
#pragma unroll
    for(int c = 0; c < D; c++)
#pragma unroll
        for(int r = 0; r < D; r++) A[r + c * D] = (r >= c) ? Cp[pidx(k, v * D + r, v * D + c)] : 0.0;
    return chol_small<D>(A, D, Lout, ok);
}

// logdet (C_jj - C_ji C_ii^-1 C_ij), i < j, from the packed C and Li = chol(C_ii): schur_logdet of spg_kernels.cuh
template <int D>
__device__ __forceinline__ double schur_logdet_packed(const double *Cp, int k, int i, int j, const double *Li) {
    double Y[D][D];
#pragma unroll
    for(int c = 0; c < D; c++) {
#pragma unroll
        for(int r = 0; r < D; r++) {
            double s = Cp[pidx(k, j * D + c, i * D + r)];
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
            double s = Cp[pidx(k, j * D + r, j * D + c)];
#pragma unroll
            for(int p = 0; p < D; p++) s -= Y[p][r] * Y[p][c];
            S[r][c] = s;
        }
    double ld_sum = 0;
    double dinv[D];
#pragma unroll
    for(int c = 0; c < D; c++) {
        double d = S[c][c];
#pragma unroll
        for(int p = 0; p < c; p++) d -= S[c][p] * S[c][p] * dinv[p];
        ld_sum += log(d);
        dinv[c] = 1.0 / d;
#pragma unroll
        for(int r = c + 1; r < D; r++) {
            double s = S[r][c];
#pragma unroll
            for(int p = 0; p < c; p++) s -= S[r][p] * S[c][p] * dinv[p];
            S[r][c] = s;
        }
    }
    return ld_sum;
}

// shared-memory plan of fast_kernel (doubles), host + device
struct FastPlan {
    int off_pose, off_T, off_C, off_col, off_h0, off_small, total;
    int tri, NP, scratch; // packed size of one k x k matrix, padded k, doubles available for the per-edge J / M blocks (T + C)
};
template <int D>
__host__ __device__ inline FastPlan fast_plan(int max_nv, int max_rec_words) {
    constexpr int PS = PoseStride<D>::value;
    FastPlan p;
    const int nk = max_nv - 1, kmax = D * (nk > 0 ? nk : 1), pairs = nk * (nk - 1) / 2;
    p.NP = kmax;
    p.tri = (kmax * (kmax + 1) / 2 + 1) & ~1;
    int o = (max_rec_words + 1) & ~1;
    p.off_pose = o; o += max_nv * PS;
    p.off_T = o;    o += p.tri;
    p.off_C = o;    o += p.tri;
    p.scratch = 2 * p.tri;
    p.off_col = o;  o += 2 * (p.NP + 2);
    p.off_h0 = o;   o += 2 * D * D + 2 * D * kmax; // H00 ping-pong, H_k0 and Y, both [D][kmax]
    p.off_small = o;
    // wt[pairs] heapw[pairs] heapab[pairs](int) tree[2 pairs + 2](int) uf[nk](int) Lfac[nk D D] logd[nk] cs[kmax] red[34] ev[max_e -> later] misc[16](int)
    o += pairs + pairs + (pairs + 1) / 2 + (pairs + 1) + (nk + 2) / 2 + nk * D * D + nk + kmax + 34 + 8;
    p.total = o;
    return p;
}

#define SPG_FT(i)                                                                                    \
    do {                                                                                             \
        if(P.prof && tid == 0) {                                                                     \
            const long long t_now = clock64();                                                       \
            atomicAdd(&P.prof[i], (unsigned long long) (t_now - t_last));                            \
            t_last = t_now;                                                                          \
        }                                                                                            \
    } while(0)

// ONEWARP: one warp per blanket (k <= 30 for SE3), __syncwarp() only. Otherwise blockDim.x = 32 * ceil(tiles / 32).
template <int D, bool ONEWARP>
__global__ void __launch_bounds__(ONEWARP ? 32 : 512) fast_kernel(const KernelParams P) {
    extern __shared__ double smem_dyn[];
    double *smem = smem_dyn;
    constexpr int PS = PoseStride<D>::value;
    constexpr int PW = (D == 6) ? 7 : 3;
    constexpr int JW = D * 2 * D;
    constexpr int TR = 3;
    constexpr int RPB = D / TR;
    const int tid = threadIdx.x, NT = blockDim.x;

    const FastPlan pl = fast_plan<D>(P.max_nv, P.max_rec_words);
    uint64_t *s_rec = reinterpret_cast<uint64_t *>(smem);
    double *s_pose = smem + pl.off_pose;
    double *Tp = smem + pl.off_T;
    double *Cp = smem + pl.off_C;
    double *colbuf = smem + pl.off_col;
    double *s_h00 = smem + pl.off_h0;         // 2 x D*D
    double *s_hk0 = s_h00 + 2 * D * D;        // [D][NP]: H_k0, p-major
    double *s_y = s_hk0 + D * pl.NP;          // [D][NP]: H_k0 H_00^-1
    const int nkmax = P.max_nv - 1, kmaxb = pl.NP, pairs_max = nkmax * (nkmax - 1) / 2;
    double *s_wt = smem + pl.off_small;
    double *s_heapw = s_wt + pairs_max;
    int *s_heapab = reinterpret_cast<int *>(s_heapw + pairs_max);
    int *s_tree = reinterpret_cast<int *>(s_heapw + pairs_max + (pairs_max + 1) / 2);
    int *s_uf = reinterpret_cast<int *>(s_heapw + pairs_max + (pairs_max + 1) / 2 + (pairs_max + 1));
    double *s_Lfac = s_heapw + pairs_max + (pairs_max + 1) / 2 + (pairs_max + 1) + (nkmax + 2) / 2;
    double *s_logd = s_Lfac + nkmax * D * D;
    double *s_cs = s_logd + nkmax;
    double *s_red = s_cs + kmaxb;
    int *s_misc = reinterpret_cast<int *>(s_red + 34);
    const int NP = pl.NP;

    for(int li = blockIdx.x; li < P.n_list; li += gridDim.x) {
        const int b = P.list ? P.list[li] : li;
        const uint64_t *grec = P.records + P.rec_off[b];
        uint64_t *gout = P.out + P.out_off[b];
        const int out_words = (int) (P.out_off[b + 1] - P.out_off[b]);
        const int32_t *gh = reinterpret_cast<const int32_t *>(grec);
        const int nv = gh[0], nrem = gh[1], ne = gh[2], rdim = gh[3], rec_words = gh[4];
        const int nk = nv - nrem, k = D * nk;

        fsync<ONEWARP>(); // the shared buffers of the previous blanket are dead
        bool refuse = (rdim != D || nv > P.max_nv || ne > P.max_e || rec_words > P.max_rec_words || nrem != 1 || nk < 0 || ne < 1);
        for(int t = tid; t < out_words; t += NT) gout[t] = 0;
        long long t_last = clock64();
        if(!refuse) {
            for(int t = tid; t < rec_words; t += NT) s_rec[t] = grec[t];
            if(tid == 0) { s_misc[0] = 0; s_misc[3] = 0; }
        }
        fsync<ONEWARP>();
        const double *r_pose = reinterpret_cast<const double *>(s_rec + spgr_poses_off(nv));
        const int32_t *r_etab = reinterpret_cast<const int32_t *>(s_rec + spgr_edgetab_off(D, nv));
        if(!refuse) {
            // every edge a relative-pose edge between two different vertices? (anything else: blanket_kernel)
            int badedge = 0;
            for(int e = tid; e < ne; e += NT) {
                const int32_t *eh = reinterpret_cast<const int32_t *>(s_rec + r_etab[e]);
                const int32_t *vi = reinterpret_cast<const int32_t *>(s_rec + r_etab[e] + 2);
                badedge |= (eh[0] != SPG_EDGE_POSE) || (eh[1] != 2) || (vi[0] == vi[1]);
            }
            refuse = fsync_or<ONEWARP>(badedge) != 0;
        }
        int n_out = 0;
        double out_kld = 0;
        int out_flags = 0;
        if(!refuse && nk >= 2) {
            // ---- tile of this thread: rows [TR rb, TR rb + TR), columns [D cb, D cb + D) of the kept block ----
            const int nbr = RPB * nk;
            const int ntiles = nbr * nk - RPB * (nk * (nk - 1) / 2);
            const bool has_tile = tid < ntiles;
            int cb = 0, rb = 0;
            if(has_tile) {
                int rem = tid;
                while(rem >= nbr - RPB * cb) { rem -= nbr - RPB * cb; cb++; }
                rb = RPB * cb + rem;
            }
            const int vr = rb / RPB, roff = (rb % RPB) * TR; // kept vertex of the rows, first row inside its block
            double a[TR][D];
#pragma unroll
            for(int r = 0; r < TR; r++)
#pragma unroll
                for(int c = 0; c < D; c++) a[r][c] = 0.0;

            // ---- S0: poses ---------------------------------------------------------------------------------
            for(int v = tid; v < nv; v += NT) {
                if constexpr(D == 6) se3_from_flat(r_pose + PW * v, s_pose + PS * v);
                else se2_from_flat(r_pose + PW * v, s_pose + PS * v);
            }
            for(int t = tid; t < (D + 2 * NP) * D; t += NT) s_h00[D * D + t] = 0.0; // second H00 buffer, H_k0, Y
            fsync<ONEWARP>();
            SPG_FT(0);
            // ---- S1: assembly. Per chunk of edges: J = [Ji Jj] and M = Omega J in shared memory (over the T / C
            // area, unused until the Schur step), then every tile gathers the edges of its vertex pair in edge
            // order (fixed summation order), and D (k + D) threads gather H_00 and H_k0. ---------------------------
            double *JM = Tp;
            const int chunk = max(1, pl.scratch / (2 * JW));
            double h0acc = 0; // entry (i, p) of [H_k0; H_00] gathered by this thread: t = i + (k + D) * p ... first entry only
            for(int e0 = 0; e0 < ne; e0 += chunk) {
                const int ce = min(chunk, ne - e0);
                for(int e = tid; e < ce; e += NT) {
                    const uint64_t *ew = s_rec + r_etab[e0 + e];
                    const int32_t *vi = reinterpret_cast<const int32_t *>(ew + 2);
                    const double *pm = reinterpret_cast<const double *>(ew + 3);
                    double Z[PS];
                    if constexpr(D == 6) se3_from_flat(pm, Z);
                    else se2_from_flat(pm, Z);
                    edge_jacobians<D>(Z, s_pose + PS * vi[0], s_pose + PS * vi[1], JM + (size_t) e * 2 * JW);
                }
                fsync<ONEWARP>();
                for(int t = tid; t < ce * JW; t += NT) { // M = Omega J
                    const int e = t / JW, q = t % JW, r = q % D, c = q / D;
                    const double *Om = reinterpret_cast<const double *>(s_rec + r_etab[e0 + e] + 3) + PW;
                    const double *J = JM + (size_t) e * 2 * JW;
                    double s = 0;
#pragma unroll
                    for(int p = 0; p < D; p++) s += Om[r + p * D] * J[p + c * D];
                    JM[(size_t) e * 2 * JW + JW + q] = s;
                }
                fsync<ONEWARP>();
                if(has_tile) {
                    const int lr = vr + 1, lc = cb + 1; // local vertex indices (the removed vertex is 0)
                    for(int e = 0; e < ce; e++) {
                        const int32_t *vi = reinterpret_cast<const int32_t *>(s_rec + r_etab[e0 + e] + 2);
                        const int va = vi[0], vb = vi[1];
                        int sr, sc; // side (0: Ji / first vertex, 1: Jj) of the row vertex and of the column vertex
                        if(va == lr) sr = 0; else if(vb == lr) sr = 1; else continue;
                        if(va == lc) sc = 0; else if(vb == lc) sc = 1; else continue;
                        const double *Jr = JM + (size_t) e * 2 * JW + (sr * D + roff) * D;
                        const double *Mc = JM + (size_t) e * 2 * JW + JW + sc * D * D;
#pragma unroll
                        for(int p = 0; p < D; p++) {
                            double jr[TR], mc[D];
#pragma unroll
                            for(int r = 0; r < TR; r++) jr[r] = Jr[p + r * D];
#pragma unroll
                            for(int c = 0; c < D; c++) mc[c] = Mc[p + c * D];
#pragma unroll
                            for(int r = 0; r < TR; r++)
#pragma unroll
                                for(int c = 0; c < D; c++) a[r][c] += jr[r] * mc[c];
                        }
                    }
                }
                // [H_k0; H_00]: entry (i, p), i in [0, k + D) (kept dims, then the removed vertex's), p in [0, D)
                for(int t = tid; t < (k + D) * D; t += NT) {
                    const int i = t % (k + D), p = t / (k + D);
                    const int lv = (i < k) ? i / D + 1 : 0, di = (i < k) ? i % D : i - k;
                    double s = 0;
                    for(int e = 0; e < ce; e++) {
                        const int32_t *vi = reinterpret_cast<const int32_t *>(s_rec + r_etab[e0 + e] + 2);
                        const int va = vi[0], vb = vi[1];
                        int sv, s0;
                        if(va == lv) sv = 0; else if(vb == lv) sv = 1; else continue;
                        if(va == 0) s0 = 0; else if(vb == 0) s0 = 1; else continue;
                        const double *Jv = JM + (size_t) e * 2 * JW + (sv * D + di) * D;
                        const double *M0 = JM + (size_t) e * 2 * JW + JW + (s0 * D + p) * D;
#pragma unroll
                        for(int q = 0; q < D; q++) s += Jv[q] * M0[q];
                    }
                    if(i < k) s_hk0[p * NP + i] += s;
                    else s_h00[D * D + di + p * D] += s;
                }
                fsync<ONEWARP>();
            }
            (void) h0acc;
            SPG_FT(1);
            // ---- S2: Schur complement Lambda_t = H_kk - H_k0 H_00^-1 H_0k (vertex_remover.cpp:443-449) ---------
            // H_00^-1 by D symmetric Gauss-Jordan steps on D*D threads (a pivot <= 0 is LLT's failure)
            {
                double *src = s_h00 + D * D, *dst = s_h00;
                for(int s0 = 0; s0 < D; s0++) {
                    if(tid < D * D) {
                        const int i = tid % D, j = tid / D;
                        const double d = src[s0 + s0 * D];
                        if(!(d > 0)) s_misc[0] = 1;
                        const double inv = 1.0 / d, bis = src[i + s0 * D], bsj = src[s0 + j * D];
                        double v = src[i + j * D] - bis * bsj * inv;
                        if(j == s0) v = bis * inv;
                        if(i == s0) v = bsj * inv;
                        if(i == s0 && j == s0) v = -inv;
                        dst[tid] = v;
                    }
                    fsync<ONEWARP>();
                    double *tmp = src; src = dst; dst = tmp;
                }
                // D even: the result (-H_00^-1) is back in s_h00 + D*D; D odd: in s_h00
                const double *Hinv = src;
                for(int t = tid; t < k * D; t += NT) { // Y[p][i] = sum_q H_k0[q][i] * H_00^-1[q][p]
                    const int i = t % k, p = t / k;
                    double s = 0;
#pragma unroll
                    for(int q = 0; q < D; q++) s -= s_hk0[q * NP + i] * Hinv[q + p * D];
                    s_y[p * NP + i] = s;
                }
                fsync<ONEWARP>();
            }
            if(s_misc[0]) refuse = true; // uniform (read after the barrier)
            if(!refuse) {
                if(has_tile) {
#pragma unroll
                    for(int p = 0; p < D; p++) {
                        double yi[TR], hl[D];
#pragma unroll
                        for(int r = 0; r < TR; r++) yi[r] = s_y[p * NP + TR * rb + r];
#pragma unroll
                        for(int c = 0; c < D; c++) hl[c] = s_hk0[p * NP + D * cb + c];
#pragma unroll
                        for(int r = 0; r < TR; r++)
#pragma unroll
                            for(int c = 0; c < D; c++) a[r][c] -= yi[r] * hl[c];
                    }
                    // Lambda_t, lower triangle, packed (the J / M blocks in this area are dead)
#pragma unroll
                    for(int r = 0; r < TR; r++)
#pragma unroll
                        for(int c = 0; c < D; c++) {
                            const int i = TR * rb + r, l = D * cb + c;
                            if(i >= l) Tp[pidx(k, i, l)] = a[r][c];
                        }
                }
                fsync<ONEWARP>();
                if(P.dbg_target) {
                    double *g = P.dbg_target + P.dbg_target_off[b];
                    if(P.dbg_target_off[b + 1] - P.dbg_target_off[b] >= (int64_t) k * k)
                        for(int t = tid; t < k * k; t += NT) g[t] = psym(Tp, k, t % k, t / k);
                }
            }
            SPG_FT(2);
            // ---- S3: Chow-Liu tree (pseudo_chow_liu.cpp:33-87) --------------------------------------------------
            if(!refuse) {
                if(nk == 2) {
                    n_out = 1;
                    if(tid == 0) s_tree[0] = pk(0, 1);
                } else {
                    n_out = nk - 1;
                    const int all = nk * (nk - 1) / 2;
                    // C = (Lambda_t + 1 I)^-1 (fillEdges, :185-190): symmetric sweep of the tiles
                    if(has_tile) {
#pragma unroll
                        for(int r = 0; r < TR; r++)
#pragma unroll
                            for(int c = 0; c < D; c++)
                                if(TR * rb + r == D * cb + c) a[r][c] += 1.0;
                    }
                    if(!fast_sweep<D, TR, ONEWARP>(a, colbuf, NP, rb, cb, k, has_tile, nullptr)) refuse = true;
                    if(!refuse) {
                        if(has_tile) {
#pragma unroll
                            for(int r = 0; r < TR; r++)
#pragma unroll
                                for(int c = 0; c < D; c++) {
                                    const int i = TR * rb + r, l = D * cb + c;
                                    if(i >= l) Cp[pidx(k, i, l)] = -a[r][c];
                                }
                        }
                        fsync<ONEWARP>();
                        SPG_FT(3);
                        for(int v = tid; v < nk; v += NT) {
                            bool ok;
                            s_logd[v] = chol_block_packed<D>(Cp, k, v, s_Lfac + v * D * D, ok);
                            if(!ok) s_misc[0] = 1;
                        }
                        fsync<ONEWARP>();
                        // weight(i,j) = logdet C_jj - logdet (C_jj - C_ji C_ii^-1 C_ij)   (:169-183)
                        for(int t = tid; t < all; t += NT) {
                            int i = 0, rem = t;
                            while(rem >= nk - 1 - i) { rem -= nk - 1 - i; i++; }
                            const int j = i + 1 + rem;
                            s_wt[t] = s_logd[j] - schur_logdet_packed<D>(Cp, k, i, j, s_Lfac + i * D * D);
                        }
                        fsync<ONEWARP>();
                        if(s_misc[0]) refuse = true;
                    }
                    if(!refuse) {
                        if(P.dbg_weights) {
                            double *g = P.dbg_weights + P.dbg_weights_off[b];
                            const int cap = (int) (P.dbg_weights_off[b + 1] - P.dbg_weights_off[b]);
                            if(P.flags & SPG_OPT_DBG_WEIGHTS_IN) {
                                for(int t = tid; t < all && t < cap; t += NT) s_wt[t] = g[t];
                                fsync<ONEWARP>();
                            } else {
                                for(int t = tid; t < cap; t += NT) g[t] = t < all ? s_wt[t] : 0.0;
                            }
                        }
                        SPG_FT(4);
                        // doKruskal (:253-289): same scheme as blanket_kernel — parallel ranking, or the replay of
                        // libstdc++'s heap when two weights are exactly equal
                        int *s_sorted = s_tree + all;
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
                        fsync<ONEWARP>();
                        if(s_misc[3]) out_flags |= 128;
                        if(tid == 0) {
                            for(int v = 0; v < nk; v++) s_uf[v] = v;
                            int nacc = 0;
                            if(!s_misc[3]) {
                                for(int q = 0; q < all && nacc < nk - 1; q++) {
                                    const int ab = s_sorted[q];
                                    int ra = pk_a(ab), rbb = pk_b(ab);
                                    while(s_uf[ra] != ra) ra = s_uf[ra];
                                    while(s_uf[rbb] != rbb) rbb = s_uf[rbb];
                                    if(ra != rbb) {
                                        s_uf[rbb] = ra;
                                        s_tree[nacc++] = ab;
                                    }
                                }
                            } else {
                                HeapView hp{s_heapw, s_heapab, 0};
                                int t = 0;
                                for(int i = 0; i < nk - 1; i++)
                                    for(int j = i + 1; j < nk; j++, t++) hp.push(s_wt[t], pk(i, j));
                                while(hp.len > 0 && nacc < nk - 1) {
                                    double w; int ab;
                                    hp.pop(w, ab);
                                    int ra = pk_a(ab), rbb = pk_b(ab);
                                    while(s_uf[ra] != ra) ra = s_uf[ra];
                                    while(s_uf[rbb] != rbb) rbb = s_uf[rbb];
                                    if(ra != rbb) {
                                        s_uf[rbb] = ra;
                                        s_tree[nacc++] = ab;
                                    }
                                }
                            }
                        }
                        fsync<ONEWARP>();
                    }
                }
            }
            SPG_FT(5);
            // ---- S4: gauge shortcut + closed form (see blanket_kernel for the argument) --------------------------
            const int kk = k - D;
            double glog = 0;
            if(!refuse) {
                // guard (ii): every diagonal entry of Lambda_t below 1e8
                int bigdiag = 0;
                for(int i = tid; i < k; i += NT) bigdiag |= !(fabs(Tp[pidx(k, i, i)]) < 1e8);
                if(fsync_or<ONEWARP>(bigdiag)) refuse = true;
            }
            if(!refuse) {
                // G = Lambda_rr^-1 (last kept vertex anchored): reload the tiles of the leading block and sweep
                const bool act = has_tile && vr < nk - 1;
                if(act) {
#pragma unroll
                    for(int r = 0; r < TR; r++)
#pragma unroll
                        for(int c = 0; c < D; c++) a[r][c] = psym(Tp, k, TR * rb + r, D * cb + c);
                }
                if(!fast_sweep<D, TR, ONEWARP>(a, colbuf, NP, rb, cb, kk, act, &glog)) refuse = true;
                if(!refuse) {
                    // guard (i): ||G||_F <= 1e5
                    double fp = 0;
                    if(act) {
#pragma unroll
                        for(int r = 0; r < TR; r++)
#pragma unroll
                            for(int c = 0; c < D; c++) {
                                const int i = TR * rb + r, l = D * cb + c;
                                const double w = (i > l) ? 2.0 : (i == l ? 1.0 : 0.0);
                                fp += w * a[r][c] * a[r][c];
                                if(i >= l) Cp[pidx(kk, i, l)] = -a[r][c]; // G over C (dead since the weights are out)
                            }
                    }
                    const double frob2 = fsum<ONEWARP>(fp, s_red);
                    if(!(frob2 <= 1e10)) refuse = true;
                    fsync<ONEWARP>();
                }
            }
            SPG_FT(6);
            if(!refuse) {
                // new-edge Jacobians at zero error (vertex_remover.cpp:466-498), Sigma blocks from G, X_e = (J Sigma J^T)^-1
                const double *Gp = Cp;
                constexpr int SW = 4 * D * D;
                double *Jn = Tp;                          // n_out * JW    (Lambda_t is dead)
                double *Sg = Jn + (size_t) n_out * JW;    // n_out * SW
                double *Tm = Sg + (size_t) n_out * SW;    // n_out * JW
                double *Bk = Tm + (size_t) n_out * JW;    // n_out * D*D
                double *Bk2 = Bk + (size_t) n_out * D * D;
                const int slot = 1 + PW + D * D;
                for(int e = tid; e < n_out; e += NT) {
                    const int ea = pk_a(s_tree[e]), eb = pk_b(s_tree[e]);
                    const double *Xa = s_pose + PS * (1 + ea), *Xb = s_pose + PS * (1 + eb);
                    double Z[PS], Ti[PS];
                    if constexpr(D == 6) { se3_inverse(Xa, Ti); se3_compose(Ti, Xb, Z); }
                    else { se2_inverse(Xa, Ti); se2_compose(Ti, Xb, Z); }
                    edge_jacobians_zero_error<D>(Z, Xa, Xb, Jn + (size_t) e * JW);
                    uint64_t *sl = gout + SPG_OUT_HEADER_WORDS + (size_t) e * slot;
                    int32_t *si = reinterpret_cast<int32_t *>(sl);
                    si[0] = ea; si[1] = eb;
                    double *sm = reinterpret_cast<double *>(sl + 1);
                    if constexpr(D == 6) se3_to_flat(Z, sm);
                    else { sm[0] = Z[0]; sm[1] = Z[1]; sm[2] = Z[2]; }
                }
                for(int t = tid; t < n_out * SW; t += NT) {
                    const int e = t / SW, q = t % SW, i = q % (2 * D), j = q / (2 * D);
                    if(i >= j) {
                        const int ea = pk_a(s_tree[e]), eb = pk_b(s_tree[e]);
                        const int ri = (i < D ? ea * D + i : eb * D + i - D);
                        const int rj = (j < D ? ea * D + j : eb * D + j - D);
                        double s = 0;
                        if(ri < kk && rj < kk) s = psym(Gp, kk, ri, rj);
                        Sg[(size_t) e * SW + i + j * 2 * D] = s;
                        Sg[(size_t) e * SW + j + i * 2 * D] = s;
                    }
                }
                fsync<ONEWARP>();
                for(int t = tid; t < n_out * JW; t += NT) { // Tm = J Sigma_e
                    const int e = t / JW, q = t % JW, rr = q % D, j = q / D;
                    const double *J = Jn + (size_t) e * JW;
                    const double *S2 = Sg + (size_t) e * SW;
                    double acc = 0;
#pragma unroll
                    for(int i = 0; i < 2 * D; i++) acc += J[rr + i * D] * S2[i + j * 2 * D];
                    Tm[t] = acc;
                }
                fsync<ONEWARP>();
                for(int t = tid; t < n_out * D * D; t += NT) { // block = Tm J^T, symmetrised (logdet_function.cpp:249-270)
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
                fsync<ONEWARP>();
                double *src = Bk, *dst = Bk2;
                for(int s0 = 0; s0 < D; s0++) { // X_e = block^-1 (:273-274), all blocks together, D Gauss-Jordan steps
                    for(int t = tid; t < n_out * D * D; t += NT) {
                        const int e = t / (D * D), q = t % (D * D), i = q % D, j = q / D;
                        const double *B = src + (size_t) e * D * D;
                        const double d = B[s0 + s0 * D];
                        if(!(d > 0)) s_misc[0] = 1;
                        if(i == s0 && j == s0) s_cs[e * D + s0] = d;
                        const double inv = 1.0 / d, bis = B[i + s0 * D], bsj = B[s0 + j * D];
                        double v = B[i + j * D] - bis * bsj * inv;
                        if(j == s0) v = bis * inv;
                        if(i == s0) v = bsj * inv;
                        if(i == s0 && j == s0) v = -inv;
                        dst[t] = v;
                    }
                    fsync<ONEWARP>();
                    double *tmp = src; src = dst; dst = tmp;
                }
                if(s_misc[0]) refuse = true;
                if(!refuse) {
                    for(int t = tid; t < n_out * D * D; t += NT) {
                        const int e = t / (D * D), q = t % (D * D);
                        double *sx = reinterpret_cast<double *>(gout + SPG_OUT_HEADER_WORDS + (size_t) e * slot + 1 + PW);
                        sx[q] = -src[t];
                    }
                    // projected KLD at the closed form = 1/2 [logdet Lambda_rr - sum_e logdet X_e] (see blanket_kernel)
                    double lp = glog;
                    for(int t = tid; t < n_out * D; t += NT) lp += log(s_cs[t]);
                    out_kld = 0.5 * fsum<ONEWARP>(lp, s_red);
                }
            }
            SPG_FT(7);
        }
        fsync<ONEWARP>();
        if(tid == 0) {
            int32_t *oh = reinterpret_cast<int32_t *>(gout);
            if(refuse) {
                oh[0] = SPG_BLANKET_UNSUPPORTED; // overwritten by blanket_kernel, which re-runs the blanket
                const int pos = atomicAdd(P.retry_count, 1);
                P.retry_list[pos] = b;
            } else {
                oh[0] = SPG_BLANKET_OK;
                oh[1] = n_out;
                oh[2] = 0;
                oh[3] = out_flags;
                reinterpret_cast<double *>(gout)[2] = out_kld;
            }
        }
    }
}

} // namespace spg
