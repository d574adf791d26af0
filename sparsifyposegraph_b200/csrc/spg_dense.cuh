// spg_dense.cuh — dense fp64 building blocks on whole-graph matrices in HBM (column-major), used by the full-graph
// evaluator and the pose-graph optimiser (spg_eval.cu): SURVEY.md §8(f) rows 1-2, reference
// src/graph_wrapper_g2o.cpp:250-269 (optimize), :472-575 (chi2 / kullbackLeibler), src/utils.cpp:70-97.
//
// The matrices here are thousands of dimensions wide (sphere.g2o: 15 000), GEMM-shaped and FP64-pipe bound — unlike
// the per-blanket kernels, which are latency bound. On B200 the FP64 vector pipe and the FP64 tensor tiles have the
// same peak, so the contraction is a register-tiled DFMA kernel (64 x 64 tile, 4 x 4 per thread) rather than DMMA.
//
//   dgemm_sub<TB>      C -= A * op(B)            (trailing updates of the blocked Cholesky, forward substitution)
//   potrf_block        Cholesky of one nb x nb diagonal block in shared memory, log-determinant accumulated
//   trsm_right_block   X * L11^T = A21           (panel below a factored diagonal block)
//   trsm_left_block    L11 * Z = B1              (one block row of a forward substitution with many right-hand sides)
// Blocked right-looking Cholesky = potrf_block, trsm_right_block, dgemm_sub<true>(lower tiles only) per panel; stopping
// after the first m columns leaves the Schur complement A22 - A21 A11^-1 A12 in the trailing block — the marginal
// information kullbackLeibler needs (graph_wrapper_g2o.cpp:539-542) without a separate solve.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace spg {

constexpr int DNB = 64; // panel width of the blocked factorisations
constexpr int DENSE_TRSM_SMEM = 2 * DNB * (DNB + 1) * (int) sizeof(double); // dynamic shared memory of the trsm kernels

// C (m x n) -= A (m x k) * B^T   with B (n x k)   [TB = true]
// C (m x n) -= A (m x k) * B     with B (k x n)   [TB = false]
// lower_only: C is square and only tiles that touch the lower triangle are computed (symmetric rank-k update).
template <bool TB>
__global__ void __launch_bounds__(256) dgemm_sub_kernel(int m, int n, int k, const double *__restrict__ A, int64_t lda,
                                                        const double *__restrict__ B, int64_t ldb, double *__restrict__ C, int64_t ldc,
                                                        int lower_only) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ double As[BK][BM + 1];
    __shared__ double Bs[BK][BN + 1];
    const int bm = blockIdx.x * BM, bn = blockIdx.y * BN;
    if(lower_only && bn > bm + BM - 1) return; // tile strictly above the diagonal
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    double acc[4][4];
#pragma unroll
    for(int i = 0; i < 4; i++)
#pragma unroll
        for(int j = 0; j < 4; j++) acc[i][j] = 0.0;
    for(int k0 = 0; k0 < k; k0 += BK) {
        // A tile: rows bm .. bm+63, columns k0 .. k0+15 (coalesced along the rows)
#pragma unroll
        for(int t = 0; t < 4; t++) {
            const int r = tid % 64, kk = tid / 64 + 4 * t;
            const int gr = bm + r, gk = k0 + kk;
            As[kk][r] = (gr < m && gk < k) ? A[gr + (int64_t) gk * lda] : 0.0;
        }
        if(TB) {
#pragma unroll
            for(int t = 0; t < 4; t++) {
                const int c = tid % 64, kk = tid / 64 + 4 * t;
                const int gc = bn + c, gk = k0 + kk;
                Bs[kk][c] = (gc < n && gk < k) ? B[gc + (int64_t) gk * ldb] : 0.0;
            }
        } else {
#pragma unroll
            for(int t = 0; t < 4; t++) {
                const int kk = tid % 16, c = tid / 16 + 16 * t;
                const int gc = bn + c, gk = k0 + kk;
                Bs[kk][c] = (gc < n && gk < k) ? B[gk + (int64_t) gc * ldb] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for(int kk = 0; kk < BK; kk++) {
            double a[4], b[4];
#pragma unroll
            for(int i = 0; i < 4; i++) a[i] = As[kk][tx + 16 * i];
#pragma unroll
            for(int j = 0; j < 4; j++) b[j] = Bs[kk][ty + 16 * j];
#pragma unroll
            for(int i = 0; i < 4; i++)
#pragma unroll
                for(int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for(int i = 0; i < 4; i++)
#pragma unroll
        for(int j = 0; j < 4; j++) {
            const int gr = bm + tx + 16 * i, gc = bn + ty + 16 * j;
            if(gr < m && gc < n) C[gr + (int64_t) gc * ldc] -= acc[i][j];
        }
}

// Cholesky of the nb x nb (nb <= 64) block at A (lower triangle read and written). *logdet += 2 sum log L_ii.
// *flag is set to 1 on a non-positive pivot (LLT's positive-definiteness test).
__global__ void __launch_bounds__(256) potrf_block_kernel(double *A, int64_t lda, int nb, double *logdet, int *flag) {
    __shared__ double S[DNB][DNB + 1];
    __shared__ int bad;
    const int tid = threadIdx.x;
    if(tid == 0) bad = 0;
    for(int t = tid; t < nb * nb; t += 256) {
        const int i = t % nb, j = t / nb;
        S[i][j] = (i >= j) ? A[i + (int64_t) j * lda] : 0.0;
    }
    __syncthreads();
    double ld = 0.0;
    for(int j = 0; j < nb; j++) {
        if(tid == 0) {
            const double d = S[j][j];
            if(!(d > 0.0)) bad = 1;
            const double r = sqrt(d > 0.0 ? d : 1.0);
            S[j][j] = r;
            ld += 2.0 * log(r);
        }
        __syncthreads();
        const double r = S[j][j];
        for(int i = j + 1 + tid; i < nb; i += 256) S[i][j] /= r;
        __syncthreads();
        const int rem = nb - j - 1; // trailing lower triangle: rows i > j, columns j < c <= i
        for(int t = tid; t < rem * rem; t += 256) {
            const int i = j + 1 + t % rem, c = j + 1 + t / rem;
            if(c <= i) S[i][c] -= S[i][j] * S[c][j];
        }
        __syncthreads();
    }
    for(int t = tid; t < nb * nb; t += 256) {
        const int i = t % nb, j = t / nb;
        if(i >= j) A[i + (int64_t) j * lda] = S[i][j];
    }
    if(tid == 0) {
        if(logdet) *logdet += ld;
        if(bad) *flag = 1;
    }
}

// X * L^T = P for the rows of a panel P (rows x nb) below a factored nb x nb diagonal block L; in place.
__global__ void __launch_bounds__(64) trsm_right_kernel(const double *__restrict__ L, int64_t ldl, int nb, double *P, int64_t ldp, int rows) {
    extern __shared__ double dense_smem[]; // DENSE_TRSM_SMEM bytes (above the 48 KB static limit: opt-in by the host)
    double (*Ls)[DNB + 1] = reinterpret_cast<double (*)[DNB + 1]>(dense_smem);
    double (*Xs)[DNB + 1] = reinterpret_cast<double (*)[DNB + 1]>(dense_smem + DNB * (DNB + 1));
    const int tid = threadIdx.x, r0 = blockIdx.x * 64;
    for(int t = tid; t < nb * nb; t += 64) {
        const int i = t % nb, j = t / nb;
        Ls[i][j] = (i >= j) ? L[i + (int64_t) j * ldl] : 0.0;
    }
    const int nr = min(64, rows - r0);
    for(int j = 0; j < nb; j++)
        if(tid < nr) Xs[tid][j] = P[(r0 + tid) + (int64_t) j * ldp]; // coalesced along the rows
    __syncthreads();
    if(tid < nr) {
        for(int j = 0; j < nb; j++) {
            double s = Xs[tid][j];
            for(int p = 0; p < j; p++) s = fma(-Xs[tid][p], Ls[j][p], s);
            Xs[tid][j] = s / Ls[j][j];
        }
    }
    __syncthreads();
    for(int j = 0; j < nb; j++)
        if(tid < nr) P[(r0 + tid) + (int64_t) j * ldp] = Xs[tid][j];
}

// L * Z = B for a block row B (nb x cols) and a factored nb x nb diagonal block L; in place, 64 columns per CTA.
__global__ void __launch_bounds__(64) trsm_left_kernel(const double *__restrict__ L, int64_t ldl, int nb, double *B, int64_t ldb, int cols) {
    extern __shared__ double dense_smem[];
    double (*Ls)[DNB + 1] = reinterpret_cast<double (*)[DNB + 1]>(dense_smem);
    double (*Bs)[64 + 1] = reinterpret_cast<double (*)[64 + 1]>(dense_smem + DNB * (DNB + 1));
    const int tid = threadIdx.x, c0 = blockIdx.x * 64;
    for(int t = tid; t < nb * nb; t += 64) {
        const int i = t % nb, j = t / nb;
        Ls[i][j] = (i >= j) ? L[i + (int64_t) j * ldl] : 0.0;
    }
    const int nc = min(64, cols - c0);
    for(int c = 0; c < nc; c++)
        if(tid < nb) Bs[tid][c] = B[tid + (int64_t) (c0 + c) * ldb]; // coalesced along the rows
    __syncthreads();
    if(tid < nc) {
        for(int i = 0; i < nb; i++) {
            double s = Bs[i][tid];
            for(int p = 0; p < i; p++) s = fma(-Ls[i][p], Bs[p][tid], s);
            Bs[i][tid] = s / Ls[i][i];
        }
    }
    __syncthreads();
    for(int c = 0; c < nc; c++)
        if(tid < nb) B[tid + (int64_t) (c0 + c) * ldb] = Bs[tid][c];
}

// strict upper triangle <- 0 (a Cholesky factor used as a matrix operand)
__global__ void zero_upper_kernel(double *A, int64_t lda, int n) {
    const int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if(t >= (int64_t) n * n) return;
    const int i = (int) (t % n), j = (int) (t / n);
    if(i < j) A[i + (int64_t) j * lda] = 0.0;
}

// upper triangle <- lower triangle (full symmetric operand from a lower-only update)
__global__ void mirror_lower_kernel(double *A, int64_t lda, int n) {
    const int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x;
    if(t >= (int64_t) n * n) return;
    const int i = (int) (t % n), j = (int) (t / n);
    if(i < j) A[i + (int64_t) j * lda] = A[j + (int64_t) i * lda];
}

// *out += sum of squares of the lower triangle (incl. diagonal) of A  — || L ||_F^2 of a triangular matrix
__global__ void __launch_bounds__(256) frob2_lower_kernel(const double *__restrict__ A, int64_t lda, int n, double *out) {
    double s = 0.0;
    for(int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; t < (int64_t) n * n; t += (int64_t) gridDim.x * blockDim.x) {
        const int i = (int) (t % n), j = (int) (t / n);
        if(i >= j) {
            const double v = A[i + (int64_t) j * lda];
            s = fma(v, v, s);
        }
    }
    for(int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ double part[8];
    if((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if(threadIdx.x == 0) {
        double tot = 0.0;
        for(int w = 0; w < 8; w++) tot += part[w];
        atomicAdd(out, tot);
    }
}

// *out += d^T A d for the symmetric matrix given by its LOWER triangle (Mahalanobis term of the KLD, utils.cpp:88)
__global__ void __launch_bounds__(256) quad_form_lower_kernel(const double *__restrict__ A, int64_t lda, int n, const double *__restrict__ d,
                                                              double *out) {
    double s = 0.0;
    for(int64_t t = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; t < (int64_t) n * n; t += (int64_t) gridDim.x * blockDim.x) {
        const int i = (int) (t % n), j = (int) (t / n);
        if(i >= j) {
            const double v = A[i + (int64_t) j * lda] * d[i] * d[j];
            s += (i == j) ? v : 2.0 * v;
        }
    }
    for(int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __shared__ double part[8];
    if((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if(threadIdx.x == 0) {
        double tot = 0.0;
        for(int w = 0; w < 8; w++) tot += part[w];
        atomicAdd(out, tot);
    }
}

} // namespace spg
