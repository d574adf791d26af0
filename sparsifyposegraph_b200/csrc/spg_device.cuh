// spg_device.cuh — device-side building blocks of the blanket kernels (sm_100a).
//
// Everything here is "group cooperative": one CTA of NT threads owns one Markov blanket, the
// matrices live in shared memory (column-major, odd leading dimension so that both column and
// row walks are bank-conflict free for fp64), and every routine is called by all NT threads.
// FP64 vector pipe only — the matrices are 6x6 ... ~100x100, far below a DMMA tile economy.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace spg {

template <int NT>
__device__ __forceinline__ void gsync() {
    if constexpr(NT <= 32) __syncwarp();
    else __syncthreads();
}

__host__ __device__ __forceinline__ int odd_ld(int n) { return n | 1; }

// ------------------------------------------------------------------------------------------
// poses
// ------------------------------------------------------------------------------------------
// SE3 pose in shared memory: R (row-major 3x3) then t  -> 12 doubles.
// SE2 pose: x y theta cos sin -> padded to 6 doubles.
template <int D> struct PoseStride { static constexpr int value = (D == 6) ? 12 : 6; };

__device__ __forceinline__ double normalize_theta(double theta) {
    if(theta >= -M_PI && theta < M_PI) return theta;
    double multiplier = floor(theta / (2 * M_PI));
    theta = theta - multiplier * 2 * M_PI;
    if(theta >= M_PI) theta -= 2 * M_PI;
    if(theta < -M_PI) theta += 2 * M_PI;
    return theta;
}

__device__ __forceinline__ void quat_to_R(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen::Quaterniond(Matrix3d) followed by g2o normalize (unit, w >= 0).
// STATIC_IDX: the rare branch (trace <= 0) written out per case, so that R and q are only ever indexed with constants
// and the caller's matrices stay in registers; without it the dynamic indices put them in local memory, which the
// register-starved multi-warp kernels prefer (fewer spills inside their sweeps). Same arithmetic either way.
template <bool STATIC_IDX = false>
__device__ __forceinline__ void R_to_quat(const double R[9], double q[4]) {
    double t = R[0] + R[4] + R[8];
    if(t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t;
        q[1] = (R[2] - R[6]) * t;
        q[2] = (R[3] - R[1]) * t;
    } else if constexpr(STATIC_IDX) {
        // largest diagonal entry i, then j = i + 1, k = i + 2 (mod 3); the three cases are written out so that R and q
        // are only ever indexed with constants (a dynamic index would put the caller's matrices in local memory)
        int i = 0;
        if(R[4] > R[0]) i = 1;
        if(R[8] > (i ? R[4] : R[0])) i = 2;
        if(i == 0) {
            t = sqrt(R[0] - R[4] - R[8] + 1.0);
            q[0] = 0.5 * t;
            t = 0.5 / t;
            q[3] = (R[7] - R[5]) * t;
            q[1] = (R[3] + R[1]) * t;
            q[2] = (R[6] + R[2]) * t;
        } else if(i == 1) {
            t = sqrt(R[4] - R[8] - R[0] + 1.0);
            q[1] = 0.5 * t;
            t = 0.5 / t;
            q[3] = (R[2] - R[6]) * t;
            q[2] = (R[7] + R[5]) * t;
            q[0] = (R[1] + R[3]) * t;
        } else {
            t = sqrt(R[8] - R[0] - R[4] + 1.0);
            q[2] = 0.5 * t;
            t = 0.5 / t;
            q[3] = (R[3] - R[1]) * t;
            q[0] = (R[2] + R[6]) * t;
            q[1] = (R[5] + R[7]) * t;
        }
    } else {
        int i = 0;
        if(R[4] > R[0]) i = 1;
        if(R[8] > R[3 * i + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[3 * i + i] - R[3 * j + j] - R[3 * k + k] + 1.0);
        double qq[4];
        qq[i] = 0.5 * t;
        t = 0.5 / t;
        qq[3] = (R[3 * k + j] - R[3 * j + k]) * t;
        qq[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        qq[k] = (R[3 * k + i] + R[3 * i + k]) * t;
        q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
    }
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double s = (q[3] < 0) ? -1.0 / n : 1.0 / n;
    q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}

// flat (t, qx qy qz qw) -> (R, t), quaternion normalised like g2o::internal::fromVectorQT
__device__ __forceinline__ void se3_from_flat(const double *v, double *p) {
    double q[4] = {v[3], v[4], v[5], v[6]};
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double s = (q[3] < 0) ? -1.0 / n : 1.0 / n;
    q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
    quat_to_R(q, p);
    p[9] = v[0]; p[10] = v[1]; p[11] = v[2];
}
template <bool STATIC_IDX = false>
__device__ __forceinline__ void se3_to_flat(const double *p, double *v) {
    double q[4];
    R_to_quat<STATIC_IDX>(p, q);
    v[0] = p[9]; v[1] = p[10]; v[2] = p[11];
    v[3] = q[0]; v[4] = q[1]; v[5] = q[2]; v[6] = q[3];
}
__device__ __forceinline__ void se3_compose(const double *a, const double *b, double *r) {
#pragma unroll
    for(int i = 0; i < 3; i++) {
#pragma unroll
        for(int j = 0; j < 3; j++) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
        r[9 + i] = a[3 * i] * b[9] + a[3 * i + 1] * b[10] + a[3 * i + 2] * b[11] + a[9 + i];
    }
}
__device__ __forceinline__ void se3_inverse(const double *a, double *r) {
#pragma unroll
    for(int i = 0; i < 3; i++)
#pragma unroll
        for(int j = 0; j < 3; j++) r[3 * i + j] = a[3 * j + i];
#pragma unroll
    for(int i = 0; i < 3; i++) r[9 + i] = -(r[3 * i] * a[9] + r[3 * i + 1] * a[10] + r[3 * i + 2] * a[11]);
}
// g2o::internal::fromVectorMQT
__device__ __forceinline__ void se3_from_mqt(const double *v, double *p) {
    double w = 1 - (v[3] * v[3] + v[4] * v[4] + v[5] * v[5]);
    if(w < 0) {
        p[0] = 1; p[1] = 0; p[2] = 0; p[3] = 0; p[4] = 1; p[5] = 0; p[6] = 0; p[7] = 0; p[8] = 1;
    } else {
        double q[4] = {v[3], v[4], v[5], sqrt(w)};
        quat_to_R(q, p);
    }
    p[9] = v[0]; p[10] = v[1]; p[11] = v[2];
}
// g2o::internal::toVectorMQT
__device__ __forceinline__ void se3_to_mqt(const double *p, double *v) {
    double q[4];
    R_to_quat(p, q);
    v[0] = p[9]; v[1] = p[10]; v[2] = p[11];
    v[3] = q[0]; v[4] = q[1]; v[5] = q[2];
}

__device__ __forceinline__ void se2_from_flat(const double *v, double *p) {
    p[0] = v[0]; p[1] = v[1];
    p[2] = normalize_theta(v[2]);
    p[3] = cos(p[2]); p[4] = sin(p[2]); p[5] = 0;
}
__device__ __forceinline__ void se2_compose(const double *a, const double *b, double *r) {
    r[0] = a[0] + a[3] * b[0] - a[4] * b[1];
    r[1] = a[1] + a[4] * b[0] + a[3] * b[1];
    r[2] = normalize_theta(a[2] + b[2]);
    r[3] = cos(r[2]); r[4] = sin(r[2]); r[5] = 0;
}
__device__ __forceinline__ void se2_inverse(const double *a, double *r) {
    r[2] = normalize_theta(-a[2]);
    r[3] = cos(r[2]); r[4] = sin(r[2]); r[5] = 0;
    r[0] = r[3] * (-a[0]) - r[4] * (-a[1]);
    r[1] = r[4] * (-a[0]) + r[3] * (-a[1]);
}

// d err / d delta_i | d err / d delta_j for one relative-pose edge; J is d x 2d column-major
// (ld = D): columns [0,D) = Ji, [D,2D) = Jj.
//   SE2: EdgeSE2ISAM::linearizeOplus, reference src/se2_compatibility.h:35-51
//   SE3: analytic derivative of toVectorMQT(Z^-1 (Xi (+) di)^-1 (Xj (+) dj)) — the quantity
//        g2o::internal::computeEdgeSE3Gradient evaluates (call site src/se3_compatibility.h:25-29);
//        closed form derived in DESIGN.md §Jacobians.
template <int D, bool STATIC_IDX = false>
__device__ __noinline__ void edge_jacobians(const double *Z, const double *Xi, const double *Xj, double *J) {
    if constexpr(D == 3) {
        const double dtx = Xj[0] - Xi[0], dty = Xj[1] - Xi[1];
        const double ci = Xi[3], si = Xi[4];
        // Ji
        J[0] = -ci; J[3] = -si; J[6] = -si * dtx + ci * dty;
        J[1] = si;  J[4] = -ci; J[7] = -ci * dtx - si * dty;
        J[2] = 0;   J[5] = 0;   J[8] = -1;
        // Jj
        J[9] = ci;   J[12] = si; J[15] = 0;
        J[10] = -si; J[13] = ci; J[16] = 0;
        J[11] = 0;   J[14] = 0;  J[17] = 1;
    } else {
        double A[12], B[12], T[12], E[12];
        se3_inverse(Z, A);
        se3_inverse(Xi, T);
        se3_compose(T, Xj, B);
        se3_compose(A, B, E);
        double qA[4], qB[4];
        R_to_quat<STATIC_IDX>(A, qA);
        R_to_quat<STATIC_IDX>(B, qB);
        const double wA = qA[3], wB = qB[3];
        double wE = wA * wB - (qA[0] * qB[0] + qA[1] * qB[1] + qA[2] * qB[2]);
        double vE[3] = {wA * qB[0] + wB * qA[0] + (qA[1] * qB[2] - qA[2] * qB[1]),
                        wA * qB[1] + wB * qA[1] + (qA[2] * qB[0] - qA[0] * qB[2]),
                        wA * qB[2] + wB * qA[2] + (qA[0] * qB[1] - qA[1] * qB[0])};
        const double nE = sqrt(wE * wE + vE[0] * vE[0] + vE[1] * vE[1] + vE[2] * vE[2]);
        const double sigma = (wE < 0) ? -1.0 : 1.0;
        wE = sigma * wE / nE;
        vE[0] = sigma * vE[0] / nE; vE[1] = sigma * vE[1] / nE; vE[2] = sigma * vE[2] / nE;
        // skew matrices (row-major)
        const double SA[9] = {0, -qA[2], qA[1], qA[2], 0, -qA[0], -qA[1], qA[0], 0};
        const double SB[9] = {0, -qB[2], qB[1], qB[2], 0, -qB[0], -qB[1], qB[0], 0};
        const double SE[9] = {0, -vE[2], vE[1], vE[2], 0, -vE[0], -vE[1], vE[0], 0};
        const double St[9] = {0, -B[11], B[10], B[11], 0, -B[9], -B[10], B[9], 0};
#pragma unroll
        for(int i = 0; i < 72; i++) J[i] = 0;
#pragma unroll
        for(int r = 0; r < 3; r++)
#pragma unroll
            for(int c = 0; c < 3; c++) {
                J[r + 6 * c] = -A[3 * r + c];
                J[r + 6 * (3 + c)] = 2 * (A[3 * r] * St[c] + A[3 * r + 1] * St[3 + c] + A[3 * r + 2] * St[6 + c]);
                J[36 + r + 6 * c] = E[3 * r + c];
                const double sbsa = SB[3 * r] * SA[c] + SB[3 * r + 1] * SA[3 + c] + SB[3 * r + 2] * SA[6 + c];
                const double M = (r == c ? wA * wB : 0.0) + wB * SA[3 * r + c] - wA * SB[3 * r + c] - sbsa - qB[r] * qA[c];
                J[(3 + r) + 6 * (3 + c)] = -sigma * M;
                J[36 + (3 + r) + 6 * (3 + c)] = (r == c ? wE : 0.0) + SE[3 * r + c];
            }
    }
}

// Jacobians of a NEW substitute edge, whose measurement was just set from the state
// (setMeasurementFromState, topology_provider_binary.hpp:43-47), i.e. at zero error E = I.
// SE3: with Z = Xa^-1 Xb = (R, p) the general formulas collapse to
//   Jj = I,   Ji = [[-R^T, 2 R^T [p]x], [0, -R^T]]     (FD-checked in tests/test_oracle.py)
// SE2: the ISAM-style Jacobians do not depend on the measurement at all.
template <int D>
__device__ __forceinline__ void edge_jacobians_zero_error(const double *Z, const double *Xi, const double *Xj, double *J) {
    if constexpr(D == 3) {
        edge_jacobians<D>(Z, Xi, Xj, J);
    } else {
#pragma unroll
        for(int i = 0; i < 72; i++) J[i] = 0;
        const double px = Z[9], py = Z[10], pz = Z[11];
        const double S[9] = {0, -pz, py, pz, 0, -px, -py, px, 0}; // [p]x row-major
#pragma unroll
        for(int r = 0; r < 3; r++)
#pragma unroll
            for(int c = 0; c < 3; c++) {
                const double rt = Z[3 * c + r]; // R^T[r][c]
                J[r + 6 * c] = -rt;
                J[(3 + r) + 6 * (3 + c)] = -rt;
                J[r + 6 * (3 + c)] = 2 * (Z[r] * S[c] + Z[3 + r] * S[3 + c] + Z[6 + r] * S[6 + c]); // 2 (R^T [p]x)[r][c]
            }
#pragma unroll
        for(int d = 0; d < 6; d++) J[36 + d + 6 * d] = 1.0;
    }
}

// ------------------------------------------------------------------------------------------
// cooperative dense linear algebra on shared-memory matrices (column-major)
// ------------------------------------------------------------------------------------------

template <int NT>
__device__ __forceinline__ int gsync_or(int pred) {
    if constexpr(NT <= 32) return __any_sync(0xffffffffu, pred);
    else return __syncthreads_or(pred);
}

// Barrier / OR-reduction over either the whole CTA (BAR == 0) or a group of GS threads on named barrier BAR
// (GS a multiple of 32; every thread of the group must call it).
template <int NT, int BAR, int GS>
__device__ __forceinline__ void group_sync() {
    if constexpr(BAR == 0) gsync<NT>();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(GS) : "memory");
}
template <int NT, int BAR, int GS>
__device__ __forceinline__ int group_or(int pred) {
    if constexpr(BAR == 0) return gsync_or<NT>(pred);
    else {
        int r;
        asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbar.red.or.pred q, %2, %3, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
                     : "=r"(r) : "r"(pred), "n"(BAR), "n"(GS) : "memory");
        return r;
    }
}

// In-place lower Cholesky A = L L^T of the leading n x n block (reads the lower triangle).
// Returns false (uniformly) if a pivot is not > 0.
// Blocked right-looking, panel width CHOL_NB: (1) the NB x NB diagonal block is factored by one
// thread, (2) the panel below it by one thread per row (triangular solve against the block),
// (3) the trailing matrix is updated by ALL threads on a 2-D (row, column) grid with NB-long dot
// products. 3 barriers per panel instead of 2 per column.
constexpr int CHOL_NB = 8;
template <int NT>
__device__ bool chol_lower(double *A, int n, int ld) {
    constexpr int TX = (NT >= 128) ? 16 : 8;
    constexpr int TY = NT / TX;
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    for(int j0 = 0; j0 < n; j0 += CHOL_NB) {
        const int jb = min(CHOL_NB, n - j0);
        int bad = 0;
        if(tid == 0) {
            for(int c = 0; c < jb; c++) {
                const int jc = j0 + c;
                double d = A[jc + jc * ld];
                for(int p = 0; p < c; p++) d -= A[jc + (j0 + p) * ld] * A[jc + (j0 + p) * ld];
                if(!(d > 0)) { bad = 1; break; }
                const double l = sqrt(d), inv = 1.0 / l;
                A[jc + jc * ld] = l;
                for(int r = c + 1; r < jb; r++) {
                    double s = A[j0 + r + jc * ld];
                    for(int p = 0; p < c; p++) s -= A[j0 + r + (j0 + p) * ld] * A[jc + (j0 + p) * ld];
                    A[j0 + r + jc * ld] = s * inv;
                }
            }
        }
        if(gsync_or<NT>(bad)) return false;
        const int i0 = j0 + jb;
        for(int i = i0 + tid; i < n; i += NT) { // panel: row i against the diagonal block
            for(int c = 0; c < jb; c++) {
                double s = A[i + (j0 + c) * ld];
                for(int p = 0; p < c; p++) s -= A[i + (j0 + p) * ld] * A[j0 + c + (j0 + p) * ld];
                A[i + (j0 + c) * ld] = s / A[j0 + c + (j0 + c) * ld];
            }
        }
        gsync<NT>();
        for(int l = i0 + ty; l < n; l += TY) // trailing update, lower triangle
            for(int i = l + tx; i < n; i += TX) {
                double s = 0;
                for(int p = 0; p < jb; p++) s += A[i + (j0 + p) * ld] * A[l + (j0 + p) * ld];
                A[i + l * ld] -= s;
            }
        gsync<NT>();
    }
    return true;
}

// X = (L L^T)^-1 for the Cholesky factor L (n x n lower, ld). The result (full symmetric matrix)
// OVERWRITES L; Y (n x ld) is scratch. Two fully parallel phases:
//   A) Y = L^-1 by blocked forward substitution (GEMM update of a row panel by all threads, then one
//      thread per column solves against the NB x NB diagonal block),
//   B) X = Y^T Y on a 2-D grid.
template <int NT>
__device__ void chol_inverse_inplace(double *L, int n, int ld, double *Y) {
    constexpr int TX = (NT >= 128) ? 16 : 8;
    constexpr int TY = NT / TX;
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    for(int j0 = 0; j0 < n; j0 += CHOL_NB) {
        const int jb = min(CHOL_NB, n - j0);
        // rhs of panel rows: -L[panel, <j0] * Y[<j0, c] for c < j0 ; identity for c in the panel
        for(int c = ty; c < j0 + jb; c += TY)
            for(int r = tx; r < jb; r += TX) {
                double s = (c == j0 + r) ? 1.0 : 0.0;
                for(int q = c; q < j0; q++) s -= L[j0 + r + q * ld] * Y[q + c * ld];
                Y[j0 + r + c * ld] = s;
            }
        gsync<NT>();
        for(int c = tid; c < j0 + jb; c += NT) { // triangular solve with the diagonal block
            double *y = Y + j0 + (size_t) c * ld;
            for(int r = 0; r < jb; r++) {
                double s = y[r];
                for(int q = 0; q < r; q++) s -= L[j0 + r + (j0 + q) * ld] * y[q];
                y[r] = s / L[j0 + r + (j0 + r) * ld];
            }
        }
        gsync<NT>();
    }
    // X = Y^T Y : X[i][j] = sum_{p >= max(i,j)} Y[p][i] Y[p][j]; written over L (L is dead)
    for(int j = ty; j < n; j += TY)
        for(int i = j + tx; i < n; i += TX) {
            double s0 = 0, s1 = 0;
            int p = i;
            for(; p + 1 < n; p += 2) {
                s0 += Y[p + i * ld] * Y[p + j * ld];
                s1 += Y[p + 1 + i * ld] * Y[p + 1 + j * ld];
            }
            if(p < n) s0 += Y[p + i * ld] * Y[p + j * ld];
            L[i + j * ld] = s0 + s1;
            L[j + i * ld] = s0 + s1;
        }
    gsync<NT>();
}

// 1/x for a normal, positive x: MUFU.RCP64H seed + two Newton steps (<= 1 ulp). The IEEE division of
// CUDA costs ~140 cycles on the critical path of a pivot step, this one ~70.
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// One pivot step of the register-tiled sweep (see sweep_spd). C0 = register row/column of the pivot.
template <int T, int TS, int C0>
__device__ __forceinline__ void sweep_step(double (&a)[TS][TS], double *colbuf, int jj, int nsweep, int tx, int ty, bool &bad,
                                           double &mypiv) {
    constexpr int NP = T * TS;
    constexpr int cstride = NP + 2;
    const int j = jj + T * C0;
    const double *col = colbuf + (j & 1) * cstride;
    double *ncol = colbuf + ((j + 1) & 1) * cstride;
    const double d = col[NP], inv = col[NP + 1];
    bad |= !(d > 0);
    if(tx + T * ty == j) mypiv = d; // group thread j keeps pivot j (log-determinant of the swept block)
    double ci[TS], cl[TS];
#pragma unroll
    for(int r = 0; r < TS; r++) {
        ci[r] = col[tx + T * r];
        cl[r] = col[ty + T * r] * inv;
    }
    const bool rowj = (tx == jj), colj = (ty == jj);
    const bool wrap = (jj + 1 == T);
    const bool more = (j + 1 < nsweep);
    constexpr int CN = (C0 + 1) % TS;
    // look-ahead: the next pivot column (register column C0, or C0 + 1 at the end of a block) is updated
    // and published first, so that the store -> barrier -> load latency hides behind the bulk of the tile
    if(!wrap) {
#pragma unroll
        for(int r = 0; r < TS; r++) a[r][C0] -= ci[r] * cl[C0];
        if(rowj) a[C0][C0] = cl[C0];
        if(colj) { // this thread's register column C0 is matrix column j itself
#pragma unroll
            for(int r = 0; r < TS; r++) a[r][C0] = ci[r] * inv;
            if(rowj) a[C0][C0] = -inv;
        }
        if(more && ty == jj + 1) {
#pragma unroll
            for(int r = 0; r < TS; r++) ncol[tx + T * r] = a[r][C0];
            if(tx == jj + 1) { ncol[NP] = a[C0][C0]; ncol[NP + 1] = fast_rcp(a[C0][C0]); }
        }
    } else if(C0 + 1 < TS) {
#pragma unroll
        for(int r = 0; r < TS; r++) a[r][CN] -= ci[r] * cl[CN];
        if(rowj) a[C0][CN] = cl[CN];
        if(more && ty == 0) {
#pragma unroll
            for(int r = 0; r < TS; r++) ncol[tx + T * r] = a[r][CN];
            if(tx == 0) { ncol[NP] = a[CN][CN]; ncol[NP + 1] = fast_rcp(a[CN][CN]); }
        }
    }
#pragma unroll
    for(int c = 0; c < TS; c++) {
        const bool done_ahead = (!wrap && c == C0) || (wrap && c == C0 + 1);
        if(!done_ahead) {
#pragma unroll
            for(int r = 0; r < TS; r++) a[r][c] -= ci[r] * cl[c];
            if(rowj) a[C0][c] = cl[c];
        }
    }
    if(wrap && colj) { // column j lives in register column C0 and was updated in the bulk
#pragma unroll
        for(int r = 0; r < TS; r++) a[r][C0] = ci[r] * inv;
        if(rowj) a[C0][C0] = -inv;
    }
}

// Blocks C0, C0 + 1, ... of T pivots each: the register index of the pivot row/column is a template
// constant, so the tile never leaves the register file; the jj loop is not unrolled (a fully unrolled
// sweep of T*TS steps thrashes the instruction cache).
template <int NT, int T, int TS, int C0, int BAR>
__device__ __forceinline__ void sweep_blocks(double (&a)[TS][TS], double *colbuf, int nsweep, int tx, int ty, bool active, bool &bad,
                                             double &mypiv) {
    if constexpr(C0 < TS) {
        const int jjmax = min(T, nsweep - T * C0); // uniform; <= 0: nothing left
#pragma unroll 1
        for(int jj = 0; jj < jjmax; jj++) {
            if(active) sweep_step<T, TS, C0>(a, colbuf, jj, nsweep, tx, ty, bad, mypiv);
            group_sync<NT, BAR, T * T>();
        }
        sweep_blocks<NT, T, TS, C0 + 1, BAR>(a, colbuf, nsweep, tx, ty, active, bad, mypiv);
    }
}

// Register-tiled symmetric sweep (Gauss-Jordan on an SPD matrix, no pivoting): sweeps the first
// `nsweep` pivots of the n x n matrix A (shared memory, ld) in place.
//   nsweep == n : A <- A^-1
//   nsweep == m : the trailing block A[m:,m:] <- A_kk - A_km A_mm^-1 A_mk (Schur complement); the
//                 first m rows/columns hold sweep by-products and must be ignored.
// Every thread of a T x T grid keeps a TS x TS cyclic tile of A in registers for the whole sweep;
// per step only the pivot column travels through shared memory (double-buffered in `colbuf`,
// 2 * (T*TS + 2) doubles) and one barrier is needed. Pivots are the Schur-complement diagonals
// (= squared Cholesky pivots), so "pivot > 0" is the same positive-definiteness test as LLT's; a
// non-positive pivot is remembered and reported at the end (the arithmetic in between is discarded).
// Requires T*T <= NT and n <= T*TS. Returns false (uniformly) on a non-positive pivot.
// Measured (tools/sweep_bench.cu, B200): ~700 cycles per pivot step at every tile shape: operand
// delivery through shared memory (24 wavefronts per warp and step) plus the FP64 issue slots, not latency —
// sharing the barriers between two independent sweeps gains 4 %, deferring the rank-T update into the
// next block's bubbles or handing the pivot columns to a dedicated warp is slower. See DESIGN.md.
// BAR > 0: the sweep is run by a group of exactly T*T threads (group-local index `tid`) that synchronises on
// named barrier BAR, so that two groups of one CTA can sweep two matrices at the same time.
template <int NT, int T, int TS, int BAR = 0>
__device__ __forceinline__ bool sweep_spd_inl(const double *S, int lds, double *A, int ld, int n, int nsweep, double *colbuf, int tid,
                                              double diag_add, bool mirror, double *logpiv = nullptr) {
    static_assert(T * T <= NT, "thread grid larger than the CTA");
    constexpr int NTG = (BAR == 0) ? NT : T * T; // threads taking part
    const bool active = tid < T * T;
    const int tx = tid % T, ty = (tid / T) % T;
    constexpr int NP = T * TS;      // padded length: loads / stores of the pivot column need no bounds checks
    constexpr int cstride = NP + 2; // the pivot column, then the pivot d and 1/d
    for(int t = n + tid; t < NP; t += NTG) { colbuf[t] = 0.0; colbuf[cstride + t] = 0.0; }
    double a[TS][TS];
#pragma unroll
    for(int c = 0; c < TS; c++)
#pragma unroll
        for(int r = 0; r < TS; r++) {
            const int i = tx + T * r, l = ty + T * c;
            a[r][c] = (active && i < n && l < n) ? S[i + l * lds] + (i == l ? diag_add : 0.0) : 0.0;
        }
    // publish column 0
    if(active && ty == 0) {
#pragma unroll
        for(int r = 0; r < TS; r++) colbuf[tx + T * r] = a[r][0];
        if(tx == 0) { colbuf[NP] = a[0][0]; colbuf[NP + 1] = fast_rcp(a[0][0]); }
    }
    group_sync<NT, BAR, T * T>();
    bool bad = false;
    double mypiv = 1.0;
    sweep_blocks<NT, T, TS, 0, BAR>(a, colbuf, nsweep, tx, ty, active, bad, mypiv);
    if(group_or<NT, BAR, T * T>(bad)) return false;
    // log of "my" pivot: summed over the group it is the log-determinant of the swept leading block
    if(logpiv) *logpiv = (active && tx + T * ty < nsweep) ? log(mypiv) : 0.0;
    // full sweep leaves -A^-1; a partial sweep leaves the Schur complement in the trailing block
    const double sgn = (nsweep >= n) ? -1.0 : 1.0;
#pragma unroll
    for(int c = 0; c < TS; c++)
#pragma unroll
        for(int r = 0; r < TS; r++) {
            const int i = tx + T * r, l = ty + T * c;
            if(active && i < n && l < n) {
                if(!mirror) A[i + l * ld] = sgn * a[r][c];
                else if(i <= l) { // upper triangle copied onto the lower one (vertex_remover.cpp:447-449)
                    A[i + l * ld] = sgn * a[r][c];
                    A[l + i * ld] = sgn * a[r][c];
                }
            }
        }
    group_sync<NT, BAR, T * T>();
    return true;
}

template <int NT, int T, int TS, int BAR = 0>
__device__ __noinline__ bool sweep_spd(const double *S, int lds, double *A, int ld, int n, int nsweep, double *colbuf, int tid, double diag_add,
                                       bool mirror, double *logpiv = nullptr) {
    return sweep_spd_inl<NT, T, TS, BAR>(S, lds, A, ld, n, nsweep, colbuf, tid, diag_add, mirror, logpiv);
}

// Two tile shapes per CTA width: the largest n each covers is T * TS.
template <int NT> struct SweepGrid;
template <> struct SweepGrid<32>  { static constexpr int T = 4,  TS0 = 2, TS1 = 3; };   // n <= 8 / 12
template <> struct SweepGrid<64>  { static constexpr int T = 8,  TS0 = 3, TS1 = 4; };   // n <= 24 / 32
template <> struct SweepGrid<128> { static constexpr int T = 8,  TS0 = 5, TS1 = 6; };   // n <= 40 / 48 (7 x 7 spills)
template <> struct SweepGrid<256> { static constexpr int T = 16, TS0 = 5, TS1 = 6; };   // n <= 80 / 96
template <> struct SweepGrid<512> { static constexpr int T = 16, TS0 = 4, TS1 = 5; };   // n <= 64 / 80: two groups of 256, 128 registers each (6 x 6 spills)

// returns 1 ok, 0 not positive definite, -1 n too large for the register tiles. scratch: 2 * (T*TS + 2) doubles.
// (Blocked variants were measured and dropped: a per-vertex rank-6 sweep took 186k vs 100k cycles at
// k = 90 and a rank-2 sweep 79k vs 78k — the extra panel operands spill next to the 6 x 6 tile, and the
// kernel is issue-latency bound at 8 warps per SM rather than barrier bound. See DESIGN.md §3.)
// LEAN: the 256-thread CTA built for two residents per SM (128 registers): the tile shapes of the 512-thread grid.
template <int D, int NT, bool LEAN = false>
__device__ __forceinline__ int sweep_spd_auto(const double *S, int lds, double *A, int ld, int n, int nsweep, double *scratch,
                                              double diag_add = 0.0, bool mirror = false, double *logpiv = nullptr) {
    using G = SweepGrid<LEAN ? 512 : NT>;
    if(n <= G::T * G::TS0) return sweep_spd<NT, G::T, G::TS0>(S, lds, A, ld, n, nsweep, scratch, threadIdx.x, diag_add, mirror, logpiv) ? 1 : 0;
    if(n <= G::T * G::TS1) return sweep_spd<NT, G::T, G::TS1>(S, lds, A, ld, n, nsweep, scratch, threadIdx.x, diag_add, mirror, logpiv) ? 1 : 0;
    return -1;
}
// does an n x n matrix fit the register tiles of this CTA shape?
template <int NT, bool LEAN = false>
__device__ __forceinline__ bool sweep_fits(int n) { return n <= SweepGrid<LEAN ? 512 : NT>::T * SweepGrid<LEAN ? 512 : NT>::TS1; }

// CTAs wide enough sweep two matrices at once, one per group of GS threads (group g = threads [g*GS, (g+1)*GS)):
//   128 threads: two 8 x 8 grids, tiles up to 6 x 6 (n <= 48)
//   512 threads: two 16 x 16 grids, tiles 4 x 4 / 5 x 5 (n <= 64 / 80), 128 registers (6 x 6 spills)
// CB = doubles of pivot-column scratch per group. (A 256-thread CTA with two 16 x 8 groups and 6 x 12 tiles
// covers n <= 96 but measured no faster than two 16 x 16 sweeps one after the other: tools/sweep_bench.cu.)
template <int NT> struct SweepDual { static constexpr bool value = false; static constexpr int GS = NT, NMAX = 0, CB = 0; };
template <> struct SweepDual<128> { static constexpr bool value = true; static constexpr int GS = 64, NMAX = 48, CB = 100; };
template <> struct SweepDual<512> { static constexpr bool value = true; static constexpr int GS = 256, NMAX = 80, CB = 164; };

// same return convention as sweep_spd_auto
template <int D, int NT, int BAR>
__device__ __forceinline__ int sweep_spd_group(const double *S, int lds, double *A, int ld, int n, int nsweep, double *scratch, int gtid,
                                               double diag_add, double *logpiv = nullptr) {
    using G = SweepGrid<NT>;
    if(n <= G::T * G::TS0) return sweep_spd<NT, G::T, G::TS0, BAR>(S, lds, A, ld, n, nsweep, scratch, gtid, diag_add, false, logpiv) ? 1 : 0;
    if(n <= G::T * G::TS1) return sweep_spd<NT, G::T, G::TS1, BAR>(S, lds, A, ld, n, nsweep, scratch, gtid, diag_add, false, logpiv) ? 1 : 0;
    return -1;
}

// X = (L L^T)^-1 written to X (n x n, ldx) from the Cholesky factor L (lower, ld). Thread per
// column: forward then backward substitution on its own column of X (no syncs inside); the dot
// products run on two independent accumulators to halve the DFMA dependency chain.
template <int NT>
__device__ void chol_inverse(const double *L, int n, int ld, double *X, int ldx) {
    for(int c = threadIdx.x; c < n; c += NT) {
        double *x = X + (size_t) c * ldx;
        for(int i = 0; i < c; i++) x[i] = 0;
        for(int i = c; i < n; i++) { // forward: L y = e_c
            double s0 = (i == c) ? 1.0 : 0.0, s1 = 0.0;
            int p = c;
            for(; p + 1 < i; p += 2) {
                s0 -= L[i + p * ld] * x[p];
                s1 -= L[i + (p + 1) * ld] * x[p + 1];
            }
            if(p < i) s0 -= L[i + p * ld] * x[p];
            x[i] = (s0 + s1) / L[i + i * ld];
        }
        for(int i = n - 1; i >= 0; i--) { // backward: L^T z = y
            double s0 = x[i], s1 = 0.0;
            int p = i + 1;
            for(; p + 1 < n; p += 2) {
                s0 -= L[p + i * ld] * x[p];
                s1 -= L[p + 1 + i * ld] * x[p + 1];
            }
            if(p < n) s0 -= L[p + i * ld] * x[p];
            x[i] = (s0 + s1) / L[i + i * ld];
        }
    }
}

// Parallel-order two-sided Jacobi eigen-decomposition of the symmetric n x n matrix A (ld),
// destroying A (eigenvalues end on its diagonal) and accumulating eigenvectors in the columns
// of V (ldv). cs: scratch of 3*((n+1)/2) doubles (c, s and the packed pair); red: one double.
// Round-robin tournament ordering: (n_even - 1) rounds of n_even/2 disjoint rotations; each round
// applies J^T A J as independent 2x2 blocks (Brent-Luk) and V J, indexed through a per-round pair
// table so the inner loops carry no integer division.
// Returns the number of sweeps used, or -1 if not converged.
template <int NT>
__device__ int jacobi_eig(double *A, int n, int ld, double *V, int ldv, double *cs, double *red) {
    constexpr int TX = (NT >= 128) ? 16 : 8;
    constexpr int TY = NT / TX;
    // NT <= 32: a single warp, which may be any warp of a wider CTA (the GLC factors of a blanket are finished by
    // several warps side by side)
    const int tid = (NT <= 32) ? (int) (threadIdx.x & 31) : (int) threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    for(int j = ty; j < n; j += TY)
        for(int i = tx; i < n; i += TX) V[i + j * ldv] = (i == j) ? 1.0 : 0.0;
    if(n < 2) {
        gsync<NT>();
        return 0;
    }
    const int ne = (n + 1) & ~1, np = ne / 2;
    int *pq = reinterpret_cast<int *>(cs + 2 * np);
    if(tid == 0) {
        double m = 0;
        for(int j = 0; j < n; j++)
            for(int i = 0; i < n; i++) m = fmax(m, fabs(A[i + j * ld]));
        red[0] = m;
    }
    gsync<NT>();
    const double scale = red[0];
    const double tol = 8.0 * 2.220446049250313e-16 * scale;
    gsync<NT>();
    const int max_sweeps = 30;
    for(int sweep = 0; sweep < max_sweeps; sweep++) {
        double mymax = 0;
        for(int r = 0; r < ne - 1; r++) {
            // phase 1: pair table + rotation parameters of this round
            for(int i = tid; i < np; i += NT) {
                int p, q;
                if(i == 0) { p = ne - 1; q = r; }
                else {
                    p = r + i; if(p >= ne - 1) p -= ne - 1;
                    q = r - i; if(q < 0) q += ne - 1;
                }
                if(p > q) { int tmp = p; p = q; q = tmp; }
                double c = 1.0, s = 0.0;
                if(q < n) {
                    const double apq = A[p + q * ld];
                    mymax = fmax(mymax, fabs(apq));
                    if(fabs(apq) > 1e-300) {
                        const double app = A[p + p * ld], aqq = A[q + q * ld];
                        const double tau = (aqq - app) / (2.0 * apq);
                        const double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                    pq[i] = p | (q << 16);
                } else {
                    pq[i] = p | (p << 16) | 0x40000000; // padding partner: identity on a single index
                }
                cs[2 * i] = c;
                cs[2 * i + 1] = s;
            }
            gsync<NT>();
            // phase 2a: A <- J^T A J on (pair, pair) blocks
            for(int i2 = ty; i2 < np; i2 += TY) {
                const int w2 = pq[i2];
                const int p2 = w2 & 0x7fff, q2 = (w2 >> 16) & 0x3fff;
                const bool v2 = !(w2 & 0x40000000);
                const double c2 = cs[2 * i2], s2 = cs[2 * i2 + 1];
                for(int i1 = tx; i1 < np; i1 += TX) {
                    const int w1 = pq[i1];
                    const int p1 = w1 & 0x7fff, q1 = (w1 >> 16) & 0x3fff;
                    const bool v1 = !(w1 & 0x40000000);
                    const double c1 = cs[2 * i1], s1 = cs[2 * i1 + 1];
                    double b00 = A[p1 + p2 * ld];
                    double b01 = v2 ? A[p1 + q2 * ld] : 0.0;
                    double b10 = v1 ? A[q1 + p2 * ld] : 0.0;
                    double b11 = (v1 && v2) ? A[q1 + q2 * ld] : 0.0;
                    const double r00 = c1 * b00 - s1 * b10, r01 = c1 * b01 - s1 * b11;
                    const double r10 = s1 * b00 + c1 * b10, r11 = s1 * b01 + c1 * b11;
                    b00 = c2 * r00 - s2 * r01; b01 = s2 * r00 + c2 * r01;
                    b10 = c2 * r10 - s2 * r11; b11 = s2 * r10 + c2 * r11;
                    if(i1 == i2 && v1) { b01 = 0.0; b10 = 0.0; } // annihilated exactly
                    A[p1 + p2 * ld] = b00;
                    if(v2) A[p1 + q2 * ld] = b01;
                    if(v1) A[q1 + p2 * ld] = b10;
                    if(v1 && v2) A[q1 + q2 * ld] = b11;
                }
                // phase 2b: V <- V J on (row, pair); rows walk a column: conflict free
                if(v2) {
                    double *vp = V + (size_t) p2 * ldv, *vq = V + (size_t) q2 * ldv;
                    for(int row = tx; row < n; row += TX) {
                        const double a = vp[row], b = vq[row];
                        vp[row] = c2 * a - s2 * b;
                        vq[row] = s2 * a + c2 * b;
                    }
                }
            }
            gsync<NT>();
        }
        unsigned conv = (mymax <= tol);
        int all;
        if constexpr(NT <= 32) all = __all_sync(0xffffffffu, conv);
        else all = __syncthreads_and((int) conv);
        if(all) return sweep + 1;
    }
    return -1;
}

} // namespace spg
