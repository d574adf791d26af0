// spg_nfr.cuh — iterative NFR information fit on device (R9): the KLD objective, gradient and
// (reference-style) Hessian of LogdetFunctionWithConstraints (reference src/logdet_function.cpp:
// 119-214, 355-427) inside the interior-point / Newton / simple-backtracking loop of
// optimizeInformation (src/optimizer.cpp:38-79), PQNOptimizer::optimize with useHessian
// (src/pqn/pqn_optimizer.cpp:29-128) and LineSearchSimpleBacktracking (src/pqn/line_search.cpp:12-36).
// Only Subgraph / Dense topologies with >= 3 kept vertices reach this (Tree and CliqueyDense have the
// closed form). One CTA per blanket; U, S and the new-edge Jacobians are in shared memory, the
// iteration state (A, A^-1, P, x, g, the q x q Hessian) lives in a per-CTA global workspace that
// stays L2 resident.
#pragma once
#include "spg_device.cuh"

namespace spg {

// sum over the CTA; every thread gets the result. red: >= 34 doubles of shared scratch.
template <int NT>
__device__ double block_sum(double v, double *red) {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if constexpr(NT <= 32) return v;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if(l == 0) red[w] = v;
    __syncthreads();
    double s = 0;
    for(int i = 0; i < NT / 32; i++) s += red[i];
    return s;
}

struct NfrWork { // per-CTA global workspace carve-up (doubles)
    double *JU, *tmp, *A, *Ai, *P, *x, *g, *d, *xn, *gn, *Xi, *H;
    int R, r, q, ldq;
};
__host__ __device__ inline int64_t nfr_work_doubles(int ne, int D, int r) {
    const int64_t R = (int64_t) ne * D, q = (int64_t) ne * D * D, ldq = q | 1;
    return 2 * R * r + 2 * (int64_t) r * r + R * R + 6 * q + q * ldq + 16;
}
__device__ inline NfrWork nfr_carve(double *ws, int ne, int D, int r) {
    NfrWork w;
    w.R = ne * D; w.r = r; w.q = ne * D * D; w.ldq = w.q | 1;
    double *p = ws;
    w.JU = p;  p += (size_t) w.R * r;
    w.tmp = p; p += (size_t) w.R * r;
    w.A = p;   p += (size_t) r * r;
    w.Ai = p;  p += (size_t) r * r;
    w.P = p;   p += (size_t) w.R * w.R;
    w.x = p;   p += w.q;
    w.g = p;   p += w.q;
    w.d = p;   p += w.q;
    w.xn = p;  p += w.q;
    w.gn = p;  p += w.q;
    w.Xi = p;  p += w.q;
    w.H = p;
    return w;
}

// LogdetFunctionWithConstraints::value at x (A's Cholesky factor is left in w.A for gradient()).
// Returns +inf if A or any X block is not positive definite.
template <int D, int NT>
__device__ double nfr_value(const NfrWork &w, const double *x, int ne, const double *S, double logdetS, double rho,
                            double *red, int *flag) {
    const int tid = threadIdx.x, r = w.r, R = w.R;
    if(tid == 0) *flag = 0;
    // tmp_e = X_e * JU_e with X_e = selfadjointView<Lower>(x_e)
    for(int t = tid; t < R * r; t += NT) {
        const int row = t % R, col = t / R, e = row / D, i = row % D;
        const double *xe = x + (size_t) e * D * D;
        double s = 0;
#pragma unroll
        for(int j = 0; j < D; j++) {
            const double xij = (i >= j) ? xe[i + j * D] : xe[j + i * D];
            s += xij * w.JU[(e * D + j) + (size_t) col * R];
        }
        w.tmp[row + (size_t) col * R] = s;
    }
    __syncthreads();
    // A = JU^T tmp, upper triangle mirrored to the lower (logdet_function.cpp:121-122)
    for(int t = tid; t < r * r; t += NT) {
        const int i = t % r, j = t / r;
        if(i <= j) {
            double s = 0;
            for(int p = 0; p < R; p++) s += w.JU[p + (size_t) i * R] * w.tmp[p + (size_t) j * R];
            w.A[i + (size_t) j * r] = s;
            w.A[j + (size_t) i * r] = s;
        }
    }
    __syncthreads();
    double part = 0;
    for(int i = tid; i < r; i += NT) part += w.A[i + (size_t) i * r] * S[i];
    const double trAS = block_sum<NT>(part, red);
    // barrier term: - rho * sum_e logdet X_e
    double lpart = 0;
    for(int e = tid; e < ne; e += NT) {
        double L[D * D];
        bool ok;
        const double ld = chol_small<D>(x + (size_t) e * D * D, D, L, ok);
        if(!ok) *flag = 1;
        lpart += ld;
    }
    const double ldX = block_sum<NT>(lpart, red);
    const bool okA = chol_lower<NT>(w.A, r, r);
    __syncthreads();
    if(!okA || (*flag && rho != 0.0)) return INFINITY; // the X blocks are only checked by the barrier (:355-370)
    double dpart = 0;
    for(int i = tid; i < r; i += NT) dpart += log(w.A[i + (size_t) i * r]);
    const double ldA = 2.0 * block_sum<NT>(dpart, red);
    return 0.5 * (trAS - ldA - logdetS - (double) r) - rho * ldX;
}

// gradient at the point of the last nfr_value call (uses the factor in w.A): fills g, w.Ai, w.Xi.
template <int D, int NT>
__device__ void nfr_gradient(const NfrWork &w, const double *x, int ne, const double *S, double rho, double *g) {
    const int tid = threadIdx.x, r = w.r, R = w.R;
    chol_inverse<NT>(w.A, r, r, w.Ai, r);
    __syncthreads();
    // tmp = JU * (diag(S) - A^-1)
    for(int t = tid; t < R * r; t += NT) {
        const int row = t % R, col = t / R;
        double s = w.JU[row + (size_t) col * R] * S[col];
        for(int p = 0; p < r; p++) s -= w.JU[row + (size_t) p * R] * w.Ai[p + (size_t) col * r];
        w.tmp[row + (size_t) col * R] = s;
    }
    for(int e = tid; e < ne; e += NT) { // X_e^-1 (logdet_function.cpp:384-389)
        double Xf[D * D], Xinv[D * D];
        const double *xe = x + (size_t) e * D * D;
        for(int j = 0; j < D; j++)
            for(int i = 0; i < D; i++) Xf[i + j * D] = (i >= j) ? xe[i + j * D] : xe[j + i * D];
        spd_inverse_small<D>(Xf, Xinv);
        for(int q2 = 0; q2 < D * D; q2++) w.Xi[(size_t) e * D * D + q2] = Xinv[q2];
    }
    __syncthreads();
    // g_e = 0.5 * sym(JU_e M JU_e^T) - rho X_e^-1
    for(int t = tid; t < ne * D * D; t += NT) {
        const int e = t / (D * D), qq = t % (D * D), i = qq % D, j = qq / D;
        double s1 = 0, s2 = 0;
        for(int p = 0; p < r; p++) {
            s1 += w.tmp[(e * D + i) + (size_t) p * R] * w.JU[(e * D + j) + (size_t) p * R];
            s2 += w.tmp[(e * D + j) + (size_t) p * R] * w.JU[(e * D + i) + (size_t) p * R];
        }
        g[t] = 0.5 * (0.5 * (s1 + s2)) - rho * w.Xi[t];
    }
    __syncthreads();
}

// Newton direction d = -LLT(H)^-1 g with the reference's Hessian (logdet_function.cpp:182-214, 398-427):
// H[(e;i,j),(f;u,v)] = P(f_u,e_i) P(e_j,f_v) + [e==f] rho Xinv(u,i) Xinv(j,v). Returns false if H is not PD.
template <int D, int NT>
__device__ bool nfr_newton_direction(const NfrWork &w, int ne, double rho, const double *g, double *d) {
    const int tid = threadIdx.x, r = w.r, R = w.R, q = w.q, ldq = w.ldq;
    // P = JU A^-1 JU^T, symmetrised
    for(int t = tid; t < R * r; t += NT) {
        const int row = t % R, col = t / R;
        double s = 0;
        for(int p = 0; p < r; p++) s += w.JU[row + (size_t) p * R] * w.Ai[p + (size_t) col * r];
        w.tmp[row + (size_t) col * R] = s;
    }
    __syncthreads();
    for(int t = tid; t < R * R; t += NT) {
        const int i = t % R, j = t / R;
        if(i <= j) {
            double s1 = 0, s2 = 0;
            for(int p = 0; p < r; p++) {
                s1 += w.tmp[i + (size_t) p * R] * w.JU[j + (size_t) p * R];
                s2 += w.tmp[j + (size_t) p * R] * w.JU[i + (size_t) p * R];
            }
            const double v = 0.5 * (s1 + s2);
            w.P[i + (size_t) j * R] = v;
            w.P[j + (size_t) i * R] = v;
        }
    }
    __syncthreads();
    constexpr int DD = D * D;
    for(int64_t t = tid; t < (int64_t) q * q; t += NT) {
        const int s = (int) (t % q), c = (int) (t / q);
        if(s < c) continue; // LLT reads the lower triangle only
        const int e = s / DD, ii = (s % DD) % D, jj = (s % DD) / D;
        const int f = c / DD, uu = (c % DD) % D, vv = (c % DD) / D;
        double v = w.P[(f * D + uu) + (size_t) (e * D + ii) * R] * w.P[(e * D + jj) + (size_t) (f * D + vv) * R];
        if(e == f) {
            const double *Xi = w.Xi + (size_t) e * DD;
            v += rho * Xi[uu + ii * D] * Xi[jj + vv * D];
        }
        w.H[s + (size_t) c * ldq] = v;
    }
    __syncthreads();
    const bool ok = chol_lower<NT>(w.H, q, ldq);
    __syncthreads();
    if(!ok) return false;
    // solve L L^T d = -g: column-oriented substitutions, one sync per column
    for(int i = tid; i < q; i += NT) d[i] = -g[i];
    __syncthreads();
    for(int j = 0; j < q; j++) {
        const double yj = d[j] / w.H[j + (size_t) j * ldq];
        __syncthreads();
        if(tid == 0) d[j] = yj;
        for(int i = j + 1 + tid; i < q; i += NT) d[i] -= w.H[i + (size_t) j * ldq] * yj;
        __syncthreads();
    }
    for(int j = q - 1; j >= 0; j--) {
        const double zj = d[j] / w.H[j + (size_t) j * ldq];
        __syncthreads();
        if(tid == 0) d[j] = zj;
        for(int i = tid; i < j; i += NT) d[i] -= w.H[j + (size_t) i * ldq] * zj;
        __syncthreads();
    }
    return true;
}

// The whole optimizeInformation loop. Inputs in shared memory: V (k x ldk eigenvectors), ord (kept
// eigen-columns), S (r), Jn (ne x [Ji Jj] D x 2D), tree (pairs). Result: x (q doubles in the workspace).
// flags out: bit0 line search failed at least once, bit1 final KLD is inf, bit2 Hessian not PD once.
template <int D, int NT>
__device__ void nfr_iterative(double *ws, int ne, int k, int r, const double *V, int ldk, const int *ord, const double *S,
                              const double *Jn, const int *tree, double *red, int *iflag, int &iters, int &flags, double &kld) {
    const int tid = threadIdx.x;
    constexpr int JW = D * 2 * D, DD = D * D;
    NfrWork w = nfr_carve(ws, ne, D, r);
    const int R = w.R, q = w.q;
    // JU_e = J_e * U[rows of (a,b), :]   (sparseJacobian() * _U; entries |J| < eps dropped, :335)
    for(int t = tid; t < R * r; t += NT) {
        const int row = t % R, col = t / R, e = row / D, i = row % D;
        const int a = tree[e] & 0xffff, b = tree[e] >> 16;
        const double *u = V + (size_t) ord[col] * ldk;
        const double *J = Jn + (size_t) e * JW;
        double s = 0;
#pragma unroll
        for(int j = 0; j < D; j++) {
            const double ja = J[i + j * D], jb = J[i + (D + j) * D];
            if(fabs(ja) >= 2.220446049250313e-16) s += ja * u[a * D + j];
            if(fabs(jb) >= 2.220446049250313e-16) s += jb * u[b * D + j];
        }
        w.JU[row + (size_t) col * R] = s;
    }
    for(int t = tid; t < q; t += NT) { // educatedGuess: identity blocks (:216-234)
        const int qq = t % DD;
        w.x[t] = (qq % D == qq / D) ? 1.0 : 0.0;
    }
    double lpart = 0;
    for(int i = tid; i < r; i += NT) lpart += log(S[i]);
    const double logdetS = block_sum<NT>(lpart, red);
    __syncthreads();
    iters = 0;
    flags = 0;
    const double stepRho = sqrt(10.0);
    double tol = 1e-4;
    for(double rho = 1.0; rho >= 5e-8; rho /= stepRho) {
        if(rho / stepRho < 5e-8) tol = 1e-12;
        // ---- PQNOptimizer::optimize ------------------------------------------------------------
        double f = nfr_value<D, NT>(w, w.x, ne, S, logdetS, rho, red, iflag);
        if(isinf(f)) {
            for(int t = tid; t < q; t += NT) w.g[t] = 0.0; // gradient() returns zeros when A is not PD
            __syncthreads();
        } else {
            nfr_gradient<D, NT>(w, w.x, ne, S, rho, w.g);
        }
        while(true) {
            if(!nfr_newton_direction<D, NT>(w, ne, rho, w.g, w.d)) { flags |= 4; break; }
            iters++;
            double p1 = 0, p2 = 0, p3 = 0;
            for(int t = tid; t < q; t += NT) { p1 += w.g[t] * w.d[t]; p2 += fabs(w.g[t]); p3 += fabs(w.d[t]); }
            const double gdotd = block_sum<NT>(p1, red);
            const double optcond = block_sum<NT>(p2, red);
            const double dsum = block_sum<NT>(p3, red);
            if(fabs(gdotd) < tol) break;
            const double f_old = f;
            // ---- LineSearchSimpleBacktracking::findStep ------------------------------------------
            double s_new = 1.0, f_new = f;
            bool failed = false;
            while(true) {
                for(int t = tid; t < q; t += NT) w.xn[t] = w.x[t] + s_new * w.d[t];
                __syncthreads();
                f_new = nfr_value<D, NT>(w, w.xn, ne, S, logdetS, rho, red, iflag);
                if(s_new < 1e-12) { failed = true; break; }
                if(isnan(f_new) || isinf(f_new) || f_new > f) s_new /= 2;
                else {
                    nfr_gradient<D, NT>(w, w.xn, ne, S, rho, w.gn);
                    break;
                }
            }
            if(failed) { flags |= 1; break; } // optimize() returns INFINITY, x unchanged; the caller ignores it
            for(int t = tid; t < q; t += NT) { w.x[t] = w.xn[t]; w.g[t] = w.gn[t]; }
            __syncthreads();
            f = f_new;
            if(optcond < tol) break;
            if(s_new * dsum < tol) break;
            if(fabs(f - f_old) < tol) break;
        }
    }
    kld = nfr_value<D, NT>(w, w.x, ne, S, logdetS, 0.0, red, iflag); // LogdetFunction::value(x) (optimizer.cpp:71)
    if(isinf(kld)) flags |= 2;
    __syncthreads();
}

} // namespace spg
