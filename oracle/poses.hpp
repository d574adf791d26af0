// ORACLE — TEST INFRASTRUCTURE ONLY (see dense.hpp header). PARITY UNPINNED.
// Restates the g2o SE2/SE3 arithmetic the reference's hot path calls. g2o is NOT in the reference
// tree (un-vendored, unpinned; cmake/FindG2O.cmake:1-6), so these follow g2o's published
// conventions and are pinned by finite-difference tests (tests/test_oracle_jacobians.py):
//   g2o::SE2 (compose/inverse/toVector, normalize_theta), VertexSE2::oplus (Euclidean add + wrap),
//   EdgeSE2ISAM error/Jacobians               -> reference src/se2_compatibility.h:26-51
//   VertexSE3::oplus  X <- X * fromVectorMQT(d), EdgeSE3 error toVectorMQT(Z^-1 Xi^-1 Xj),
//   computeEdgeSE3Gradient (analytic)         -> call sites src/se3_compatibility.h:25-29,
//                                                 src/glc_reparam_binary.hpp:50-64,97-116
#pragma once
#include "dense.hpp"

namespace orc {

inline double normalize_theta(double theta) {
    if(theta >= -M_PI && theta < M_PI) return theta;
    double multiplier = std::floor(theta / (2 * M_PI));
    theta = theta - multiplier * 2 * M_PI;
    if(theta >= M_PI) theta -= 2 * M_PI;
    if(theta < -M_PI) theta += 2 * M_PI;
    return theta;
}

// A pose of either group. SE2: (x, y, theta). SE3: rotation matrix R (row-major) + t, like
// Eigen::Isometry3d which g2o stores.
struct Pose {
    int dim = 3; // 3 = SE2, 6 = SE3
    double x = 0, y = 0, th = 0;
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double t[3] = {0, 0, 0};

    static Pose identity(int dim) {
        Pose p;
        p.dim = dim;
        return p;
    }
    static Pose se2(double x, double y, double th) {
        Pose p;
        p.dim = 3;
        p.x = x; p.y = y; p.th = normalize_theta(th);
        return p;
    }
};

// ---- quaternion helpers (Eigen conventions, coefficient order x y z w) ----
inline void quatToR(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
inline void RToQuat(const double R[9], double q[4]) {
    // Eigen::Quaterniond(Matrix3d)
    auto m = [&](int i, int j) { return R[3 * i + j]; };
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if(t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m(2, 1) - m(1, 2)) * t;
        q[1] = (m(0, 2) - m(2, 0)) * t;
        q[2] = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if(m(1, 1) > m(0, 0)) i = 1;
        if(m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (m(k, j) - m(j, k)) * t;
        q[j] = (m(j, i) + m(i, j)) * t;
        q[k] = (m(k, i) + m(i, k)) * t;
    }
}
// g2o::internal::normalize(Quaterniond&): unit norm and w >= 0
inline void quatNormalize(double q[4]) {
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for(int i = 0; i < 4; i++) q[i] /= n;
    if(q[3] < 0)
        for(int i = 0; i < 4; i++) q[i] = -q[i];
}

// g2o::internal::fromVectorQT (file form: tx ty tz qx qy qz qw)
inline Pose se3FromQT(const double v[7]) {
    Pose p;
    p.dim = 6;
    double q[4] = {v[3], v[4], v[5], v[6]};
    quatNormalize(q);
    quatToR(q, p.R);
    p.t[0] = v[0]; p.t[1] = v[1]; p.t[2] = v[2];
    return p;
}
inline void se3ToQT(const Pose &p, double v[7]) {
    double q[4];
    RToQuat(p.R, q);
    quatNormalize(q);
    v[0] = p.t[0]; v[1] = p.t[1]; v[2] = p.t[2];
    v[3] = q[0]; v[4] = q[1]; v[5] = q[2]; v[6] = q[3];
}
// g2o::internal::toVectorMQT: [t; vector part of the normalised (w>=0) quaternion]
inline void se3ToMQT(const Pose &p, double v[6]) {
    double q[4];
    RToQuat(p.R, q);
    quatNormalize(q);
    v[0] = p.t[0]; v[1] = p.t[1]; v[2] = p.t[2];
    v[3] = q[0]; v[4] = q[1]; v[5] = q[2];
}
// g2o::internal::fromVectorMQT
inline Pose se3FromMQT(const double v[6]) {
    Pose p;
    p.dim = 6;
    double w = 1 - (v[3] * v[3] + v[4] * v[4] + v[5] * v[5]);
    if(w < 0) {
        // fromCompactQuaternion returns identity
    } else {
        double q[4] = {v[3], v[4], v[5], std::sqrt(w)};
        quatToR(q, p.R);
    }
    p.t[0] = v[0]; p.t[1] = v[1]; p.t[2] = v[2];
    return p;
}

inline Pose compose(const Pose &a, const Pose &b) {
    Pose r;
    r.dim = a.dim;
    if(a.dim == 3) {
        double c = std::cos(a.th), s = std::sin(a.th);
        r.x = a.x + c * b.x - s * b.y;
        r.y = a.y + s * b.x + c * b.y;
        r.th = normalize_theta(a.th + b.th);
    } else {
        for(int i = 0; i < 3; i++) {
            for(int j = 0; j < 3; j++) {
                double s = 0;
                for(int k = 0; k < 3; k++) s += a.R[3 * i + k] * b.R[3 * k + j];
                r.R[3 * i + j] = s;
            }
            r.t[i] = a.R[3 * i] * b.t[0] + a.R[3 * i + 1] * b.t[1] + a.R[3 * i + 2] * b.t[2] + a.t[i];
        }
    }
    return r;
}
inline Pose inverse(const Pose &a) {
    Pose r;
    r.dim = a.dim;
    if(a.dim == 3) {
        r.th = normalize_theta(-a.th);
        double c = std::cos(r.th), s = std::sin(r.th);
        r.x = c * (-a.x) - s * (-a.y);
        r.y = s * (-a.x) + c * (-a.y);
    } else {
        for(int i = 0; i < 3; i++)
            for(int j = 0; j < 3; j++) r.R[3 * i + j] = a.R[3 * j + i];
        for(int i = 0; i < 3; i++)
            r.t[i] = -(r.R[3 * i] * a.t[0] + r.R[3 * i + 1] * a.t[1] + r.R[3 * i + 2] * a.t[2]);
    }
    return r;
}

// pose <-> the flat storage used in records / files: SE2 (x y theta), SE3 (t, qx qy qz qw)
inline int poseWords(int dim) { return dim == 3 ? 3 : 7; }
inline Pose poseFromFlat(int dim, const double *v) {
    if(dim == 3) return Pose::se2(v[0], v[1], v[2]);
    return se3FromQT(v);
}
inline void poseToFlat(const Pose &p, double *v) {
    if(p.dim == 3) {
        v[0] = p.x; v[1] = p.y; v[2] = p.th;
    } else {
        se3ToQT(p, v);
    }
}

// Vertex oplus: VertexSE2::oplusImpl (Euclidean add, angle wrapped), VertexSE3::oplusImpl
// (X <- X * fromVectorMQT(delta)).
inline Pose oplus(const Pose &x, const double *delta) {
    if(x.dim == 3) return Pose::se2(x.x + delta[0], x.y + delta[1], x.th + delta[2]);
    return compose(x, se3FromMQT(delta));
}

// Edge error. SE2: EdgeSE2ISAM::computeError (se2_compatibility.h:26-33). SE3: g2o::EdgeSE3
// (EdgeSE3ISAM inherits it under G2S_QUATERNIONS, se3_compatibility.h:30, CMakeLists.txt:18).
inline void edgeError(const Pose &Z, const Pose &Xi, const Pose &Xj, double *err) {
    if(Z.dim == 3) {
        Pose delta = compose(inverse(Xi), Xj);
        err[0] = delta.x - Z.x;
        err[1] = delta.y - Z.y;
        err[2] = normalize_theta(delta.th - Z.th);
    } else {
        Pose E = compose(inverse(Z), compose(inverse(Xi), Xj));
        se3ToMQT(E, err);
    }
}

static inline void skew(const double v[3], double S[9]) {
    S[0] = 0;     S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2];  S[4] = 0;     S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0];  S[8] = 0;
}

// Edge Jacobians d err / d delta_i, d err / d delta_j (d x d each, as Mat).
// SE2: explicit formulas of se2_compatibility.h:35-51 (independent of the measurement).
// SE3: analytic derivative of toVectorMQT(Z^-1 (Xi*D(di))^-1 (Xj*D(dj))) at 0 — the quantity
// g2o::internal::computeEdgeSE3Gradient evaluates (generated code in g2o; re-derived here):
//   A = Z^-1, B = Xi^-1 Xj, E = A B, q_E = sigma * (q_A (x) q_B) with sigma making w_E >= 0
//   Jj = [[R_E, 0], [0, w_E I + [v_E]x]]
//   Ji = [[-R_A, 2 R_A [t_B]x], [0, -sigma * M]],
//   M  = wA wB I + wB [vA]x - wA [vB]x - [vB]x [vA]x - vB vA^T
inline void edgeJacobians(const Pose &Z, const Pose &Xi, const Pose &Xj, Mat &Ji, Mat &Jj) {
    if(Z.dim == 3) {
        Ji = Mat(3, 3);
        Jj = Mat(3, 3);
        double thetai = Xi.th;
        double dtx = Xj.x - Xi.x, dty = Xj.y - Xi.y;
        double si = std::sin(thetai), ci = std::cos(thetai);
        Ji(0, 0) = -ci; Ji(0, 1) = -si; Ji(0, 2) = -si * dtx + ci * dty;
        Ji(1, 0) = si;  Ji(1, 1) = -ci; Ji(1, 2) = -ci * dtx - si * dty;
        Ji(2, 0) = 0;   Ji(2, 1) = 0;   Ji(2, 2) = -1;
        Jj(0, 0) = ci;  Jj(0, 1) = si;  Jj(0, 2) = 0;
        Jj(1, 0) = -si; Jj(1, 1) = ci;  Jj(1, 2) = 0;
        Jj(2, 0) = 0;   Jj(2, 1) = 0;   Jj(2, 2) = 1;
        return;
    }
    Ji = Mat(6, 6);
    Jj = Mat(6, 6);
    Pose A = inverse(Z);
    Pose B = compose(inverse(Xi), Xj);
    Pose E = compose(A, B);
    double qA[4], qB[4];
    RToQuat(A.R, qA);
    quatNormalize(qA);
    RToQuat(B.R, qB);
    quatNormalize(qB);
    // q_E = qA (x) qB (Hamilton), then normalised to w >= 0
    double vA[3] = {qA[0], qA[1], qA[2]}, wA = qA[3];
    double vB[3] = {qB[0], qB[1], qB[2]}, wB = qB[3];
    double wE = wA * wB - (vA[0] * vB[0] + vA[1] * vB[1] + vA[2] * vB[2]);
    double vE[3] = {wA * vB[0] + wB * vA[0] + (vA[1] * vB[2] - vA[2] * vB[1]),
                    wA * vB[1] + wB * vA[1] + (vA[2] * vB[0] - vA[0] * vB[2]),
                    wA * vB[2] + wB * vA[2] + (vA[0] * vB[1] - vA[1] * vB[0])};
    double nE = std::sqrt(wE * wE + vE[0] * vE[0] + vE[1] * vE[1] + vE[2] * vE[2]);
    double sigma = (wE < 0) ? -1.0 : 1.0;
    wE = sigma * wE / nE;
    for(int i = 0; i < 3; i++) vE[i] = sigma * vE[i] / nE;

    double SA[9], SB[9], SE[9], StB[9];
    skew(vA, SA);
    skew(vB, SB);
    skew(vE, SE);
    skew(B.t, StB);
    for(int r = 0; r < 3; r++)
        for(int c = 0; c < 3; c++) {
            // translation rows
            Ji(r, c) = -A.R[3 * r + c];
            double s = 0;
            for(int k = 0; k < 3; k++) s += A.R[3 * r + k] * StB[3 * k + c];
            Ji(r, 3 + c) = 2 * s;
            Jj(r, c) = E.R[3 * r + c];
            // rotation rows
            double sbsa = 0;
            for(int k = 0; k < 3; k++) sbsa += SB[3 * r + k] * SA[3 * k + c];
            double M = (r == c ? wA * wB : 0.0) + wB * SA[3 * r + c] - wA * SB[3 * r + c] - sbsa -
                       vB[r] * vA[c];
            Ji(3 + r, 3 + c) = -sigma * M;
            Jj(3 + r, 3 + c) = (r == c ? wE : 0.0) + SE[3 * r + c];
        }
}

} // namespace orc
