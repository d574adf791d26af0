// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product; never linked by libspg_b200.so.
// PARITY UNPINNED: the reference ships no golden vectors and cannot be built here (needs Eigen,
// g2o, iSAM, CHOLMOD). This file restates the Eigen dense primitives the reference's hot path
// calls (call sites cited per function) as plain C++ so the CPU restatement is self-contained.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

// Column-major dense matrix (Eigen::MatrixXd storage order, so informationData() layouts match,
// reference src/vertex_remover.cpp:531-533).
struct Mat {
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t) r_ * c_, 0.0) {}
    static Mat identity(int n) {
        Mat m(n, n);
        for(int i = 0; i < n; i++) m(i, i) = 1;
        return m;
    }
    double &operator()(int i, int j) { return a[(size_t) j * r + i]; }
    double operator()(int i, int j) const { return a[(size_t) j * r + i]; }
    int rows() const { return r; }
    int cols() const { return c; }
    Mat transpose() const {
        Mat t(c, r);
        for(int j = 0; j < c; j++)
            for(int i = 0; i < r; i++) t(j, i) = (*this)(i, j);
        return t;
    }
    Mat block(int i0, int j0, int nr, int nc) const {
        Mat b(nr, nc);
        for(int j = 0; j < nc; j++)
            for(int i = 0; i < nr; i++) b(i, j) = (*this)(i0 + i, j0 + j);
        return b;
    }
    void setBlock(int i0, int j0, const Mat &b) {
        for(int j = 0; j < b.c; j++)
            for(int i = 0; i < b.r; i++) (*this)(i0 + i, j0 + j) = b(i, j);
    }
    void addBlock(int i0, int j0, const Mat &b) {
        for(int j = 0; j < b.c; j++)
            for(int i = 0; i < b.r; i++) (*this)(i0 + i, j0 + j) += b(i, j);
    }
    double frob() const {
        double s = 0;
        for(double v : a) s += v * v;
        return std::sqrt(s);
    }
};

inline Mat operator*(const Mat &A, const Mat &B) {
    assert(A.c == B.r);
    Mat C(A.r, B.c);
    for(int j = 0; j < B.c; j++)
        for(int k = 0; k < A.c; k++) {
            double b = B(k, j);
            if(b == 0) continue;
            for(int i = 0; i < A.r; i++) C(i, j) += A(i, k) * b;
        }
    return C;
}
inline Mat operator+(const Mat &A, const Mat &B) {
    Mat C = A;
    for(size_t i = 0; i < C.a.size(); i++) C.a[i] += B.a[i];
    return C;
}
inline Mat operator-(const Mat &A, const Mat &B) {
    Mat C = A;
    for(size_t i = 0; i < C.a.size(); i++) C.a[i] -= B.a[i];
    return C;
}
inline Mat operator*(double s, const Mat &A) {
    Mat C = A;
    for(double &v : C.a) v *= s;
    return C;
}

// m.selfadjointView<Eigen::Upper>() -> full (pseudo_chow_liu.cpp:137, logdet_function.cpp:322)
inline Mat selfadjointUpper(const Mat &A) {
    Mat C = A;
    for(int j = 0; j < A.c; j++)
        for(int i = j + 1; i < A.r; i++) C(i, j) = A(j, i);
    return C;
}
// m.selfadjointView<Eigen::Lower>() -> full (logdet_function.cpp:96)
inline Mat selfadjointLower(const Mat &A) {
    Mat C = A;
    for(int j = 0; j < A.c; j++)
        for(int i = 0; i < j; i++) C(i, j) = A(j, i);
    return C;
}
// X.triangularView<StrictlyLower>() = X.triangularView<StrictlyUpper>().transpose()
// (vertex_remover.cpp:448-449, logdet_function.cpp:122)
inline void mirrorUpperToLower(Mat &A) {
    for(int j = 0; j < A.c; j++)
        for(int i = j + 1; i < A.r; i++) A(i, j) = A(j, i);
}
inline void mirrorLowerToUpper(Mat &A) {
    for(int j = 0; j < A.c; j++)
        for(int i = 0; i < j; i++) A(i, j) = A(j, i);
}

// selectVariables (utils.cpp:27-41)
inline Mat selectVariables(const Mat &o, const std::vector<int> &v1, const std::vector<int> &v2) {
    Mat ret((int) v1.size(), (int) v2.size());
    for(size_t i = 0; i < v1.size(); i++)
        for(size_t j = 0; j < v2.size(); j++) ret((int) i, (int) j) = o(v1[i], v2[j]);
    return ret;
}
inline Mat selectVariables(const Mat &o, const std::vector<int> &v) { return selectVariables(o, v, v); }

// Eigen::LLT<MatrixXd> (reads the lower triangle). Call sites: vertex_remover.cpp:444,
// pseudo_chow_liu.cpp:134,189, logdet_function.cpp:246,273, pqn/pqn_optimizer.cpp:52.
struct LLT {
    Mat L;
    bool ok = true;
    explicit LLT(const Mat &A) : L(A) {
        int n = A.r;
        for(int j = 0; j < n; j++) {
            double d = L(j, j);
            for(int k = 0; k < j; k++) d -= L(j, k) * L(j, k);
            if(!(d > 0)) {
                ok = false; // Eigen reports NumericalIssue and keeps going with garbage
                d = std::fabs(d) > 0 ? std::fabs(d) : 1.0;
            }
            double ljj = std::sqrt(d);
            L(j, j) = ljj;
            for(int i = j + 1; i < n; i++) {
                double s = L(i, j);
                for(int k = 0; k < j; k++) s -= L(i, k) * L(j, k);
                L(i, j) = s / ljj;
            }
        }
        for(int j = 0; j < n; j++)
            for(int i = 0; i < j; i++) L(i, j) = 0;
    }
    Mat solve(const Mat &B) const {
        int n = L.r;
        Mat X = B;
        for(int c = 0; c < B.c; c++) {
            for(int i = 0; i < n; i++) {
                double s = X(i, c);
                for(int k = 0; k < i; k++) s -= L(i, k) * X(k, c);
                X(i, c) = s / L(i, i);
            }
            for(int i = n - 1; i >= 0; i--) {
                double s = X(i, c);
                for(int k = i + 1; k < n; k++) s -= L(k, i) * X(k, c);
                X(i, c) = s / L(i, i);
            }
        }
        return X;
    }
    double logdet() const {
        double s = 0;
        for(int i = 0; i < L.r; i++) s += std::log(L(i, i));
        return 2 * s;
    }
};

// Eigen::LDLT<MatrixXd> with diagonal pivoting (reads the lower triangle). Only vectorD(),
// isPositive() and solve() are used by the reference: pseudo_chow_liu.cpp:178-182,
// logdet_function.cpp:123-131,137-138, 361-369, utils.cpp:74-81.
// The Eigen release is unpinned; the early-termination cutoff of some 3.2.x releases is not
// reproduced (it only matters for exactly rank-deficient inputs).
struct LDLT {
    Mat M;                 // unit-lower L below the diagonal, D on the diagonal (permuted)
    std::vector<int> tr;   // transpositions
    bool allPositive = true;
    explicit LDLT(const Mat &A) : M(A) {
        int n = A.r;
        tr.resize(n);
        // work on the lower triangle
        for(int k = 0; k < n; k++) {
            int p = k;
            double big = std::fabs(M(k, k));
            for(int i = k + 1; i < n; i++)
                if(std::fabs(M(i, i)) > big) { big = std::fabs(M(i, i)); p = i; }
            tr[k] = p;
            if(p != k) {
                // symmetric swap of rows/cols k and p in the lower triangle
                for(int j = 0; j < k; j++) std::swap(M(k, j), M(p, j));
                for(int i = p + 1; i < n; i++) std::swap(M(i, k), M(i, p));
                std::swap(M(k, k), M(p, p));
                for(int i = k + 1; i < p; i++) std::swap(M(i, k), M(p, i));
            }
            // M(k,k) -= sum_j L(k,j)^2 D(j); column below likewise (left-looking, as Eigen's unblocked kernel)
            double dk = M(k, k);
            for(int j = 0; j < k; j++) dk -= M(k, j) * M(k, j) * M(j, j);
            M(k, k) = dk;
            for(int i = k + 1; i < n; i++) {
                double s = M(i, k);
                for(int j = 0; j < k; j++) s -= M(i, j) * M(j, j) * M(k, j);
                M(i, k) = (dk != 0) ? s / dk : s;
            }
            if(!(dk > 0)) allPositive = false;
        }
    }
    std::vector<double> vectorD() const {
        std::vector<double> d(M.r);
        for(int i = 0; i < M.r; i++) d[i] = M(i, i);
        return d;
    }
    // `chol.isPositive() && (chol.vectorD().array() > 0).all()` collapses to "every pivot > 0"
    bool positive() const { return allPositive; }
    double sumLogD() const {
        double s = 0;
        for(int i = 0; i < M.r; i++) s += std::log(M(i, i));
        return s;
    }
    Mat solve(const Mat &B) const {
        int n = M.r;
        Mat X = B;
        for(int c = 0; c < B.c; c++) {
            for(int k = 0; k < n; k++)
                if(tr[k] != k) std::swap(X(k, c), X(tr[k], c));
            for(int i = 0; i < n; i++) {
                double s = X(i, c);
                for(int k = 0; k < i; k++) s -= M(i, k) * X(k, c);
                X(i, c) = s;
            }
            for(int i = 0; i < n; i++) X(i, c) /= M(i, i);
            for(int i = n - 1; i >= 0; i--) {
                double s = X(i, c);
                for(int k = i + 1; k < n; k++) s -= M(k, i) * X(k, c);
                X(i, c) = s;
            }
            for(int k = n - 1; k >= 0; k--)
                if(tr[k] != k) std::swap(X(k, c), X(tr[k], c));
        }
        return X;
    }
};

// Eigen::SelfAdjointEigenSolver<MatrixXd>: eigenvalues ascending, orthonormal eigenvectors in
// columns. Householder tridiagonalisation + implicit QL (tred2/tql2). Call sites:
// topology_provider_glc.cpp:45,66, logdet_function.cpp:19.
struct SymEig {
    std::vector<double> w;
    Mat V;
    bool ok = true;
    explicit SymEig(const Mat &A) {
        int n = A.r;
        V = A;
        // use the lower triangle like Eigen
        for(int j = 0; j < n; j++)
            for(int i = 0; i < j; i++) V(i, j) = V(j, i);
        w.assign(n, 0.0);
        std::vector<double> e(n, 0.0);
        if(n == 0) return;
        tred2(n, e);
        tql2(n, e);
    }

private:
    void tred2(int n, std::vector<double> &e) {
        std::vector<double> &d = w;
        for(int j = 0; j < n; j++) d[j] = V(n - 1, j);
        for(int i = n - 1; i > 0; i--) {
            double scale = 0.0, h = 0.0;
            for(int k = 0; k < i; k++) scale += std::fabs(d[k]);
            if(scale == 0.0) {
                e[i] = d[i - 1];
                for(int j = 0; j < i; j++) {
                    d[j] = V(i - 1, j);
                    V(i, j) = 0.0;
                    V(j, i) = 0.0;
                }
            } else {
                for(int k = 0; k < i; k++) {
                    d[k] /= scale;
                    h += d[k] * d[k];
                }
                double f = d[i - 1];
                double g = std::sqrt(h);
                if(f > 0) g = -g;
                e[i] = scale * g;
                h -= f * g;
                d[i - 1] = f - g;
                for(int j = 0; j < i; j++) e[j] = 0.0;
                for(int j = 0; j < i; j++) {
                    f = d[j];
                    V(j, i) = f;
                    g = e[j] + V(j, j) * f;
                    for(int k = j + 1; k <= i - 1; k++) {
                        g += V(k, j) * d[k];
                        e[k] += V(k, j) * f;
                    }
                    e[j] = g;
                }
                f = 0.0;
                for(int j = 0; j < i; j++) {
                    e[j] /= h;
                    f += e[j] * d[j];
                }
                double hh = f / (h + h);
                for(int j = 0; j < i; j++) e[j] -= hh * d[j];
                for(int j = 0; j < i; j++) {
                    f = d[j];
                    g = e[j];
                    for(int k = j; k <= i - 1; k++) V(k, j) -= (f * e[k] + g * d[k]);
                    d[j] = V(i - 1, j);
                    V(i, j) = 0.0;
                }
            }
            d[i] = h;
        }
        for(int i = 0; i < n - 1; i++) {
            V(n - 1, i) = V(i, i);
            V(i, i) = 1.0;
            double h = d[i + 1];
            if(h != 0.0) {
                for(int k = 0; k <= i; k++) d[k] = V(k, i + 1) / h;
                for(int j = 0; j <= i; j++) {
                    double g = 0.0;
                    for(int k = 0; k <= i; k++) g += V(k, i + 1) * V(k, j);
                    for(int k = 0; k <= i; k++) V(k, j) -= g * d[k];
                }
            }
            for(int k = 0; k <= i; k++) V(k, i + 1) = 0.0;
        }
        for(int j = 0; j < n; j++) {
            d[j] = V(n - 1, j);
            V(n - 1, j) = 0.0;
        }
        V(n - 1, n - 1) = 1.0;
        e[0] = 0.0;
    }
    void tql2(int n, std::vector<double> &e) {
        std::vector<double> &d = w;
        for(int i = 1; i < n; i++) e[i - 1] = e[i];
        e[n - 1] = 0.0;
        double f = 0.0, tst1 = 0.0;
        const double eps = std::numeric_limits<double>::epsilon();
        for(int l = 0; l < n; l++) {
            tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
            int m = l;
            while(m < n) {
                if(std::fabs(e[m]) <= eps * tst1) break;
                m++;
            }
            if(m >= n) m = n - 1;
            if(m > l) {
                int iter = 0;
                do {
                    if(++iter > 60 * n) { ok = false; break; }
                    double g = d[l];
                    double p = (d[l + 1] - g) / (2.0 * e[l]);
                    double r = std::hypot(p, 1.0);
                    if(p < 0) r = -r;
                    d[l] = e[l] / (p + r);
                    d[l + 1] = e[l] * (p + r);
                    double dl1 = d[l + 1];
                    double h = g - d[l];
                    for(int i = l + 2; i < n; i++) d[i] -= h;
                    f += h;
                    p = d[m];
                    double c = 1.0, c2 = c, c3 = c;
                    double el1 = e[l + 1];
                    double s = 0.0, s2 = 0.0;
                    for(int i = m - 1; i >= l; i--) {
                        c3 = c2;
                        c2 = c;
                        s2 = s;
                        g = c * e[i];
                        h = c * p;
                        r = std::hypot(p, e[i]);
                        e[i + 1] = s * r;
                        s = e[i] / r;
                        c = p / r;
                        p = c * d[i] - s * g;
                        d[i + 1] = h + s * (c * g + s * d[i]);
                        for(int k = 0; k < n; k++) {
                            h = V(k, i + 1);
                            V(k, i + 1) = s * V(k, i) + c * h;
                            V(k, i) = c * V(k, i) - s * h;
                        }
                    }
                    p = -s * s2 * c3 * el1 * e[l] / dl1;
                    e[l] = s * p;
                    d[l] = c * p;
                } while(std::fabs(e[l]) > eps * tst1);
            }
            d[l] = d[l] + f;
            e[l] = 0.0;
        }
        // ascending sort (Eigen sorts eigenvalues in increasing order)
        for(int i = 0; i < n - 1; i++) {
            int k = i;
            double p = d[i];
            for(int j = i + 1; j < n; j++)
                if(d[j] < p) { k = j; p = d[j]; }
            if(k != i) {
                d[k] = d[i];
                d[i] = p;
                for(int j = 0; j < n; j++) std::swap(V(j, i), V(j, k));
            }
        }
    }
};

// Eigen::PartialPivLU<MatrixXd>::solve(Identity) (topology_provider_glc.cpp:63-64)
inline Mat luInverse(const Mat &A) {
    int n = A.r;
    Mat LU = A;
    std::vector<int> piv(n);
    for(int k = 0; k < n; k++) {
        int p = k;
        double big = std::fabs(LU(k, k));
        for(int i = k + 1; i < n; i++)
            if(std::fabs(LU(i, k)) > big) { big = std::fabs(LU(i, k)); p = i; }
        piv[k] = p;
        if(p != k)
            for(int j = 0; j < n; j++) std::swap(LU(k, j), LU(p, j));
        double d = LU(k, k);
        for(int i = k + 1; i < n; i++) {
            LU(i, k) /= d;
            double l = LU(i, k);
            if(l != 0)
                for(int j = k + 1; j < n; j++) LU(i, j) -= l * LU(k, j);
        }
    }
    Mat X = Mat::identity(n);
    for(int c = 0; c < n; c++) {
        for(int k = 0; k < n; k++)
            if(piv[k] != k) std::swap(X(k, c), X(piv[k], c));
        for(int i = 0; i < n; i++) {
            double s = X(i, c);
            for(int k = 0; k < i; k++) s -= LU(i, k) * X(k, c);
            X(i, c) = s;
        }
        for(int i = n - 1; i >= 0; i--) {
            double s = X(i, c);
            for(int k = i + 1; k < n; k++) s -= LU(i, k) * X(k, c);
            X(i, c) = s / LU(i, i);
        }
    }
    return X;
}

} // namespace orc
