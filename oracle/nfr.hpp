// ORACLE — TEST INFRASTRUCTURE ONLY (see dense.hpp header). PARITY UNPINNED.
// Restatement of the NFR information fit: reference src/logdet_function.{h,cpp},
// src/optimizer.{h,cpp}, src/pqn/pqn_optimizer.cpp:29-128 (Newton branch, the only one
// optimizer.cpp:54 selects), src/pqn/line_search.cpp:12-36 (LineSearchSimpleBacktracking).
#pragma once
#include <list>
#include <memory>
#include <set>
#include <utility>
#include "dense.hpp"

namespace orc {

typedef std::pair<Mat, int> JacobianEntry;          // optimizer.h:17
typedef std::list<JacobianEntry> MeasurementJacobian;
typedef std::list<MeasurementJacobian> JacobianMapping;

typedef std::vector<double> Vec;

class LogdetFunction {
public:
    // logdet_function.cpp:14-64 (the #else branch: G2S_MORE_GENERIC_LESS_WORKING is not defined)
    LogdetFunction(const JacobianMapping &mapping, const Mat &target)
        : _mapping(mapping), _target(target), _jacsize(0) {
        SymEig eig(_target);
        eigOk = eig.ok;
        for(auto &jaclist : _mapping) _jacsize += jaclist.front().first.rows();

        static const double cutoff = 1e-5;
        int n = _target.rows();
        int smalleigs = 0;
        for(int i = 0; i < n; i++)
            if(eig.w[i] < cutoff) smalleigs++;
        int dim = _mapping.front().front().first.cols();
        smallEigs = smalleigs;

        if(smalleigs <= dim) {
            _S.assign(eig.w.begin() + dim, eig.w.end());
            for(double &s : _S) s = 1.0 / s;
            _U = eig.V.block(0, dim, n, n - dim);
        } else {
            Mat candidates = eig.V.block(0, 0, n, smalleigs);
            std::set<int> toDrop = chooseDimensions(candidates);
            _S.assign(n - dim, 0.0);
            _U = Mat(n, n - dim);
            for(int i = 0, j = 0; i < n; i++) {
                if(toDrop.count(i) == 0) {
                    _S[j] = std::min(std::fabs(1 / eig.w[i]), 1e6 / eig.w[n - 1]);
                    for(int r = 0; r < n; r++) _U(r, j) = eig.V(r, i);
                    j++;
                }
            }
        }
        _logdet = 0;
        for(double s : _S) _logdet += std::log(s);
    }
    virtual ~LogdetFunction() {}

    bool eigOk = true;
    int smallEigs = 0;

    // logdet_function.cpp:66-81
    std::set<int> chooseDimensions(const Mat &candidates) const {
        int dim = _mapping.front().front().first.cols();
        Mat JU = sparseJacobian() * candidates;
        std::vector<std::pair<double, int>> norms;
        for(int i = 0; i < JU.cols(); i++) {
            double s = 0;
            for(int r = 0; r < JU.rows(); r++) s += JU(r, i) * JU(r, i);
            norms.push_back(std::make_pair(std::sqrt(s), i));
        }
        std::sort(norms.begin(), norms.end());
        std::set<int> toDrop;
        for(int i = 0; i < dim; i++) toDrop.insert(norms[i].second);
        return toDrop;
    }

    bool hasClosedFormSolution() const { return _jacsize == (int) _S.size(); } // :83-86

    int xsize() const {
        int s = 0;
        for(auto &jaclist : _mapping) {
            int n = jaclist.front().first.rows();
            s += n * n;
        }
        return s;
    }

    // :88-99
    std::list<Mat> decondense(const Vec &x) const {
        std::list<Mat> ret;
        int k = 0;
        for(auto &jaclist : _mapping) {
            int n = jaclist.front().first.rows();
            Mat m(n, n);
            std::memcpy(m.a.data(), x.data() + k, sizeof(double) * n * n);
            k += n * n;
            ret.push_back(selfadjointLower(m));
        }
        return ret;
    }
    // :101-117
    Vec condense(const std::list<Mat> &X) const {
        Vec ret;
        for(const Mat &e : X) ret.insert(ret.end(), e.a.begin(), e.a.end());
        return ret;
    }

    // :119-133
    virtual double value(const Vec &x) {
        Mat UJXJU = _U.transpose() * informationProduct(x) * _U;
        mirrorUpperToLower(UJXJU);
        _chol.reset(new LDLT(UJXJU));
        if(_chol->positive()) {
            double tr = 0;
            for(int i = 0; i < UJXJU.rows(); i++) tr += UJXJU(i, i) * _S[i];
            return 0.5 * (tr - _chol->sumLogD() - _logdet - (double) _S.size());
        }
        return INFINITY;
    }

    // :135-180
    virtual void gradient(const Vec &x, Vec &g) {
        if(_chol && _chol->positive()) {
            int r = (int) _S.size();
            _xinv = _chol->solve(Mat::identity(r));
            Mat middle = -1.0 * _xinv;
            for(int i = 0; i < r; i++) middle(i, i) += _S[i];
            Mat Y = _U * middle * _U.transpose();
            g.assign(x.size(), 0.0);
            int k = 0;
            for(auto &jaclist : _mapping) {
                Mat block = blockSandwich(jaclist, Y);
                int n = block.rows();
                for(int i = 0; i < n * n; i++) g[k + i] = 0.5 * block.a[i];
                k += n * n;
            }
        } else {
            g.assign(x.size(), 0.0);
        }
    }

    // :182-214
    virtual void hessian(const Vec &x, Mat &H) {
        Mat JU = sparseJacobian() * _U;
        Mat P = JU * _xinv * JU.transpose();
        P = 0.5 * (P + P.transpose());
        int s = 0, k = 0;
        for(auto &jaclist : _mapping) {
            int n = jaclist.front().first.rows();
            for(int jj = 0; jj < n; jj++)
                for(int ii = 0; ii < n; ii++, s++) {
                    // singleVariableHessian(P, k + ii, k + jj, hij) -> H.row(s)
                    int i = k + ii, j = k + jj;
                    int kk = 0, t = 0;
                    for(auto &jl2 : _mapping) {
                        int n2 = jl2.front().first.rows();
                        for(int vv = 0; vv < n2; vv++)
                            for(int uu = 0; uu < n2; uu++) H(s, t++) = P(kk + uu, i) * P(j, kk + vv);
                        kk += n2;
                    }
                }
            k += n;
        }
    }

    // :216-234
    Vec educatedGuess() const {
        Vec ret(xsize(), 0.0);
        int k = 0;
        for(auto &jaclist : _mapping) {
            int n = jaclist.front().first.rows();
            for(int i = 0; i < n; i++) ret[k + i * n + i] = 1.0;
            k += n * n;
        }
        return ret;
    }

    // :236-279
    std::list<Mat> closedFormSolution(bool *okp = nullptr) const {
        std::list<Mat> ret;
        int n = _target.rows(), r = (int) _S.size();
        Mat US(n, r);
        for(int j = 0; j < r; j++)
            for(int i = 0; i < n; i++) US(i, j) = _U(i, j) * _S[j];
        Mat Sigma = US * _U.transpose();
        mirrorLowerToUpper(Sigma);
        bool ok = true;
        if(_mapping.size() == 1) {
            Mat J = sparseJacobian();
            Mat piece = Sigma * J.transpose();
            LLT chol(J * piece);
            ok = ok && chol.ok;
            ret.push_back(chol.solve(Mat::identity(J.rows())));
        } else {
            for(auto &jaclist : _mapping) {
                Mat block = blockSandwich(jaclist, Sigma);
                LLT chol(block);
                ok = ok && chol.ok;
                ret.push_back(chol.solve(Mat::identity(block.rows())));
            }
        }
        if(okp) *okp = ok;
        return ret;
    }

    Mat informationProduct(const Vec &x) const { return informationProduct(decondense(x)); }

    // :287-323
    Mat informationProduct(const std::list<Mat> &X) const {
        Mat JXJ(_target.cols(), _target.rows());
        if(X.size() == 1) {
            Mat J = sparseJacobian();
            Mat piece = J.transpose() * (*X.begin());
            JXJ = piece * J;
        } else {
            std::list<Mat>::const_iterator it = X.begin();
            for(auto &jaclist : _mapping) {
                for(auto s1 = jaclist.begin(); s1 != jaclist.end(); ++s1) {
                    for(auto s2 = s1; s2 != jaclist.end(); ++s2) {
                        if(s1->second <= s2->second) {
                            JXJ.addBlock(s1->second, s2->second, s1->first.transpose() * (*it) * s2->first);
                        } else {
                            JXJ.addBlock(s2->second, s1->second, s2->first.transpose() * (*it) * s1->first);
                        }
                    }
                }
                ++it;
            }
        }
        return selfadjointUpper(JXJ);
    }

    // :325-346 — dense matrix with the |J| < eps entries dropped (setFromTriplets sums duplicates)
    Mat sparseJacobian() const {
        Mat J(_jacsize, _target.cols());
        int k = 0;
        for(auto &jaclist : _mapping) {
            for(auto &Jpair : jaclist)
                for(int ii = 0; ii < Jpair.first.rows(); ii++)
                    for(int jj = 0; jj < Jpair.first.cols(); jj++)
                        if(std::fabs(Jpair.first(ii, jj)) >= std::numeric_limits<double>::epsilon())
                            J(k + ii, Jpair.second + jj) += Jpair.first(ii, jj);
            k += jaclist.front().first.rows();
        }
        return J;
    }

    const Vec &S() const { return _S; }
    const Mat &U() const { return _U; }

protected:
    // the sweep1/sweep2 double loop shared by gradient (:146-167) and closedFormSolution (:249-270)
    static Mat blockSandwich(const MeasurementJacobian &jaclist, const Mat &Y) {
        int n = jaclist.front().first.rows();
        Mat block(n, n);
        for(auto s1 = jaclist.begin(); s1 != jaclist.end(); ++s1) {
            int m = s1->first.cols();
            Mat thisBlock = s1->first * Y.block(s1->second, s1->second, m, m) * s1->first.transpose();
            block = block + 0.5 * (thisBlock + thisBlock.transpose());
            auto s2 = s1;
            for(++s2; s2 != jaclist.end(); ++s2) {
                int p = s2->first.cols();
                thisBlock = s1->first * Y.block(s1->second, s2->second, m, p) * s2->first.transpose();
                block = block + (thisBlock + thisBlock.transpose());
            }
        }
        return block;
    }

    const JacobianMapping &_mapping;
    const Mat &_target;
    double _logdet;
    int _jacsize;
    Vec _S;
    Mat _U, _xinv;
    std::unique_ptr<LDLT> _chol;
};

// logdet_function.cpp:348-427
class LogdetFunctionWithConstraints : public LogdetFunction {
public:
    LogdetFunctionWithConstraints(const JacobianMapping &mapping, const Mat &target)
        : LogdetFunction(mapping, target), _rho(0) {}
    void setRho(double rho) { _rho = rho; }

    double value(const Vec &x) override {
        double original = LogdetFunction::value(x);
        std::list<Mat> X = decondense(x);
        for(const Mat &Xblock : X) {
            LDLT chol(Xblock);
            if(chol.positive()) original -= _rho * chol.sumLogD();
            else return INFINITY;
        }
        return original;
    }
    void gradient(const Vec &x, Vec &g) override {
        LogdetFunction::gradient(x, g);
        std::list<Mat> X = decondense(x);
        int k = 0;
        _invXblocks.clear();
        if(_chol && _chol->positive()) {
            for(const Mat &Xblock : X) {
                LDLT chol(Xblock);
                if(chol.positive()) {
                    int n = Xblock.rows();
                    Mat inv = chol.solve(Mat::identity(n));
                    _invXblocks.push_back(inv);
                    for(int i = 0; i < n * n; i++) g[k + i] -= _rho * inv.a[i];
                    k += n * n;
                } else {
                    g.assign(g.size(), 0.0);
                    return;
                }
            }
        }
    }
    void hessian(const Vec &x, Mat &H) override {
        LogdetFunction::hessian(x, H);
        if(_chol && _chol->positive()) {
            int s = 0, t = 0;
            for(const Mat &inv : _invXblocks) {
                int n = inv.rows();
                for(int j = 0; j < n; j++)
                    for(int i = 0; i < n; i++, s++) {
                        int q = 0;
                        for(int v = 0; v < n; v++)
                            for(int u = 0; u < n; u++, q++) H(s, t + q) += _rho * inv(u, i) * inv(j, v);
                    }
                t += n * n;
            }
        }
    }

private:
    double _rho;
    std::list<Mat> _invXblocks;
};

struct NfrStats {
    int newtonIters = 0;   // Newton directions computed over all 15 barrier stages
    int funEvals = 0;
    bool lineSearchFailed = false;
    bool closedForm = false;
    bool kldInf = false;
    bool notPd = false;
    double kld = 0;
};

static inline bool invalid(double v) { return std::isnan(v) || std::isinf(v); }

// pqn/line_search.cpp:12-36
static inline int simpleBacktracking(LogdetFunctionWithConstraints &fun, const Vec &d, double s, double f,
                                     const Vec &x, double &s_new, double &f_new, Vec &g_new, Vec &x_new) {
    int funevals = 0;
    s_new = s;
    while(1) {
        for(size_t i = 0; i < x.size(); i++) x_new[i] = x[i] + s_new * d[i];
        double fold = fun.value(x);
        (void) fold;
        f_new = fun.value(x_new);
        if(s_new < 1e-12) return -1;
        funevals++;
        if(invalid(f_new) || f_new > f) {
            s_new /= 2;
        } else {
            fun.gradient(x_new, g_new);
            return funevals;
        }
    }
}

// pqn/pqn_optimizer.cpp:29-128 with useHessian = true, verbose = false, maxIters = 0
static inline double pqnOptimize(LogdetFunctionWithConstraints &fun, Vec &x, double tol, NfrStats &st) {
    int n = (int) x.size();
    Vec g(n), g_old, x_old;
    double f = fun.value(x), f_old = 0;
    fun.gradient(x, g);
    while(true) {
        double step;
        Mat H(n, n);
        fun.hessian(x, H);
        LLT chol(H);
        if(!chol.ok) st.notPd = true;
        Mat rhs(n, 1);
        for(int i = 0; i < n; i++) rhs(i, 0) = -g[i];
        Mat dm = chol.solve(rhs);
        Vec d(dm.a);
        st.newtonIters++;

        double gdotd = 0;
        for(int i = 0; i < n; i++) gdotd += g[i] * d[i];
        if(std::fabs(gdotd) < tol) return f;

        step = 1;
        f_old = f;
        g_old = g;
        x_old = x;

        double f_new = f;
        Vec x_new(n), g_new(n);
        int ret = simpleBacktracking(fun, d, step, f, x, step, f_new, g_new, x_new);
        if(ret < 0) {
            st.lineSearchFailed = true;
            return INFINITY;
        }
        st.funEvals += ret;

        double optcond = 0;
        for(int i = 0; i < n; i++) optcond += std::fabs(g[i]); // previous gradient ("TODO: Check")
        x = x_new;
        f = f_new;
        g = g_new;
        if(optcond < tol) return f;
        double dsum = 0;
        for(int i = 0; i < n; i++) dsum += std::fabs(d[i]);
        if(step * dsum < tol) return f;
        if(std::fabs(f - f_old) < tol) return f;
    }
}

// optimizer.cpp:16-81
static inline std::list<Mat> optimizeInformation(const JacobianMapping &mapping, const Mat &target,
                                                 NfrStats &st) {
    LogdetFunctionWithConstraints fun(mapping, target);
    if(fun.hasClosedFormSolution()) {
        st.closedForm = true;
        bool ok = true;
        std::list<Mat> sol = fun.closedFormSolution(&ok);
        st.notPd = !ok;
        st.kld = fun.LogdetFunction::value(fun.condense(sol));
        return sol;
    }
    Vec x = fun.educatedGuess();
    const double startRho = 1, endRho = 5e-8, stepRho = std::sqrt(10);
    double tol = 1e-4;
    for(double rho = startRho; rho >= endRho; rho /= stepRho) {
        fun.setRho(rho);
        if(rho / stepRho < endRho) tol = 1e-12;
        pqnOptimize(fun, x, tol, st);
    }
    double final = fun.LogdetFunction::value(x);
    st.kld = final;
    if(std::isinf(final)) st.kldInf = true; // the reference calls exit(0) here
    return fun.decondense(x);
}

} // namespace orc
