// ORACLE — TEST INFRASTRUCTURE ONLY (see dense.hpp header). PARITY UNPINNED.
// Restatement of PseudoChowLiu, reference src/pseudo_chow_liu.{h,cpp}.
#pragma once
#include <list>
#include <queue>
#include <set>
#include "dense.hpp"

namespace orc {

// reference src/sparsity_options.h:11-30
struct SparsityOptions {
    enum SparsityTopology { Tree, Subgraph, CliqueySubgraph, Dense, CliqueyDense };
    enum LinearizationPoint { Local, Global };
    SparsityTopology topology = Tree;
    double chordRatio = 1;
    LinearizationPoint linPoint = Local;
    bool includeIntraClique = true;
};

// Test hook (tie-break tests): when set, fillEdges() takes the mutual-information weights from here (pairs in
// (i<j) lexicographic order) instead of computing them, so that the CUDA path and the oracle can be fed the SAME
// weights with exact ties and compared on the pop order of std::priority_queue alone.
static thread_local const double *g_overrideWeights = nullptr;

// Vertices are the blanket's kept vertices 0..n-1 (all of dimension d).
class PseudoChowLiu {
public:
    typedef std::list<std::pair<int, int>> CorrelatedSkeletonTree;
    typedef std::list<CorrelatedSkeletonTree> SparsityPattern;

    PseudoChowLiu(const SparsityOptions &opts, const Mat &info, int n, int d)
        : _opts(opts), _information(info), _n(n), _d(d) {}

    bool ok = true;                      // LLT(Lambda + I) succeeded
    std::vector<double> weights;         // MI of every pair in (i<j) lexicographic order (debug)

    // pseudo_chow_liu.cpp:33-87
    void computeSparsityPattern() {
        _pattern.clear();
        int n = _n;
        int m = int((1 + _opts.chordRatio) * (n - 1));
        bool full = (m >= n * (n - 1) / 2);

        if(n == 2) {
            _pattern.push_back(CorrelatedSkeletonTree(1, std::make_pair(0, 1)));
        } else if(_opts.topology == SparsityOptions::Dense ||
                  (_opts.topology == SparsityOptions::Subgraph && full)) {
            for(int i = 0; i < n - 1; i++)
                for(int j = i + 1; j < n; j++)
                    _pattern.push_back(CorrelatedSkeletonTree(1, std::make_pair(i, j)));
        } else {
            fillEdges();
            doKruskal();
            if(_opts.topology == SparsityOptions::Tree || _opts.topology == SparsityOptions::Subgraph) {
                int nedges = (_opts.topology == SparsityOptions::Tree) ? n - 1 : m;
                for(int i = 0; i < nedges; i++)
                    _pattern.push_back(CorrelatedSkeletonTree(
                            1, std::make_pair(_edgeBin[i].vert1, _edgeBin[i].vert2)));
            } else if(_opts.topology == SparsityOptions::CliqueyDense ||
                      (_opts.topology == SparsityOptions::CliqueySubgraph && full)) {
                CorrelatedSkeletonTree tree;
                for(int i = 0; i < n - 1; i++)
                    tree.push_back(std::make_pair(_edgeBin[i].vert1, _edgeBin[i].vert2));
                _pattern.push_back(tree);
            } else {
                fillCliques();
            }
        }
    }
    const SparsityPattern &getSparsityPattern() const { return _pattern; }

    // pseudo_chow_liu.cpp:130-167
    Mat marginal(int vert, bool *okp = nullptr) const {
        return marginalKeep(range(vert * _d, vert * _d + _d), okp);
    }
    Mat jointMarginal(int v1, int v2, bool *okp = nullptr) const {
        return marginalKeep(join(range(v1 * _d, v1 * _d + _d), range(v2 * _d, v2 * _d + _d)), okp);
    }

private:
    struct WeightedEdge {
        double weight;
        int vert1, vert2;
        bool operator<(const WeightedEdge &e) const { return weight < e.weight; }
    };

    static std::vector<int> range(int min, int lessthan) {
        std::vector<int> ret;
        for(; min < lessthan; min++) ret.push_back(min);
        return ret;
    }
    static std::vector<int> join(const std::vector<int> &a, const std::vector<int> &b) {
        std::vector<int> ret(a);
        ret.insert(ret.end(), b.begin(), b.end());
        return ret;
    }
    // assumes sorted a (pseudo_chow_liu.cpp:106-117)
    static std::vector<int> complement(const std::vector<int> &a, int bound) {
        std::vector<int> ret;
        for(int i = 0, j = 0; i < bound; i++) {
            if(j < int(a.size()) && a[j] == i) j++;
            else ret.push_back(i);
        }
        return ret;
    }
    Mat marginalKeep(const std::vector<int> &keep, bool *okp) const {
        std::vector<int> marginalize = complement(keep, _information.rows());
        Mat kk = selectVariables(_information, keep);
        if(marginalize.empty()) {
            if(okp) *okp = true;
            return selfadjointUpper(kk);
        }
        Mat mixed = selectVariables(_information, keep, marginalize);
        LLT chol(selectVariables(_information, marginalize));
        if(okp) *okp = chol.ok;
        Mat schur = kk - mixed * chol.solve(mixed.transpose());
        return selfadjointUpper(schur);
    }

    // pseudo_chow_liu.cpp:169-183
    double weight(int v1, int v2) {
        const int d = _d;
        Mat jointCov = selectVariables(_pseudoCovariance, join(range(v1 * d, v1 * d + d), range(v2 * d, v2 * d + d)));
        LDLT x(_pseudoCovariance.block(v1 * d, v1 * d, d, d));
        LDLT y(_pseudoCovariance.block(v2 * d, v2 * d, d, d));
        LDLT xy(jointCov);
        return x.sumLogD() + y.sumLogD() - xy.sumLogD();
    }

    // pseudo_chow_liu.cpp:185-196
    void fillEdges() {
        static const double tikhonov_eps = 1;
        _edges = std::priority_queue<WeightedEdge>();
        Mat reg = _information;
        for(int i = 0; i < reg.r; i++) reg(i, i) += tikhonov_eps;
        LLT llt(reg);
        ok = llt.ok;
        _pseudoCovariance = llt.solve(Mat::identity(reg.r));
        weights.clear();
        for(int i = 0; i < _n - 1; i++)
            for(int j = i + 1; j < _n; j++) {
                double w = weight(i, j);
                if(g_overrideWeights) w = g_overrideWeights[weights.size()];
                weights.push_back(w);
                _edges.push({w, i, j});
            }
    }

    // pseudo_chow_liu.cpp:253-289
    void doKruskal() {
        std::vector<std::set<int>> connectivity;
        std::vector<WeightedEdge> rejectBin, acceptBin;
        for(int i = 0; i < _n; i++) {
            connectivity.push_back(std::set<int>());
            connectivity.back().insert(i);
        }
        while(_edges.size() > 0) {
            int set1 = 0, set2 = 0;
            WeightedEdge e = _edges.top();
            _edges.pop();
            for(int i = 0; i < (int) connectivity.size(); i++) {
                if(connectivity[i].count(e.vert1) > 0) set1 = i;
                if(connectivity[i].count(e.vert2) > 0) set2 = i;
            }
            if(set1 != set2) {
                acceptBin.push_back(e);
                connectivity[set1].insert(connectivity[set2].begin(), connectivity[set2].end());
                connectivity.erase(connectivity.begin() + set2);
            } else {
                rejectBin.push_back(e);
            }
        }
        _edgeBin = acceptBin;
        _edgeBin.insert(_edgeBin.end(), rejectBin.begin(), rejectBin.end());
    }

    static size_t intersectionSize(const std::set<int> &a, const std::set<int> &b) {
        size_t c = 0;
        for(int x : a) c += b.count(x);
        return c;
    }

    // pseudo_chow_liu.cpp:198-251
    void fillCliques() {
        int n = _n;
        int m = int((1 + _opts.chordRatio) * (n - 1));
        std::vector<std::set<int>> cliques;
        for(int i = 0; i < n - 1; i++) cliques.push_back({_edgeBin[i].vert1, _edgeBin[i].vert2});

        bool joined = true;
        for(int nedges = n - 1, maxfill = 1; nedges < m && joined; maxfill++) {
            joined = false;
            int minfill = std::numeric_limits<int>::max();
            for(int i = 0; i < (int) cliques.size(); i++) {
                for(int j = i + 1; j < (int) cliques.size(); j++) {
                    if(intersectionSize(cliques[i], cliques[j]) > 0) {
                        int thisfill = (int(cliques[i].size()) - 1) * (int(cliques[j].size()) - 1);
                        minfill = std::min(thisfill, minfill);
                        if(thisfill <= maxfill && nedges + thisfill <= m) {
                            cliques[i].insert(cliques[j].begin(), cliques[j].end());
                            nedges += thisfill;
                            cliques.erase(cliques.begin() + j);
                            joined = true;
                            j--;
                        }
                    }
                }
            }
            if(!joined && minfill > maxfill) {
                joined = true;
                maxfill = minfill - 1;
            }
        }
        std::vector<CorrelatedSkeletonTree> pat(cliques.size());
        for(int i = 0; i < n - 1; i++)
            for(int j = 0; j < (int) cliques.size(); j++)
                if(cliques[j].count(_edgeBin[i].vert1) > 0 && cliques[j].count(_edgeBin[i].vert2))
                    pat[j].push_back(std::make_pair(_edgeBin[i].vert1, _edgeBin[i].vert2));
        for(auto &t : pat) _pattern.push_back(t);
    }

    SparsityOptions _opts;
    const Mat &_information;
    int _n, _d;
    Mat _pseudoCovariance;
    std::priority_queue<WeightedEdge> _edges;
    std::vector<WeightedEdge> _edgeBin;
    SparsityPattern _pattern;
};

} // namespace orc
