// spg_host.h — host-side mirror of the reference's plugin surface for the node-removal path,
// over a g2o-free pose-graph container. Names follow the reference so the adapters read like it:
//   SparsityOptions        reference src/sparsity_options.h:11-30
//   DecimateOptions / *Decimate   src/decimation.{h,cpp}
//   TopologyProvider{,GLC,SE2,SE3}  src/topology_provider*.h — here a provider is a *descriptor*: it
//        decides applicability per blanket on the host and selects the device code path
//        (algorithm + dimension); the numerical body of topology()/optimizeInformation() runs in the
//        fused CUDA kernel (spg_kernels.cuh).
//   VertexRemover          src/vertex_remover.{h,cpp} — same public calls; remove() groups the
//        removal list into wavefront rounds of non-interfering blankets (result identical to the
//        reference's one-at-a-time loop) and sends each round through spg_remove_round().
//   computeSubstituteEdge  src/compute_substitute_edge.cpp:13-96
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>
#include <initializer_list>
#include <list>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <memory>
#include <vector>

#include "../../include/spg_capi.h"

namespace spg {

struct SparsityOptions {
    enum SparsityTopology { Tree, Subgraph, CliqueySubgraph, Dense, CliqueyDense };
    enum LinearizationPoint { Local, Global };
    SparsityTopology topology = Tree;
    double chordRatio = 1;
    LinearizationPoint linPoint = Local;
    bool includeIntraClique = true;
    int flags = 0;
};

struct DecimateOptions {
    int sparsity;
    int clusterSize;
};
std::vector<int> clusterDecimate(int last, int endvert, const DecimateOptions &opts);
std::vector<int> onlineDecimate(int last, int endvert, const DecimateOptions &opts);
std::vector<int> globalDecimate(int lastid, int endvert, const DecimateOptions &opts);
typedef std::vector<int> (*DecimateFunction)(int last, int endvert, const DecimateOptions &opts);

// ---- pose-graph container ----------------------------------------------------------------------
// vertex-id list of an edge: two ids inline (POSE edges never touch the heap), more on the heap
class IdList {
public:
    IdList() {}
    IdList(std::initializer_list<int> il) { for(int x : il) push_back(x); }
    size_t size() const { return _n; }
    bool empty() const { return _n == 0; }
    void clear() { _n = 0; _more.clear(); }
    void push_back(int x) {
        if(_n < 2) _in[_n] = x;
        else {
            if(_n == 2) { _more.assign(_in, _in + 2); }
            _more.push_back(x);
        }
        _n++;
    }
    const int *begin() const { return _n <= 2 ? _in : _more.data(); }
    const int *end() const { return begin() + _n; }
    int operator[](size_t i) const { return begin()[i]; }
private:
    int _in[2] = {0, 0};
    unsigned _n = 0;
    std::vector<int> _more;
};

struct GraphEdge {
    int kind = SPG_EDGE_POSE;
    IdList v;                    // vertex ids, in the edge's own order (from, to, ...)
    int rows = 0;                // error dimension
    // one payload block per edge: the measurement (POSE: P; GLC: d*nv; MULTI: nmeas*P doubles) followed by the
    // information (POSE/MULTI: rows*rows column-major; GLC: W rows x d*nv row-major). Freed when the edge dies.
    std::vector<double> payload;
    int nMeas = 0;
    double *meas() { return payload.data(); }
    const double *meas() const { return payload.data(); }
    double *info() { return payload.data() + nMeas; }
    const double *info() const { return payload.data() + nMeas; }
    int nInfo() const { return (int) payload.size() - nMeas; }
    void allocPayload(int nm, int ni) { nMeas = nm; payload.resize((size_t) nm + ni); }
    void setPayload(const double *m, int nm, const double *inf, int ni) {
        nMeas = nm;
        payload.resize((size_t) nm + ni);
        std::memcpy(payload.data(), m, sizeof(double) * nm);
        std::memcpy(payload.data() + nm, inf, sizeof(double) * ni);
    }
    std::vector<int> pairs;      // MULTI: 2*nmeas indices into v
    int uidMajor = -1, uidMinor = 0; // canonical order: file edges (-1, file index); new edges (list index, k)
    bool alive = true;
    // Graph::verts indices of the vertices (set by Graph::addEdge): the first two inline, all of them in vxn when the
    // edge has more than two (the scheduler walks adjacency by index, without id lookups or heap hops for POSE edges)
    int vx0 = -1, vx1 = -1;
    std::vector<int> vxn;
    int nv() const { return (int) v.size(); }
    int vx(int q) const { return v.size() <= 2 ? (q == 0 ? vx0 : vx1) : vxn[q]; }
    unsigned long long uidKey() const { return ((unsigned long long) (unsigned) (uidMajor + 1) << 32) | (unsigned) uidMinor; }
};
// edge store: fixed-size blocks, so growing never moves an edge (a std::vector of ~150-byte edges re-touches gigabytes on
// every doubling of a million-pose graph) and references stay valid while other threads append. The table of block
// pointers has a fixed capacity too, so reserve() may run on a helper thread (the remover grows the store for the
// substitute edges of a round while the GPU works on it) next to threads that read existing edges.
template <class T, int LOG2 = 14>
class ChunkedVector {
public:
    ChunkedVector() : _blocks(new T *[MAXB]()) {}
    ChunkedVector(const ChunkedVector &) = delete;
    ChunkedVector &operator=(const ChunkedVector &) = delete;
    ~ChunkedVector() { for(size_t b = 0; b < _nb; b++) delete[] _blocks[b]; }
    size_t size() const { return _n; }
    T &operator[](size_t i) { return _blocks[i >> LOG2][i & MASK]; }
    const T &operator[](size_t i) const { return _blocks[i >> LOG2][i & MASK]; }
    T &back() { return (*this)[_n - 1]; }
    void reserve(size_t n) { reserve(n, [](T &) {}); }
    // blocks for n elements (size() unchanged); every element of a NEW block is default-constructed, then init(element)
    template <class F>
    void reserve(size_t n, F init) {
        while((_nb << LOG2) < n && _nb < MAXB) {
            T *b = new T[(size_t) 1 << LOG2];
            for(size_t i = 0; i < ((size_t) 1 << LOG2); i++) init(b[i]);
            _blocks[_nb] = b;
            _nb++;
        }
    }
    size_t blockEnd(size_t i) const { return ((i >> LOG2) + 1) << LOG2; } // first index behind the block of element i
    void resize(size_t n) { // grow only
        reserve(n);
        if(n > _n) _n = n;
    }
    void push_back(T &&x) {
        resize(_n + 1);
        (*this)[_n - 1] = std::move(x);
    }
private:
    static constexpr size_t MASK = ((size_t) 1 << LOG2) - 1;
    static constexpr size_t MAXB = (size_t) 1 << 17; // 2^31 edges: the edge indices are ints
    std::unique_ptr<T *[]> _blocks;
    size_t _nb = 0, _n = 0;
};

struct GraphVertex {
    int id = 0;
    double pose[7] = {0, 0, 0, 0, 0, 0, 1};
    std::vector<int> edges;      // indices into Graph::edges (alive ones only)
    // same positions as `edges`: for a two-vertex edge (other endpoint's verts index << 1) | (this vertex is the
    // edge's first one), -1 for any other edge — blanket extraction then walks binary edges without touching them
    std::vector<int> peer;
    bool alive = true;
};

class Graph {
public:
    explicit Graph(int dim_) : dim(dim_) {
        static std::atomic<unsigned long long> next(1);
        serial = next++;
    }
    int dim;
    unsigned long long serial = 0, version = 0; // identity and mutation count (caches of derived orders key on these)
    int poseWords() const { return dim == 3 ? 3 : 7; }
    std::vector<GraphVertex> verts;
    ChunkedVector<GraphEdge> edges;
    // id -> verts index: a dense table for ids in [0, 2^26), a hash map for anything else
    std::vector<int> dense;
    std::unordered_map<int, int> sparse;
    int indexOf(int id) const {
        if((unsigned) id < dense.size()) return dense[(unsigned) id];
        if((unsigned) id < (1u << 26)) return -1;
        auto it = sparse.find(id);
        return it == sparse.end() ? -1 : it->second;
    }
    int fileEdges = 0;
    int aliveVertices = 0, aliveEdges = 0;

    static Graph *loadG2o(const std::string &path, std::string *err);
    bool saveG2o(const std::string &path, std::string *err) const;
    bool hasVertex(int id) const;
    GraphVertex *vertex(int id);
    const GraphVertex *vertex(int id) const;
    bool addVertex(int id, const double *pose);
    int addPoseEdge(int from, int to, const double *meas, const double *info);
    int addEdge(GraphEdge &&e);      // generic (uid must be set)
    int addEdge(const GraphEdge &e) { return addEdge(GraphEdge(e)); }
    void removeEdge(int ei);
    void removeVertex(int id);
    int maxVertexId() const;
    std::vector<int> vertexIds() const;          // ascending
    std::vector<int> edgeOrder() const;          // alive edges in canonical order
};

// ---- topology providers (descriptors) --------------------------------------------------------------
class TopologyProvider {
public:
    virtual ~TopologyProvider() {}
    // src/topology_provider_base.h:20-24 — every vertex/edge kind of the blanket must be supported
    virtual bool applicable(int dim, const std::set<int> &edgeKinds) const = 0;
    virtual int algorithm() const = 0;                 // SPG_ALG_*
    virtual bool requiresOptimization() const = 0;     // NFR: true, GLC: false
    virtual void setSparsityOptions(const SparsityOptions &o) { _opts = o; }
protected:
    SparsityOptions _opts;
};
class TopologyProviderGLC : public TopologyProvider { // src/topology_provider_glc.{h,cpp}
public:
    bool applicable(int dim, const std::set<int> &) const override { return dim == 3 || dim == 6; } // checks vertices only (:26-30)
    int algorithm() const override { return SPG_ALG_GLC; }
    bool requiresOptimization() const override { return false; }
};
template <int DIM>
class TopologyProviderBinary : public TopologyProvider { // src/topology_provider_binary.{h,hpp}
public:
    bool applicable(int dim, const std::set<int> &kinds) const override {
        if(dim != DIM) return false;
        for(int k : kinds)
            if(k != SPG_EDGE_POSE && k != SPG_EDGE_MULTI) return false; // E and MultiEdgeCorrelated<E> (:15-19)
        return true;
    }
    int algorithm() const override { return SPG_ALG_NFR; }
    bool requiresOptimization() const override { return true; }
};
typedef TopologyProviderBinary<3> TopologyProviderSE2ISAM; // the graphs built here carry the *ISAM edge types
typedef TopologyProviderBinary<6> TopologyProviderSE3ISAM; // (graph_wrapper_g2o.cpp:120-146)
typedef TopologyProviderBinary<3> TopologyProviderSE2;
typedef TopologyProviderBinary<6> TopologyProviderSE3;

// small int list with N entries inline: the scheduler's per-unit vertex lists are read every round for every unit of the
// window, and a heap hop per list is a cache miss per unit
template <int N>
class InlineInts {
public:
    size_t size() const { return _n; }
    bool empty() const { return _n == 0; }
    void clear() { _n = 0; _more.clear(); }
    void push_back(int x) {
        if(_n < (unsigned) N) _in[_n] = x;
        else {
            if(_n == (unsigned) N) _more.assign(_in, _in + N);
            _more.push_back(x);
        }
        _n++;
    }
    const int *begin() const { return _n <= (unsigned) N ? _in : _more.data(); }
    const int *end() const { return begin() + _n; }
    int operator[](size_t i) const { return begin()[i]; }
private:
    int _in[N];
    unsigned _n = 0;
    std::vector<int> _more;
};

// ---- VertexRemover ---------------------------------------------------------------------------------
struct RemovalUnit {
    int listIndex = 0;               // index of the root in the removal list
    std::vector<int> removed;        // ascending id (toRemoveNow)
    std::vector<int> kept;           // ascending id
    std::vector<int> edges;          // blanket edges, canonical order
    InlineInts<2> ridx;              // Graph::verts indices of removed / kept (scheduler: no id lookups per round)
    InlineInts<20> kidx;
};

// the cached blankets of a removal list: a million units with three heap vectors each are constructed (first touch of
// ~0.2 GB) and destroyed (millions of frees) by a few threads instead of one
class UnitCache {
public:
    UnitCache() {}
    UnitCache(const UnitCache &) = delete;
    UnitCache &operator=(const UnitCache &) = delete;
    ~UnitCache() { reset(0); }
    void reset(size_t n); // destroys the old units, default-constructs n new ones
    RemovalUnit &operator[](size_t i) { return _p[i]; }
    const RemovalUnit &operator[](size_t i) const { return _p[i]; }
    size_t size() const { return _n; }
private:
    RemovalUnit *_p = nullptr;
    size_t _n = 0;
};

// growable host buffer of 8-byte words: page-locked when a CUDA device is present, plain memory otherwise
struct HostBuf {
    uint64_t *p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    HostBuf() {}
    HostBuf(const HostBuf &) = delete;
    HostBuf &operator=(const HostBuf &) = delete;
    ~HostBuf() { release(); }
    uint64_t *reserve(size_t words); // contents are NOT preserved on growth
    void release();
};

class VertexRemover {
public:
    VertexRemover();
    ~VertexRemover();
    void setSparsityOptions(const SparsityOptions &opts);
    void setGraph(Graph *graph) { _graph = graph; }
    void setContext(spg_ctx *ctx) { _ctx = ctx; }
    void registerTopologyProvider(TopologyProvider *topology) { _topologies.push_back(topology); } // takes ownership

    // Returns the indices (into Graph::edges) of the edges added, in removal order then provider order.
    // status: SPG_OK or the first error (graph is left consistent up to the failing round).
    std::vector<int> remove(int toRemove, spg_status *status = nullptr);
    std::vector<int> remove(const std::vector<int> &toRemove, spg_status *status = nullptr);

    std::vector<int> markovBlanketVertices(int root) const;
    std::vector<int> extendedMarkovBlanketVertices(int root, const std::set<int> &pickBin, std::vector<int> &picked) const;
    std::vector<int> markovBlanketEdges(const std::vector<int> &mbVertices, const std::vector<int> &hubs) const;

    // round-by-round interface (what remove() loops over): beginRemoval, then planRound /
    // roundDescriptor / applyRound until planRound leaves an empty round
    struct Round {
        std::vector<int> sel;        // selected units: indices into the removal list, ascending
        HostBuf records;             // used when the remover has no context (rounds planned for a caller-run engine)
        uint64_t *rec = nullptr;     // the packed records of this round (context staging buffer, or `records`)
        std::vector<std::vector<double>> linPoses; // Local lin. point: optimised subgraph poses of non-star blankets
        std::vector<int64_t> recOff, outOff;
        int algorithm = SPG_ALG_NFR;
        bool poseOnly = false;       // every blanket edge of the round is a POSE edge (SPG_OPT_POSE_EDGES_ONLY)
        int64_t maxNewEdges = 0;     // substitute edges the round can add at most (the edge store is grown ahead)
    };
    spg_status beginRemoval(const std::vector<int> &toRemove);
    spg_status planRound(bool packNow = true);
    bool _packFailed = false;
    bool _extended = false;                 // Dense / CliqueyDense: a unit may remove other list entries
    std::vector<unsigned char> _dupInList;  // list entry repeats an earlier one
    bool packRange(size_t q0, size_t q1); // pack the selected units [q0, q1) of the planned round (planRound(false))
    spg_round_in roundDescriptor() const;
    void applyRound(const uint64_t *out);
    // the same in steps (remove() splices every pipeline chunk as its records arrive, see spg_host.cpp)
    void applyBegin();
    void applyRange(const uint64_t *out, size_t q0, size_t q1);
    void applyEnd();
    spg_status failureStatus(); // SPG_OK, or SPG_ERR_BLANKET_FAILED (+ error text) once a blanket failed

    spg_marginalize_stats stats{};
    std::string error;

private:
    bool buildUnit(int root, int listIndex, const std::set<int> &toRemoveSet, RemovalUnit &u) const;
    TopologyProvider *chooseTopologyProvider(const RemovalUnit &u) const;
    int64_t unitWords(const RemovalUnit &u) const;
    bool packUnit(const RemovalUnit &u, uint64_t *rec, int64_t words, const double *linPoses) const;
    bool needsSubgraphOptimisation(const RemovalUnit &u) const;
    spg_status localLinearise(const RemovalUnit &u, std::vector<double> &poses) const;

    Round _round;
    struct ApplyState {              // splice of the round in flight (applyBegin .. applyEnd)
        std::vector<int> base;       // substitutes of unit ui go to edges[e0 + base[ui] ...): slots by the provider's upper bound
        std::vector<int> used;       // how many of its slots the unit filled
        std::vector<char> okUnit;
        int e0 = 0, removedEdges = 0, removedVerts = 0;
        size_t spliced = 0;
        bool sized = false;          // the edge store holds the round's slots
    } _apply;
    void applySizeEdgeStore();
    std::vector<int> _pending, _added;
    std::set<int> _toRemoveSet;
    std::vector<char> _done;
    size_t _remaining = 0;
    // scheduler state: cached blankets of the pending vertices (re-extracted when stale) and per-round scratch
    UnitCache _unitCache;
    std::vector<int> _rootIdx;              // Graph::verts index of every list entry (-1: not in the graph)
    std::vector<int> _unitBuilt;            // planning pass that extracted the cached blanket (0: never)
    std::vector<int> _stamp;                // per vertex index: last planning pass whose round touched it
    // selection scratch. Per vertex (one cache line for both): head of its list of touching regions of this round, and
    // the region that removes it; the list nodes (next, region) side by side as well
    struct VTouch { int head = -1, removedBy = -1; };
    struct TouchNode { int next, region; };
    std::vector<VTouch> _vtouch;
    std::vector<TouchNode> _touchNodes;
    int _planNo = 0;
    // the window a round is drawn from: units deferred by earlier rounds (list order) + fresh entries from _cursor on
    std::vector<int> _leftover;
    size_t _cursor = 0, _window = 0;   // _window = 0: adaptive (env SPG_PLAN_WINDOW overrides; tests)
    HostBuf _outBuf;                   // output records of the round in flight (remove())
    std::vector<std::atomic<unsigned char>> _vlock; // per-vertex byte locks of the parallel splice
    SparsityOptions _opts;
    Graph *_graph = nullptr;
    spg_ctx *_ctx = nullptr;
    std::list<TopologyProvider *> _topologies;
};

// src/compute_substitute_edge.cpp:13-96 (meas: P doubles, info: d*d column-major)
void computeSubstituteEdge(const Graph *gw, const std::set<int> &marginalized, int maxid, int &from, int &to,
                           double *edgemeas, double *edgeinfo);

// small pose helpers shared by the host code (packer / substitute edge); not the hot path
void poseCompose(int dim, const double *a, const double *b, double *out);
void poseInverse(int dim, const double *a, double *out);

} // namespace spg

struct spg_graph {
    spg::Graph *g = nullptr;
    spg::VertexRemover *session = nullptr; // round-by-round removal in progress
    spg_marginalize_stats stats{};
};
