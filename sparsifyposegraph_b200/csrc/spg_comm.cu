// spg_comm.cu — multi-GPU side of the C ABI (include/spg_capi.h, "sharded rounds").
//
// The blankets of a wavefront round are independent, so a round is split into contiguous,
// cost-balanced shards, one per rank / GPU. The only exchange of the path is the gather of the substitute-edge
// records: NCCL over NVLink directly on the device output buffers (every rank keeps the round's output buffer at the
// same offsets, so the gather is in place: rank r's slice is broadcast from r — an all-gather with ragged counts —
// or sent to the root that holds the graph). The gather runs on its own stream behind the kernels of a chunk and
// overlaps the kernels of the next one.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already carries, e.g. PyTorch's, else the
// system one): libspg_b200.so itself has no link-time dependency on it and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "spg_ctx.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;

bool load_nccl() {
    if(g_nccl.handle) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); // the copy this process already uses
    if(!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if(!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if(!h) {
        spg_set_err(std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror());
        return false;
    }
    NcclApi a;
    a.handle = h;
#define SPG_SYM(field, name)                                                  \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name));            \
    if(!a.field) {                                                            \
        spg_set_err(std::string("NCCL symbol missing: ") + name);             \
        return false;                                                         \
    }
    SPG_SYM(GetUniqueId, "ncclGetUniqueId")
    SPG_SYM(CommInitRank, "ncclCommInitRank")
    SPG_SYM(CommDestroy, "ncclCommDestroy")
    SPG_SYM(Broadcast, "ncclBroadcast")
    SPG_SYM(Send, "ncclSend")
    SPG_SYM(Recv, "ncclRecv")
    SPG_SYM(GroupStart, "ncclGroupStart")
    SPG_SYM(GroupEnd, "ncclGroupEnd")
    SPG_SYM(GetErrorString, "ncclGetErrorString")
    SPG_SYM(GetVersion, "ncclGetVersion")
#undef SPG_SYM
    g_nccl = a;
    return true;
}

#define SPG_NCCL(call)                                                                           \
    do {                                                                                         \
        ncclResult_t r_ = (call);                                                                \
        if(r_ != ncclSuccess) {                                                                  \
            spg_set_err(std::string(#call) + ": " + g_nccl.GetErrorString(r_));                  \
            return SPG_ERR_COMM;                                                                 \
        }                                                                                        \
    } while(0)

// per-blanket cost model for the shard balance (only ratios matter). NFR tree: SM-time per blanket measured on the C4
// sweep (profiles/README.md: 1.0 / 1.6 / 4.4 / 6.8 / 10.6 / 19 / 33.6 / 45.6 ms per 1e5 blankets for n = 2 ... 16) is
// close to linear in n, because wider blankets also get wider thread groups; the other paths (GLC factors, Newton
// iterations) are cubic in the blanket size with a per-edge assembly term.
inline double blanket_cost(int algorithm, int topology, int nv, int ne) {
    if(algorithm == SPG_ALG_NFR && topology == SPG_TOPO_TREE) return nv <= 2 ? 1.0 : (nv == 3 ? 1.6 : 3.4 * (nv - 2.6));
    return 40.0 + 12.0 * ne + (double) nv * nv * (nv + 6.0);
}

// cum[b] = modelled cost of blankets [0, b): the headers are read by several host threads (one cache miss per blanket
// in a multi-GB record buffer), the prefix sum is sequential
void cost_prefix(const spg_round_in *in, std::vector<double> &cum) {
    const int nb = in->n_blankets;
    cum.assign((size_t) nb + 1, 0.0);
    auto fill = [&](int lo, int hi) {
        for(int b = lo; b < hi; b++) {
            const int32_t *h = reinterpret_cast<const int32_t *>(in->records + in->rec_off[b]);
            cum[(size_t) b + 1] = blanket_cost(in->algorithm, in->opts.topology, h[0], h[2]);
        }
    };
    const unsigned nthr = nb < 16384 ? 1u : std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 8u);
    if(nthr <= 1) fill(0, nb);
    else {
        std::vector<std::thread> pool;
        const int per = (nb + (int) nthr - 1) / (int) nthr;
        for(unsigned t = 0; t < nthr; t++) {
            const int lo = std::min(nb, (int) t * per), hi = std::min(nb, lo + per);
            if(lo < hi) pool.emplace_back(fill, lo, hi);
        }
        for(auto &th : pool) th.join();
    }
    for(int b = 0; b < nb; b++) cum[(size_t) b + 1] += cum[b];
}

// contiguous split of blankets [b0, b1) into R parts of about equal modelled cost: bounds[0..R]
void cost_bounds(const std::vector<double> &cum, int b0, int b1, int R, int32_t *bounds) {
    bounds[0] = b0;
    const double base = cum[b0], total = cum[b1] - base;
    for(int r = 1; r < R; r++) {
        const double goal = base + total * r / R;
        int b = (int) (std::lower_bound(cum.begin() + b0, cum.begin() + b1 + 1, goal) - cum.begin());
        if(b > b0 && goal - cum[b - 1] < cum[b] - goal) b--; // nearer boundary
        bounds[r] = std::max(bounds[r - 1], std::min(b, b1));
    }
    bounds[R] = b1;
}

spg_status ensure_comm_stream(spg_ctx *ctx, int n_events) {
    if(!ctx->s_comm) SPG_CUDA(cudaStreamCreateWithFlags(&ctx->s_comm, cudaStreamNonBlocking));
    if(!ctx->ev_g0) {
        SPG_CUDA(cudaEventCreate(&ctx->ev_g0));
        SPG_CUDA(cudaEventCreate(&ctx->ev_g1));
    }
    while((int) ctx->ev_comm.size() < n_events) {
        cudaEvent_t e;
        SPG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_comm.push_back(e);
    }
    return SPG_OK;
}

// In-place ragged gather on stream s_comm: rank r contributes the output words [w0[r], w0[r] + cnt[r]).
// root < 0: every rank receives everything (grouped broadcasts = an all-gather with ragged counts);
// root >= 0: only the root receives (grouped send / recv).
spg_status gather_outputs(spg_ctx *ctx, uint64_t *d_out, const int64_t *w0, const int64_t *cnt, int root, int64_t *bytes_moved) {
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    const int R = ctx->nranks, me = ctx->rank;
    if(R == 1) return SPG_OK;
    SPG_NCCL(g_nccl.GroupStart());
    for(int r = 0; r < R; r++) {
        if(cnt[r] <= 0) continue;
        uint64_t *p = d_out + w0[r];
        if(root < 0) {
            SPG_NCCL(g_nccl.Broadcast(p, p, (size_t) cnt[r], ncclUint64, r, comm, ctx->s_comm));
            if(r != me) *bytes_moved += cnt[r] * 8;
        } else if(r != root) {
            if(me == r) {
                SPG_NCCL(g_nccl.Send(p, (size_t) cnt[r], ncclUint64, root, comm, ctx->s_comm));
                *bytes_moved += cnt[r] * 8;
            } else if(me == root) {
                SPG_NCCL(g_nccl.Recv(p, (size_t) cnt[r], ncclUint64, r, comm, ctx->s_comm));
                *bytes_moved += cnt[r] * 8;
            }
        }
    }
    SPG_NCCL(g_nccl.GroupEnd());
    return SPG_OK;
}

} // namespace

void spg_comm_release(spg_ctx *ctx) {
    if(ctx->comm && g_nccl.handle) g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
    ctx->nranks = 1;
    ctx->rank = 0;
    for(cudaEvent_t e : ctx->ev_comm) cudaEventDestroy(e);
    ctx->ev_comm.clear();
    if(ctx->ev_g0) cudaEventDestroy(ctx->ev_g0);
    if(ctx->ev_g1) cudaEventDestroy(ctx->ev_g1);
    ctx->ev_g0 = ctx->ev_g1 = nullptr;
    if(ctx->s_comm) cudaStreamDestroy(ctx->s_comm);
    ctx->s_comm = nullptr;
}

extern "C" {

spg_status spg_comm_unique_id(uint8_t *id) {
    if(!id) return SPG_ERR_INVALID;
    if(!load_nccl()) return SPG_ERR_COMM;
    static_assert(SPG_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    ncclUniqueId u;
    SPG_NCCL(g_nccl.GetUniqueId(&u));
    std::memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return SPG_OK;
}

spg_status spg_comm_init(spg_ctx *ctx, int32_t nranks, int32_t rank, const uint8_t *id) {
    if(!ctx || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id)) {
        spg_set_err("spg_comm_init: bad arguments");
        return SPG_ERR_INVALID;
    }
    SPG_CUDA(cudaSetDevice(ctx->device));
    spg_comm_release(ctx);
    if(nranks > 1) {
        if(!load_nccl()) return SPG_ERR_COMM;
        ncclUniqueId u;
        std::memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
        ncclComm_t comm = nullptr;
        SPG_NCCL(g_nccl.CommInitRank(&comm, nranks, u, rank));
        ctx->comm = comm;
    }
    ctx->nranks = nranks;
    ctx->rank = rank;
    return ensure_comm_stream(ctx, 2);
}

spg_status spg_comm_destroy(spg_ctx *ctx) {
    if(!ctx) return SPG_ERR_INVALID;
    cudaSetDevice(ctx->device);
    spg_comm_release(ctx);
    return SPG_OK;
}

spg_status spg_comm_join(spg_ctx *ctx) {
    if(!ctx) return SPG_ERR_INVALID;
    if(!ctx->s_comm || ctx->nranks <= 1) return SPG_OK;
    SPG_CUDA(cudaSetDevice(ctx->device));
    SPG_CUDA(cudaEventRecord(ctx->ev_comm[1], ctx->s_comm));
    SPG_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm[1], 0));
    return SPG_OK;
}

int32_t spg_comm_nranks(const spg_ctx *ctx) { return ctx ? ctx->nranks : 0; }
int32_t spg_comm_rank(const spg_ctx *ctx) { return ctx ? ctx->rank : -1; }
int32_t spg_comm_nccl_version(void) {
    int v = 0;
    if(!load_nccl() || g_nccl.GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

// pure host code (CPU tests cover it): contiguous shards of the round, balanced by the per-blanket cost model
spg_status spg_shard_bounds(const spg_round_in *in, int32_t nranks, int32_t *bounds) {
    if(!in || !bounds || nranks < 1 || in->n_blankets < 0) return SPG_ERR_INVALID;
    std::vector<double> cum;
    cost_prefix(in, cum);
    cost_bounds(cum, 0, in->n_blankets, nranks, bounds);
    return SPG_OK;
}

spg_status spg_remove_round_sharded_device(spg_ctx *ctx, const spg_round_in *in, spg_round_out *out,
                                           const int32_t *bounds, const int64_t *out_word_bounds, int32_t max_n_vert,
                                           int32_t max_n_edges, int32_t max_rec_words, int32_t root) {
    if(!ctx || !in || !out || !bounds || !out_word_bounds || root >= ctx->nranks) {
        spg_set_err("spg_remove_round_sharded_device: bad arguments");
        return SPG_ERR_INVALID;
    }
    if(ctx->nranks > 1 && !ctx->comm) {
        spg_set_err("spg_comm_init has not been called on this context");
        return SPG_ERR_INVALID;
    }
    SPG_CUDA(cudaSetDevice(ctx->device));
    spg_status st = ensure_comm_stream(ctx, 2);
    if(st != SPG_OK) return st;
    const int b0 = bounds[ctx->rank], b1 = bounds[ctx->rank + 1];
    if(b1 > b0) {
        spg_round_in sub = *in; // the offsets are absolute, so the shard is the same round seen from blanket b0 on
        sub.n_blankets = b1 - b0;
        sub.rec_off = in->rec_off + b0;
        sub.out_off = in->out_off + b0;
        spg_round_out so = *out;
        so.dbg_target = nullptr;
        so.dbg_weights = nullptr;
        st = spg_remove_round_device(ctx, &sub, &so, max_n_vert, max_n_edges, max_rec_words);
        if(st != SPG_OK) return st;
    }
    if(ctx->nranks > 1) {
        // gather behind this call's kernels; the next call's kernels run beside it
        SPG_CUDA(cudaEventRecord(ctx->ev_comm[0], ctx->stream));
        SPG_CUDA(cudaStreamWaitEvent(ctx->s_comm, ctx->ev_comm[0], 0));
        int64_t moved = 0;
        std::vector<int64_t> cnt((size_t) ctx->nranks);
        for(int r = 0; r < ctx->nranks; r++) cnt[r] = out_word_bounds[r + 1] - out_word_bounds[r];
        st = gather_outputs(ctx, out->out, out_word_bounds, cnt.data(), root, &moved);
        if(st != SPG_OK) return st;
        ctx->last_gather_bytes = moved;
    }
    return SPG_OK;
}

spg_status spg_remove_round_sharded(spg_ctx *ctx, const spg_round_in *in, spg_round_out *out, int32_t root,
                                    spg_shard_info *info) {
    if(!ctx || !in || !out || !out->out || (in->dim != 3 && in->dim != 6) || in->n_blankets < 0 || root >= ctx->nranks ||
       root < SPG_ROOT_SHARED_HOST) {
        spg_set_err("spg_remove_round_sharded: bad arguments");
        return SPG_ERR_INVALID;
    }
    if(ctx->nranks > 1 && !ctx->comm) {
        spg_set_err("spg_comm_init has not been called on this context");
        return SPG_ERR_INVALID;
    }
    if(info) *info = spg_shard_info{};
    const int nb = in->n_blankets, R = ctx->nranks, me = ctx->rank;
    if(nb == 0) return SPG_OK;
    if((out->dbg_target && out->dbg_target_off) || (out->dbg_weights && out->dbg_weights_off)) {
        spg_set_err("debug outputs are not gathered: use spg_remove_round");
        return SPG_ERR_INVALID;
    }
    SPG_CUDA(cudaSetDevice(ctx->device));

    const bool prof = getenv("SPG_SHARD_PROF") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tp0 = now();
    // ---- every rank derives the same plan from the (replicated) round description --------------------------------
    // Step-major: the round is cut into `steps` contiguous chunks of about equal bytes, and every chunk is split over
    // the ranks by modelled cost. Every rank then moves 1/R of the bytes and does 1/R of the work of every chunk,
    // whatever the order of blanket sizes inside the round, and step s of all ranks is one gather.
    const int64_t total_bytes = (in->rec_off[nb] + in->out_off[nb]) * 8;
    const int steps = (int) std::max<int64_t>(1, std::min<int64_t>(32, total_bytes / std::max<int64_t>(ctx->chunk_bytes, 4096)));
    std::vector<int> gcb;
    spg_split_by_bytes(in, 0, nb, steps, gcb);
    std::vector<int32_t> sb((size_t) R + 1);
    std::vector<std::vector<int32_t>> lo((size_t) steps, std::vector<int32_t>((size_t) R + 1)); // lo[s][r] .. lo[s][r+1]: rank r's blankets of step s
    int mine = 0;
    // Inside a step the split is by BYTES (records + outputs), read off the offset tables alone: every rank derives the
    // plan of the whole round on every call, so the plan must not touch the records (a header read per blanket is a cache
    // miss per blanket, times the number of ranks). A pipeline step is a short contiguous run of the round, and for the
    // NFR-tree path cost per byte is flat (both grow linearly with the blanket size), so bytes balance the work too;
    // spg_shard_bounds offers the cost-model split for callers that shard by themselves.
    std::vector<int> sbi;
    for(int s2 = 0; s2 < steps; s2++) {
        spg_split_by_bytes(in, gcb[s2], gcb[s2 + 1], R, sbi);
        for(int r = 0; r <= R; r++) sb[r] = sbi[r];
        lo[s2] = sb;
        mine += sb[me + 1] - sb[me];
    }
    spg_status st = SPG_OK;
    st = ensure_comm_stream(ctx, steps);
    if(st != SPG_OK) return st;

    const double tp1 = now();
    spg::RoundRun run;
    st = spg_round_prepare(ctx, in, out, steps, run);
    if(st != SPG_OK) return st;
    const double tp2 = now();
    const bool shared_host = root == SPG_ROOT_SHARED_HOST; // every rank writes its own records into one shared host buffer
    const bool receiver = !shared_host && (root < 0 || root == me);
    uint64_t *d_out = reinterpret_cast<uint64_t *>(ctx->d_out.p);
    int64_t moved = 0, h2d = 2 * (int64_t) (nb + 1) * 8, d2h = 0;
    std::vector<int64_t> w0((size_t) R), cnt((size_t) R);
    SPG_CUDA(cudaEventRecord(ctx->ev_g0, ctx->s_comm));
    for(int s = 0; s < steps; s++) {
        const int b0 = lo[s][me], b1 = lo[s][me + 1];
        bool launched = false;
        if(b1 > b0) {
            st = spg_round_enqueue_chunk(ctx, in, run, b0, b1, s);
            if(st != SPG_OK) return st;
            launched = true;
        }
        if(R > 1 && !shared_host) {
            // the slices of step s of all ranks are disjoint word ranges of the one output buffer
            if(launched) SPG_CUDA(cudaStreamWaitEvent(ctx->s_comm, ctx->ev_pool[2 * s + 1], 0));
            for(int r = 0; r < R; r++) {
                w0[r] = in->out_off[lo[s][r]];
                cnt[r] = in->out_off[lo[s][r + 1]] - w0[r];
            }
            st = gather_outputs(ctx, d_out, w0.data(), cnt.data(), root, &moved);
            if(st != SPG_OK) return st;
            SPG_CUDA(cudaEventRecord(ctx->ev_comm[s], ctx->s_comm));
            if(receiver) SPG_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_comm[s], 0));
        } else if(launched && !shared_host) {
            SPG_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_pool[2 * s + 1], 0));
        }
        // ---- D2H: a receiver reads the whole step (the slices of the ranks are adjacent), the others only their own
        // slice (they keep a copy of what they produced; the graph-holding root gets everything)
        {
            const int r0 = receiver ? 0 : me, r1 = receiver ? R : me + 1;
            const int64_t o0 = in->out_off[lo[s][r0]], oc = in->out_off[lo[s][r1]] - o0;
            if(oc > 0) {
                if(!receiver) SPG_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_pool[2 * s + 1], 0));
                SPG_CUDA(cudaMemcpyAsync(out->out + o0, d_out + o0, (size_t) oc * 8, cudaMemcpyDeviceToHost, ctx->s_out));
                d2h += oc * 8;
            }
            h2d += (in->rec_off[b1] - in->rec_off[b0]) * 8 + (int64_t) (b1 - b0) * 4;
        }
    }
    SPG_CUDA(cudaEventRecord(ctx->ev_g1, ctx->s_comm));
    const double tp3 = now();
    st = spg_round_finish(ctx, out, run);
    if(st != SPG_OK) return st;
    if(prof)
        fprintf(stderr, "[spg shard r%d] plan %.1f ms, prepare %.1f ms, enqueue loop (validate + bucket + launches) %.1f ms, drain %.1f ms\n", me,
                tp1 - tp0, tp2 - tp1, tp3 - tp2, now() - tp3);
    SPG_CUDA(cudaStreamSynchronize(ctx->s_comm));
    float gms = 0;
    SPG_CUDA(cudaEventElapsedTime(&gms, ctx->ev_g0, ctx->ev_g1));
    ctx->last_gather_ms = gms;
    ctx->last_gather_bytes = moved;
    if(info) {
        info->nranks = R;
        info->rank = me;
        info->n_blankets_mine = mine;
        info->reserved0 = 0;
        info->steps = steps;
        info->kernel_ms = ctx->last_ms;
        info->gather_window_ms = gms;
        info->gather_bytes = moved;
        info->h2d_bytes = h2d;
        info->d2h_bytes = d2h;
    }
    return SPG_OK;
}

} // extern "C"
