// spg_ctx.h — internal: context object and launch plumbing shared by the translation units of
// libspg_b200.so (one .cu per kernel instantiation so they compile in parallel).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/spg_capi.h"

void spg_set_err(const std::string &s);

#define SPG_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if(e_ != cudaSuccess) {                                                                   \
            spg_set_err(std::string(#call) + ": " + cudaGetErrorString(e_));                      \
            return SPG_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while(0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if(bytes <= cap) return cudaSuccess;
        if(p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if(e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if(p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct spg_ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;               // kernels (exposed by spg_stream)
    cudaStream_t s_in = nullptr, s_out = nullptr; // H2D / D2H of the chunked host-buffer path
    std::vector<cudaEvent_t> ev_pool;            // two per chunk (copy-in done, kernels done)
    // size buckets of one chunk run side by side (a bucket of a few hundred blankets fills a fraction of the SMs)
    static constexpr int N_SIDE = 4;
    cudaStream_t s_side[N_SIDE] = {};
    cudaEvent_t ev_side[N_SIDE] = {};
    int64_t chunk_bytes = (int64_t) 48 << 20;    // records + outputs per chunk of spg_remove_round (env SPG_CHUNK_BYTES)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0;
    double last_ms = 0;
    DevBuf d_rec, d_recoff, d_outoff, d_out, d_list, d_tgt, d_tgtoff, d_wts, d_wtsoff, d_ws, d_gws, d_prof;
    DevBuf d_retry, d_retry_cnt;                 // fast_kernel's refused blankets (per launch: a slice and a counter)
    int retry_used = 0;                          // counters used by the last call
    int fast_ctas_per_sm = 0;                    // resident CTAs per SM of the last fast_kernel launch (diagnostics)
    bool no_fast = false;                        // env SPG_NO_FAST=1: blanket_kernel only (A/B measurements, tests)
    bool profiling = false;
    // multi-GPU (spg_comm.cu): NCCL communicator of this context, its own stream, events compute -> gather -> D2H
    void *comm = nullptr; // ncclComm_t
    int nranks = 1, rank = 0;
    cudaStream_t s_comm = nullptr;
    std::vector<cudaEvent_t> ev_comm;
    cudaEvent_t ev_g0 = nullptr, ev_g1 = nullptr; // timing of the gathers of the last call
    double last_gather_ms = 0;
    int64_t last_gather_bytes = 0;
    // page-locked staging of the graph level (round records, round outputs): kept for the life of the context, since
    // cudaHostAlloc / cudaFreeHost of hundreds of MB per spg_graph_marginalize call cost more than a small graph does
    void *h_pinned[2] = {nullptr, nullptr};
    size_t h_pinned_words[2] = {0, 0};
};
// grow-only page-locked buffer `slot` (0: records, 1: outputs) of at least `words` 8-byte words; NULL on failure
uint64_t *spg_ctx_pinned(spg_ctx *ctx, int slot, size_t words);

// host-buffer round, shared between spg_remove_round (spg_capi.cu) and spg_remove_round_sharded (spg_comm.cu)
namespace spg {
struct RoundRun {
    std::vector<int32_t> flat; // bucket-ordered blanket indices of every chunk; must outlive the async copies
    int counter_next = 0;
    bool first = true;
    int64_t tgt_n = 0, wts_n = 0;
};
} // namespace spg
void spg_split_by_bytes(const spg_round_in *in, int b0, int b1, int parts, std::vector<int> &cb);
spg_status spg_round_prepare(spg_ctx *ctx, const spg_round_in *in, spg_round_out *out, int nchunks, spg::RoundRun &run);
spg_status spg_round_enqueue_chunk(spg_ctx *ctx, const spg_round_in *in, spg::RoundRun &run, int b0, int b1, int c);
spg_status spg_round_finish(spg_ctx *ctx, spg_round_out *out, spg::RoundRun &run);
void spg_comm_release(spg_ctx *ctx);

namespace spg { struct KernelParams; }
// one per kernel instantiation (spg_inst_<D>_<NT>.cu)
spg_status spg_launch_3_32(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_3_64(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_3_128(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_3_256(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_3_512(spg_ctx *, spg::KernelParams &); // two sweep groups of 256, 128 registers: N <= 80
spg_status spg_launch_6_32(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_6_64(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_6_128(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_6_256(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_6_512(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_3_256l(spg_ctx *, spg::KernelParams &);  // lean: 256 threads, 128 registers, two CTAs per SM
spg_status spg_launch_6_256l(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_3_spill(spg_ctx *, spg::KernelParams &); // 256 threads, buffers in global memory
spg_status spg_launch_6_spill(spg_ctx *, spg::KernelParams &);
// fast_kernel<D, MAXW> (spg_fast.cuh): NFR tree rounds; SPG_ERR_UNSUPPORTED when the bucket does not fit
spg_status spg_launch_fast_6_g8(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_fast_6_g16(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_fast_6_g32(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_fast_6_c4(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_fast_6_c8(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_fast_3_g8(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_fast_3_g16(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_fast_3_g32(spg_ctx *, spg::KernelParams &);
spg_status spg_launch_fast_3_c8(spg_ctx *, spg::KernelParams &);
