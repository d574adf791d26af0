// spg_inst.cuh — launch wrapper of one blanket_kernel<D, NT> instantiation.
#pragma once
#include <algorithm>
#include <mutex>

#include "spg_ctx.h"
#include "spg_kernels.cuh"

namespace {
#define set_err spg_set_err
template <int D, int NT, bool SPILL = false, bool LEAN = false>
spg_status launch_bucket(spg_ctx *ctx, spg::KernelParams &kp) {
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: one flag per device ordinal
    static std::mutex mu;
    static bool configured[64] = {};
    size_t smem = (size_t) kp.total_doubles * sizeof(double);
    if(SPILL) smem = 0;
    if(smem > ctx->smem_optin) {
        set_err("bucket needs more shared memory than the device offers");
        return SPG_ERR_INVALID;
    }
    if(!SPILL) {
        std::lock_guard<std::mutex> lk(mu);
        const int dv = ctx->device & 63;
        if(!configured[dv]) {
            SPG_CUDA(cudaFuncSetAttribute(spg::blanket_kernel<D, NT, SPILL, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int) ctx->smem_optin));
            configured[dv] = true;
        }
    }
    int per_sm = 0;
    SPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spg::blanket_kernel<D, NT, SPILL, LEAN>, NT, smem));
    if(LEAN && per_sm < 2) return SPG_ERR_UNSUPPORTED; // the caller falls back to the one-CTA-per-SM variant
    if(per_sm < 1) per_sm = 1;
    int grid = std::min<int64_t>(kp.n_list, (int64_t) per_sm * ctx->sm_count);
    if(grid < 1) return SPG_OK;
    kp.gws = nullptr;
    kp.gws_stride = 0;
    if(SPILL) {
        const int64_t stride = ((int64_t) kp.total_doubles + 15) & ~(int64_t) 15;
        const int64_t budget = (int64_t) 8 << 27; // 8 GiB of doubles at most
        if(stride > budget) {
            set_err("blanket too large for the spill workspace");
            return SPG_ERR_INVALID;
        }
        if((int64_t) grid * stride > budget) grid = (int) std::max<int64_t>(1, budget / stride);
        SPG_CUDA(ctx->d_gws.reserve((size_t) grid * stride * sizeof(double)));
        kp.gws = reinterpret_cast<double *>(ctx->d_gws.p);
        kp.gws_stride = stride;
    }
    kp.prof = ctx->profiling ? reinterpret_cast<unsigned long long *>(ctx->d_prof.p) : nullptr;
    // iterative NFR (Subgraph / Dense, >= 3 kept vertices): per-CTA global workspace
    kp.nfr_ws = nullptr;
    kp.nfr_ws_stride = 0;
    const int nkmax = kp.max_nv - 1;
    if(kp.algorithm == SPG_ALG_NFR && nkmax >= 3 && (kp.topology == SPG_TOPO_SUBGRAPH || kp.topology == SPG_TOPO_DENSE)) {
        const int ne = spgr_out_edge_count(SPG_ALG_NFR, kp.topology, kp.chord_ratio, nkmax);
        if(ne > nkmax - 1) {
            const int64_t stride = (spg::nfr_work_doubles(ne, D, D * (nkmax - 1)) + 1) & ~(int64_t) 1;
            const int64_t budget = (int64_t) 6 << 27; // 6 GiB of doubles workspace at most
            if(stride <= budget) {
                if((int64_t) grid * stride > budget) grid = (int) std::max<int64_t>(1, budget / stride);
                SPG_CUDA(ctx->d_ws.reserve((size_t) grid * stride * sizeof(double)));
                kp.nfr_ws = reinterpret_cast<double *>(ctx->d_ws.p);
                kp.nfr_ws_stride = stride;
            }
        }
    }
    spg::blanket_kernel<D, NT, SPILL, LEAN><<<grid, NT, smem, ctx->stream>>>(kp);
    SPG_CUDA(cudaGetLastError());
    ctx->launches++;
    return SPG_OK;
}

#undef set_err
} // namespace
