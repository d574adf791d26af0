// spg_fast_inst.cuh — launch wrapper of one fast_kernel<D, MAXW> instantiation (spg_fast.cuh).
#pragma once
#include <algorithm>
#include <mutex>

#include "spg_ctx.h"
#include "spg_fast.cuh"

namespace {
// Returns SPG_ERR_UNSUPPORTED when the bucket does not fit this instantiation (the caller runs blanket_kernel alone).
template <int D, int MAXW>
spg_status launch_fast(spg_ctx *ctx, spg::KernelParams &kp) {
    static std::mutex mu;
    static bool configured[64] = {};
    // shared memory per CTA that still lets the register file's CTA count (MINB) be resident
    constexpr int MINB = MAXW == 1 ? (D == 6 ? 12 : 16) : (MAXW <= 4 ? 3 : (D == 6 ? 1 : 2));
    const size_t budget = std::min<size_t>(ctx->smem_optin, (size_t) (228 * 1024) / MINB - 1024);
    const spg::FastPlan pl = spg::fast_plan<D>(kp.max_nv, kp.max_e, kp.max_rec_words, budget);
    kp.fast = pl;
    const size_t smem = (size_t) pl.total * sizeof(double);
    const int threads = spg::fast_threads(kp.max_nv);
    if(smem > ctx->smem_optin || threads > 32 * MAXW) return SPG_ERR_UNSUPPORTED;
    {
        std::lock_guard<std::mutex> lk(mu);
        const int dv = ctx->device & 63;
        if(!configured[dv]) {
            SPG_CUDA(cudaFuncSetAttribute(spg::fast_kernel<D, MAXW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) ctx->smem_optin));
            SPG_CUDA(cudaFuncSetAttribute(spg::fast_kernel<D, MAXW>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          (int) cudaSharedmemCarveoutMaxShared));
            configured[dv] = true;
        }
    }
    int per_sm = 0;
    SPG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, spg::fast_kernel<D, MAXW>, threads, smem));
    if(per_sm < 1) return SPG_ERR_UNSUPPORTED;
    const int grid = (int) std::min<int64_t>(kp.n_list, (int64_t) per_sm * ctx->sm_count);
    if(grid < 1) return SPG_OK;
    kp.prof = ctx->profiling ? reinterpret_cast<unsigned long long *>(ctx->d_prof.p) + 16 : nullptr;
    spg::fast_kernel<D, MAXW><<<grid, threads, smem, ctx->stream>>>(kp);
    SPG_CUDA(cudaGetLastError());
    ctx->launches++;
    ctx->fast_ctas_per_sm = per_sm;
    return SPG_OK;
}
} // namespace
