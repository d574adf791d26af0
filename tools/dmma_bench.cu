// dmma_bench.cu — FP64 tensor-core (mma.sync.m8n8k4.f64) issue rate on B200 vs warps per SM and independent chains.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/dmma_bench tools/dmma_bench.cu
#include <cstdio>
#include <vector>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k(double *sink, int iters, double seed, long long *cyc) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for(int i = 0; i < ILP; i++) { c0[i] = seed + i; c1[i] = seed - i; }
    const double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.25;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for(int it = 0; it < iters; it++) {
#pragma unroll
        for(int u = 0; u < 4; u++)
#pragma unroll
            for(int i = 0; i < ILP; i++) dmma(c0[i], c1[i], a, b);
    }
    const long long t1 = clock64();
    double r = 0;
#pragma unroll
    for(int i = 0; i < ILP; i++) r += c0[i] + c1[i];
    if(r == 12345.678) sink[0] = r;
    if(threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps) {
    long long *dc;
    double *ds;
    cudaMalloc(&dc, 148 * 8);
    cudaMalloc(&ds, 8);
    const int iters = 1000;
    cudaFuncSetAttribute(k<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<ILP><<<148, warps * 32, 200 * 1024>>>(ds, iters, 1.0, dc);
    k<ILP><<<148, warps * 32, 200 * 1024>>>(ds, iters, 1.0, dc);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
    double mean = 0;
    for(auto v : h) mean += (double) v / 148;
    const double per_warp = (double) iters * 4 * ILP;
    const double cyc_per_dmma_smsp = mean / (per_warp * warps / 4.0);
    // one DMMA = 8x8x4 = 256 FMA; the DFMA pipe does 16 FMA / cycle / sub-partition
    printf("warps/SM %2d chains %2d: %7.2f cycles per DMMA per warp, %6.2f per sub-partition -> %5.1f FMA/cycle/SM  (%s)\n", warps, ILP, mean / per_warp,
           cyc_per_dmma_smsp, 4 * 256.0 / cyc_per_dmma_smsp, cudaGetErrorString(e));
    cudaFree(dc); cudaFree(ds);
}

int main() {
    for(int w : {4, 8, 16}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); run<18>(w); }
    return 0;
}
