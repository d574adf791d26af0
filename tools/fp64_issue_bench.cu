// fp64_issue_bench.cu — how many warps / how much ILP does the B200 FP64 pipe need? One CTA per SM,
// W warps, ILP independent DFMA chains per thread. Prints cycles per warp-level DFMA per SM sub-partition.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/build/fp64_issue_bench tools/fp64_issue_bench.cu
#include <cstdio>
#include <vector>

template <int ILP>
__global__ void k(double *sink, int iters, double seed, long long *cyc) {
    double a[ILP];
#pragma unroll
    for(int i = 0; i < ILP; i++) a[i] = seed + threadIdx.x + i;
    const double m = 1.0000001, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for(int it = 0; it < iters; it++) {
#pragma unroll
        for(int u = 0; u < 4; u++)
#pragma unroll
            for(int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    }
    const long long t1 = clock64();
    double r = 0;
#pragma unroll
    for(int i = 0; i < ILP; i++) r += a[i];
    if(r == 12345.678) sink[0] = r;
    if(threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps) {
    long long *dc;
    double *ds;
    cudaMalloc(&dc, 148 * 8);
    cudaMalloc(&ds, 8);
    const int iters = 2000;
    cudaFuncSetAttribute(k<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<ILP><<<148, warps * 32, 200 * 1024>>>(ds, iters, 1.0, dc);
    k<ILP><<<148, warps * 32, 200 * 1024>>>(ds, iters, 1.0, dc);
    cudaDeviceSynchronize();
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), dc, 148 * 8, cudaMemcpyDeviceToHost);
    double mean = 0;
    for(auto v : h) mean += (double) v / 148;
    const double dfma_per_warp = (double) iters * 4 * ILP;
    const double warps_per_smsp = warps / 4.0;
    printf("warps/SM %2d ILP %2d: %7.2f cycles per DFMA per warp, %6.2f cycles per DFMA per SMSP (%.1f%% of 2-cycle rate)\n", warps, ILP,
           mean / dfma_per_warp, mean / (dfma_per_warp * warps_per_smsp), 100.0 * 2.0 / (mean / (dfma_per_warp * warps_per_smsp)));
    cudaFree(dc); cudaFree(ds);
}

int main() {
    for(int w : {4, 8, 16, 32}) {
        run<1>(w); run<2>(w); run<4>(w); run<8>(w); run<16>(w); run<36>(w);
    }
    return 0;
}
