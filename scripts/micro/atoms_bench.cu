// micro-benchmark: shared-memory atomic add vs match.any + private RMW vs global RED, random addresses
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
template <int MODE>
__global__ void k(unsigned* g, int iters, int span, unsigned long long* sink) {
    extern __shared__ unsigned sm[];
    for (int i = threadIdx.x; i < span * (MODE == 1 ? (blockDim.x >> 5) : 1); i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t s = blockIdx.x * 1315423911u + threadIdx.x * 2654435761u + 7u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        const uint32_t a = rng(s) % span;
        if (MODE == 0) atomicAdd(&sm[a], 1u);
        if (MODE == 1) {
            unsigned* u = sm + warp * span;
            const unsigned grp = __match_any_sync(0xffffffffu, a);
            if (lane == __ffs(grp) - 1) u[a] += __popc(grp);
            __syncwarp();
        }
        if (MODE == 2) atomicAdd(&g[(size_t)a * 37u % (1u << 22)], 1u);
        if (MODE == 3) acc += sm[a];
        if (MODE == 4) { unsigned* u = sm + 0; u[a] += 1; }  // racy plain RMW (cost reference only)
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = sm[0] + acc;
}
template <int MODE> void run(const char* name, int threads, size_t smem, int span) {
    unsigned* g; unsigned long long* sink;
    cudaMalloc(&g, sizeof(unsigned) << 22); cudaMemset(g, 0, sizeof(unsigned) << 22);
    cudaMalloc(&sink, 8 * 1024);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 2;
    k<MODE><<<blocks, threads, smem>>>(g, 64, span, sink);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads, smem>>>(g, iters, span, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters;
    printf("%-28s %8.3f ms  %8.1f Gop/s  (%.2f cyc/lane/SM at 1.9 GHz) err=%s\n", name, ms, ops / ms / 1e6,
           ms * 1e-3 * 1.9e9 * 148 / ops, cudaGetErrorString(cudaGetLastError()));
    cudaFree(g); cudaFree(sink);
}
int main() {
    const int span = 1664;
    run<3>("LDS random", 512, span * 4, span);
    run<4>("LDS+STS random (racy)", 512, span * 4, span);
    run<0>("ATOMS add random", 512, span * 4, span);
    run<1>("match.any + private RMW", 512, span * 4 * 16, span);
    run<2>("REDG random 16MB", 512, span * 4, span);
    return 0;
}
