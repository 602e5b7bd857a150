// Micro-benchmark: cycles per tcgen05.mma (M = 128, N = 256, one K step = 32 bytes of K) for kind::tf32 with A from shared
// memory / from tensor memory and kind::f16 (bf16 operands), R back-to-back instructions accumulating into one D.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/micro/mma_rate scripts/micro/mma_rate.cu && scripts/micro/mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fffu);
    d |= (uint64_t)(128u >> 4) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

template <int KIND>  // 0: tf32 SS, 1: tf32 TS, 2: bf16 SS
__global__ void k(int R, int N, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < 40 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bar)) : "memory");
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(sa(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = s_tmem;
    if (tid == 0) {
        const uint32_t fmt = KIND == 2 ? 1u : 2u;  // bf16 : tf32
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t ad = desc(sa(smem)), bd = desc(sa(smem) + 8192);
        long long c0 = clock64();
        for (int r = 0; r < R; ++r) {
            if (KIND == 0)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(t0), "l"(ad), "l"(bd), "r"(idesc), "r"(r) : "memory");
            else if (KIND == 1)
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(t0), "r"(t0 + 256), "l"(bd), "r"(idesc), "r"(r) : "memory");
            else
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(t0), "l"(ad), "l"(bd), "r"(idesc), "r"(r) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sa(&bar)) : "memory");
        asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(sa(&bar)) : "memory");
        long long c1 = clock64();
        if (blockIdx.x == 0) out[0] = c1 - c0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(t0) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    const char* names[3] = {"tf32 A smem", "tf32 A tmem", "bf16 A smem"};
    for (int N : {32, 64, 128, 256}) {
        for (int kind = 0; kind < 3; ++kind) {
            for (int grid : {1, 148}) {
                const int R = 512;
                auto launch = [&](int r) {
                    if (kind == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024); k<0><<<grid, 128, 48 * 1024>>>(r, N, d); }
                    if (kind == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024); k<1><<<grid, 128, 48 * 1024>>>(r, N, d); }
                    if (kind == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024); k<2><<<grid, 128, 48 * 1024>>>(r, N, d); }
                };
                launch(R);
                cudaDeviceSynchronize();
                long long c1 = 0, c2 = 0;
                launch(R); cudaMemcpy(&c1, d, 8, cudaMemcpyDeviceToHost);
                launch(2 * R); cudaMemcpy(&c2, d, 8, cudaMemcpyDeviceToHost);
                cudaError_t e = cudaGetLastError();
                const double per = double(c2 - c1) / R;
                const int kelem = kind == 2 ? 16 : 8;
                printf("N=%3d %-12s grid=%3d: %7.1f cycles per MMA (M128 x N%d x K%d) = %6.0f MAC/clk/SM  %s\n", N, names[kind], grid, per, N, kelem,
                       128.0 * N * kelem / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
        }
    }
    return 0;
}
