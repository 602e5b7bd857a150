// fp32-accurate Linear layers on the 5th-generation tensor cores, for the predictor heads wider than csrc/head_tc.cu
// serves (hidden 64 / 128 / 256: the ppa / ddi and the Cora / Pubmed / collab configs, model.py:2192-2235):
//
//     v = A[rows x K] . W[N x K]^T + bias ;  v = LayerNorm(v) ;  v = ReLU(v)          (each optional)
//     out = v ;  z = (z +) z_scale * v ;  out_final[:, o] = <v, Wo[o]> + bo[o]         (each optional)
//
// One call is one layer with its element-wise tail fused, so a head is 11 - 13 launches that read and write each
// [rows x N] activation once, instead of the ~40 GEMM / bias / LayerNorm / ReLU launches of the torch modules.
//
//   * a CTA of 128 threads owns 128 rows (thread = row, TMEM lane = row); K is walked in chunks of 32:
//     the thread loads its row's 32 floats (the next chunk's loads are issued before the current chunk is consumed),
//     splits them x = hi + lo (hi = the 19 bits a tf32 keeps) and writes both halves into TENSOR MEMORY with tcgen05.st
//     -- the activations are the A operand from TMEM, they never touch shared memory;
//   * the weights were split and laid out once (ocn_linear_tc_prep: per K chunk, hi then lo, N x 32 in the K-major
//     core-matrix layout the descriptor names), so a chunk is ONE contiguous block of N x 256 bytes that one thread
//     hands to the copy engine (cp.async.bulk -> shared memory, completion on an mbarrier): no thread touches a weight;
//   * tcgen05.mma.cta_group::1.kind::tf32, M = 128, N, K = 8: hi.hi + lo.hi + hi.lo per K step (fp32 accuracy: the
//     dropped lo.lo term is 2^-22 relative), accumulated in TMEM; two A buffers and a ring of 3 - 6 weight chunks, a fifth
//     warp issues the copies and every MMA, so staging chunk c + 1 and the copies of chunks c + 2 .. overlap the MMAs of
//     chunk c (tcgen05.commit -> mbarrier frees a buffer);
//   * the tail runs out of TMEM with tcgen05.ld.32x32b.x32, 32 columns at a time, one thread per row: LayerNorm is two
//     passes over the row's N accumulators, nothing is shuffled.
#include "common.cuh"

namespace ocn {
namespace ltc {

constexpr int kRows = 128;
constexpr int kKc = 32;                        // K chunk (floats)
constexpr uint32_t kLbo = 128, kSbo = 1024;    // core-matrix layout of an [N x 32] chunk: next 16-byte K chunk, next 8 rows

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((kLbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((kSbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100; base offset 0, no swizzle
    return d;
}

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}

// W[N x K] row-major (nn.Linear.weight) -> per K chunk c: [hi | lo], each N x 32 floats at
// (n >> 3) * 256 + (kk >> 2) * 32 + (n & 7) * 4 + (kk & 3)
__global__ void k_linear_tc_prep(const float* __restrict__ w, int n, int k, float* __restrict__ prepped) {
    const int64_t total = (int64_t)n * k, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int row = (int)(e / k), col = (int)(e - (int64_t)row * k);
        const int c = col / kKc, kk = col - c * kKc;
        const float x = __ldg(w + e);
        const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        float* dst = prepped + (size_t)c * 2 * n * kKc + (row >> 3) * 256 + (kk >> 2) * 32 + (row & 7) * 4 + (kk & 3);
        dst[0] = hi;
        dst[(size_t)n * kKc] = x - hi;
    }
}

__device__ __forceinline__ void bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Stages of the weight ring (chunks of N x 256 bytes): N = 256 owns the SM (3 x 64 KB, all 512 tensor-memory columns);
// narrower outputs leave room for two CTAs per SM (<= 96 KB and 256 columns each)
template <int N>
struct Ring {
    static constexpr int kStages = N == 256 ? 3 : ((96 * 1024) / (N * 256) > 6 ? 6 : (96 * 1024) / (N * 256));
    static constexpr int kPerSm = N == 256 ? 1 : 2;
};

// Warps 0-3: one thread per row (stage the A chunk into tensor memory, run the tail).  Warp 4, lane 0: the weight ring (bulk
// copies two to five chunks ahead) and every MMA.  Measured and NOT kept (profiles/r02_ab_head_wide.txt): two threads per
// row (column halves), stores through a shared-memory tile (4 rows x 128 bytes per instruction), every CTA starting K at
// its own chunk -- none moved the layer time; with parts switched off (temporary switches) a 65 536 x 256 x 256 layer is
// 29 us of staging + 20 us of MMAs + 36 us of tail that do not overlap yet (one accumulator, one set of row threads).
// The roles meet on mbarriers only:
//   wfull[s]  copy engine -> MMA issuer   (chunk landed)          wfree[s]  tcgen05.commit -> ring        (slot reusable)
//   aready[b] 128 row threads -> issuer   (A chunk in TMEM)        afree[b]  tcgen05.commit -> row threads (A buffer reusable)
//   dfull     tcgen05.commit -> row threads (tile accumulated)     dfree     128 row threads -> issuer     (tail has read D)
template <int N>
__global__ void __launch_bounds__(kRows + 32)
k_linear_tc(const float* __restrict__ a, int64_t rows, int K, const float* __restrict__ prepped, const float* __restrict__ bias,
            const float* __restrict__ ln_g, const float* __restrict__ ln_b, int relu, float* __restrict__ out,
            float* __restrict__ z, float z_scale, int z_accumulate, const float* __restrict__ wo, const float* __restrict__ bo,
            int out_ch, float* __restrict__ out_final) {
    constexpr int S = Ring<N>::kStages;
    constexpr int kCols = (N + 128 <= 256) ? 256 : 512;       // D (N) + two A buffers of hi 32 + lo 32
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
    constexpr uint32_t kChunkBytes = 2u * N * kKc * 4u;       // hi + lo of one K chunk
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_bar[2 * S + 6];   // wfull[S], wfree[S], aready[2], afree[2], dfull, dfree
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid < 2 * S + 6) {
        const bool many = (tid == 2 * S || tid == 2 * S + 1 || tid == 2 * S + 5);     // aready[0..1], dfree: one arrival per row thread
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&s_bar[tid])), "r"(many ? kRows : 1) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "r"(kCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const uint32_t d_tmem = tmem_base;
    const uint32_t bar0 = smem_addr(&s_bar[0]);
    auto wfull = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto wfree = [&](int s) { return bar0 + 8u * (uint32_t)(S + s); };
    auto aready = [&](int b) { return bar0 + 8u * (uint32_t)(2 * S + b); };
    auto afree = [&](int b) { return bar0 + 8u * (uint32_t)(2 * S + 2 + b); };
    const uint32_t dfull = bar0 + 8u * (uint32_t)(2 * S + 4), dfree = bar0 + 8u * (uint32_t)(2 * S + 5);
    const uint32_t wsm0 = smem_addr(smem_raw);
    const int kch = K / kKc;
    const int64_t ntiles = (rows + kRows - 1) / kRows;
    const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total = my_tiles * kch;     // chunks this CTA consumes, in order

    if (warp == 4) {
        if ((tid & 31) == 0) {
            auto issue = [&](int64_t g) {
                const int s = (int)(g % S), c = (int)(g % kch);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wfull(s)), "r"(kChunkBytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 wsm0 + (uint32_t)s * kChunkBytes),
                             "l"(prepped + (size_t)c * 2 * N * kKc), "r"(kChunkBytes), "r"(wfull(s))
                             : "memory");
            };
            int64_t gi = 0;
            for (; gi < S && gi < total; ++gi) issue(gi);
            for (int64_t g = 0; g < total; ++g) {
                const int s = (int)(g % S), c = (int)(g % kch), b = (int)(g & 1);
                if (c == 0 && g > 0) bar_wait(dfree, (uint32_t)((g / kch - 1) & 1));      // the previous tile's tail has read D
                bar_wait(wfull(s), (uint32_t)((g / S) & 1));
                bar_wait(aready(b), (uint32_t)((g >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi_t = tmem_base + N + 64 * b, a_lo_t = a_hi_t + 32;
                const uint32_t ws = wsm0 + (uint32_t)s * kChunkBytes;
                const uint64_t whi = make_desc(ws), wlo = make_desc(ws + N * kKc * 4u);
#pragma unroll
                for (int j = 0; j < kKc / 8; ++j) mma_ts(d_tmem, a_hi_t + 8 * j, whi + j * 16, kIdesc, (c > 0 || j > 0) ? 1u : 0u);
#pragma unroll
                for (int j = 0; j < kKc / 8; ++j) mma_ts(d_tmem, a_lo_t + 8 * j, whi + j * 16, kIdesc, 1u);
#pragma unroll
                for (int j = 0; j < kKc / 8; ++j) mma_ts(d_tmem, a_hi_t + 8 * j, wlo + j * 16, kIdesc, 1u);
                mma_commit(wfree(s));
                mma_commit(afree(b));
                if (c == kch - 1) mma_commit(dfull);
                // refill the slot of the chunk BEFORE this one (its MMAs are done or about to be): the ring stays S - 1 ahead
                if (g >= 1 && gi < total) {
                    bar_wait(wfree((int)((g - 1) % S)), (uint32_t)(((g - 1) / S) & 1));
                    issue(gi);
                    ++gi;
                }
            }
        }
    } else {
        const uint32_t lanes = (uint32_t)((warp & 3) * 32) << 16;
        int64_t g = 0;
        // 32 consecutive parameters (bias / gamma / beta / a row of Wo) as eight 16-byte loads: every thread of the CTA asks
        // for the same addresses, the loads are broadcasts
        auto load32 = [&](const float* p32, float (&dst)[32]) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p32) + q);
                dst[4 * q] = t.x; dst[4 * q + 1] = t.y; dst[4 * q + 2] = t.z; dst[4 * q + 3] = t.w;
            }
        };
        float bufA[kKc], bufB[kKc];
        auto load_chunk = [&](int64_t tile, int c, float (&dst)[kKc]) {
            const int64_t row = tile * kRows + tid;
            const bool ok = row < rows && tile < ntiles;
            const float4* arow = reinterpret_cast<const float4*>(a + (ok ? row : 0) * K);
#pragma unroll
            for (int q = 0; q < kKc / 4; ++q) {
                const float4 t = ok ? __ldg(arow + c * (kKc / 4) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                dst[4 * q] = t.x; dst[4 * q + 1] = t.y; dst[4 * q + 2] = t.z; dst[4 * q + 3] = t.w;
            }
        };
        // one chunk: registers -> hi / lo -> tensor memory, then tell the issuer
        auto stage = [&](const float (&cur)[kKc]) {
            const int b = (int)(g & 1);
            if (g >= 2) bar_wait(afree(b), (uint32_t)(((g >> 1) - 1) & 1));       // the MMAs that read this A buffer are done
            const uint32_t a_hi_t = tmem_base + N + 64 * b, a_lo_t = a_hi_t + 32;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    hi[j] = __uint_as_float(__float_as_uint(cur[16 * half + j]) & 0xffffe000u);
                    lo[j] = cur[16 * half + j] - hi[j];
                }
                tmem_st16(a_hi_t + lanes + 16 * half, hi);
                tmem_st16(a_lo_t + lanes + 16 * half, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(aready(b));
            ++g;
        };
        load_chunk(blockIdx.x, 0, bufA);
        for (int64_t tile = blockIdx.x, it = 0; tile < ntiles; tile += gridDim.x, ++it) {
            const int64_t row = tile * kRows + tid;
            const bool valid = row < rows;
            // chunks alternate between two register buffers: the loads of chunk c + 1 are issued before chunk c is staged and
            // consumed one whole stage later
            for (int c = 0; c < kch; c += 2) {
                if (c + 1 < kch) load_chunk(tile, c + 1, bufB);
                stage(bufA);
                if (c + 1 < kch) {
                    if (c + 2 < kch) load_chunk(tile, c + 2, bufA);
                    else load_chunk(tile + gridDim.x, 0, bufA);          // the next tile's first chunk, in flight during the tail
                    stage(bufB);
                } else {
                    load_chunk(tile + gridDim.x, 0, bufA);
                }
            }
            bar_wait(dfull, (uint32_t)(it & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // ---- the tail, one thread per row, 32 accumulators at a time
            float mean = 0.f, rstd = 1.f;
            if (ln_g != nullptr) {
                // mean and variance in ONE pass over the accumulators, sums taken about the row's first value (shifted data:
                // no cancellation for rows whose mean is far from zero)
                float sm = 0.f, q = 0.f, shift = 0.f;
                for (int p = 0; p < N / 32; ++p) {
                    float v[32], bb[32];
                    tmem_ld32(d_tmem + lanes + 32 * p, v);
                    load32(bias + 32 * p, bb);
                    if (p == 0) shift = v[0] + bb[0];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { const float d = v[j] + bb[j] - shift; sm += d; q = fmaf(d, d, q); }
                }
                const float md = sm * (1.0f / N);
                mean = shift + md;
                rstd = rsqrtf(fmaxf(q * (1.0f / N) - md * md, 0.f) + 1e-5f);
            }
            float dots[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) dots[o] = 0.f;
            for (int p = 0; p < N / 32; ++p) {
                float v[32], bb[32];
                tmem_ld32(d_tmem + lanes + 32 * p, v);
                if (p == N / 32 - 1) {      // D is read: the issuer may start the next tile while the stores go out
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    bar_arrive(dfree);
                }
                load32(bias + 32 * p, bb);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += bb[j];
                if (ln_g != nullptr) {
                    load32(ln_g + 32 * p, bb);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * bb[j];
                    load32(ln_b + 32 * p, bb);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += bb[j];
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (valid) {
                    if (out != nullptr) {
                        float4* o4 = reinterpret_cast<float4*>(out + row * N + 32 * p);
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) o4[q4] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
                    }
                    if (z != nullptr) {
                        float4* z4 = reinterpret_cast<float4*>(z + row * N + 32 * p);
#pragma unroll
                        for (int q4 = 0; q4 < 8; ++q4) {
                            float4 t = z_accumulate ? z4[q4] : make_float4(0.f, 0.f, 0.f, 0.f);
                            t.x = fmaf(z_scale, v[4 * q4], t.x); t.y = fmaf(z_scale, v[4 * q4 + 1], t.y);
                            t.z = fmaf(z_scale, v[4 * q4 + 2], t.z); t.w = fmaf(z_scale, v[4 * q4 + 3], t.w);
                            z4[q4] = t;
                        }
                    }
                }
                if (wo != nullptr) {
                    for (int o = 0; o < out_ch; ++o) {
                        load32(wo + o * N + 32 * p, bb);
                        float sd = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) sd = fmaf(v[j], bb[j], sd);
                        dots[o] += sd;
                    }
                }
            }
            if (wo != nullptr && valid)
                for (int o = 0; o < out_ch; ++o) out_final[row * out_ch + o] = dots[o] + __ldg(bo + o);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols) : "memory");
    }
}

template <int N>
static int launch(const float* a, int64_t rows, int k, const float* prepped, const float* bias, const float* ln_g, const float* ln_b,
                  int relu, float* out, float* z, float z_scale, int z_acc, const float* wo, const float* bo, int out_ch,
                  float* out_final, cudaStream_t st) {
    const size_t smem = (size_t)Ring<N>::kStages * 2 * N * kKc * sizeof(float);
    OCN_CUDA(cudaFuncSetAttribute(k_linear_tc<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = Ring<N>::kPerSm;
    const int64_t ntiles = (rows + kRows - 1) / kRows;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    k_linear_tc<N><<<(int)(ntiles < cap ? ntiles : cap), kRows + 32, smem, st>>>(a, rows, k, prepped, bias, ln_g, ln_b, relu, out, z, z_scale,
                                                                           z_acc, wo, bo, out_ch, out_final);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // namespace ltc
}  // namespace ocn

using namespace ocn;

extern "C" {

int64_t ocn_linear_tc_prep_floats(int n, int k) {
    if (n <= 0 || k <= 0 || (n != 32 && n != 64 && n != 128 && n != 256) || k % 32 != 0 || k > 1024) return -1;
    return (int64_t)2 * n * k;
}

int ocn_linear_tc_prep(const float* w, int n, int k, float* prepped, void* stream) {
    OCN_CHECK_ARG(w && prepped, "ocn_linear_tc_prep: null pointer");
    OCN_CHECK_ARG(ocn_linear_tc_prep_floats(n, k) > 0, "ocn_linear_tc_prep: out features must be 32 / 64 / 128 / 256 and in features a multiple of 32 (got %d x %d)", n, k);
    const int64_t total = (int64_t)n * k;
    ltc::k_linear_tc_prep<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, n, k, prepped);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_linear_tc(const float* a, int64_t rows, int k, int n, const float* prepped, const float* bias, const float* ln_gamma,
                  const float* ln_beta, int relu, float* out, float* z, float z_scale, int z_accumulate, const float* wo,
                  const float* bo, int out_ch, float* out_final, void* stream) {
    OCN_RANGE("ocn_linear_tc");
    OCN_CHECK_ARG(rows >= 0, "ocn_linear_tc: bad sizes");
    if (rows == 0) return OCN_OK;
    OCN_CHECK_ARG(a && prepped && bias, "ocn_linear_tc: null pointer");
    OCN_CHECK_ARG(ocn_linear_tc_prep_floats(n, k) > 0, "ocn_linear_tc: out features must be 32 / 64 / 128 / 256 and in features a multiple of 32 (got %d x %d)", n, k);
    OCN_CHECK_ARG((ln_gamma == nullptr) == (ln_beta == nullptr), "ocn_linear_tc: LayerNorm needs both gamma and beta");
    OCN_CHECK_ARG(out || z || wo, "ocn_linear_tc: no output requested");
    OCN_CHECK_ARG(wo == nullptr || (bo && out_final && out_ch >= 1 && out_ch <= 8), "ocn_linear_tc: the fused final Linear serves 1..8 outputs");
    OCN_CHECK_ARG((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(prepped) & 15) == 0,
                  "ocn_linear_tc: a and prepped must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    switch (n) {
        case 32: return ltc::launch<32>(a, rows, k, prepped, bias, ln_gamma, ln_beta, relu, out, z, z_scale, z_accumulate, wo, bo, out_ch, out_final, st);
        case 64: return ltc::launch<64>(a, rows, k, prepped, bias, ln_gamma, ln_beta, relu, out, z, z_scale, z_accumulate, wo, bo, out_ch, out_final, st);
        case 128: return ltc::launch<128>(a, rows, k, prepped, bias, ln_gamma, ln_beta, relu, out, z, z_scale, z_accumulate, wo, bo, out_ch, out_final, st);
        default: return ltc::launch<256>(a, rows, k, prepped, bias, ln_gamma, ln_beta, relu, out, z, z_scale, z_accumulate, wo, bo, out_ch, out_final, st);
    }
}

}  // extern "C"
