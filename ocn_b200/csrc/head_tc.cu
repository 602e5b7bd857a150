// The predictor head (csrc/head.cu: same operator, same packed parameters) on the 5th-generation tensor cores for
// the width of the large-graph configs, in = hidden = 32 (citation2 / ppa, README.md:53):
//
//   * a CTA runs two independent pipelines of 128 links (4 warps each, thread = link); every Linear of the head is
//     D[128 x 32] = A[128 x 32] . W^T with A written to shared memory by its own threads and D accumulated in TENSOR
//     MEMORY by tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = 32, K = 8 per instruction, issued by one thread);
//   * fp32 accuracy out of tf32 products: every operand is split x = hi + lo (hi = the 19 bits a tf32 keeps, lo = the
//     exact remainder) and three MMAs accumulate hi.hi + lo.hi + hi.lo -- the dropped lo.lo term is 2^-22 relative;
//   * tcgen05.commit arrives on an mbarrier when the 12 MMAs of a layer are done; tcgen05.ld.32x32b.x32 then hands every
//     thread the 32 outputs of ITS link, so bias, ReLU, LayerNorm, the branch mix and the final Linear(32 -> out) are
//     plain register code without a shuffle;
//   * all weight matrices sit in shared memory as hi / lo pairs in the K-major core-matrix layout the descriptors
//     name (8 rows x 16 bytes contiguous; leading byte offset = next 16-byte chunk of K, stride byte offset = next 8 rows).
//
// The CUDA-core kernel (k_cn_head) stays for the other served widths and as the A/B partner (OCN_OPT_HEAD_TC = 2).
#include "common.cuh"

namespace ocn {
namespace htc {

constexpr int kTile = 128;
constexpr int H = 32;
constexpr int kMat = H * H;                    // floats of one matrix
constexpr int kMaxMats = 12;                   // 3 branches x 3 + xijlin 2 + lin 1 (+ 1 with twolayerlin) <= 13: see head_tc_mats
constexpr uint32_t kLbo = 128, kSbo = 1024;    // bytes: next K chunk (16 B x 8 rows), next 8 rows (8 chunks of K = 32 floats)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);

__device__ __forceinline__ int core_off(int row, int k) { return (row >> 3) * (H * 8) + (k >> 2) * 32 + (row & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((kLbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((kSbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version of sm_100
    return d;                // base offset 0, no swizzle
}

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
        : "memory");
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

// offsets (floats, in the packed parameter buffer of csrc/head.cu) of the H x H matrices in compute order; returns how many
__host__ __device__ inline int head_tc_mats(int flags, int branches, int* off, int* vec_end) {
    const bool ln = flags & 1, tailact = flags & 2, two = flags & 4;
    int p = 0, n = 0;
    for (int br = 0; br < branches; ++br) {
        off[n++] = p; p += kMat + H;
        off[n++] = p; p += kMat + H + (ln ? 2 * H : 0);
        off[n++] = p; p += kMat + H;
    }
    off[n++] = p; p += kMat + H + (ln ? 2 * H : 0);
    if (!tailact) { off[n++] = p; p += kMat + H; }
    off[n++] = p; p += kMat + H + (ln ? 2 * H : 0);
    if (two) { off[n++] = p; p += kMat + H + (ln ? 2 * H : 0); }
    *vec_end = p;  // Wo[out, H] and bo[out] follow
    return n;
}

// kGroups independent 128-link pipelines per CTA.  kATmem: the activations are the A operand FROM TENSOR MEMORY (written
// with tcgen05.st, no shared-memory tile, no proxy fence), which leaves room for four pipelines per SM.
template <int kGroups, bool kATmem>
__global__ void __launch_bounds__(kGroups * kTile, 1)
k_cn_head_tc(const float* __restrict__ xcn1, const float* __restrict__ xcn2, const float* __restrict__ xcn3,
             const float* __restrict__ xij, int64_t B, int out_ch, int flags, const float* __restrict__ params,
             const float* __restrict__ mix, float* __restrict__ out) {
    constexpr int kColsPerGroup = kATmem ? 3 * H : H;                    // D (+ A hi, A lo)
    constexpr int kCols = kGroups * kColsPerGroup <= 64 ? 64 : (kGroups * kColsPerGroup <= 128 ? 128 : (kGroups * kColsPerGroup <= 256 ? 256 : 512));
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_bar[kGroups];
    const bool ln = flags & 1, tailact = flags & 2, two = flags & 4;
    const int branches = xcn3 ? 3 : 2;
    int moff[kMaxMats + 1], wo_off;
    const int nmat = head_tc_mats(flags, branches, moff, &wo_off);
    float* Wsm = reinterpret_cast<float*>(smem_raw);                  // [nmat][hi, lo][kMat] in core-matrix layout
    float* Asm = Wsm + (size_t)nmat * 2 * kMat;                       // [group][hi, lo][kTile * H]
    const int tid = threadIdx.x, warp = tid >> 5, g = tid / kTile, tg = tid - g * kTile;

    // ---- one-time setup: weights -> shared memory (W[n][k] = packed W^T[k][n], split), barriers, tensor memory
    // (eight loads in flight per thread: the conversion is latency-bound otherwise -- 16 us per CTA measured)
    for (int e0 = tid; e0 < nmat * kMat; e0 += kGroups * kTile * 8) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * kGroups * kTile;
            w[u] = e < nmat * kMat ? __ldg(params + moff[e / kMat] + (e % kMat)) : 0.f;   // (k * H + n: consecutive threads, consecutive n)
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * kGroups * kTile;
            if (e >= nmat * kMat) break;
            const int m = e / kMat, r = e - m * kMat, k = r / H, n = r - k * H;
            const float hi = __uint_as_float(__float_as_uint(w[u]) & 0xffffe000u);
            float* dst = Wsm + (size_t)m * 2 * kMat + core_off(n, k);
            dst[0] = hi;
            dst[kMat] = w[u] - hi;
        }
    }
    if (tid < kGroups) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&s_bar[tid])), "r"(1) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "r"(kCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const uint32_t d_tmem = tmem_base + (uint32_t)(g * kColsPerGroup);                  // this group's accumulator columns
    const uint32_t lanes = (uint32_t)((warp & 3) * 32) << 16;                           // this warp's 32 lanes
    const uint32_t ld_addr = d_tmem + lanes;
    const uint32_t a_hi_t = d_tmem + H, a_lo_t = d_tmem + 2 * H;                        // (kATmem) the A operand's columns
    float* A_hi = Asm + (kATmem ? (size_t)0 : (size_t)g * 2 * kTile * H);   // (unused with kATmem: no tile is allocated)
    float* A_lo = A_hi + kTile * H;
    // descriptors are built once; a K step (8 tf32 = two 16-byte chunks) and a matrix advance the start-address field only
    const uint64_t a_hi_d = make_desc(smem_addr(A_hi)), a_lo_d = make_desc(smem_addr(A_lo)), w_d0 = make_desc(smem_addr(Wsm));
    constexpr uint64_t kStepD = (2 * kLbo) >> 4, kMatD = (kMat * 4) >> 4;
    const uint32_t bar = smem_addr(&s_bar[g]);
    uint32_t phase = 0;
    const float mixw[4] = {__ldg(mix), __ldg(mix + 1), __ldg(mix + 2), __ldg(mix + 3)};
    const float* ins[4] = {xcn1, xcn2, xcn3, xij};

    // v <- v . W_m^T + bias (bias at params[moff[m] + kMat ..])
    auto linear = [&](float (&v)[32], int m) {
        if (kATmem) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    hi[j] = __uint_as_float(__float_as_uint(v[16 * half + j]) & 0xffffe000u);
                    lo[j] = v[16 * half + j] - hi[j];
                }
                tmem_st16(a_hi_t + lanes + 16 * half, hi);
                tmem_st16(a_lo_t + lanes + 16 * half, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 hi, lo;
                hi.x = __uint_as_float(__float_as_uint(v[4 * c]) & 0xffffe000u);     lo.x = v[4 * c] - hi.x;
                hi.y = __uint_as_float(__float_as_uint(v[4 * c + 1]) & 0xffffe000u); lo.y = v[4 * c + 1] - hi.y;
                hi.z = __uint_as_float(__float_as_uint(v[4 * c + 2]) & 0xffffe000u); lo.z = v[4 * c + 2] - hi.z;
                hi.w = __uint_as_float(__float_as_uint(v[4 * c + 3]) & 0xffffe000u); lo.w = v[4 * c + 3] - hi.w;
                const int o = core_off(tg, 4 * c);
                *reinterpret_cast<float4*>(A_hi + o) = hi;
                *reinterpret_cast<float4*>(A_lo + o) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(kTile) : "memory");
        if (tg == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t whi = w_d0 + (uint64_t)m * 2u * kMatD, wlo = whi + kMatD;
            if (kATmem) {
#pragma unroll
                for (int j = 0; j < H / 8; ++j) mma_tf32_ts(d_tmem, a_hi_t + 8 * j, whi + j * kStepD, j > 0 ? 1u : 0u);
#pragma unroll
                for (int j = 0; j < H / 8; ++j) mma_tf32_ts(d_tmem, a_lo_t + 8 * j, whi + j * kStepD, 1u);
#pragma unroll
                for (int j = 0; j < H / 8; ++j) mma_tf32_ts(d_tmem, a_hi_t + 8 * j, wlo + j * kStepD, 1u);
            } else {
#pragma unroll
                for (int j = 0; j < H / 8; ++j) mma_tf32(d_tmem, a_hi_d + j * kStepD, whi + j * kStepD, j > 0 ? 1u : 0u);
#pragma unroll
                for (int j = 0; j < H / 8; ++j) mma_tf32(d_tmem, a_lo_d + j * kStepD, whi + j * kStepD, 1u);
#pragma unroll
                for (int j = 0; j < H / 8; ++j) mma_tf32(d_tmem, a_hi_d + j * kStepD, wlo + j * kStepD, 1u);
            }
            mma_commit(bar);
        }
        bar_wait(bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_ld32(ld_addr, v);
        const float* b = params + moff[m] + kMat;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __ldg(b + j);
    };
    auto layer_norm = [&](float (&v)[32], const float* gb) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) s += v[j];
        const float mean = s * (1.0f / 32);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
        const float rstd = rsqrtf(q * (1.0f / 32) + 1e-5f);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * __ldg(gb + j) + __ldg(gb + 32 + j);
    };
    auto relu = [&](float (&v)[32]) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    };

    const int64_t ntiles = (B + kTile - 1) / kTile;
    for (int64_t tile = (int64_t)blockIdx.x * kGroups + g; tile < ntiles; tile += (int64_t)gridDim.x * kGroups) {
        const int64_t b = tile * kTile + tg;
        const bool valid = b < B;
        // the rows this thread reads later (the other branches of this link, the first branch of its next link) start
        // their way from DRAM to L2 now: with two warps per scheduler a cold 128-byte row is an exposed microsecond
        {
            const int64_t bn = b + (int64_t)gridDim.x * kGroups * kTile;
#pragma unroll
            for (int br = 0; br < 4; ++br) {
                const int64_t row = br == 0 ? bn : b;
                if (ins[br] != nullptr && row < B)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ins[br] + row * H));
            }
        }
        float z[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = 0.f;
        int m = 0;
#pragma unroll
        for (int br = 0; br < 4; ++br) {
            const float* src = ins[br];
            if (src == nullptr) continue;
            float v[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 t = valid ? __ldg(reinterpret_cast<const float4*>(src + b * H) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
            }
            if (br < 3) {  // CN branch: Linear ReLU Linear (LN) ReLU Linear
                linear(v, m); ++m;
                relu(v);
                linear(v, m);
                if (ln) layer_norm(v, params + moff[m] + kMat + H);
                ++m;
                relu(v);
                linear(v, m); ++m;
            } else {       // xijlin: Linear (LN) ReLU [Linear]
                linear(v, m);
                if (ln) layer_norm(v, params + moff[m] + kMat + H);
                ++m;
                relu(v);
                if (!tailact) { linear(v, m); ++m; }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = fmaf(mixw[br], v[j], z[j]);
        }
        linear(z, m);
        if (ln) layer_norm(z, params + moff[m] + kMat + H);
        ++m;
        relu(z);
        if (two) {
            linear(z, m);
            if (ln) layer_norm(z, params + moff[m] + kMat + H);
            ++m;
            relu(z);
        }
        const float* Wo = params + wo_off;
        const float* bo = Wo + (int64_t)out_ch * H;
        for (int o = 0; o < out_ch; ++o) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) s = fmaf(z[j], __ldg(Wo + o * H + j), s);
            if (valid) out[b * out_ch + o] = s + __ldg(bo + o);
        }
    }
    // ---- teardown: every tensor-memory access is done before the columns are handed back
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols) : "memory");
    }
}

}  // namespace htc

// in = hidden = 32 on the tensor cores; false when the shape is not served (the caller launches k_cn_head)
template <int kGroups, bool kATmem>
static int launch_head_tc_t(const float* xcn1, const float* xcn2, const float* xcn3, const float* xij, int64_t num_links,
                            int out_ch, int flags, const float* params, const float* mix, float* out, cudaStream_t st) {
    using namespace htc;
    int moff[kMaxMats + 1], wo_off;
    const int nmat = head_tc_mats(flags, xcn3 ? 3 : 2, moff, &wo_off);
    const size_t smem = sizeof(float) * ((size_t)nmat * 2 * kMat + (kATmem ? (size_t)0 : (size_t)kGroups * 2 * kTile * H));
    OCN_CUDA(cudaFuncSetAttribute(k_cn_head_tc<kGroups, kATmem>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (num_links + kTile - 1) / kTile;
    int64_t blocks = (ntiles + kGroups - 1) / kGroups;
    if (blocks > sm_count()) blocks = sm_count();
    k_cn_head_tc<kGroups, kATmem><<<(int)blocks, kGroups * kTile, smem, st>>>(xcn1, xcn2, xcn3, xij, num_links, out_ch, flags,
                                                                              params, mix, out);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

// in = hidden = 32 on the tensor cores.  variant 1: two pipelines per SM, activations through shared memory; 3: four
// pipelines, activations through tensor memory
int launch_head_tc(const float* xcn1, const float* xcn2, const float* xcn3, const float* xij, int64_t num_links, int out_ch,
                   int flags, const float* params, const float* mix, float* out, cudaStream_t st, int variant) {
    if (variant == 3)
        return launch_head_tc_t<4, true>(xcn1, xcn2, xcn3, xij, num_links, out_ch, flags, params, mix, out, st);
    return launch_head_tc_t<2, false>(xcn1, xcn2, xcn3, xij, num_links, out_ch, flags, params, mix, out, st);
}

}  // namespace ocn
