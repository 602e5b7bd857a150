// The step AFTER the aggregation: the predictor's MLP head in inference mode,
//
//     out = lin( a0 * xcn1lin(xcn1) + a1 * xcn2lin(xcn2) [+ a2 * xcn3lin(xcn3)] + beta * xijlin(xij) )
//         model.py:2192-2235 (modules), :2429-2437 (cn5), :2935-2950 (cn6), :3216-3226 (cn7)
//
// as ONE kernel for the narrow widths the large-graph configs use (citation2 / ppa: hidden 32 or 64):
// the reference runs ~30 separate [B, F] x [F, F] GEMM / bias / ReLU / LayerNorm launches that each re-read
// and re-write the [B, F] activations; here every parameter (<= 200 KB) sits in shared memory, transposed so
// that lane o reads W[o, k] conflict-free, a warp carries kLinks links through all layers in registers
// (lane o owns hidden feature o, inputs are broadcast with shuffles) and only the four [B, F] inputs are read
// and [B, out] written.  Widths above 64 are GEMM-shaped and stay with torch / cuBLAS (host mirror decides).
//
// Parameter buffer (fp32), packed by ocn_b200/head.py in exactly this order; every weight is stored
// TRANSPOSED ([in, out] row-major), LayerNorm entries are present only with the `ln` flag:
//   for each CN branch (xcn1lin, xcn2lin[, xcn3lin]):  W1t[in,h] b1[h]  W2t[h,h] b2[h] (g[h] beta[h])  W3t[h,h] b3[h]
//   xijlin:                                             W1t[in,h] b1[h] (g[h] beta[h])  (W2t[h,h] b2[h] unless tailact)
//   lin:   W1t[h,h] b1[h] (g beta)  (W2t[h,h] b2[h] (g beta) if twolayerlin)  Wo[out,h] (NOT transposed) bo[out]
#include "common.cuh"

namespace ocn {

constexpr int kLinks = 4;       // links a warp carries at once (register blocking of the weight reads)
constexpr int kHeadThreads = 256;

template <int KL, int HL>
__device__ __forceinline__ void matvec(const float* __restrict__ Wt, const float* __restrict__ b, const float (&x)[kLinks][KL],
                                       float (&y)[kLinks][HL], int lane) {
    constexpr int H = 32 * HL;
#pragma unroll
    for (int q = 0; q < kLinks; ++q)
#pragma unroll
        for (int h = 0; h < HL; ++h) y[q][h] = b[lane + 32 * h];
#pragma unroll
    for (int kl = 0; kl < KL; ++kl) {
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) {
            const int k = kl * 32 + kk;
            float w[HL];
#pragma unroll
            for (int h = 0; h < HL; ++h) w[h] = Wt[k * H + lane + 32 * h];
#pragma unroll
            for (int q = 0; q < kLinks; ++q) {
                const float xk = __shfl_sync(0xffffffffu, x[q][kl], kk);
#pragma unroll
                for (int h = 0; h < HL; ++h) y[q][h] = fmaf(w[h], xk, y[q][h]);
            }
        }
    }
}

template <int HL>
__device__ __forceinline__ void layer_norm(float (&y)[kLinks][HL], const float* __restrict__ g, const float* __restrict__ be,
                                           int lane) {
    constexpr float invH = 1.0f / (32 * HL);
#pragma unroll
    for (int q = 0; q < kLinks; ++q) {
        float s = 0.f;
#pragma unroll
        for (int h = 0; h < HL; ++h) s += y[q][h];
        const float mean = warp_sum(s) * invH;
        float v = 0.f;
#pragma unroll
        for (int h = 0; h < HL; ++h) { const float d = y[q][h] - mean; v = fmaf(d, d, v); }
        const float rstd = rsqrtf(warp_sum(v) * invH + 1e-5f);
#pragma unroll
        for (int h = 0; h < HL; ++h) y[q][h] = (y[q][h] - mean) * rstd * g[lane + 32 * h] + be[lane + 32 * h];
    }
}

template <int HL>
__device__ __forceinline__ void relu(float (&y)[kLinks][HL]) {
#pragma unroll
    for (int q = 0; q < kLinks; ++q)
#pragma unroll
        for (int h = 0; h < HL; ++h) y[q][h] = fmaxf(y[q][h], 0.f);
}

template <int IL, int HL>
__global__ void __launch_bounds__(kHeadThreads)
k_cn_head(const float* __restrict__ xcn1, const float* __restrict__ xcn2, const float* __restrict__ xcn3,
          const float* __restrict__ xij, int64_t B, int out_ch, int flags, const float* __restrict__ params,
          int64_t params_len, const float* __restrict__ mix, float* __restrict__ out) {
    extern __shared__ float sp[];
    constexpr int IN = 32 * IL, H = 32 * HL;
    const bool ln = flags & 1, tailact = flags & 2, two = flags & 4;
    for (int64_t k = threadIdx.x * 4; k < params_len; k += kHeadThreads * 4) {
        if (k + 4 <= params_len) *reinterpret_cast<float4*>(sp + k) = __ldg(reinterpret_cast<const float4*>(params + k));
        else for (int64_t r = k; r < params_len; ++r) sp[r] = params[r];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (kHeadThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kHeadThreads / 32);
    const float mixw[4] = {mix[0], mix[1], mix[2], mix[3]};
    const float* ins[4] = {xcn1, xcn2, xcn3, xij};
    for (int64_t b0 = warp * kLinks; b0 < B; b0 += nwarps * kLinks) {
        float z[kLinks][HL];
#pragma unroll
        for (int q = 0; q < kLinks; ++q)
#pragma unroll
            for (int h = 0; h < HL; ++h) z[q][h] = 0.f;
        const float* p = sp;
#pragma unroll
        for (int br = 0; br < 4; ++br) {
            const float* src = ins[br];
            if (src == nullptr) continue;  // no third CN branch (its parameters are not packed either)
            float x[kLinks][IL];
#pragma unroll
            for (int q = 0; q < kLinks; ++q)
#pragma unroll
                for (int i = 0; i < IL; ++i) x[q][i] = (b0 + q < B) ? __ldg(src + (b0 + q) * IN + lane + 32 * i) : 0.f;
            float a[kLinks][HL], c[kLinks][HL];
            matvec<IL, HL>(p, p + IN * H, x, a, lane);
            p += IN * H + H;
            if (br < 3) {  // CN branch: Linear ReLU Linear (LN) ReLU Linear
                relu<HL>(a);
                matvec<HL, HL>(p, p + H * H, a, c, lane);
                p += H * H + H;
                if (ln) { layer_norm<HL>(c, p, p + H, lane); p += 2 * H; }
                relu<HL>(c);
                matvec<HL, HL>(p, p + H * H, c, a, lane);
                p += H * H + H;
            } else {  // xijlin: Linear (LN) ReLU [Linear]
                if (ln) { layer_norm<HL>(a, p, p + H, lane); p += 2 * H; }
                relu<HL>(a);
                if (!tailact) {
                    matvec<HL, HL>(p, p + H * H, a, c, lane);
                    p += H * H + H;
#pragma unroll
                    for (int q = 0; q < kLinks; ++q)
#pragma unroll
                        for (int h = 0; h < HL; ++h) a[q][h] = c[q][h];
                }
            }
#pragma unroll
            for (int q = 0; q < kLinks; ++q)
#pragma unroll
                for (int h = 0; h < HL; ++h) z[q][h] = fmaf(mixw[br], a[q][h], z[q][h]);
        }
        // lin
        float a[kLinks][HL], c[kLinks][HL];
        matvec<HL, HL>(p, p + H * H, z, a, lane);
        p += H * H + H;
        if (ln) { layer_norm<HL>(a, p, p + H, lane); p += 2 * H; }
        relu<HL>(a);
        if (two) {
            matvec<HL, HL>(p, p + H * H, a, c, lane);
            p += H * H + H;
            if (ln) { layer_norm<HL>(c, p, p + H, lane); p += 2 * H; }
            relu<HL>(c);
#pragma unroll
            for (int q = 0; q < kLinks; ++q)
#pragma unroll
                for (int h = 0; h < HL; ++h) a[q][h] = c[q][h];
        }
        const float* Wo = p;
        const float* bo = p + (int64_t)out_ch * H;
        for (int o = 0; o < out_ch; ++o) {
#pragma unroll
            for (int q = 0; q < kLinks; ++q) {
                float s = 0.f;
#pragma unroll
                for (int h = 0; h < HL; ++h) s = fmaf(a[q][h], Wo[o * H + lane + 32 * h], s);
                s = warp_sum(s);
                if (lane == 0 && b0 + q < B) out[(b0 + q) * out_ch + o] = s + bo[o];
            }
        }
    }
}

int launch_head_tc(const float* xcn1, const float* xcn2, const float* xcn3, const float* xij, int64_t num_links, int out_ch,
                   int flags, const float* params, const float* mix, float* out, cudaStream_t st, int variant);  // head_tc.cu

static int64_t head_params(int in_ch, int hid, int out_ch, int flags, int branches) {
    const bool ln = flags & 1, tailact = flags & 2, two = flags & 4;
    const int64_t I = in_ch, H = hid;
    int64_t n = branches * (I * H + H + 2 * (H * H + H) + (ln ? 2 * H : 0));
    n += I * H + H + (ln ? 2 * H : 0) + (tailact ? 0 : H * H + H);
    n += H * H + H + (ln ? 2 * H : 0) + (two ? H * H + H + (ln ? 2 * H : 0) : 0) + (int64_t)out_ch * H + out_ch;
    return n;
}

}  // namespace ocn

using namespace ocn;

extern "C" {

int64_t ocn_cn_head_params(int in_ch, int hid, int out_ch, int flags, int branches) {
    if (in_ch <= 0 || hid <= 0 || out_ch <= 0 || branches < 2 || branches > 3) return -1;
    if ((in_ch != 32 && in_ch != 64) || (hid != 32 && hid != 64)) return -1;  // wider heads are GEMM-shaped: torch / cuBLAS
    const int64_t n = head_params(in_ch, hid, out_ch, flags, branches);
    return (n * (int64_t)sizeof(float) <= 200 * 1024) ? n : -1;
}

int ocn_cn_head(const float* xcn1, const float* xcn2, const float* xcn3, const float* xij, int64_t num_links, int in_ch,
                int hid, int out_ch, int flags, const float* params, int64_t params_len, const float* mix, float* out,
                void* stream) {
    OCN_RANGE("ocn_cn_head");
    OCN_CHECK_ARG(num_links >= 0, "ocn_cn_head: bad sizes");
    if (num_links == 0) return OCN_OK;
    OCN_CHECK_ARG(xcn1 && xcn2 && xij && params && mix && out, "ocn_cn_head: null pointer");
    const int64_t want = ocn_cn_head_params(in_ch, hid, out_ch, flags, xcn3 ? 3 : 2);
    OCN_CHECK_ARG(want > 0, "ocn_cn_head: widths %d -> %d are not served by the fused head (32 or 64, <= 200 KB of parameters)",
                  in_ch, hid);
    OCN_CHECK_ARG(params_len == want, "ocn_cn_head: parameter buffer holds %lld floats, expected %lld", (long long)params_len,
                  (long long)want);
    cudaStream_t st = (cudaStream_t)stream;
    // in = hidden = 32: tcgen05 kernel (head_tc.cu; OCN_OPT_HEAD_TC 0 / 3: four pipelines per SM, activations through
    // tensor memory; 1: two pipelines, activations through shared memory) unless OCN_OPT_HEAD_TC == 2 (CUDA-core kernel below)
    if (in_ch == 32 && hid == 32 && option(OCN_OPT_HEAD_TC, 3) != 2)
        return launch_head_tc(xcn1, xcn2, xcn3, xij, num_links, out_ch, flags, params, mix, out, st, (int)option(OCN_OPT_HEAD_TC, 3));
    const size_t smem = sizeof(float) * (size_t)((params_len + 3) & ~int64_t(3));
    const int64_t groups = (num_links + kLinks - 1) / kLinks;
    int64_t blocks = (groups + (kHeadThreads / 32) - 1) / (kHeadThreads / 32);
    const int per_sm = smem > 100 * 1024 ? 1 : (smem > 64 * 1024 ? 2 : 3);
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
#define OCN_HEAD_LAUNCH(IL, HL)                                                                                          \
    do {                                                                                                                 \
        OCN_CUDA(cudaFuncSetAttribute(k_cn_head<IL, HL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        k_cn_head<IL, HL><<<(int)blocks, kHeadThreads, smem, st>>>(xcn1, xcn2, xcn3, xij, num_links, out_ch, flags, params, \
                                                                   params_len, mix, out);                              \
    } while (0)
    if (in_ch == 32 && hid == 32) OCN_HEAD_LAUNCH(1, 1);
    else if (in_ch == 64 && hid == 32) OCN_HEAD_LAUNCH(2, 1);
    else if (in_ch == 32 && hid == 64) OCN_HEAD_LAUNCH(1, 2);
    else OCN_HEAD_LAUNCH(2, 2);
#undef OCN_HEAD_LAUNCH
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // extern "C"
