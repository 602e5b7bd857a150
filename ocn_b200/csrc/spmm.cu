// CSR SpMM family: spmm_add / spmm_mean / spmm_max (torch_sparse.matmul, model.py:6,45-53,2426-2427),
// its transpose-by-scatter backward, and the GCN-normalised aggregation of PureConv / PureConv3 /
// GCNConv (model.py:42-55, 128-142, 58-71).
//
// One warp per output row; a feature row is covered by LPR lanes with VPL float4 each and
// 32/LPR neighbour rows are gathered at once (128-bit coalesced loads), partial sums are
// combined with a fixed butterfly, so results are run-to-run deterministic.
#include <float.h>

#include "common.cuh"

namespace ocn {

enum { kSum = 0, kMean = 1, kMax = 2, kGcnSelf = 3, kGcnNoSelf = 4 };

template <int VPL>
__global__ void __launch_bounds__(256)
k_spmm(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
       const float* __restrict__ norm, int64_t num_rows, const float* __restrict__ x, int nvec, int lpr, int mode,
       float* __restrict__ out) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    const int rpw = 32 / lpr, grp = lane / lpr, sub = lane - grp * lpr;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    const bool is_max = mode == kMax;
    // Software pipeline over the rows of this warp: the row pointers of the row after next and the first 32
    // columns of the next row are in flight while the current row's neighbour rows are gathered, so a row costs
    // one exposed memory latency (the gather) instead of three dependent ones (rowptr -> col -> x).
    auto load_ptr = [&](int64_t rr, int64_t& s, int64_t& e) {
        s = 0; e = 0;
        if (rr < num_rows) { s = ldg_i64(rowptr + rr); e = ldg_i64(rowptr + rr + 1); }
    };
    auto load_cols = [&](int64_t at, int64_t e, int32_t& c, float& w) {
        c = 0; w = 0.f;
        if (at + lane < e) { c = ldg_i32(col + at + lane); w = val ? __ldg(val + at + lane) : 1.0f; }
    };
    int64_t s1, e1, s2, e2;
    int32_t c1;
    float w1;
    load_ptr(warp, s1, e1);
    load_ptr(warp + nwarps, s2, e2);
    load_cols(s1, e1, c1, w1);
    for (int64_t r = warp; r < num_rows; r += nwarps) {
        int64_t s3, e3;
        load_ptr(r + 2 * nwarps, s3, e3);
        int32_t c2;
        float w2;
        load_cols(s2, e2, c2, w2);
        const int64_t s = s1, e = e1;
        const float nr = (mode >= kGcnSelf) ? norm[r] : 1.0f;
        float4 acc[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            acc[v] = is_max ? make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t base = s; base < e; base += 32) {
            int32_t c = c1;
            float w = w1;
            if (base != s) load_cols(base, e, c, w);
            const int cnt = (int)((e - base) < 32 ? (e - base) : 32);
            // kGather neighbour rows per lane group are requested before any of them is consumed: a lane keeps
            // kGather * VPL independent 16-byte loads in flight (one per iteration left the gather latency-bound)
            constexpr int kGather = 2;
            for (int q = 0; q < cnt; q += rpw * kGather) {
                float4 xv[kGather][VPL];
                float ww[kGather], pre[kGather];
                bool on[kGather];
#pragma unroll
                for (int u = 0; u < kGather; ++u) {
                    const int sl = q + u * rpw + grp;
                    on[u] = sl < cnt;
                    const int srcl = on[u] ? sl : 0;
                    const int32_t cc = __shfl_sync(0xffffffffu, c, srcl);
                    ww[u] = __shfl_sync(0xffffffffu, w, srcl);
                    pre[u] = 1.0f;
                    if (on[u]) {
                        if (mode >= kGcnSelf) pre[u] = norm[cc];  // (consumed after the x loads are out: no stall here)
#pragma unroll
                        for (int v = 0; v < VPL; ++v) {
                            const int k = sub + v * lpr;
                            xv[u][v] = (k < nvec) ? __ldg(x4 + (int64_t)cc * nvec + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kGather; ++u) {
                    if (!on[u]) continue;
                    if (mode == kGcnNoSelf) ww[u] = ww[u] * (nr * pre[u]);
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (sub + v * lpr >= nvec) continue;
                        float4 t = xv[u][v];
                        if (mode == kGcnSelf) { t.x *= pre[u]; t.y *= pre[u]; t.z *= pre[u]; t.w *= pre[u]; }
                        if (is_max) {
                            acc[v].x = fmaxf(acc[v].x, ww[u] * t.x); acc[v].y = fmaxf(acc[v].y, ww[u] * t.y);
                            acc[v].z = fmaxf(acc[v].z, ww[u] * t.z); acc[v].w = fmaxf(acc[v].w, ww[u] * t.w);
                        } else {
                            acc[v].x = fmaf(ww[u], t.x, acc[v].x); acc[v].y = fmaf(ww[u], t.y, acc[v].y);
                            acc[v].z = fmaf(ww[u], t.z, acc[v].z); acc[v].w = fmaf(ww[u], t.w, acc[v].w);
                        }
                    }
                }
            }
        }
        for (int o = lpr; o < 32; o <<= 1) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const float ox = __shfl_xor_sync(0xffffffffu, acc[v].x, o), oy = __shfl_xor_sync(0xffffffffu, acc[v].y, o);
                const float oz = __shfl_xor_sync(0xffffffffu, acc[v].z, o), ow = __shfl_xor_sync(0xffffffffu, acc[v].w, o);
                if (is_max) {
                    acc[v].x = fmaxf(acc[v].x, ox); acc[v].y = fmaxf(acc[v].y, oy);
                    acc[v].z = fmaxf(acc[v].z, oz); acc[v].w = fmaxf(acc[v].w, ow);
                } else {
                    acc[v].x += ox; acc[v].y += oy; acc[v].z += oz; acc[v].w += ow;
                }
            }
        }
        if (grp == 0) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int k = sub + v * lpr;
                if (k < nvec) {
                    float4 a = acc[v];
                    if (is_max && e == s) a = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (mode == kMean) {
                        const float inv = 1.0f / (float)((e - s) > 0 ? (e - s) : 1);
                        a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
                    } else if (mode == kGcnSelf) {
                        float4 xs = __ldg(x4 + r * nvec + k);
                        a.x = nr * (a.x + nr * xs.x); a.y = nr * (a.y + nr * xs.y);
                        a.z = nr * (a.z + nr * xs.z); a.w = nr * (a.w + nr * xs.w);
                    }
                    reinterpret_cast<float4*>(out)[r * nvec + k] = a;
                }
            }
        }
        s1 = s2; e1 = e2; c1 = c2; w1 = w2;
        s2 = s3; e2 = e3;
    }
}

// scalar fallback for feature widths that are not a multiple of 4
__global__ void k_spmm_scalar(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                              const float* __restrict__ val, const float* __restrict__ norm, int64_t num_rows,
                              const float* __restrict__ x, int64_t F, int mode, float* __restrict__ out) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (int64_t r = warp; r < num_rows; r += nwarps) {
        const int64_t s = rowptr[r], e = rowptr[r + 1];
        const float nr = (mode >= kGcnSelf) ? norm[r] : 1.0f;
        for (int64_t f = lane; f < F; f += 32) {
            float acc = mode == kMax ? -FLT_MAX : 0.f;
            for (int64_t o = s; o < e; ++o) {
                const int32_t c = col[o];
                float w = val ? val[o] : 1.0f;
                float xv = x[(int64_t)c * F + f];
                if (mode == kGcnSelf) xv *= norm[c];
                else if (mode == kGcnNoSelf) w = w * (nr * norm[c]);
                acc = mode == kMax ? fmaxf(acc, w * xv) : fmaf(w, xv, acc);
            }
            if (mode == kMax && e == s) acc = 0.f;
            if (mode == kMean) acc *= 1.0f / (float)((e - s) > 0 ? (e - s) : 1);
            if (mode == kGcnSelf) acc = nr * (acc + nr * x[r * F + f]);
            out[r * F + f] = acc;
        }
    }
}

__global__ void k_spmm_bwd(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                           const float* __restrict__ val, int64_t num_rows, const float* __restrict__ g, int64_t F,
                           int mode, float* __restrict__ grad_x) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (int64_t r = warp; r < num_rows; r += nwarps) {
        const int64_t s = rowptr[r], e = rowptr[r + 1];
        const float rowscale = mode == kMean ? 1.0f / (float)((e - s) > 0 ? (e - s) : 1) : 1.0f;
        for (int64_t o = s; o < e; ++o) {
            const int32_t c = ldg_i32(col + o);
            const float w = (val ? __ldg(val + o) : 1.0f) * rowscale;
            for (int64_t f = lane; f < F; f += 32) atomicAdd(grad_x + (int64_t)c * F + f, w * g[r * F + f]);
        }
    }
}

// spmm_max backward: the gradient of out[r, f] goes to the FIRST column (in row order) that attains the
// maximum (torch_sparse keeps the first arg-max: its running comparison is a strict ">"); empty rows
// receive nothing.  One warp per row, lanes over the features; the forward value is recomputed.
__global__ void k_spmm_max_bwd(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                               const float* __restrict__ val, int64_t num_rows, const float* __restrict__ x,
                               const float* __restrict__ g, int64_t F, float* __restrict__ grad_x) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (int64_t r = warp; r < num_rows; r += nwarps) {
        const int64_t s = rowptr[r], e = rowptr[r + 1];
        if (e == s) continue;
        for (int64_t f = lane; f < F; f += 32) {
            float best = -FLT_MAX;
            int64_t arg = s;
            for (int64_t o = s; o < e; ++o) {
                const float w = val ? __ldg(val + o) : 1.0f;
                const float v = w * __ldg(x + (int64_t)ldg_i32(col + o) * F + f);
                if (v > best) { best = v; arg = o; }
            }
            const float w = val ? __ldg(val + arg) : 1.0f;
            atomicAdd(grad_x + (int64_t)ldg_i32(col + arg) * F + f, w * g[r * F + f]);
        }
    }
}

__global__ void k_gcn_norm(const int64_t* __restrict__ rowptr, const float* __restrict__ ew, int64_t n,
                           float* __restrict__ out) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (int64_t r = warp; r < n; r += nwarps) {
        const int64_t s = rowptr[r], e = rowptr[r + 1];
        float sum;
        if (ew == nullptr) {
            sum = (float)(e - s);
        } else {
            float a = 0.f;
            for (int64_t o = s + lane; o < e; o += 32) a += ew[o];
            sum = warp_sum(a);
        }
        if (lane == 0) out[r] = rsqrtf(1.0f + sum);
    }
}

// ---- lane-per-feature variant (F = 32 / 64 / 128 / 256) -----------------------------------------------------------------
// k_spmm above covers a feature row with few lanes (float4 each) and gathers several neighbour rows per load
// instruction; at F = 32 that costs shuffles, a butterfly reduction per output row and idle lane groups on short rows
// (ncu, citation2 shape: 21 warp instructions per neighbour row, 42 % issue slots at 33 % occupancy), at F = 128 it
// leaves two loads in flight per lane.  Here a lane owns VEC = F / 32 consecutive features of EVERY neighbour row:
// a neighbour costs one coalesced load instruction per warp (128 x VEC bytes) and VEC FMAs, kU neighbour rows are
// requested before the first is consumed, there is no cross-lane reduction and the output row is one coalesced store.
template <int VEC, int kU>
__global__ void __launch_bounds__(256)
k_spmm_lane(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
            const float* __restrict__ norm, int64_t num_rows, const float* __restrict__ x, int mode, float* __restrict__ out) {
    constexpr int F = 32 * VEC;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    const bool is_max = mode == kMax, gcn = mode >= kGcnSelf;
    auto load_ptr = [&](int64_t rr, int64_t& s, int64_t& e) {
        s = 0; e = 0;
        if (rr < num_rows) { s = ldg_i64(rowptr + rr); e = ldg_i64(rowptr + rr + 1); }
    };
    auto load_cols = [&](int64_t at, int64_t e, int32_t& c, float& w) {
        c = 0; w = 0.f;
        if (at + lane < e) { c = ldg_i32(col + at + lane); w = val ? __ldg(val + at + lane) : 1.0f; }
    };
    auto load_vec = [&](int32_t c, float (&v)[VEC]) {
        const float* p = x + (int64_t)c * F + lane * VEC;
        if (VEC == 1) { v[0] = __ldg(p); }
        else if (VEC == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
        else {
#pragma unroll
            for (int q = 0; q < VEC; q += 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p + q));
                v[q] = t.x; v[q + 1] = t.y; v[q + 2] = t.z; v[q + 3] = t.w;
            }
        }
    };
    int64_t s1, e1, s2, e2;
    int32_t c1;
    float w1;
    load_ptr(warp, s1, e1);
    load_ptr(warp + nwarps, s2, e2);
    load_cols(s1, e1, c1, w1);
    for (int64_t r = warp; r < num_rows; r += nwarps) {
        int64_t s3, e3;
        load_ptr(r + 2 * nwarps, s3, e3);
        int32_t c2;
        float w2;
        load_cols(s2, e2, c2, w2);
        const int64_t s = s1, e = e1;
        const float nr = gcn ? __ldg(norm + r) : 1.0f;
        float self[VEC];
        if (mode == kGcnSelf) load_vec((int32_t)r, self);   // requested now, used after the neighbours
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = is_max ? -FLT_MAX : 0.f;
        for (int64_t base = s; base < e; base += 32) {
            int32_t c = c1;
            float w = w1;
            if (base != s) load_cols(base, e, c, w);
            const int cnt = (int)((e - base) < 32 ? (e - base) : 32);
            const float nc = (gcn && lane < cnt) ? __ldg(norm + c) : 1.0f;   // D^-1/2 of this lane's neighbour: used after the x loads are out
            for (int j = 0; j < cnt; j += kU) {
                float xv[kU][VEC];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int32_t cc = __shfl_sync(0xffffffffu, c, (j + u) & 31);
                    if (j + u < cnt) load_vec(cc, xv[u]);
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const float ww = __shfl_sync(0xffffffffu, w * nc, (j + u) & 31);
                    if (j + u < cnt) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[v] = is_max ? fmaxf(acc[v], ww * xv[u][v]) : fmaf(ww, xv[u][v], acc[v]);
                    }
                }
            }
        }
        float* o = out + r * F + lane * VEC;
        float res[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float a = acc[v];
            if (is_max && e == s) a = 0.f;
            if (mode == kMean) a *= 1.0f / (float)((e - s) > 0 ? (e - s) : 1);
            else if (mode == kGcnSelf) a = nr * (a + nr * self[v]);
            else if (mode == kGcnNoSelf) a = nr * a;
            res[v] = a;
        }
        if (VEC == 1) o[0] = res[0];
        else if (VEC == 2) *reinterpret_cast<float2*>(o) = make_float2(res[0], res[1]);
        else {
#pragma unroll
            for (int q = 0; q < VEC; q += 4) *reinterpret_cast<float4*>(o + q) = make_float4(res[q], res[q + 1], res[q + 2], res[q + 3]);
        }
        s1 = s2; e1 = e2; c1 = c2; w1 = w2;
        s2 = s3; e2 = e3;
    }
}

// ---- TMA-gather variant -------------------------------------------------------------------------------------------
// The gather of neighbour rows is latency-bound on a randomly labelled graph (ncu, citation2 shape, F = 32: DRAM 15 % busy,
// L2 hit 36 %): what limits it is the number of bytes a warp keeps in flight.  Here every lane hands ONE whole neighbour
// row (4 F bytes, 128 B at F = 32) to the copy engine with cp.async.bulk (SASS: UBLKCP) into a per-warp ring of
// kTmaStages x 4 KB in shared memory, completion counted by one mbarrier per stage (expect_tx = rows x 4 F); the warp
// then reduces a landed stage out of shared memory (lane = feature, conflict-free) while the other stages are still in
// flight.  Up to kTmaStages x 4 KB per warp and 12 warps per SM are outstanding without holding a register for them.
// Weights (edge value x D^-1/2 of the neighbour) are fetched into registers at issue time and parked in shared memory
// one iteration later, so no load result is consumed in the iteration that requested it.  The GCN self term rides along
// as one more row of the gather.  Deterministic: a row is reduced by one warp in stage order.
constexpr int kTmaWarps = 4;
constexpr int kTmaStages = 4;
constexpr int kTmaStageBytes = 4096;
constexpr int kTmaWarpBytes = kTmaStages * (kTmaStageBytes + 128 + 32 + 8);  // ring + weights + meta + barriers
struct __align__(16) TmaMeta {
    long long row;
    int cnt, flags;
    float nr;
    int deg, pad0, pad1;
};
static_assert(sizeof(TmaMeta) == 32, "meta slot");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <int VEC>
__global__ void __launch_bounds__(kTmaWarps * 32)
k_spmm_tma(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
           const float* __restrict__ norm, int64_t num_rows, const float* __restrict__ x, int mode, float* __restrict__ out) {
    constexpr int F = 32 * VEC;
    constexpr int RPS = kTmaStageBytes / (F * 4);  // neighbour rows per stage: 32 / 16 / 8 / 4
    constexpr int kFirst = 1, kLast = 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int wib = threadIdx.x >> 5, lane = lane_id();
    unsigned char* base = smem_raw + (size_t)wib * kTmaWarpBytes;
    float* ring = reinterpret_cast<float*>(base);
    float* sw = reinterpret_cast<float*>(base + kTmaStages * kTmaStageBytes);
    TmaMeta* meta = reinterpret_cast<TmaMeta*>(base + kTmaStages * (kTmaStageBytes + 128));
    const uint32_t bars = smem_u32(base + kTmaStages * (kTmaStageBytes + 128 + 32));
    const uint32_t ring_s = smem_u32(ring);
    if (lane == 0) {
        for (int k = 0; k < kTmaStages; ++k) mbar_init(bars + 8 * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t warp = (int64_t)blockIdx.x * kTmaWarps + wib, nwarps = (int64_t)gridDim.x * kTmaWarps;
    const bool gcn = mode >= kGcnSelf;
    const int self = mode == kGcnSelf ? 1 : 0;

    // cursor over (row, chunk) of this warp's rows; the pointers of the following row are prefetched
    int64_t cur = warp, rs = 0, re = 0, ns = 0, ne = 0, pos = 0;
    auto load_ptr = [&](int64_t rr, int64_t& s, int64_t& e) {
        s = 0; e = 0;
        if (rr < num_rows) { s = ldg_i64(rowptr + rr); e = ldg_i64(rowptr + rr + 1); }
    };
    load_ptr(cur, rs, re);
    load_ptr(cur + nwarps, ns, ne);
    pos = rs;
    // P: the chunk that will be issued next (its columns are already requested)
    bool p_valid = false;
    int64_t p_row = 0;
    int p_cnt = 0, p_flags = 0, p_deg = 0;
    int32_t p_c = 0;
    float p_w = 1.0f;
    auto fetch_next = [&]() {
        p_valid = cur < num_rows;
        if (!p_valid) return;
        const int64_t vend = re + self;
        const int64_t left = vend - pos;
        const int cnt = (int)(left < RPS ? left : RPS);
        p_row = cur; p_cnt = cnt; p_deg = (int)(re - rs);
        p_flags = (pos == rs ? kFirst : 0) | (pos + cnt >= vend ? kLast : 0);
        const int64_t idx = pos + lane;
        const bool inrow = lane < cnt && idx < re;
        p_c = lane < cnt ? (inrow ? ldg_i32(col + idx) : (int32_t)cur) : 0;
        p_w = (inrow && val) ? __ldg(val + idx) : 1.0f;
        pos += cnt;
        if (pos >= vend) {
            cur += nwarps; rs = ns; re = ne; pos = rs;
            load_ptr(cur + nwarps, ns, ne);
        }
    };
    // D: weights / meta of the stage issued last, parked in shared memory one iteration later
    bool d_valid = false;
    int d_k = 0, d_cnt = 0, d_flags = 0, d_deg = 0;
    int64_t d_row = 0;
    float d_w = 0.f, d_n = 1.f, d_nr = 1.f;
    auto flush = [&]() {
        if (d_valid) {
            sw[d_k * 32 + lane] = d_w * d_n;
            if (lane == 0) {
                TmaMeta m;
                m.row = d_row; m.cnt = d_cnt; m.flags = d_flags; m.nr = d_nr; m.deg = d_deg; m.pad0 = 0; m.pad1 = 0;
                meta[d_k] = m;
            }
            d_valid = false;
        }
        __syncwarp();
    };
    auto issue = [&](int k) {
        flush();
        d_n = (gcn && lane < p_cnt) ? __ldg(norm + p_c) : 1.0f;
        d_nr = gcn ? __ldg(norm + p_row) : 1.0f;
        if (lane == 0) mbar_expect_tx(bars + 8 * k, (uint32_t)p_cnt * (uint32_t)(F * 4));
        __syncwarp();
        if (lane < p_cnt)
            bulk_g2s(ring_s + (uint32_t)k * kTmaStageBytes + (uint32_t)lane * (F * 4), x + (int64_t)p_c * F, F * 4, bars + 8 * k);
        d_valid = true; d_k = k; d_cnt = p_cnt; d_flags = p_flags; d_deg = p_deg; d_row = p_row; d_w = p_w;
        fetch_next();
    };
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    auto consume = [&](int k, uint32_t parity) {
        mbar_wait(bars + 8 * k, parity);
        const TmaMeta m = meta[k];
        if (m.flags & kFirst) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        }
        const float* st = ring + k * (kTmaStageBytes / 4) + lane * VEC;
        const float* wk = sw + k * 32;
        for (int j = 0; j < m.cnt; ++j) {
            const float wj = wk[j];
            if (VEC == 1) {
                acc[0] = fmaf(wj, st[j * F], acc[0]);
            } else if (VEC == 2) {
                const float2 t = *reinterpret_cast<const float2*>(st + j * F);
                acc[0] = fmaf(wj, t.x, acc[0]); acc[1] = fmaf(wj, t.y, acc[1]);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; v += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(st + j * F + v);
                    acc[v] = fmaf(wj, t.x, acc[v]); acc[v + 1] = fmaf(wj, t.y, acc[v + 1]);
                    acc[v + 2] = fmaf(wj, t.z, acc[v + 2]); acc[v + 3] = fmaf(wj, t.w, acc[v + 3]);
                }
            }
        }
        if (m.flags & kLast) {
            float sc = 1.0f;
            if (mode == kMean) sc = 1.0f / (float)(m.deg > 0 ? m.deg : 1);
            else if (gcn) sc = m.nr;
            float* o = out + m.row * F + lane * VEC;
#pragma unroll
            for (int v = 0; v < VEC; ++v) o[v] = acc[v] * sc;
        }
        __syncwarp();  // every lane is done with the stage before it is handed to the copy engine again
    };
    int64_t issued = 0, consumed = 0;
    fetch_next();
    while (issued < kTmaStages && p_valid) { issue((int)(issued % kTmaStages)); ++issued; }
    while (consumed < issued) {
        if (d_valid && d_k == (int)(consumed % kTmaStages)) flush();  // (only when the newest stage is the next one read)
        consume((int)(consumed % kTmaStages), (uint32_t)((consumed / kTmaStages) & 1));
        ++consumed;
        if (p_valid) { issue((int)(issued % kTmaStages)); ++issued; }
    }
}

// ---- cp.async variant of the ring above ------------------------------------------------------------------------------------
// Same per-warp ring, weights and row order; the copies are per-thread cp.async.cg (16 bytes each, 4 / VEC whole neighbour rows
// per instruction, completion by commit / wait groups) instead of one bulk copy per row: a bulk copy is issued one lane at a
// time (ELECT / R2UR / UBLKCP), which is what kept the ring behind the register gather at F = 32.
template <int VEC>
__global__ void __launch_bounds__(kTmaWarps * 32)
k_spmm_async(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
           const float* __restrict__ norm, int64_t num_rows, const float* __restrict__ x, int mode, float* __restrict__ out) {
    constexpr int F = 32 * VEC;
    constexpr int RPS = kTmaStageBytes / (F * 4);  // neighbour rows per stage: 32 / 16 / 8 / 4
    constexpr int kFirst = 1, kLast = 2;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int wib = threadIdx.x >> 5, lane = lane_id();
    unsigned char* base = smem_raw + (size_t)wib * kTmaWarpBytes;
    float* ring = reinterpret_cast<float*>(base);
    float* sw = reinterpret_cast<float*>(base + kTmaStages * kTmaStageBytes);
    TmaMeta* meta = reinterpret_cast<TmaMeta*>(base + kTmaStages * (kTmaStageBytes + 128));
    const uint32_t ring_s = smem_u32(ring);
    constexpr int kLpr = 8 * VEC;          // 16-byte pieces (lanes) per neighbour row
    const int64_t warp = (int64_t)blockIdx.x * kTmaWarps + wib, nwarps = (int64_t)gridDim.x * kTmaWarps;
    const bool gcn = mode >= kGcnSelf;
    const int self = mode == kGcnSelf ? 1 : 0;

    // cursor over (row, chunk) of this warp's rows; the pointers of the following row are prefetched
    int64_t cur = warp, rs = 0, re = 0, ns = 0, ne = 0, pos = 0;
    auto load_ptr = [&](int64_t rr, int64_t& s, int64_t& e) {
        s = 0; e = 0;
        if (rr < num_rows) { s = ldg_i64(rowptr + rr); e = ldg_i64(rowptr + rr + 1); }
    };
    load_ptr(cur, rs, re);
    load_ptr(cur + nwarps, ns, ne);
    pos = rs;
    // P: the chunk that will be issued next (its columns are already requested)
    bool p_valid = false;
    int64_t p_row = 0;
    int p_cnt = 0, p_flags = 0, p_deg = 0;
    int32_t p_c = 0;
    float p_w = 1.0f;
    auto fetch_next = [&]() {
        p_valid = cur < num_rows;
        if (!p_valid) return;
        const int64_t vend = re + self;
        const int64_t left = vend - pos;
        const int cnt = (int)(left < RPS ? left : RPS);
        p_row = cur; p_cnt = cnt; p_deg = (int)(re - rs);
        p_flags = (pos == rs ? kFirst : 0) | (pos + cnt >= vend ? kLast : 0);
        const int64_t idx = pos + lane;
        const bool inrow = lane < cnt && idx < re;
        p_c = lane < cnt ? (inrow ? ldg_i32(col + idx) : (int32_t)cur) : 0;
        p_w = (inrow && val) ? __ldg(val + idx) : 1.0f;
        pos += cnt;
        if (pos >= vend) {
            cur += nwarps; rs = ns; re = ne; pos = rs;
            load_ptr(cur + nwarps, ns, ne);
        }
    };
    // D: weights / meta of the stage issued last, parked in shared memory one iteration later
    bool d_valid = false;
    int d_k = 0, d_cnt = 0, d_flags = 0, d_deg = 0;
    int64_t d_row = 0;
    float d_w = 0.f, d_n = 1.f, d_nr = 1.f;
    auto flush = [&]() {
        if (d_valid) {
            sw[d_k * 32 + lane] = d_w * d_n;
            if (lane == 0) {
                TmaMeta m;
                m.row = d_row; m.cnt = d_cnt; m.flags = d_flags; m.nr = d_nr; m.deg = d_deg; m.pad0 = 0; m.pad1 = 0;
                meta[d_k] = m;
            }
            d_valid = false;
        }
        __syncwarp();
    };
    auto issue = [&](int k) {
        flush();
        d_n = (gcn && lane < p_cnt) ? __ldg(norm + p_c) : 1.0f;
        d_nr = gcn ? __ldg(norm + p_row) : 1.0f;
        // the stage is 256 pieces of 16 bytes (row r = pieces r * kLpr ..): piece q is copied by lane q % 32 with cp.async --
        // a per-thread copy with nothing to wait for in registers, 4 / VEC whole rows per instruction
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int q = lane + 32 * i, r = q / kLpr, part = q - r * kLpr;
            const int32_t cc = __shfl_sync(0xffffffffu, p_c, r & 31);
            if (r < p_cnt) {
                const float* src = x + (int64_t)cc * F + part * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_s + (uint32_t)k * kTmaStageBytes + (uint32_t)q * 16u),
                             "l"(src)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        d_valid = true; d_k = k; d_cnt = p_cnt; d_flags = p_flags; d_deg = p_deg; d_row = p_row; d_w = p_w;
        fetch_next();
    };
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    auto consume = [&](int k, uint32_t parity) {
        // this lane's copies of the stage are done once at most kTmaStages - 1 newer groups are pending; the warp barrier
        // makes every lane's copies visible to every lane
        asm volatile("cp.async.wait_group %0;" ::"n"(kTmaStages - 1) : "memory");
        __syncwarp();
        const TmaMeta m = meta[k];
        if (m.flags & kFirst) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        }
        const float* st = ring + k * (kTmaStageBytes / 4) + lane * VEC;
        const float* wk = sw + k * 32;
        for (int j = 0; j < m.cnt; ++j) {
            const float wj = wk[j];
            if (VEC == 1) {
                acc[0] = fmaf(wj, st[j * F], acc[0]);
            } else if (VEC == 2) {
                const float2 t = *reinterpret_cast<const float2*>(st + j * F);
                acc[0] = fmaf(wj, t.x, acc[0]); acc[1] = fmaf(wj, t.y, acc[1]);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; v += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(st + j * F + v);
                    acc[v] = fmaf(wj, t.x, acc[v]); acc[v + 1] = fmaf(wj, t.y, acc[v + 1]);
                    acc[v + 2] = fmaf(wj, t.z, acc[v + 2]); acc[v + 3] = fmaf(wj, t.w, acc[v + 3]);
                }
            }
        }
        if (m.flags & kLast) {
            float sc = 1.0f;
            if (mode == kMean) sc = 1.0f / (float)(m.deg > 0 ? m.deg : 1);
            else if (gcn) sc = m.nr;
            float* o = out + m.row * F + lane * VEC;
#pragma unroll
            for (int v = 0; v < VEC; ++v) o[v] = acc[v] * sc;
        }
        __syncwarp();  // every lane is done with the stage before it is handed to the copy engine again
    };
    int64_t issued = 0, consumed = 0;
    fetch_next();
    while (issued < kTmaStages && p_valid) { issue((int)(issued % kTmaStages)); ++issued; }
    for (int64_t pad = issued; pad < kTmaStages; ++pad) asm volatile("cp.async.commit_group;" ::: "memory");   // (group count stays aligned)
    while (consumed < issued) {
        if (d_valid && d_k == (int)(consumed % kTmaStages)) flush();  // (only when the newest stage is the next one read)
        consume((int)(consumed % kTmaStages), 0u);
        ++consumed;
        if (p_valid) { issue((int)(issued % kTmaStages)); ++issued; }
        else asm volatile("cp.async.commit_group;" ::: "memory");       // an empty group keeps "kTmaStages - 1 newer" true
    }
}

template <int VEC>
static int launch_spmm_async(const int64_t* rowptr, const int32_t* col, const float* val, const float* norm, int64_t num_rows,
                             const float* x, int mode, float* out, cudaStream_t st) {
    const size_t smem = (size_t)kTmaWarps * kTmaWarpBytes;
    OCN_CUDA(cudaFuncSetAttribute(k_spmm_async<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)((227 * 1024) / (smem + 1024));
    int64_t want = (num_rows + kTmaWarps - 1) / kTmaWarps;
    const int64_t cap = (int64_t)sm_count() * (per_sm > 0 ? per_sm : 1);
    k_spmm_async<VEC><<<(int)(want < cap ? (want < 1 ? 1 : want) : cap), kTmaWarps * 32, smem, st>>>(rowptr, col, val, norm, num_rows,
                                                                                                     x, mode, out);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

template <int VEC>
static int launch_spmm_tma(const int64_t* rowptr, const int32_t* col, const float* val, const float* norm, int64_t num_rows,
                           const float* x, int mode, float* out, cudaStream_t st) {
    const size_t smem = (size_t)kTmaWarps * kTmaWarpBytes;
    OCN_CUDA(cudaFuncSetAttribute(k_spmm_tma<VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)((227 * 1024) / (smem + 1024));
    int64_t want = (num_rows + kTmaWarps - 1) / kTmaWarps;
    const int64_t cap = (int64_t)sm_count() * (per_sm > 0 ? per_sm : 1);
    k_spmm_tma<VEC><<<(int)(want < cap ? (want < 1 ? 1 : want) : cap), kTmaWarps * 32, smem, st>>>(rowptr, col, val, norm, num_rows,
                                                                                                   x, mode, out);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

static int launch_spmm(const int64_t* rowptr, const int32_t* col, const float* val, const float* norm,
                       int64_t num_rows, const float* x, int64_t feat, int mode, float* out, cudaStream_t st) {
    // OCN_OPT_SPMM_TMA: 0 automatic, 1 bulk (TMA) gather where it applies, 2 register gather, 3 lane-per-feature gather
    // (measured behind both: profiles/r02_ab_spmm_v3.txt), 4 the ring fed by per-thread cp.async (profiles/r02_ab_spmm_v4.txt).  Bulk gather: whole feature rows of 32 / 64 / 128 / 256
    // floats, 16-byte aligned, every reduction but max
    int64_t tma = option(OCN_OPT_SPMM_TMA, 0);
    if (tma == 0)  // automatic: where the one-GPU A/B (profiles/r02_ab_spmm_tma.txt) has the bulk gather ahead
        tma = (num_rows >= 100000 && (mode == kSum || mode == kMean)) ? (feat == 128 ? 1 : ((feat == 32 || feat == 64) ? 4 : 2)) : 2;
    if (tma == 1 && mode != kMax && (feat == 32 || feat == 64 || feat == 128 || feat == 256) &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        if (feat == 32) return launch_spmm_tma<1>(rowptr, col, val, norm, num_rows, x, mode, out, st);
        if (feat == 64) return launch_spmm_tma<2>(rowptr, col, val, norm, num_rows, x, mode, out, st);
        if (feat == 128) return launch_spmm_tma<4>(rowptr, col, val, norm, num_rows, x, mode, out, st);
        return launch_spmm_tma<8>(rowptr, col, val, norm, num_rows, x, mode, out, st);
    }
    int64_t want = (num_rows + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    if (tma == 4 && mode != kMax && (feat == 32 || feat == 64 || feat == 128 || feat == 256) &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        if (feat == 32) return launch_spmm_async<1>(rowptr, col, val, norm, num_rows, x, mode, out, st);
        if (feat == 64) return launch_spmm_async<2>(rowptr, col, val, norm, num_rows, x, mode, out, st);
        if (feat == 128) return launch_spmm_async<4>(rowptr, col, val, norm, num_rows, x, mode, out, st);
        return launch_spmm_async<8>(rowptr, col, val, norm, num_rows, x, mode, out, st);
    }
    if (tma == 3 && (feat == 32 || feat == 64 || feat == 128 || feat == 256) && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        if (feat == 32) k_spmm_lane<1, 8><<<grid, 256, 0, st>>>(rowptr, col, val, norm, num_rows, x, mode, out);
        else if (feat == 64) k_spmm_lane<2, 8><<<grid, 256, 0, st>>>(rowptr, col, val, norm, num_rows, x, mode, out);
        else if (feat == 128) k_spmm_lane<4, 4><<<grid, 256, 0, st>>>(rowptr, col, val, norm, num_rows, x, mode, out);
        else k_spmm_lane<8, 2><<<grid, 256, 0, st>>>(rowptr, col, val, norm, num_rows, x, mode, out);
        OCN_LAUNCH_CHECK();
        return OCN_OK;
    }
    if (feat % 4 != 0 || feat > 1024) {
        k_spmm_scalar<<<grid, 256, 0, st>>>(rowptr, col, val, norm, num_rows, x, feat, mode, out);
    } else {
        const int nvec = (int)(feat / 4);
        int lpr = 1;
        while (lpr < nvec && lpr < 32) lpr <<= 1;
        const int vpl = (nvec + 31) / 32;
        if (vpl <= 1) k_spmm<1><<<grid, 256, 0, st>>>(rowptr, col, val, norm, num_rows, x, nvec, lpr, mode, out);
        else if (vpl <= 2) k_spmm<2><<<grid, 256, 0, st>>>(rowptr, col, val, norm, num_rows, x, nvec, lpr, mode, out);
        else if (vpl <= 4) k_spmm<4><<<grid, 256, 0, st>>>(rowptr, col, val, norm, num_rows, x, nvec, lpr, mode, out);
        else k_spmm<8><<<grid, 256, 0, st>>>(rowptr, col, val, norm, num_rows, x, nvec, lpr, mode, out);
    }
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // namespace ocn

using namespace ocn;

extern "C" {

int ocn_spmm_csr(const int64_t* rowptr, const int32_t* col, const float* val, int64_t num_rows, const float* x,
                 int64_t feat, int reduce, float* out, void* stream) {
    OCN_RANGE("ocn_spmm_csr");
    OCN_CHECK_ARG(rowptr && x && out, "ocn_spmm_csr: null pointer");
    OCN_CHECK_ARG(num_rows >= 0 && feat > 0, "ocn_spmm_csr: bad sizes");
    OCN_CHECK_ARG(reduce >= 0 && reduce <= 2, "ocn_spmm_csr: reduce must be 0 (sum), 1 (mean) or 2 (max)");
    if (num_rows == 0) return OCN_OK;
    return launch_spmm(rowptr, col, val, nullptr, num_rows, x, feat, reduce, out, (cudaStream_t)stream);
}

int ocn_spmm_csr_bwd(const int64_t* rowptr, const int32_t* col, const float* val, int64_t num_rows,
                     const float* grad_out, int64_t feat, int reduce, float* grad_x, void* stream) {
    OCN_RANGE("ocn_spmm_csr_bwd");
    OCN_CHECK_ARG(rowptr && grad_out && grad_x, "ocn_spmm_csr_bwd: null pointer");
    OCN_CHECK_ARG(num_rows >= 0 && feat > 0, "ocn_spmm_csr_bwd: bad sizes");
    OCN_CHECK_ARG(reduce == 0 || reduce == 1, "ocn_spmm_csr_bwd: reduce must be 0 (sum) or 1 (mean)");
    if (num_rows == 0) return OCN_OK;
    int64_t want = (num_rows + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    k_spmm_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(rowptr, col, val, num_rows, grad_out, feat, reduce, grad_x);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_spmm_csr_max_bwd(const int64_t* rowptr, const int32_t* col, const float* val, int64_t num_rows, const float* x,
                         const float* grad_out, int64_t feat, float* grad_x, void* stream) {
    OCN_CHECK_ARG(rowptr && x && grad_out && grad_x, "ocn_spmm_csr_max_bwd: null pointer");
    OCN_CHECK_ARG(num_rows >= 0 && feat > 0, "ocn_spmm_csr_max_bwd: bad sizes");
    if (num_rows == 0) return OCN_OK;
    int64_t want = (num_rows + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    k_spmm_max_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>(rowptr, col, val, num_rows, x, grad_out, feat, grad_x);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_gcn_norm(const int64_t* rowptr, const float* edge_w, int64_t n, float* out_norm, void* stream) {
    OCN_CHECK_ARG(rowptr && out_norm && n >= 0, "ocn_gcn_norm: bad arguments");
    if (n == 0) return OCN_OK;
    int64_t want = (n + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? want : cap);
    k_gcn_norm<<<grid, 256, 0, (cudaStream_t)stream>>>(rowptr, edge_w, n, out_norm);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_gcn_spmm(const int64_t* rowptr, const int32_t* col, const float* edge_w, int64_t n, const float* norm,
                 int mode, const float* x, int64_t feat, float* out, void* stream) {
    OCN_RANGE("ocn_gcn_spmm");
    OCN_CHECK_ARG(rowptr && norm && x && out, "ocn_gcn_spmm: null pointer");
    OCN_CHECK_ARG(mode == kGcnSelf || mode == kGcnNoSelf, "ocn_gcn_spmm: mode must be 3 (self term) or 4 (no self term)");
    OCN_CHECK_ARG(n >= 0 && feat > 0, "ocn_gcn_spmm: bad sizes");
    if (n == 0) return OCN_OK;
    return launch_spmm(rowptr, col, edge_w, norm, n, x, feat, mode, out, (cudaStream_t)stream);
}

}  // extern "C"
