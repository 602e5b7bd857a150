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
                        if (mode == kGcnSelf) pre[u] = norm[cc];
                        else if (mode == kGcnNoSelf) ww[u] = ww[u] * (nr * norm[cc]);
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

static int launch_spmm(const int64_t* rowptr, const int32_t* col, const float* val, const float* norm,
                       int64_t num_rows, const float* x, int64_t feat, int mode, float* out, cudaStream_t st) {
    int64_t want = (num_rows + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 16;
    const int grid = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
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
    OCN_CHECK_ARG(rowptr && x && out, "ocn_spmm_csr: null pointer");
    OCN_CHECK_ARG(num_rows >= 0 && feat > 0, "ocn_spmm_csr: bad sizes");
    OCN_CHECK_ARG(reduce >= 0 && reduce <= 2, "ocn_spmm_csr: reduce must be 0 (sum), 1 (mean) or 2 (max)");
    if (num_rows == 0) return OCN_OK;
    return launch_spmm(rowptr, col, val, nullptr, num_rows, x, feat, reduce, out, (cudaStream_t)stream);
}

int ocn_spmm_csr_bwd(const int64_t* rowptr, const int32_t* col, const float* val, int64_t num_rows,
                     const float* grad_out, int64_t feat, int reduce, float* grad_x, void* stream) {
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
    OCN_CHECK_ARG(rowptr && norm && x && out, "ocn_gcn_spmm: null pointer");
    OCN_CHECK_ARG(mode == kGcnSelf || mode == kGcnNoSelf, "ocn_gcn_spmm: mode must be 3 (self term) or 4 (no self term)");
    OCN_CHECK_ARG(n >= 0 && feat > 0, "ocn_gcn_spmm: bad sizes");
    if (n == 0) return OCN_OK;
    return launch_spmm(rowptr, col, edge_w, norm, n, x, feat, mode, out, (cudaStream_t)stream);
}

}  // extern "C"
