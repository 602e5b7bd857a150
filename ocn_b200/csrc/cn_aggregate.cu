// Batch statistics, orthogonalised weights, CN-indicator SpMM (forward / backward), sparse
// extraction and release of the per-batch column statistics.
//
// Weight algebra (cn5: model.py:2261-2423, order 3: model.py:2546-2933, cn7: model.py:3114-3126).
// For node k of batch b with c1 = #links having k in CN1, S2 = sum_b C2[b,k], S3 = sum_b C3[b,k]:
//   w1(k)      = 1/c1 if c1 >= 2 else fill                   (singletons get `fill`)
//   C1h[e,k]   = w1(k) on CN1(e)
//   scale      = max |C1h| = 1 / min{c1 : c1 >= 2}           (0 if there is none)
//   ipn_x      = ip_x / scale if scale > 0 else ip_x
//   C2'[e,k]   = C2[e,k] - ipn_a * C1h[e,k]     on CN1(e) u CN2(e)
//   c2(k)      = sum_e C2'[e,k] = S2 - (ipn_a * w1(k)) * c1   (c1 >= 2; exact integer part, one
//                rounding for the correction -- the reference adds the per-link terms one by
//                one in fp32, which differs by rounding only)          ; 0 -> 1
//   C2h        = C2' * (1 / c2(k))
//   C3'[e,k]   = C3 - ipn_b * C1h - ipn_c * C2h  on CN1 u CN2 u CN3
//   c3(k)      = S3 - (ipn_b * w1) * c1 - ipn_c * (c2raw * (1/c2))          ; 0 -> 1
//   C3h        = C3' * (1 / c3(k))
#include "common.cuh"
#include "cn_weights.cuh"

namespace ocn {

// ---- batch scalars -------------------------------------------------------------------------
// batch_scalars[b*8 + {0: scale, 1: s12, 2: s13, 3: s23, 4: (u32) min c1>=2}]
__global__ void k_stats_init(float* __restrict__ bscal, int64_t num_batches, int stage) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= num_batches) return;
    if (stage == 0) {
        bscal[b * 8 + 0] = 0.0f;
        bscal[b * 8 + 1] = 0.0f;
        bscal[b * 8 + 2] = 0.0f;
        bscal[b * 8 + 3] = 0.0f;
        reinterpret_cast<uint32_t*>(bscal)[b * 8 + 4] = 0xffffffffu;
    }
}

// one warp per link (a whole CTA per link with a heavy source: per-warp sums combined in warp order, so the
// result stays run-to-run deterministic): min over CN1 members of the column count, and the link's partial sums
__global__ void k_stats(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                        const int64_t* __restrict__ src, int64_t T, int64_t batch_size, int order, int weighted,
                        int variant, float fill, const float* __restrict__ ip, int stage,
                        const int64_t* __restrict__ rec_off, const Record* __restrict__ records,
                        const ColStat* __restrict__ colstat, float* __restrict__ bscal, float* __restrict__ partial,
                        int64_t min_deg /* links whose source has at most this many neighbours were done by cn_grouped.cu */) {
    __shared__ float sh_a[32], sh_b[32];
    __shared__ uint32_t sh_m[32];
    const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int64_t warp = (int64_t)blockIdx.x * wpb + wib;
    int64_t nwarps = (int64_t)gridDim.x * wpb;
    int lane = lane_id();
    // warp-level sums over the 32-position chunks base0, base0 + stride, ... of link t
    auto walk = [&](int64_t t, int64_t rs, int64_t d, int64_t base0, int64_t stride, float& s_a, float& s_b, uint32_t& minc) {
        const int64_t b = t / batch_size, ro = rec_off[t];
        const ColStat* cs = colstat + b * n;
        WeightParams P = make_params(order, weighted, variant, fill, ip, stage == 0 ? nullptr : bscal + b * 8);
        minc = 0xffffffffu;
        s_a = 0.0f;
        s_b = 0.0f;
        for (int64_t base = base0; base < d; base += stride) {
            const int64_t p = base + lane;
            if (p < d) {
                const Record rec = records[ro + p];
                if (rec.x | rec.y) {
                    const int32_t k = ldg_i32(col + rs + p);
                    uint32_t c1;
                    unsigned long long s2, s3;
                    load_colstat(cs + k, c1, s2, s3);
                    const EntryWeights W = entry_weights(rec, c1, s2, s3, P);
                    const uint32_t C2 = rec.x & 0x7fffffffu, C3 = rec.y;
                    if (stage == 0) {
                        if (W.in1 && c1 >= 2u) minc = c1 < minc ? c1 : minc;
                        if (W.in1 && C2) s_a += (weighted ? (float)C2 : 1.0f) * W.w1;
                    } else {
                        if (C3) {
                            const float c3v = weighted ? (float)C3 : 1.0f;
                            if (W.in1) s_a += c3v * W.w1;
                            if (W.in2) s_b += c3v * W.w2;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t other = __shfl_xor_sync(0xffffffffu, minc, o);
            minc = other < minc ? other : minc;
        }
        s_a = warp_sum(s_a);
        s_b = warp_sum(s_b);
    };
    auto store = [&](int64_t t, float s_a, float s_b, uint32_t minc) {
        const int64_t b = t / batch_size;
        if (stage == 0) {
            if (minc != 0xffffffffu) atomicMin(reinterpret_cast<uint32_t*>(bscal) + b * 8 + 4, minc);
            partial[t] = s_a;
        } else {
            partial[T + 1 + t] = s_a;
            partial[2 * (T + 1) + t] = s_b;
        }
    };
    for (int64_t t = warp; t < T; t += nwarps) {
        const int64_t i = src[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        if (d > kHeavyLink || d <= min_deg) continue;
        float s_a, s_b;
        uint32_t minc;
        walk(t, rs, d, 0, 32, s_a, s_b, minc);
        if (lane == 0) store(t, s_a, s_b, minc);
    }
    for_each_heavy_link(rowptr, src, T, rec_off, [&](int64_t t) {
        const int64_t i = src[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        float s_a, s_b;
        uint32_t minc;
        walk(t, rs, d, 32 * (int64_t)wib, 32 * (int64_t)wpb, s_a, s_b, minc);
        if (lane == 0) { sh_a[wib] = s_a; sh_b[wib] = s_b; sh_m[wib] = minc; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float a = 0.0f, c = 0.0f;
            uint32_t m = 0xffffffffu;
            for (int w = 0; w < wpb; ++w) { a += sh_a[w]; c += sh_b[w]; m = sh_m[w] < m ? sh_m[w] : m; }
            store(t, a, c, m);
        }
        __syncthreads();
    });
}

// one CTA per batch: fixed-order tree sum of the link partials (run-to-run deterministic)
__global__ void k_stats_finalize(int64_t T, int64_t batch_size, int stage, const float* __restrict__ partial,
                                 float* __restrict__ bscal) {
    __shared__ float sh[2][256];
    const int64_t b = blockIdx.x;
    const int64_t t0 = b * batch_size, t1 = (t0 + batch_size < T) ? t0 + batch_size : T;
    float a = 0.0f, c = 0.0f;
    for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        if (stage == 0) a += partial[t];
        else { a += partial[T + 1 + t]; c += partial[2 * (T + 1) + t]; }
    }
    sh[0][threadIdx.x] = a;
    sh[1][threadIdx.x] = c;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (stage == 0) {
            const uint32_t minc = reinterpret_cast<const uint32_t*>(bscal)[b * 8 + 4];
            bscal[b * 8 + 0] = (minc == 0xffffffffu) ? 0.0f : __fdiv_rn(1.0f, (float)minc);
            bscal[b * 8 + 1] = sh[0][0];
        } else {
            bscal[b * 8 + 2] = sh[0][0];
            bscal[b * 8 + 3] = sh[1][0];
        }
    }
}

// for_each_heavy_link (common.cuh) with the threshold as an argument: the links whose source has more than `above`
// neighbours among t = blockIdx.x, blockIdx.x + gridDim.x, ..., body(t) called with all the CTA's threads.
template <typename Body>
__device__ __forceinline__ void for_each_link_above(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ src,
                                                    int64_t T, int64_t above, Body&& body) {
    __shared__ int s_n;
    __shared__ long long s_list[256];
    for (int64_t c0 = 0; (int64_t)blockIdx.x + c0 * gridDim.x < T; c0 += blockDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const int64_t t = (int64_t)blockIdx.x + (c0 + threadIdx.x) * gridDim.x;
        if (t < T) {
            const int64_t i = src[t];
            if (rowptr[i + 1] - rowptr[i] > above) s_list[atomicAdd(&s_n, 1)] = t;
        }
        __syncthreads();
        const int n = s_n;
        for (int k = 0; k < n; ++k) body((int64_t)s_list[k]);
    }
}

// ---- CN-indicator SpMM ---------------------------------------------------------------------
// One warp per link.  A feature row is covered by LPR lanes (VPL float4 each); 32/LPR rows are
// gathered at once.  The three weighted sums share one gather of x[k,:] because all three CN
// sets live on N(src).  Summation order is fixed (ascending position), so results are
// run-to-run deterministic.
template <int VPL>
__global__ void __launch_bounds__(256)
k_cn_aggregate(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
               const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t T, int64_t batch_size,
               int order, int weighted, int variant, float fill, const float* __restrict__ ip,
               const int64_t* __restrict__ rec_off, const Record* __restrict__ records,
               const ColStat* __restrict__ colstat, const float* __restrict__ bscal,
               const float* __restrict__ x, int nvec, int lpr,
               float* __restrict__ xcn1, float* __restrict__ xcn2, float* __restrict__ xcn3, float* __restrict__ xij,
               int64_t min_deg, int64_t wide_deg) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    const int rpw = 32 / lpr;          // rows gathered at once
    const int grp = lane / lpr;        // which of them this lane works on
    const int sub = lane - grp * lpr;  // position inside the row
    const int64_t F = (int64_t)nvec * 4;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    // weighted sums over the 32-position chunks base0, base0 + stride, ... of link t; afterwards the lanes of row
    // group 0 hold the sums (fixed butterfly order)
    auto walk = [&](int64_t t, int64_t rs, int64_t d, int64_t base0, int64_t stride, float4 (&a1)[VPL], float4 (&a2)[VPL],
                    float4 (&a3)[VPL]) {
        const int64_t b = t / batch_size, ro = rec_off[t];
        const ColStat* cs = colstat + b * n;
        const WeightParams P = make_params(order, weighted, variant, fill, ip, bscal + b * 8);
#pragma unroll
        for (int v = 0; v < VPL; ++v) a1[v] = a2[v] = a3[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t base = base0; base < d; base += stride) {
            const int64_t p = base + lane;
            int32_t k = 0;
            float w1 = 0.f, w2 = 0.f, w3 = 0.f;
            bool nz = false;
            if (p < d) {
                const Record rec = records[ro + p];
                if (rec.x | rec.y) {
                    k = ldg_i32(col + rs + p);
                    uint32_t c1;
                    unsigned long long s2, s3;
                    load_colstat(cs + k, c1, s2, s3);
                    const EntryWeights W = entry_weights(rec, c1, s2, s3, P);
                    nz = W.in1 || W.in2 || W.in3;
                    w1 = W.w1; w2 = W.w2; w3 = W.w3;
                }
            }
            unsigned active = __ballot_sync(0xffffffffu, nz);
            // kU rounds of row gathers are requested before the first is consumed (a round = one row per lane group): with a
            // whole warp per 1 KB row (F = 256) one round at a time left a single row in flight per warp -- the collab-shape
            // aggregate ran at an eighth of its bytes' worth.  The sums are still taken in ascending position order.
            constexpr int kU = VPL == 2 ? 4 : (VPL == 1 ? 2 : 1);
            while (active) {
                int sls[kU];
                float u1s[kU], u2s[kU], u3s[kU];
                float4 xv[kU][VPL];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    unsigned tmp = active;
                    for (int q = 0; q < grp; ++q) tmp &= tmp - 1;
                    const int sl = tmp ? (__ffs(tmp) - 1) : -1;
                    for (int q = 0; q < rpw && active; ++q) active &= active - 1;
                    const int srcl = sl < 0 ? 0 : sl;
                    const int32_t kk = __shfl_sync(0xffffffffu, k, srcl);
                    u1s[u] = __shfl_sync(0xffffffffu, w1, srcl);
                    u2s[u] = __shfl_sync(0xffffffffu, w2, srcl);
                    u3s[u] = __shfl_sync(0xffffffffu, w3, srcl);
                    sls[u] = sl;
                    if (sl >= 0) {
#pragma unroll
                        for (int v = 0; v < VPL; ++v) {
                            const int c = sub + v * lpr;
                            xv[u][v] = (c < nvec) ? __ldg(x4 + (int64_t)kk * nvec + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    if (sls[u] < 0) continue;
                    const float u1 = u1s[u], u2 = u2s[u], u3 = u3s[u];
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (sub + v * lpr >= nvec) continue;
                        const float4 t4 = xv[u][v];
                        a1[v].x = fmaf(u1, t4.x, a1[v].x); a1[v].y = fmaf(u1, t4.y, a1[v].y);
                        a1[v].z = fmaf(u1, t4.z, a1[v].z); a1[v].w = fmaf(u1, t4.w, a1[v].w);
                        a2[v].x = fmaf(u2, t4.x, a2[v].x); a2[v].y = fmaf(u2, t4.y, a2[v].y);
                        a2[v].z = fmaf(u2, t4.z, a2[v].z); a2[v].w = fmaf(u2, t4.w, a2[v].w);
                        a3[v].x = fmaf(u3, t4.x, a3[v].x); a3[v].y = fmaf(u3, t4.y, a3[v].y);
                        a3[v].z = fmaf(u3, t4.z, a3[v].z); a3[v].w = fmaf(u3, t4.w, a3[v].w);
                    }
                }
            }
        }
        // combine the row groups (fixed butterfly order)
        for (int o = lpr; o < 32; o <<= 1) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                a1[v].x += __shfl_xor_sync(0xffffffffu, a1[v].x, o); a1[v].y += __shfl_xor_sync(0xffffffffu, a1[v].y, o);
                a1[v].z += __shfl_xor_sync(0xffffffffu, a1[v].z, o); a1[v].w += __shfl_xor_sync(0xffffffffu, a1[v].w, o);
                a2[v].x += __shfl_xor_sync(0xffffffffu, a2[v].x, o); a2[v].y += __shfl_xor_sync(0xffffffffu, a2[v].y, o);
                a2[v].z += __shfl_xor_sync(0xffffffffu, a2[v].z, o); a2[v].w += __shfl_xor_sync(0xffffffffu, a2[v].w, o);
                a3[v].x += __shfl_xor_sync(0xffffffffu, a3[v].x, o); a3[v].y += __shfl_xor_sync(0xffffffffu, a3[v].y, o);
                a3[v].z += __shfl_xor_sync(0xffffffffu, a3[v].z, o); a3[v].w += __shfl_xor_sync(0xffffffffu, a3[v].w, o);
            }
        }
    };
    // accumulate == false: the sums become the output rows; true: they are added to them (same lane, same address)
    auto emit = [&](int64_t t, const float4 (&a1)[VPL], const float4 (&a2)[VPL], const float4 (&a3)[VPL], bool accumulate) {
        if (grp != 0) return;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = sub + v * lpr;
            if (c < nvec) {
                float4* o1 = reinterpret_cast<float4*>(xcn1 + t * F) + c;
                float4* o2 = xcn2 ? reinterpret_cast<float4*>(xcn2 + t * F) + c : nullptr;
                float4* o3 = xcn3 ? reinterpret_cast<float4*>(xcn3 + t * F) + c : nullptr;
                float4 r1 = a1[v], r2 = a2[v], r3 = a3[v];
                if (accumulate) {
                    const float4 p1 = *o1;
                    r1 = make_float4(p1.x + r1.x, p1.y + r1.y, p1.z + r1.z, p1.w + r1.w);
                    if (o2) { const float4 p2 = *o2; r2 = make_float4(p2.x + r2.x, p2.y + r2.y, p2.z + r2.z, p2.w + r2.w); }
                    if (o3) { const float4 p3 = *o3; r3 = make_float4(p3.x + r3.x, p3.y + r3.y, p3.z + r3.z, p3.w + r3.w); }
                }
                *o1 = r1;
                if (o2) *o2 = r2;
                if (o3) *o3 = r3;
            }
        }
    };
    auto pair_term = [&](int64_t t, int64_t i, int64_t j) {
        if (!xij) return;
        for (int c = lane; c < nvec; c += 32) {
            const float4 xa = __ldg(x4 + i * nvec + c), xb = __ldg(x4 + j * nvec + c);
            reinterpret_cast<float4*>(xij + t * F)[c] = make_float4(xa.x * xb.x, xa.y * xb.y, xa.z * xb.z, xa.w * xb.w);
        }
    };
    for (int64_t t = warp; t < T; t += nwarps) {  // one warp per link ...
        const int64_t i = src[t], j = dst[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        if (d > wide_deg || d <= min_deg) continue;
        float4 a1[VPL], a2[VPL], a3[VPL];
        walk(t, rs, d, 0, 32, a1, a2, a3);
        emit(t, a1, a2, a3, false);
        pair_term(t, i, j);
    }
    // ... a whole CTA per link with a wide source (more than wide_deg neighbours: a warp gathers 32 / lpr rows at a time,
    // so at F >= 128 a few links of some hundred neighbours each set the kernel's duration when a warp walks them alone):
    // every warp sums its share of the chunks; the shares are added in warp order (run-to-run deterministic), through
    // shared memory up to F = 256 and into the output rows one warp after the other beyond that
    constexpr bool kSmemCombine = VPL <= 2;
    __shared__ float4 s_part[kSmemCombine ? 8 : 1][3][kSmemCombine ? VPL * 32 : 1];
    for_each_link_above(rowptr, src, T, wide_deg > min_deg ? wide_deg : min_deg, [&](int64_t t) {
        const int64_t i = src[t], j = dst[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        float4 a1[VPL], a2[VPL], a3[VPL];
        walk(t, rs, d, 32 * (int64_t)wib, 32 * (int64_t)wpb, a1, a2, a3);
        if constexpr (kSmemCombine) {
            if (grp == 0) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int c = sub + v * lpr;
                    if (c < nvec) { s_part[wib][0][c] = a1[v]; s_part[wib][1][c] = a2[v]; s_part[wib][2][c] = a3[v]; }
                }
            }
            __syncthreads();
            for (int e = threadIdx.x; e < 3 * nvec; e += blockDim.x) {
                const int which = e / nvec, c = e - which * nvec;
                float* o = which == 0 ? xcn1 : (which == 1 ? xcn2 : xcn3);
                if (!o) continue;
                float4 r = s_part[0][which][c];
                for (int w = 1; w < wpb; ++w) {
                    const float4 q = s_part[w][which][c];
                    r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
                }
                reinterpret_cast<float4*>(o + t * F)[c] = r;
            }
            __syncthreads();
        } else {
            for (int w = 0; w < wpb; ++w) {
                if (wib == w) emit(t, a1, a2, a3, w != 0);
                __syncthreads();
            }
        }
        if (wib == 0) pair_term(t, i, j);
    });
}

// backward: grad_x[k,:] += w1*g1[t,:] + w2*g2[t,:] + w3*g3[t,:]; pair term by the product rule.
__global__ void __launch_bounds__(256)
k_cn_aggregate_bwd(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                   const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t T, int64_t batch_size,
                   int order, int weighted, int variant, float fill, const float* __restrict__ ip,
                   const int64_t* __restrict__ rec_off, const Record* __restrict__ records,
                   const ColStat* __restrict__ colstat, const float* __restrict__ bscal,
                   const float* __restrict__ x, int64_t F,
                   const float* __restrict__ g1, const float* __restrict__ g2, const float* __restrict__ g3,
                   const float* __restrict__ gij, float* __restrict__ grad_x) {
    const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int64_t warp = (int64_t)blockIdx.x * wpb + wib;
    int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int lane = lane_id();
    const bool vec = (F % 4 == 0) && F <= 128;  // 16-byte vector reductions (rows of grad_x and of g1..g3 are 16-byte aligned)
    const int nvec = (int)(F / 4);
    int lpr = 1;
    while (lpr < nvec && lpr < 32) lpr <<= 1;
    const int rpw = 32 / lpr, grp = lane / lpr, sub = lane - grp * lpr;
    // 32-position chunks base0, base0 + stride, ... of link t (all sums are atomic: any split of the chunks is fine)
    auto walk = [&](int64_t t, int64_t i, int64_t j, int64_t rs, int64_t d, int64_t base0, int64_t stride, bool pair_term) {
        const int64_t b = t / batch_size, ro = rec_off[t];
        const ColStat* cs = colstat + b * n;
        const WeightParams P = make_params(order, weighted, variant, fill, ip, bscal + b * 8);
        float4 G1 = make_float4(0.f, 0.f, 0.f, 0.f), G2 = G1, G3 = G1;  // this lane's slice of the link's output gradients
        if (vec && sub < nvec) {
            if (g1) G1 = __ldg(reinterpret_cast<const float4*>(g1) + t * nvec + sub);
            if (g2) G2 = __ldg(reinterpret_cast<const float4*>(g2) + t * nvec + sub);
            if (g3) G3 = __ldg(reinterpret_cast<const float4*>(g3) + t * nvec + sub);
        }
        for (int64_t base = base0; base < d; base += stride) {
            const int64_t p = base + lane;
            int32_t k = 0;
            float w1 = 0.f, w2 = 0.f, w3 = 0.f;
            bool nz = false;
            if (p < d) {
                const Record rec = records[ro + p];
                if (rec.x | rec.y) {
                    k = ldg_i32(col + rs + p);
                    uint32_t c1;
                    unsigned long long s2, s3;
                    load_colstat(cs + k, c1, s2, s3);
                    const EntryWeights W = entry_weights(rec, c1, s2, s3, P);
                    nz = W.in1 || W.in2 || W.in3;
                    w1 = W.w1; w2 = W.w2; w3 = W.w3;
                }
            }
            unsigned active = __ballot_sync(0xffffffffu, nz);
            if (vec) {
                // F / 4 <= 32: a row of grad_x is covered by lpr lanes with one 16-byte vector reduction each, and
                // 32 / lpr records are scattered at once (4x fewer reductions than one float per lane)
                while (active) {
                    unsigned tmp = active;
                    for (int q = 0; q < grp; ++q) tmp &= tmp - 1;
                    const int sl = tmp ? (__ffs(tmp) - 1) : -1;
                    for (int q = 0; q < rpw && active; ++q) active &= active - 1;
                    const int srcl = sl < 0 ? 0 : sl;
                    const int32_t kk = __shfl_sync(0xffffffffu, k, srcl);
                    const float u1 = __shfl_sync(0xffffffffu, w1, srcl);
                    const float u2 = __shfl_sync(0xffffffffu, w2, srcl);
                    const float u3 = __shfl_sync(0xffffffffu, w3, srcl);
                    if (sl >= 0 && sub < nvec) {
                        float4 g;
                        g.x = fmaf(u1, G1.x, fmaf(u2, G2.x, u3 * G3.x));
                        g.y = fmaf(u1, G1.y, fmaf(u2, G2.y, u3 * G3.y));
                        g.z = fmaf(u1, G1.z, fmaf(u2, G2.z, u3 * G3.z));
                        g.w = fmaf(u1, G1.w, fmaf(u2, G2.w, u3 * G3.w));
                        if (g.x != 0.f || g.y != 0.f || g.z != 0.f || g.w != 0.f)
                            atomicAdd(reinterpret_cast<float4*>(grad_x) + (int64_t)kk * nvec + sub, g);
                    }
                }
                continue;
            }
            while (active) {
                const int sl = __ffs(active) - 1;
                active &= active - 1;
                const int32_t kk = __shfl_sync(0xffffffffu, k, sl);
                const float u1 = __shfl_sync(0xffffffffu, w1, sl);
                const float u2 = __shfl_sync(0xffffffffu, w2, sl);
                const float u3 = __shfl_sync(0xffffffffu, w3, sl);
                for (int64_t c = lane; c < F; c += 32) {
                    float g = 0.f;
                    if (g1) g = fmaf(u1, g1[t * F + c], g);
                    if (g2) g = fmaf(u2, g2[t * F + c], g);
                    if (g3) g = fmaf(u3, g3[t * F + c], g);
                    if (g != 0.f) atomicAdd(grad_x + (int64_t)kk * F + c, g);
                }
            }
        }
        if (gij && pair_term) {
            for (int64_t c = lane; c < F; c += 32) {
                const float g = gij[t * F + c];
                atomicAdd(grad_x + i * F + c, g * x[j * F + c]);
                atomicAdd(grad_x + j * F + c, g * x[i * F + c]);
            }
        }
    };
    for (int64_t t = warp; t < T; t += nwarps) {  // one warp per link ...
        const int64_t i = src[t], j = dst[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        if (d <= kHeavyLink) walk(t, i, j, rs, d, 0, 32, true);
    }
    for_each_heavy_link(rowptr, src, T, rec_off, [&](int64_t t) {  // ... a whole CTA per link with a heavy source
        const int64_t i = src[t], j = dst[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        walk(t, i, j, rs, d, 32 * (int64_t)wib, 32 * (int64_t)wpb, wib == 0);
    });
}

// ---- sparse extraction ---------------------------------------------------------------------
__device__ __forceinline__ bool in_which(Record rec, int which) {
    const bool has1 = (rec.x >> 31) != 0u;
    const bool has2 = (rec.x & 0x7fffffffu) != 0u, has3 = rec.y != 0u;
    switch (which) {
        case 1: case 11: return has1;
        case 2: return has2;
        case 3: return has3;
        case 12: return has1 || has2;
        case 13: return has1 || has2 || has3;
    }
    return false;
}

template <bool FILL>
__global__ void k_cn_extract(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                             const int64_t* __restrict__ src, int64_t T, int64_t batch_size, int which, int order,
                             int weighted, int variant, float fill, const float* __restrict__ ip,
                             const int64_t* __restrict__ rec_off, const Record* __restrict__ records,
                             const ColStat* __restrict__ colstat, const float* __restrict__ bscal,
                             int64_t* __restrict__ out_counts, const int64_t* __restrict__ out_rowptr,
                             int64_t* __restrict__ out_col, float* __restrict__ out_val) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (int64_t t = warp; t < T; t += nwarps) {
        const int64_t b = t / batch_size;
        const int64_t i = src[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs, ro = rec_off[t];
        WeightParams P;
        if (FILL && which > 10) P = make_params(order, weighted, variant, fill, ip, bscal + b * 8);
        int64_t count = 0;
        const int64_t obase = FILL ? out_rowptr[t] : 0;
        for (int64_t base = 0; base < d; base += 32) {
            const int64_t p = base + lane;
            Record rec = make_uint2(0u, 0u);
            if (p < d) rec = records[ro + p];
            const bool hit = in_which(rec, which);
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (FILL && hit) {
                const int64_t o = obase + count + __popc(m & ((1u << lane) - 1));
                const int32_t k = ldg_i32(col + rs + p);
                out_col[o] = k;
                if (out_val) {
                    float v;
                    if (which == 1) v = 1.0f;
                    else if (which == 2) v = weighted ? (float)(rec.x & 0x7fffffffu) : 1.0f;
                    else if (which == 3) v = weighted ? (float)rec.y : 1.0f;
                    else {
                        uint32_t c1;
                        unsigned long long s2, s3;
                        load_colstat(colstat + b * n + k, c1, s2, s3);
                        const EntryWeights W = entry_weights(rec, c1, s2, s3, P);
                        v = which == 11 ? W.w1 : (which == 12 ? W.w2 : W.w3);
                    }
                    out_val[o] = v;
                }
            }
            count += __popc(m);
        }
        if (!FILL && lane == 0) out_counts[t] = count;
    }
}

// ---- release -------------------------------------------------------------------------------
__global__ void k_cn_release(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                             const int64_t* __restrict__ src, int64_t T, int64_t batch_size,
                             const int64_t* __restrict__ rec_off, const Record* __restrict__ records,
                             ColStat* __restrict__ colstat) {
    const int wpb = blockDim.x >> 5;
    int64_t warp = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int lane = lane_id();
    auto walk = [&](int64_t t, int64_t rs, int64_t d, int64_t p0, int64_t stride) {
        const int64_t b = t / batch_size, ro = rec_off[t];
        for (int64_t p = p0; p < d; p += stride) {
            const Record rec = records[ro + p];
            if (rec.x | rec.y) {
                uint4* c = reinterpret_cast<uint4*>(colstat + b * n + ldg_i32(col + rs + p));
                c[0] = make_uint4(0u, 0u, 0u, 0u);
                c[1] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    };
    for (int64_t t = warp; t < T; t += nwarps) {  // one warp per link ...
        const int64_t i = src[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        if (d <= kHeavyLink) walk(t, rs, d, lane, 32);
    }
    for_each_heavy_link(rowptr, src, T, rec_off, [&](int64_t t) {  // ... a whole CTA per link with a heavy source
        const int64_t i = src[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        walk(t, rs, d, threadIdx.x, blockDim.x);
    });
}

static int grid_for_warps(int64_t items) {
    int64_t want = (items + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace ocn

using namespace ocn;

#define PLAN_PTRS()                                                        \
    PlanLayout L = plan_layout(num_edges);                                 \
    const char* pbase = (const char*)plan_scratch;                         \
    const int64_t* rec_off = (const int64_t*)(pbase + L.rec_off);          \
    float* partial = (float*)(const_cast<char*>(pbase) + L.partial);       \
    (void)partial; (void)rec_off;

extern "C" {

int ocn_cn_stats(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, int64_t num_edges,
                 int64_t batch_size, int order, int weighted, int variant, float fill, const float* ip, int stage,
                 const void* plan_scratch, const void* records, const void* colstat, float* batch_scalars,
                 const int64_t* plan_host, void* stream) {
    OCN_RANGE("ocn_cn_stats");
    OCN_CHECK_ARG(rowptr && col && src && plan_scratch && colstat && batch_scalars, "ocn_cn_stats: null pointer");
    OCN_CHECK_ARG(num_edges > 0 && batch_size > 0 && n > 0, "ocn_cn_stats: sizes must be positive");
    OCN_CHECK_ARG(stage == 0 || stage == 1, "ocn_cn_stats: stage must be 0 or 1");
    OCN_CHECK_ARG(variant == 5 || variant == 7, "ocn_cn_stats: variant must be 5 or 7");
    cudaStream_t st = (cudaStream_t)stream;
    PLAN_PTRS();
    const int64_t nb = (num_edges + batch_size - 1) / batch_size;
    k_stats_init<<<(int)((nb + 255) / 256), 256, 0, st>>>(batch_scalars, nb, stage);
    OCN_LAUNCH_CHECK();
    const bool grouped = use_grouped(num_edges, plan_host);
    if (grouped)
        if (int rc = grouped_stats(rowptr, col, n, src, num_edges, batch_size, order, weighted, variant, fill, ip, stage,
                                   plan_scratch, (const Record*)records, (const ColStat*)colstat, batch_scalars, partial, st))
            return rc;
    if (!grouped || plan_host[OCN_PLAN_WIDE_LINKS] > 0) {  // every link, or the ones the grouped kernel left out
        k_stats<<<grid_for_warps(num_edges), 256, 0, st>>>(rowptr, col, n, src, num_edges, batch_size, order, weighted,
                                                           variant, fill, ip, stage, rec_off, (const Record*)records,
                                                           (const ColStat*)colstat, batch_scalars, partial,
                                                           grouped ? (int64_t)kGroupedMaxDeg : (int64_t)-1);
        OCN_LAUNCH_CHECK();
    }
    k_stats_finalize<<<(int)nb, 256, 0, st>>>(num_edges, batch_size, stage, partial, batch_scalars);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_cn_aggregate(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, const int64_t* dst,
                     int64_t num_edges, int64_t batch_size, int order, int weighted, int variant, float fill,
                     const float* ip, const void* plan_scratch, const void* records, const void* colstat,
                     const float* batch_scalars, const float* x, int64_t feat, float* xcn1, float* xcn2, float* xcn3,
                     float* xij, const int64_t* plan_host, void* stream) {
    OCN_RANGE("ocn_cn_aggregate");
    OCN_CHECK_ARG(rowptr && col && src && dst && plan_scratch && colstat && batch_scalars && x && xcn1,
                  "ocn_cn_aggregate: null pointer");
    OCN_CHECK_ARG(num_edges > 0 && batch_size > 0 && n > 0, "ocn_cn_aggregate: sizes must be positive");
    OCN_CHECK_ARG(variant == 5 || variant == 7, "ocn_cn_aggregate: variant must be 5 (cn5/cn6) or 7 (cn7)");
    OCN_CHECK_ARG(order >= 1 && order <= 3, "ocn_cn_aggregate: order must be 1..3");
    OCN_CHECK_ARG(feat > 0 && feat % 4 == 0 && feat <= 1024,
                  "ocn_cn_aggregate: feature width must be a multiple of 4 and <= 1024 (got %lld)", (long long)feat);
    cudaStream_t st = (cudaStream_t)stream;
    PLAN_PTRS();
    const int nvec = (int)(feat / 4);
    const bool grouped = use_grouped(num_edges, plan_host) && grouped_aggregate_smem(nvec) <= 96 * 1024;
    if (grouped) {
        if (int rc = grouped_aggregate(rowptr, col, n, src, dst, num_edges, batch_size, order, weighted, variant, fill, ip,
                                       plan_scratch, (const Record*)records, (const ColStat*)colstat, batch_scalars, x, nvec,
                                       xcn1, xcn2, xcn3, xij, st))
            return rc;
        if (plan_host[OCN_PLAN_WIDE_LINKS] == 0) return OCN_OK;
    }
    const int64_t min_deg = grouped ? kGroupedMaxDeg : -1;  // grouped: only the links the grouped kernel left out
    int lpr = 1;
    while (lpr < nvec && lpr < 32) lpr <<= 1;
    const int vpl = (nvec + 31) / 32;
    const int grid = grid_for_warps(num_edges);
    // A warp walks a link alone while that link is a small share of the warp's work: beyond an eighth of the mean number
    // of positions per resident warp (16 a multiprocessor) a whole CTA takes the link.  Streams of uniformly wide links
    // (ddi: the CTA walk costs 30 % more there) keep the warp walk up to kHeavyLink; a stream with a skewed tail (collab:
    // mean 30, maximum 751) hands the tail over (1190 -> 700 us; profiles/r02_ab_aggregate_wide.txt).
    int64_t wide_deg = plan_host ? plan_host[OCN_PLAN_NUM_RECORDS] / ((int64_t)sm_count() * 16 * 8) : kHeavyLink;
    if (wide_deg < 64) wide_deg = 64;
    if (wide_deg > kHeavyLink) wide_deg = kHeavyLink;
#define AGG(V)                                                                                                       \
    k_cn_aggregate<V><<<grid, 256, 0, st>>>(rowptr, col, n, src, dst, num_edges, batch_size, order, weighted, variant, \
                                            fill, ip, rec_off, (const Record*)records, (const ColStat*)colstat,      \
                                            batch_scalars, x, nvec, lpr, xcn1, xcn2, xcn3, xij, min_deg, wide_deg)
    if (vpl <= 1) AGG(1);
    else if (vpl <= 2) AGG(2);
    else if (vpl <= 4) AGG(4);
    else AGG(8);
#undef AGG
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_cn_aggregate_bwd(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src,
                         const int64_t* dst, int64_t num_edges, int64_t batch_size, int order, int weighted,
                         int variant, float fill, const float* ip, const void* plan_scratch, const void* records,
                         const void* colstat, const float* batch_scalars, const float* x, int64_t feat,
                         const float* g_xcn1, const float* g_xcn2, const float* g_xcn3, const float* g_xij,
                         float* grad_x, void* stream) {
    OCN_RANGE("ocn_cn_aggregate_bwd");
    OCN_CHECK_ARG(rowptr && col && src && dst && plan_scratch && colstat && batch_scalars && grad_x,
                  "ocn_cn_aggregate_bwd: null pointer");
    OCN_CHECK_ARG(num_edges > 0 && batch_size > 0 && n > 0 && feat > 0, "ocn_cn_aggregate_bwd: sizes must be positive");
    OCN_CHECK_ARG(g_xij == nullptr || x != nullptr, "ocn_cn_aggregate_bwd: x is needed for the pair term");
    cudaStream_t st = (cudaStream_t)stream;
    PLAN_PTRS();
    k_cn_aggregate_bwd<<<grid_for_warps(num_edges), 256, 0, st>>>(
        rowptr, col, n, src, dst, num_edges, batch_size, order, weighted, variant, fill, ip, rec_off,
        (const Record*)records, (const ColStat*)colstat, batch_scalars, x, feat, g_xcn1, g_xcn2, g_xcn3, g_xij, grad_x);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_cn_extract_count(const int64_t* rowptr, int64_t n, const int64_t* src, int64_t num_edges, int which,
                         int weighted, const void* plan_scratch, const void* records, int64_t* out_counts,
                         void* stream) {
    OCN_CHECK_ARG(rowptr && src && plan_scratch && records && out_counts, "ocn_cn_extract_count: null pointer");
    OCN_CHECK_ARG(which == 1 || which == 2 || which == 3 || which == 11 || which == 12 || which == 13,
                  "ocn_cn_extract_count: which must be 1,2,3,11,12,13");
    OCN_CHECK_ARG(num_edges > 0, "ocn_cn_extract_count: num_edges must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    PLAN_PTRS();
    k_cn_extract<false><<<grid_for_warps(num_edges), 256, 0, st>>>(
        rowptr, nullptr, n, src, num_edges, 1, which, 3, weighted, 5, 0.f, nullptr, rec_off, (const Record*)records,
        nullptr, nullptr, out_counts, nullptr, nullptr, nullptr);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_cn_extract_fill(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, int64_t num_edges,
                        int64_t batch_size, int which, int weighted, int variant, float fill, const float* ip,
                        const void* plan_scratch, const void* records, const void* colstat,
                        const float* batch_scalars, const int64_t* out_rowptr, int64_t* out_col, float* out_val,
                        void* stream) {
    OCN_CHECK_ARG(rowptr && col && src && plan_scratch && records && out_rowptr && out_col,
                  "ocn_cn_extract_fill: null pointer");
    OCN_CHECK_ARG(which == 1 || which == 2 || which == 3 || which == 11 || which == 12 || which == 13,
                  "ocn_cn_extract_fill: which must be 1,2,3,11,12,13");
    OCN_CHECK_ARG(which < 10 || (colstat && batch_scalars) || !out_val,
                  "ocn_cn_extract_fill: normalised values need colstat and batch_scalars");
    OCN_CHECK_ARG(num_edges > 0 && batch_size > 0, "ocn_cn_extract_fill: sizes must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    PLAN_PTRS();
    const int order = which > 10 ? which - 10 : which;
    k_cn_extract<true><<<grid_for_warps(num_edges), 256, 0, st>>>(
        rowptr, col, n, src, num_edges, batch_size, which, order, weighted, variant, fill, ip, rec_off,
        (const Record*)records, (const ColStat*)colstat, batch_scalars, nullptr, out_rowptr, out_col, out_val);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_cn_release(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, int64_t num_edges,
                   int64_t batch_size, const void* plan_scratch, const void* records, void* colstat,
                   const int64_t* plan_host, void* stream) {
    OCN_RANGE("ocn_cn_release");
    OCN_CHECK_ARG(rowptr && col && src && plan_scratch && records && colstat, "ocn_cn_release: null pointer");
    OCN_CHECK_ARG(num_edges > 0 && batch_size > 0, "ocn_cn_release: sizes must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    PLAN_PTRS();
    // a stream with more positions than the statistics have entries (ddi: 21 M positions, 16 batches x 4267 columns) is
    // cleaned by clearing the whole table instead of walking the records again
    const int64_t num_batches = (num_edges + batch_size - 1) / batch_size;
    if (plan_host != nullptr && plan_host[OCN_PLAN_NUM_RECORDS] >= num_batches * n) {
        OCN_CUDA(cudaMemsetAsync(colstat, 0, sizeof(ColStat) * (size_t)num_batches * (size_t)n, st));
        return OCN_OK;
    }
    if (use_grouped(num_edges, plan_host))
        return grouped_release(rowptr, col, n, src, num_edges, batch_size, plan_scratch, (ColStat*)colstat, st);
    k_cn_release<<<grid_for_warps(num_edges), 256, 0, st>>>(rowptr, col, n, src, num_edges, batch_size, rec_off,
                                                             (const Record*)records, (ColStat*)colstat);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // extern "C"
