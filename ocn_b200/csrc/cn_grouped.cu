// Run-grouped variants of the kernels that follow the build: column statistics, batch scalars, CN-indicator
// SpMM (forward) and release, for streams whose links come in long runs of one source (the citation2 evaluation
// stream: every source against 1000 destinations, NeighborOverlapCitation2.py:248-252).
//
// The per-link kernels of cn_build.cu / cn_aggregate.cu give every link a warp that chases
//     src -> rowptr -> records -> col -> column statistics (32 B out of a 94 MB-per-batch array) -> x[k] (128 B)
// i.e. five dependent, mostly DRAM-latency loads for ~17 records; 65 536 links took 21 + 45 + 122 + 27 us
// (profiles/r02_launches_a.txt).  But all links of a run share N(src): the same 17 columns k_p, the same 17 rows of x
// and -- inside one link batch -- the same 17 column statistics, and their records are one dense
// [links x positions] matrix (rec_off grows by deg(src) per link).  So here a CTA takes a WINDOW of kWinLinks
// consecutive links, cuts it into segments of one (run, batch), and per segment and tile of 32 positions
//   * loads the columns, their statistics and their rows of x ONCE into shared memory,
//   * computes the weights of all (link, position) records of the segment (coalesced record reads),
//   * accumulates the three weighted sums per link from shared memory; the only gather left per link is x[dst] for
//     the pair term.
// The divisions of the weight algebra depend on the column only (cn_weights.cuh: node_weights): they are evaluated
// once per (batch, position) instead of once per record.
// Sums run over ascending positions: run-to-run deterministic (and equal for any window size).
// The host side picks these kernels when the stream averages >= kGroupedMinRun links per run.  Segments whose
// source has more than kGroupedMaxDeg neighbours are skipped here and handled by the per-link kernels (launched
// with min_deg = kGroupedMaxDeg when the plan counted such links): a window of a source with 900 neighbours would
// run 29 tiles and be the tail of the launch (first version: SMs busy 42 % of the kernel's duration,
// profiles/r02_grouped_v1_ncu_summary.txt).
#include "cn_weights.cuh"

namespace ocn {

constexpr int kWinLinks = 64;   // links per window
constexpr int kPosTile = 32;    // positions per tile (one lane per position in the weight phase)
constexpr int kGroupedThreads = 256;
constexpr int kGroupedWarps = kGroupedThreads / 32;
constexpr int kLinksPerWarp = kWinLinks / kGroupedWarps;

struct Segment {
    int64_t t0, t1;  // links [t0, t1): one run, one batch
    int64_t b;       // batch
    int64_t i;       // source
    int64_t rs;      // rowptr[i]
    int64_t d;       // deg(i)
    int64_t ro;      // rec_off[t0]; link t0 + l has its records at ro + l * d
};

// the segment that starts at link t inside the window that ends at tw1
__device__ __forceinline__ Segment segment_at(int64_t t, int64_t tw1, int64_t batch_size, const int64_t* __restrict__ rowptr,
                                              const int64_t* __restrict__ src, const int32_t* __restrict__ run_id,
                                              const int32_t* __restrict__ run_start, const int64_t* __restrict__ rec_off) {
    Segment s;
    s.t0 = t;
    s.b = t / batch_size;
    const int64_t rend = run_start[run_id[t]];  // run_id is the run index + 1: the first link of the next run
    const int64_t bend = (s.b + 1) * batch_size;
    s.t1 = tw1 < rend ? tw1 : rend;
    if (bend < s.t1) s.t1 = bend;
    s.i = src[t];
    s.rs = rowptr[s.i];
    s.d = rowptr[s.i + 1] - s.rs;
    s.ro = rec_off[t];
    return s;
}

// ---- column statistics ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGroupedThreads)
k_cn_colstat_grouped(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                     const int64_t* __restrict__ src, int64_t T, int64_t batch_size, int weighted,
                     const int64_t* __restrict__ rec_off, const int32_t* __restrict__ run_id,
                     const int32_t* __restrict__ run_start, const Record* __restrict__ records,
                     ColStat* __restrict__ colstat) {
    __shared__ unsigned int s_c1[kPosTile];
    __shared__ unsigned long long s_s2[kPosTile], s_s3[kPosTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_win = (T + kWinLinks - 1) / kWinLinks;
    for (int64_t w = blockIdx.x; w < n_win; w += gridDim.x) {
        const int64_t tw1 = (w + 1) * kWinLinks < T ? (w + 1) * kWinLinks : T;
        for (int64_t t = w * kWinLinks; t < tw1;) {
            const Segment sg = segment_at(t, tw1, batch_size, rowptr, src, run_id, run_start, rec_off);
            t = sg.t1;
            if (sg.d > kGroupedMaxDeg) continue;  // per-link kernel
            const int nl = (int)(sg.t1 - sg.t0);
            for (int64_t p0 = 0; p0 < sg.d; p0 += kPosTile) {
                if (threadIdx.x < kPosTile) { s_c1[threadIdx.x] = 0u; s_s2[threadIdx.x] = 0ull; s_s3[threadIdx.x] = 0ull; }
                __syncthreads();
                // lane = position of the tile; the warp's links of the segment, summed in registers
                unsigned int c1 = 0u;
                unsigned long long s2 = 0ull, s3 = 0ull;
                const int64_t p = p0 + lane;
                if (p < sg.d) {
                    for (int l = warp; l < nl; l += kGroupedWarps) {
                        const Record rec = records[sg.ro + (int64_t)l * sg.d + p];
                        const unsigned c2 = rec.x & 0x7fffffffu;
                        c1 += rec.x >> 31;
                        s2 += weighted ? c2 : (c2 ? 1u : 0u);
                        s3 += weighted ? rec.y : (rec.y ? 1u : 0u);
                    }
                }
                if (c1) atomicAdd(&s_c1[lane], c1);
                if (s2) atomicAdd(&s_s2[lane], s2);
                if (s3) atomicAdd(&s_s3[lane], s3);
                __syncthreads();
                if (threadIdx.x < kPosTile && p0 + threadIdx.x < sg.d) {
                    const unsigned int a = s_c1[threadIdx.x];
                    const unsigned long long b2 = s_s2[threadIdx.x], b3 = s_s3[threadIdx.x];
                    if (a | b2 | b3) {
                        ColStat* c = colstat + sg.b * n + ldg_i32(col + sg.rs + p0 + threadIdx.x);
                        if (a) atomicAdd(&c->c1, a);
                        if (b2) atomicAdd(&c->s2, b2);
                        if (b3) atomicAdd(&c->s3, b3);
                    }
                }
                __syncthreads();
            }
        }
    }
}

// ---- batch scalars (same outputs as k_stats: per-link partial sums + the minimum column count) -------------------
__global__ void __launch_bounds__(kGroupedThreads)
k_stats_grouped(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                const int64_t* __restrict__ src, int64_t T, int64_t batch_size, int order, int weighted, int variant,
                float fill, const float* __restrict__ ip, int stage, const int64_t* __restrict__ rec_off,
                const int32_t* __restrict__ run_id, const int32_t* __restrict__ run_start,
                const Record* __restrict__ records, const ColStat* __restrict__ colstat, float* __restrict__ bscal,
                float* __restrict__ partial) {
    __shared__ unsigned int s_c1[kPosTile];
    __shared__ NodeWeights s_nw[kPosTile];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_win = (T + kWinLinks - 1) / kWinLinks;
    for (int64_t w = blockIdx.x; w < n_win; w += gridDim.x) {
        const int64_t tw1 = (w + 1) * kWinLinks < T ? (w + 1) * kWinLinks : T;
        for (int64_t t = w * kWinLinks; t < tw1;) {
            const Segment sg = segment_at(t, tw1, batch_size, rowptr, src, run_id, run_start, rec_off);
            t = sg.t1;
            if (sg.d > kGroupedMaxDeg) continue;  // per-link kernels
            const int nl = (int)(sg.t1 - sg.t0);
            const WeightParams P = make_params(order, weighted, variant, fill, ip, stage == 0 ? nullptr : bscal + sg.b * 8);
            float sa[kLinksPerWarp], sb[kLinksPerWarp];
            uint32_t mc[kLinksPerWarp];
#pragma unroll
            for (int q = 0; q < kLinksPerWarp; ++q) { sa[q] = 0.f; sb[q] = 0.f; mc[q] = 0xffffffffu; }
            for (int64_t p0 = 0; p0 < sg.d; p0 += kPosTile) {
                __syncthreads();
                if (threadIdx.x < kPosTile && p0 + threadIdx.x < sg.d) {
                    uint32_t c1;
                    unsigned long long s2, s3;
                    load_colstat(colstat + sg.b * n + ldg_i32(col + sg.rs + p0 + threadIdx.x), c1, s2, s3);
                    s_c1[threadIdx.x] = c1;
                    s_nw[threadIdx.x] = node_weights(c1, s2, s3, P);
                }
                __syncthreads();
                const int64_t p = p0 + lane;
                if (p < sg.d) {
                    const uint32_t c1 = s_c1[lane];
                    const NodeWeights nw = s_nw[lane];
#pragma unroll
                    for (int q = 0; q < kLinksPerWarp; ++q) {
                        const int l = warp + q * kGroupedWarps;
                        if (l < nl) {
                            const Record rec = records[sg.ro + (int64_t)l * sg.d + p];
                            if (rec.x | rec.y) {
                                const EntryWeights W = record_weights(rec, nw, P);
                                const uint32_t C2 = rec.x & 0x7fffffffu, C3 = rec.y;
                                if (stage == 0) {
                                    if (W.in1 && c1 >= 2u) mc[q] = c1 < mc[q] ? c1 : mc[q];
                                    if (W.in1 && C2) sa[q] += (weighted ? (float)C2 : 1.0f) * W.w1;
                                } else if (C3) {
                                    const float c3v = weighted ? (float)C3 : 1.0f;
                                    if (W.in1) sa[q] += c3v * W.w1;
                                    if (W.in2) sb[q] += c3v * W.w2;
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kLinksPerWarp; ++q) {
                const int l = warp + q * kGroupedWarps;
                if (l < nl) {  // warp-uniform
                    uint32_t m = mc[q];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const uint32_t other = __shfl_xor_sync(0xffffffffu, m, o);
                        m = other < m ? other : m;
                    }
                    const float a = warp_sum(sa[q]), c = warp_sum(sb[q]);
                    if (lane == 0) {
                        const int64_t tt = sg.t0 + l;
                        if (stage == 0) {
                            if (m != 0xffffffffu) atomicMin(reinterpret_cast<uint32_t*>(bscal) + sg.b * 8 + 4, m);
                            partial[tt] = a;
                        } else {
                            partial[T + 1 + tt] = a;
                            partial[2 * (T + 1) + tt] = c;
                        }
                    }
                }
            }
        }
    }
}

// ---- CN-indicator SpMM, forward ----------------------------------------------------------------------------
// dynamic shared memory: float4 Xs[kPosTile * nvec] | float4 xi[nvec] | float W[3][kWinLinks * kPosTile]
__global__ void __launch_bounds__(kGroupedThreads)
k_cn_aggregate_grouped(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                       const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t T, int64_t batch_size,
                       int order, int weighted, int variant, float fill, const float* __restrict__ ip,
                       const int64_t* __restrict__ rec_off, const int32_t* __restrict__ run_id,
                       const int32_t* __restrict__ run_start, const Record* __restrict__ records,
                       const ColStat* __restrict__ colstat, const float* __restrict__ bscal,
                       const float* __restrict__ x, int nvec, float* __restrict__ xcn1, float* __restrict__ xcn2,
                       float* __restrict__ xcn3, float* __restrict__ xij) {
    extern __shared__ float4 g_smem[];
    __shared__ NodeWeights s_nw[kPosTile];
    __shared__ int32_t s_k[kPosTile];
    __shared__ uint32_t s_nz[kWinLinks];
    float4* Xs = g_smem;
    float4* xi = g_smem + (size_t)kPosTile * nvec;
    float* W1 = reinterpret_cast<float*>(xi + nvec);
    float* W2 = W1 + kWinLinks * kPosTile;
    float* W3 = W2 + kWinLinks * kPosTile;
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t F = (int64_t)nvec * 4;
    const int64_t n_win = (T + kWinLinks - 1) / kWinLinks;
    for (int64_t w = blockIdx.x; w < n_win; w += gridDim.x) {
        const int64_t tw1 = (w + 1) * kWinLinks < T ? (w + 1) * kWinLinks : T;
        for (int64_t t = w * kWinLinks; t < tw1;) {
            const Segment sg = segment_at(t, tw1, batch_size, rowptr, src, run_id, run_start, rec_off);
            t = sg.t1;
            if (sg.d > kGroupedMaxDeg) continue;  // per-link kernel
            const int nl = (int)(sg.t1 - sg.t0);
            const WeightParams P = make_params(order, weighted, variant, fill, ip, bscal + sg.b * 8);
            const int64_t dd = sg.d > 0 ? sg.d : 1;  // a source without neighbours still writes its (zero) rows
            for (int64_t p0 = 0; p0 < dd; p0 += kPosTile) {
                __syncthreads();  // the previous tile / segment is done with the shared arrays
                if (threadIdx.x < kPosTile) {
                    int32_t k = -1;
                    if (p0 + threadIdx.x < sg.d) {
                        k = ldg_i32(col + sg.rs + p0 + threadIdx.x);
                        uint32_t c1;
                        unsigned long long s2, s3;
                        load_colstat(colstat + sg.b * n + k, c1, s2, s3);
                        s_nw[threadIdx.x] = node_weights(c1, s2, s3, P);
                    }
                    s_k[threadIdx.x] = k;
                }
                if (p0 == 0 && xij != nullptr)
                    for (int c = threadIdx.x; c < nvec; c += kGroupedThreads) xi[c] = __ldg(x4 + sg.i * nvec + c);
                __syncthreads();
                for (int idx = threadIdx.x; idx < kPosTile * nvec; idx += kGroupedThreads) {
                    const int row = idx / nvec, c = idx - row * nvec;
                    const int32_t k = s_k[row];
                    if (k >= 0) Xs[idx] = __ldg(x4 + (int64_t)k * nvec + c);
                }
                // weights of the (link, position) records of the tile: one warp per link, lane = position
                {
                    const int64_t p = p0 + lane;
                    const NodeWeights nw = s_nw[lane];
                    for (int l = warp; l < nl; l += kGroupedWarps) {
                        float w1 = 0.f, w2 = 0.f, w3 = 0.f;
                        bool nz = false;
                        if (p < sg.d) {
                            const Record rec = records[sg.ro + (int64_t)l * sg.d + p];
                            if (rec.x | rec.y) {
                                const EntryWeights Wt = record_weights(rec, nw, P);
                                nz = Wt.in1 || Wt.in2 || Wt.in3;
                                w1 = Wt.w1; w2 = Wt.w2; w3 = Wt.w3;
                            }
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, nz);
                        if (nz) {
                            W1[l * kPosTile + lane] = w1;
                            W2[l * kPosTile + lane] = w2;
                            W3[l * kPosTile + lane] = w3;
                        }
                        if (lane == 0) s_nz[l] = m;
                    }
                }
                __syncthreads();
                // the three sums of every link, one float4 column chunk per thread at a time
                for (int u = threadIdx.x; u < nl * nvec; u += kGroupedThreads) {
                    const int l = u / nvec, c = u - l * nvec;
                    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1, a3 = a1;
                    unsigned m = s_nz[l];
                    while (m) {
                        const int pp = __ffs(m) - 1;
                        m &= m - 1;
                        const float u1 = W1[l * kPosTile + pp], u2 = W2[l * kPosTile + pp], u3 = W3[l * kPosTile + pp];
                        const float4 xv = Xs[pp * nvec + c];
                        a1.x = fmaf(u1, xv.x, a1.x); a1.y = fmaf(u1, xv.y, a1.y); a1.z = fmaf(u1, xv.z, a1.z); a1.w = fmaf(u1, xv.w, a1.w);
                        a2.x = fmaf(u2, xv.x, a2.x); a2.y = fmaf(u2, xv.y, a2.y); a2.z = fmaf(u2, xv.z, a2.z); a2.w = fmaf(u2, xv.w, a2.w);
                        a3.x = fmaf(u3, xv.x, a3.x); a3.y = fmaf(u3, xv.y, a3.y); a3.z = fmaf(u3, xv.z, a3.z); a3.w = fmaf(u3, xv.w, a3.w);
                    }
                    const int64_t tt = sg.t0 + l;
                    float4* o1 = reinterpret_cast<float4*>(xcn1 + tt * F) + c;
                    float4* o2 = xcn2 ? reinterpret_cast<float4*>(xcn2 + tt * F) + c : nullptr;
                    float4* o3 = xcn3 ? reinterpret_cast<float4*>(xcn3 + tt * F) + c : nullptr;
                    if (p0 > 0) {  // later tiles of a long row add to what the same thread wrote for the earlier ones
                        const float4 q1 = *o1;
                        a1 = make_float4(q1.x + a1.x, q1.y + a1.y, q1.z + a1.z, q1.w + a1.w);
                        if (o2) { const float4 q2 = *o2; a2 = make_float4(q2.x + a2.x, q2.y + a2.y, q2.z + a2.z, q2.w + a2.w); }
                        if (o3) { const float4 q3 = *o3; a3 = make_float4(q3.x + a3.x, q3.y + a3.y, q3.z + a3.z, q3.w + a3.w); }
                    }
                    *o1 = a1;
                    if (o2) *o2 = a2;
                    if (o3) *o3 = a3;
                    if (p0 == 0 && xij != nullptr) {
                        const float4 xa = xi[c], xb = __ldg(x4 + dst[tt] * nvec + c);
                        reinterpret_cast<float4*>(xij + tt * F)[c] = make_float4(xa.x * xb.x, xa.y * xb.y, xa.z * xb.z, xa.w * xb.w);
                    }
                }
            }
        }
    }
}

// ---- release: the statistics of a (run, batch) live on the columns of N(src) ----------------------------------
__global__ void __launch_bounds__(kGroupedThreads)
k_cn_release_grouped(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                     const int64_t* __restrict__ src, int64_t T, int64_t batch_size,
                     const int64_t* __restrict__ rec_off, const int32_t* __restrict__ run_id,
                     const int32_t* __restrict__ run_start, ColStat* __restrict__ colstat) {
    const int64_t n_win = (T + kWinLinks - 1) / kWinLinks;
    for (int64_t w = blockIdx.x; w < n_win; w += gridDim.x) {
        const int64_t tw1 = (w + 1) * kWinLinks < T ? (w + 1) * kWinLinks : T;
        for (int64_t t = w * kWinLinks; t < tw1;) {
            const Segment sg = segment_at(t, tw1, batch_size, rowptr, src, run_id, run_start, rec_off);
            for (int64_t p = threadIdx.x; p < sg.d; p += kGroupedThreads) {
                uint4* c = reinterpret_cast<uint4*>(colstat + sg.b * n + ldg_i32(col + sg.rs + p));
                c[0] = make_uint4(0u, 0u, 0u, 0u);
                c[1] = make_uint4(0u, 0u, 0u, 0u);
            }
            t = sg.t1;
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------
bool use_grouped(int64_t T, const int64_t* plan_host) {
    if (plan_host == nullptr) return false;
    const int64_t num_runs = plan_host[OCN_PLAN_NUM_RUNS];
    return num_runs > 0 && T >= (int64_t)kGroupedMinRun * num_runs && option(OCN_OPT_GROUPED_OFF, 0) != 1;
}

static int grouped_grid(int64_t T) {
    const int64_t n_win = (T + kWinLinks - 1) / kWinLinks;
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)(n_win < cap ? n_win : cap);
}

int grouped_colstat(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, int64_t T, int64_t batch_size,
                    int weighted, const void* plan_scratch, const Record* records, ColStat* colstat, cudaStream_t st) {
    PlanLayout L = plan_layout(T);
    const char* pb = (const char*)plan_scratch;
    k_cn_colstat_grouped<<<grouped_grid(T), kGroupedThreads, 0, st>>>(
        rowptr, col, n, src, T, batch_size, weighted, (const int64_t*)(pb + L.rec_off), (const int32_t*)(pb + L.run_id),
        (const int32_t*)(pb + L.run_start), records, colstat);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int grouped_stats(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, int64_t T, int64_t batch_size,
                  int order, int weighted, int variant, float fill, const float* ip, int stage, const void* plan_scratch,
                  const Record* records, const ColStat* colstat, float* bscal, float* partial, cudaStream_t st) {
    PlanLayout L = plan_layout(T);
    const char* pb = (const char*)plan_scratch;
    k_stats_grouped<<<grouped_grid(T), kGroupedThreads, 0, st>>>(
        rowptr, col, n, src, T, batch_size, order, weighted, variant, fill, ip, stage, (const int64_t*)(pb + L.rec_off),
        (const int32_t*)(pb + L.run_id), (const int32_t*)(pb + L.run_start), records, colstat, bscal, partial);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

size_t grouped_aggregate_smem(int nvec) {
    return sizeof(float4) * ((size_t)kPosTile * nvec + nvec) + sizeof(float) * 3 * kWinLinks * kPosTile;
}

int grouped_aggregate(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, const int64_t* dst, int64_t T,
                      int64_t batch_size, int order, int weighted, int variant, float fill, const float* ip,
                      const void* plan_scratch, const Record* records, const ColStat* colstat, const float* bscal,
                      const float* x, int nvec, float* xcn1, float* xcn2, float* xcn3, float* xij, cudaStream_t st) {
    PlanLayout L = plan_layout(T);
    const char* pb = (const char*)plan_scratch;
    const size_t smem = grouped_aggregate_smem(nvec);
    OCN_CUDA(cudaFuncSetAttribute(k_cn_aggregate_grouped, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cn_aggregate_grouped<<<grouped_grid(T), kGroupedThreads, smem, st>>>(
        rowptr, col, n, src, dst, T, batch_size, order, weighted, variant, fill, ip, (const int64_t*)(pb + L.rec_off),
        (const int32_t*)(pb + L.run_id), (const int32_t*)(pb + L.run_start), records, colstat, bscal, x, nvec, xcn1, xcn2,
        xcn3, xij);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int grouped_release(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, int64_t T, int64_t batch_size,
                    const void* plan_scratch, ColStat* colstat, cudaStream_t st) {
    PlanLayout L = plan_layout(T);
    const char* pb = (const char*)plan_scratch;
    k_cn_release_grouped<<<grouped_grid(T), kGroupedThreads, 0, st>>>(
        rowptr, col, n, src, T, batch_size, (const int64_t*)(pb + L.rec_off), (const int32_t*)(pb + L.run_id),
        (const int32_t*)(pb + L.run_start), colstat);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // namespace ocn
