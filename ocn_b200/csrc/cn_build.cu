// ocn_cn_build: higher-order common-neighbour sets without materialising A^2 / A^3.
//
// For a target link e = (i, j) every CN set of the reference is a subset of N(i):
//     CN_k(e) = A[i] (*) A^k[j]                  (NeighborOverlapCitation2.py:78-85, utils.py:248-285)
// and, because A is symmetric, with mask_i(l) = { p : l in N(N(i)[p]) } (a bit per position p):
//     C1[p] = bit p of mask_i(j)
//     C2[p] = #{ m in N(j)               : bit p of mask_i(m) }      (= A^2[j, N(i)[p]])
//     C3[p] = #{ m in N(j), l in N(m)    : bit p of mask_i(l) }      (= A^3[j, N(i)[p]])
// so one shared-memory hash table  l -> mask_i(l)  (built once per run of links that share the
// source i) turns the whole computation into a 3-level walk from j with one table probe per
// visited node.
//
// Work unit = (run, 32 positions of N(i), <= kEdgeSub links), scheduled dynamically (cn_plan.cu).
// The 32 rows N(N(i)[p]) are inserted as one flattened list; if it exceeds the table capacity the
// unit makes several passes (table = next slice of the list, j-side walked again) and sums the
// per-position counts in shared memory, so a hub next to the source needs no special case.
//
// The j-side frontier is streamed row by row, 8 x 32 columns per warp iteration (8 independent
// 128-byte loads in flight per warp); the 32 per-position counters of every lane are kept
// bit-sliced in registers and updated with carry-save adders (~9 ALU ops per probed column, no
// shared-memory atomics on the hot path).  Links with a large frontier are walked by all warps of
// the CTA, the others by one warp each.
#include "common.cuh"

namespace ocn {

constexpr int kBuildThreads = 512;
constexpr int kBuildWarps = kBuildThreads / 32;
constexpr int kSlots = 13312;                  // 104 KB of (key, mask) pairs -> 2 CTAs / SM
constexpr int kCap = (kSlots * 5) / 8;         // keys inserted per pass (load factor 0.625)
constexpr uint32_t kEmpty = 0xffffffffu;
constexpr int kHiPlanes = 17;                  // bit-sliced counter: 3 + 17 planes (< 2^20 per lane)
constexpr int kShortRow = 8;                   // rows this short are walked one per lane
constexpr int kHeavyFrontier = 12288;          // links above this frontier size use the whole CTA

struct BuildSmem {
    uint2 table[kSlots];
    unsigned acc2[kEdgeSub][32];
    unsigned acc3[kEdgeSub][32];
    unsigned long long tot2[32];
    unsigned long long tot3[32];
    long long krs[32];
    unsigned tot1[32];
    unsigned m1[kEdgeSub];
    int heavy[kEdgeSub];
    int kp[32];
    int kdeg[32];
    int kpre[33];
    long long unit;
    int n_heavy;
};

__device__ __forceinline__ uint32_t ht_home(uint32_t key) {
    return __umulhi(key * 2654435769u, (uint32_t)kSlots);
}

__device__ __forceinline__ void ht_insert(uint2* table, uint32_t key, uint32_t bit) {
    uint32_t slot = ht_home(key);
    while (true) {
        uint32_t prev = atomicCAS(&table[slot].x, kEmpty, key);
        if (prev == kEmpty || prev == key) {
            atomicOr(&table[slot].y, bit);
            return;
        }
        slot = (slot + 1 == (uint32_t)kSlots) ? 0u : slot + 1;
    }
}

__device__ __forceinline__ uint32_t ht_lookup(const uint2* table, uint32_t key) {
    uint32_t slot = ht_home(key);
    while (true) {
        const uint2 e = table[slot];
        if (e.x == key) return e.y;
        if (e.x == kEmpty) return 0u;
        slot = (slot + 1 == (uint32_t)kSlots) ? 0u : slot + 1;
    }
}

// ---- bit-sliced per-position counters ---------------------------------------------------------
struct Sliced {
    uint32_t ones, twos, fours;
    uint32_t hi[kHiPlanes];  // hi[b] has weight 8 << b
};

__device__ __forceinline__ void sliced_zero(Sliced& s) {
    s.ones = s.twos = s.fours = 0u;
#pragma unroll
    for (int b = 0; b < kHiPlanes; ++b) s.hi[b] = 0u;
}

// carry-save adder: (h, l) = a + b + c bitwise
__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}

// add eight 32-position masks to the counters
__device__ __forceinline__ void sliced_add8(Sliced& s, const uint32_t (&m)[8]) {
    uint32_t t0, t1, t2, t3, f0, f1, e0;
    csa(t0, s.ones, s.ones, m[0], m[1]);
    csa(t1, s.ones, s.ones, m[2], m[3]);
    csa(f0, s.twos, s.twos, t0, t1);
    csa(t2, s.ones, s.ones, m[4], m[5]);
    csa(t3, s.ones, s.ones, m[6], m[7]);
    csa(f1, s.twos, s.twos, t2, t3);
    csa(e0, s.fours, s.fours, f0, f1);
    uint32_t carry = e0;
#pragma unroll
    for (int b = 0; b < kHiPlanes; ++b) {
        const uint32_t t = s.hi[b] & carry;
        s.hi[b] ^= carry;
        carry = t;
    }
}

// lane p receives the sum over all lanes of position p's counter
__device__ __forceinline__ uint32_t plane_sum(uint32_t word, int lane) {
    uint32_t tot = 0;
    if (__any_sync(0xffffffffu, word != 0u)) {
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            const uint32_t bal = __ballot_sync(0xffffffffu, (word >> p) & 1u);
            if (lane == p) tot = __popc(bal);
        }
    }
    return tot;
}

__device__ __forceinline__ uint32_t sliced_flush(const Sliced& s, int lane) {
    uint32_t total = plane_sum(s.ones, lane) + 2u * plane_sum(s.twos, lane) + 4u * plane_sum(s.fours, lane);
#pragma unroll
    for (int b = 0; b < kHiPlanes; ++b) total += plane_sum(s.hi[b], lane) << (3 + b);
    return total;
}

__device__ __forceinline__ void add_bits(unsigned* acc, uint32_t mask) {
    while (mask) {
        const int b = __ffs(mask) - 1;
        atomicAdd(&acc[b], 1u);
        mask &= mask - 1;
    }
}

// Walk the frontier of link slot `e` (destination j) with `nw` cooperating warps, this one being
// `rank`.  Adds into S.acc2[e], S.acc3[e], S.m1[e].
__device__ __forceinline__ void walk_link(BuildSmem& S, const int64_t* __restrict__ rowptr,
                                          const int32_t* __restrict__ col, int64_t j, int order, int e, int rank,
                                          int nw, int lane) {
    const uint2* table = S.table;
    if (rank == 0) {
        const uint32_t m1 = ht_lookup(table, (uint32_t)j);
        if (lane == 0 && m1) atomicOr(&S.m1[e], m1);
    }
    if (order < 2) return;
    const int64_t rs_j = ldg_i64(rowptr + j);
    const int64_t dj = ldg_i64(rowptr + j + 1) - rs_j;
    Sliced cnt;
    sliced_zero(cnt);
    int blk = 0;
    for (int64_t base = 0; base < dj; base += 32, ++blk) {
        const int64_t o = base + lane;
        int32_t m = -1;
        int64_t rs_m = 0;
        int dm = 0;
        if (o < dj) {
            m = ldg_i32(col + rs_j + o);
            if (order >= 3) {
                rs_m = ldg_i64(rowptr + m);
                dm = (int)(ldg_i64(rowptr + m + 1) - rs_m);
            }
        }
        const bool mine = (blk % nw) == rank;
        if (mine && m >= 0) add_bits(S.acc2[e], ht_lookup(table, (uint32_t)m));
        if (order < 3) continue;
        // short rows: one row per lane, all of them at once
        if (mine) {
            const bool is_short = m >= 0 && dm <= kShortRow;
            if (__any_sync(0xffffffffu, is_short)) {
                uint32_t mk[8];
                int32_t l[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) l[u] = (is_short && u < dm) ? ldg_i32(col + rs_m + u) : -1;
#pragma unroll
                for (int u = 0; u < 8; ++u) mk[u] = l[u] >= 0 ? ht_lookup(table, (uint32_t)l[u]) : 0u;
                sliced_add8(cnt, mk);
            }
        }
        // long rows: the cooperating warps stride over each row, 8 x 32 columns per iteration
        unsigned longmask = __ballot_sync(0xffffffffu, m >= 0 && dm > kShortRow);
        while (longmask) {
            const int u = __ffs(longmask) - 1;
            longmask &= longmask - 1;
            const int64_t rs = __shfl_sync(0xffffffffu, rs_m, u);
            const int d = __shfl_sync(0xffffffffu, dm, u);
            const int first = (rank + nw - (u % nw)) % nw;  // rotate so medium rows spread over the warps
            for (int b2 = first * 256; b2 < d; b2 += nw * 256) {
                uint32_t mk[8];
                int32_t l[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int idx = b2 + k * 32 + lane;
                    l[k] = idx < d ? ldg_i32(col + rs + idx) : -1;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) mk[k] = l[k] >= 0 ? ht_lookup(table, (uint32_t)l[k]) : 0u;
                sliced_add8(cnt, mk);
            }
        }
    }
    if (order >= 3) {
        const uint32_t total = sliced_flush(cnt, lane);
        if (total) atomicAdd(&S.acc3[e][lane], total);
    }
}

__global__ void __launch_bounds__(kBuildThreads, 2)
k_cn_build(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
           const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t batch_size, int order,
           int weighted, const int64_t* __restrict__ rec_off, const int32_t* __restrict__ run_start,
           const int64_t* __restrict__ run_unit_off, int64_t* __restrict__ plan, Record* __restrict__ records,
           ColStat* __restrict__ colstat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BuildSmem& S = *reinterpret_cast<BuildSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_units = plan[OCN_PLAN_NUM_UNITS];
    const int64_t n_runs = plan[OCN_PLAN_NUM_RUNS];

    while (true) {
        __syncthreads();
        if (tid == 0) S.unit = (long long)atomicAdd((unsigned long long*)&plan[4], 1ull);
        __syncthreads();
        const int64_t unit = S.unit;
        if (unit >= n_units) break;
        int64_t lo = 0, hi = n_runs;  // run = last r with run_unit_off[r] <= unit
        while (hi - lo > 1) {
            const int64_t mid = (lo + hi) >> 1;
            if (run_unit_off[mid] <= unit) lo = mid; else hi = mid;
        }
        const int64_t r = lo;
        const int64_t local = unit - run_unit_off[r];
        const int64_t t0 = run_start[r], len = run_start[r + 1] - t0;
        const int64_t i = src[t0];
        const int64_t rs_i = rowptr[i];
        const int64_t d = rowptr[i + 1] - rs_i;
        const int64_t n_es = (len + kEdgeSub - 1) / kEdgeSub;
        const int64_t pc = local / n_es, es = local - pc * n_es;
        const int64_t p0 = pc * kPChunk;
        const int np = (int)((d - p0) < kPChunk ? (d - p0) : kPChunk);
        const int64_t e0 = t0 + es * kEdgeSub;
        const int ne = (int)((t0 + len - e0) < kEdgeSub ? (t0 + len - e0) : kEdgeSub);
        const int64_t batch = t0 / batch_size;

        if (warp == 0) {
            S.tot1[lane] = 0;
            S.tot2[lane] = 0;
            S.tot3[lane] = 0;
            int kd = 0;
            if (lane < np) {
                const int32_t k = col[rs_i + p0 + lane];
                const int64_t krs = rowptr[k];
                kd = (int)(rowptr[k + 1] - krs);
                S.kp[lane] = k;
                S.krs[lane] = krs;
                S.kdeg[lane] = kd;
            }
            int incl = kd;  // inclusive warp scan of the row lengths
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            S.kpre[lane] = incl - kd;
            if (lane == 31) S.kpre[32] = incl;
        }
        for (int s = tid; s < kEdgeSub * 32; s += kBuildThreads) {
            (&S.acc2[0][0])[s] = 0u;
            (&S.acc3[0][0])[s] = 0u;
        }
        if (tid < kEdgeSub) S.m1[tid] = 0u;
        __syncthreads();
        const int total_keys = S.kpre[32];
        const int npass = total_keys > 0 ? (total_keys + kCap - 1) / kCap : 1;

        for (int pass = 0; pass < npass; ++pass) {
            __syncthreads();
            for (int s = tid; s < kSlots / 2; s += kBuildThreads)
                reinterpret_cast<uint4*>(S.table)[s] = make_uint4(kEmpty, 0u, kEmpty, 0u);
            if (tid == 0) S.n_heavy = 0;
            __syncthreads();
            const int k_lo = pass * kCap, k_hi = (k_lo + kCap < total_keys) ? k_lo + kCap : total_keys;
            for (int idx = k_lo + tid; idx < k_hi; idx += kBuildThreads) {
                int a = 0, b = np;  // last position a with kpre[a] <= idx
                while (b - a > 1) {
                    const int mid = (a + b) >> 1;
                    if (S.kpre[mid] <= idx) a = mid; else b = mid;
                }
                const int32_t l = ldg_i32(col + S.krs[a] + (idx - S.kpre[a]));
                ht_insert(S.table, (uint32_t)l, 1u << a);
            }
            __syncthreads();
            if (ne >= kBuildWarps && order >= 3) {
                // one warp per link; links with a large frontier are deferred to the whole CTA
                for (int e = warp; e < ne; e += kBuildWarps) {
                    const int64_t j = dst[e0 + e];
                    const int64_t rs_j = ldg_i64(rowptr + j);
                    const int64_t dj = ldg_i64(rowptr + j + 1) - rs_j;
                    long long fr = 0;
                    for (int64_t o = lane; o < dj; o += 32) {
                        const int32_t m = ldg_i32(col + rs_j + o);
                        fr += ldg_i64(rowptr + m + 1) - ldg_i64(rowptr + m);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) fr += __shfl_xor_sync(0xffffffffu, fr, o);
                    if (fr > kHeavyFrontier) {
                        if (lane == 0) S.heavy[atomicAdd(&S.n_heavy, 1)] = e;
                    } else {
                        walk_link(S, rowptr, col, j, order, e, 0, 1, lane);
                    }
                }
            } else if (order < 3) {
                for (int e = warp; e < ne; e += kBuildWarps) walk_link(S, rowptr, col, dst[e0 + e], order, e, 0, 1, lane);
            } else {
                if (tid < ne) S.heavy[tid] = tid;
                if (tid == 0) S.n_heavy = ne;
            }
            __syncthreads();
            const int nh = S.n_heavy;
            for (int h = 0; h < nh; ++h) {
                const int e = S.heavy[h];
                walk_link(S, rowptr, col, dst[e0 + e], order, e, warp, kBuildWarps, lane);
            }
        }
        __syncthreads();
        // records and per-unit column totals
        for (int s = tid; s < ne * 32; s += kBuildThreads) {
            const int e = s >> 5, p = s & 31;
            if (p < np) {
                const unsigned c1 = (S.m1[e] >> p) & 1u;
                const unsigned c2 = S.acc2[e][p], c3 = S.acc3[e][p];
                records[rec_off[e0 + e] + p0 + p] = make_uint2(c2 | (c1 << 31), c3);
                if (colstat != nullptr) {
                    if (c1) atomicAdd(&S.tot1[p], 1u);
                    const unsigned long long v2 = weighted ? c2 : (c2 ? 1u : 0u);
                    const unsigned long long v3 = weighted ? c3 : (c3 ? 1u : 0u);
                    if (v2) atomicAdd(&S.tot2[p], v2);
                    if (v3) atomicAdd(&S.tot3[p], v3);
                }
            }
        }
        __syncthreads();
        if (colstat != nullptr && tid < np) {
            ColStat* cs = colstat + batch * n + S.kp[tid];
            if (S.tot1[tid]) atomicAdd(&cs->c1, S.tot1[tid]);
            if (S.tot2[tid]) atomicAdd(&cs->s2, S.tot2[tid]);
            if (S.tot3[tid]) atomicAdd(&cs->s3, S.tot3[tid]);
        }
    }
}

}  // namespace ocn

using namespace ocn;

extern "C" int ocn_cn_build(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src,
                            const int64_t* dst, int64_t num_edges, int64_t batch_size, int order, int weighted,
                            const void* plan_scratch, const int64_t* plan, void* records, int64_t records_capacity,
                            void* colstat, void* stream) {
    OCN_CHECK_ARG(rowptr && col && src && dst && plan_scratch && plan, "ocn_cn_build: null pointer");
    OCN_CHECK_ARG(order >= 1 && order <= 3, "ocn_cn_build: order must be 1, 2 or 3 (got %d)", order);
    OCN_CHECK_ARG(n > 0 && num_edges > 0 && batch_size > 0, "ocn_cn_build: sizes must be positive");
    OCN_CHECK_ARG(records || records_capacity == 0, "ocn_cn_build: records is null");
    cudaStream_t st = (cudaStream_t)stream;
    PlanLayout L = plan_layout(num_edges);
    const char* base = (const char*)plan_scratch;
    OCN_CUDA(cudaFuncSetAttribute(k_cn_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BuildSmem)));
    // restart the dynamic unit counter (plan[4])
    OCN_CUDA(cudaMemsetAsync((void*)(plan + 4), 0, sizeof(int64_t), st));
    const int blocks = sm_count() * 2;
    k_cn_build<<<blocks, kBuildThreads, sizeof(BuildSmem), st>>>(
        rowptr, col, n, src, dst, batch_size, order, weighted, (const int64_t*)(base + L.rec_off),
        (const int32_t*)(base + L.run_start), (const int64_t*)(base + L.run_unit_off), (int64_t*)plan,
        (Record*)records, (ColStat*)colstat);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}
