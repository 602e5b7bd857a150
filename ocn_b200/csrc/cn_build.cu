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
// visited node.  The j-side frontier is streamed from HBM exactly once per (link, table).
//
// A work unit = (run, 32 positions of N(i), <= 64 links); see cn_plan.cu.  If the 32 rows do not
// fit the table the unit processes them in greedy sub-chunks; a single row longer than the
// table capacity is probed by binary search instead ("direct" mode).
#include "common.cuh"

namespace ocn {

constexpr int kBuildThreads = 256;
constexpr int kBuildWarps = kBuildThreads / 32;
constexpr int kSlotBits = 13;
constexpr int kSlots = 1 << kSlotBits;        // 8192 x 8 B = 64 KB
constexpr int kCap = (kSlots * 3) / 4;        // max keys inserted per table
constexpr uint32_t kEmpty = 0xffffffffu;

struct BuildSmem {
    uint2 table[kSlots];
    unsigned acc2[kBuildWarps][32];
    unsigned acc3[kBuildWarps][32];
    unsigned long long tot2[32];
    unsigned long long tot3[32];
    long long krs[32];
    unsigned tot1[32];
    int kp[32];
    int kdeg[32];
    int kpre[33];
    long long unit;
    int q1;
    int direct;
};

__device__ __forceinline__ uint32_t ht_hash(uint32_t key) { return (key * 2654435769u) >> (32 - kSlotBits); }

__device__ __forceinline__ void ht_insert(uint2* table, uint32_t key, uint32_t bit) {
    uint32_t slot = ht_hash(key);
    while (true) {
        uint32_t prev = atomicCAS(&table[slot].x, kEmpty, key);
        if (prev == kEmpty || prev == key) {
            atomicOr(&table[slot].y, bit);
            return;
        }
        slot = (slot + 1) & (kSlots - 1);
    }
}

__device__ __forceinline__ uint32_t ht_lookup(const uint2* table, uint32_t key) {
    uint32_t slot = ht_hash(key);
    while (true) {
        uint2 e = table[slot];
        if (e.x == key) return e.y;
        if (e.x == kEmpty) return 0u;
        slot = (slot + 1) & (kSlots - 1);
    }
}

template <bool DIRECT>
__device__ __forceinline__ uint32_t probe(const BuildSmem& S, const int32_t* __restrict__ col, int q0, uint32_t key) {
    if (DIRECT) return row_contains(col + S.krs[q0], S.kdeg[q0], (int32_t)key) ? (1u << q0) : 0u;
    return ht_lookup(S.table, key);
}

__device__ __forceinline__ void add_bits(unsigned* acc, uint32_t mask) {
    while (mask) {
        int b = __ffs(mask) - 1;
        atomicAdd(&acc[b], 1u);
        mask &= mask - 1;
    }
}

// one warp, one link, against the table currently in shared memory
template <bool DIRECT>
__device__ __forceinline__ void walk_link(BuildSmem& S, const int64_t* __restrict__ rowptr,
                                          const int32_t* __restrict__ col, int64_t j, int order, int q0, int warp,
                                          int lane, uint32_t& m1_out) {
    unsigned* acc2 = S.acc2[warp];
    unsigned* acc3 = S.acc3[warp];
    acc2[lane] = 0;
    acc3[lane] = 0;
    __syncwarp();
    m1_out = probe<DIRECT>(S, col, q0, (uint32_t)j);
    if (order >= 2) {
        int64_t rs_j = ldg_i64(rowptr + j);
        int64_t dj = ldg_i64(rowptr + j + 1) - rs_j;
        for (int64_t base = 0; base < dj; base += 32) {
            int64_t o = base + lane;
            int32_t m = -1;
            int64_t rs_m = 0;
            int dm = 0;
            if (o < dj) {
                m = ldg_i32(col + rs_j + o);
                add_bits(acc2, probe<DIRECT>(S, col, q0, (uint32_t)m));
                if (order >= 3) {
                    rs_m = ldg_i64(rowptr + m);
                    dm = (int)(ldg_i64(rowptr + m + 1) - rs_m);
                }
            }
            if (order >= 3) {
                int cnt = (int)((dj - base) < 32 ? (dj - base) : 32);
                for (int u = 0; u < cnt; ++u) {
                    int64_t rs = __shfl_sync(0xffffffffu, rs_m, u);
                    int d = __shfl_sync(0xffffffffu, dm, u);
                    for (int oo = lane; oo < d; oo += 32) {
                        int32_t l = ldg_i32(col + rs + oo);
                        add_bits(acc3, probe<DIRECT>(S, col, q0, (uint32_t)l));
                    }
                }
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kBuildThreads)
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
        // run = last r with run_unit_off[r] <= unit
        int64_t lo = 0, hi = n_runs;
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
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

        if (tid < 32) {
            S.tot1[tid] = 0;
            S.tot2[tid] = 0;
            S.tot3[tid] = 0;
            if (tid < np) {
                int32_t k = col[rs_i + p0 + tid];
                int64_t krs = rowptr[k];
                S.kp[tid] = k;
                S.krs[tid] = krs;
                S.kdeg[tid] = (int)(rowptr[k + 1] - krs);
            }
        }
        int q0 = 0;
        while (q0 < np) {
            __syncthreads();
            if (tid == 0) {
                int q = q0, sum = 0, direct = 0;
                if (S.kdeg[q0] > kCap) {
                    direct = 1;
                    q = q0 + 1;
                } else {
                    while (q < np && sum + S.kdeg[q] <= kCap) {
                        S.kpre[q] = sum;
                        sum += S.kdeg[q];
                        ++q;
                    }
                    S.kpre[q] = sum;
                }
                S.q1 = q;
                S.direct = direct;
            }
            for (int s = tid; s < kSlots; s += kBuildThreads) S.table[s] = make_uint2(kEmpty, 0u);
            __syncthreads();
            const int q1 = S.q1;
            const bool direct = S.direct != 0;
            if (!direct) {
                const int total = S.kpre[q1];
                for (int idx = tid; idx < total; idx += kBuildThreads) {
                    int a = q0, b = q1;  // last q in [q0,q1) with kpre[q] <= idx
                    while (b - a > 1) {
                        int mid = (a + b) >> 1;
                        if (S.kpre[mid] <= idx) a = mid; else b = mid;
                    }
                    int32_t l = ldg_i32(col + S.krs[a] + (idx - S.kpre[a]));
                    ht_insert(S.table, (uint32_t)l, 1u << a);
                }
            }
            __syncthreads();
            for (int e = warp; e < ne; e += kBuildWarps) {
                const int64_t t = e0 + e;
                const int64_t j = dst[t];
                uint32_t m1;
                if (direct) walk_link<true>(S, rowptr, col, j, order, q0, warp, lane, m1);
                else walk_link<false>(S, rowptr, col, j, order, q0, warp, lane, m1);
                if (lane >= q0 && lane < q1) {
                    unsigned c1 = (m1 >> lane) & 1u;
                    unsigned c2 = S.acc2[warp][lane];
                    unsigned c3 = S.acc3[warp][lane];
                    records[rec_off[t] + p0 + lane] = make_uint2(c2 | (c1 << 31), c3);
                    if (colstat != nullptr) {
                        if (c1) atomicAdd(&S.tot1[lane], 1u);
                        unsigned long long v2 = weighted ? c2 : (c2 ? 1u : 0u);
                        unsigned long long v3 = weighted ? c3 : (c3 ? 1u : 0u);
                        if (v2) atomicAdd(&S.tot2[lane], v2);
                        if (v3) atomicAdd(&S.tot3[lane], v3);
                    }
                }
                __syncwarp();
            }
            q0 = q1;
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
    int blocks = sm_count() * 3;
    k_cn_build<<<blocks, kBuildThreads, sizeof(BuildSmem), st>>>(
        rowptr, col, n, src, dst, batch_size, order, weighted, (const int64_t*)(base + L.rec_off),
        (const int32_t*)(base + L.run_start), (const int64_t*)(base + L.run_unit_off), (int64_t*)plan,
        (Record*)records, (ColStat*)colstat);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}
