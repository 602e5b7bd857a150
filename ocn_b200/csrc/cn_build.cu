// ocn_cn_build: higher-order common-neighbour sets without materialising A^2 / A^3.
//
// For a target link e = (i, j) every CN set of the reference is a subset of N(i):
//     CN_k(e) = A[i] (*) A^k[j]                  (NeighborOverlapCitation2.py:78-85, utils.py:248-285)
// and, because A is symmetric, with mask_i(l) = { p : l in N(N(i)[p]) } (a bit per position p):
//     C1[p] = bit p of mask_i(j)
//     C2[p] = #{ m in N(j)               : bit p of mask_i(m) }      (= A^2[j, N(i)[p]])
//     C3[p] = #{ m in N(j), l in N(m)    : bit p of mask_i(l) }      (= A^3[j, N(i)[p]])
// so a shared-memory hash table  l -> mask_i(l)  (built once per group of links that share the
// source i) turns the whole computation into a 3-level walk from j with one probe per visited node.
//
// Work unit = (run of links with one source, 32 positions of N(i), table pass, cost window of the run),
// handed out dynamically (cn_plan.cu).  A cost window holds about unit_budget() probed columns: many
// light links (walked kEdgeSub at a time against the same table) or a slice of the rows of one heavy link.  The 32 rows N(N(i)[p]) form one flattened key list; pass q
// holds keys [q*kCap, (q+1)*kCap) of it, so hubs next to the source need no special case.  Units
// add their partial counts into the records with global atomics (records are zeroed first).
//
// Inside a unit
//   * the table is paired with a 256 Kbit one-hash bit filter: > 90 % of the frontier misses the
//     table (measured hit rate 0-8 %), and a miss costs one branch-free shared-memory bit test;
//   * columns that pass the filter are compacted into a per-warp queue and looked up 32 at a time,
//     so the divergent probe loops run with all lanes busy;
//   * the rows of the j-side frontier are described in shared memory, cut into chunks of kChunk columns
//     and pulled by the warps from a shared counter, so the heavy-tailed row lengths do not stall the
//     CTA at a barrier; rows of <= kTinyRow columns are walked by one thread each;
//   * frontier rows are streamed 8 x 32 columns per warp iteration: eight independent 128-byte
//     loads in flight per warp.
#include "common.cuh"

namespace ocn {

constexpr int kBuildThreads = 1024;
constexpr int kBuildWarps = kBuildThreads / 32;
constexpr uint32_t kEmpty = 0xffffffffu;
constexpr int kFilterBits = 18;                       // 262144-bit filter (32 KB)
constexpr int kFilterWords = 1 << (kFilterBits - 5);
constexpr int kQueue = 96;                            // per-warp queue of filter survivors
constexpr int kChunk = 2048;                          // columns of a frontier row handed to a warp at a time
constexpr int kTinyRow = 8;                           // rows this short are walked by one thread
constexpr int kRowBatch = 768;                        // frontier rows described in shared memory at a time

struct BuildSmem {
    uint2 table[kSlots + 7];
    uint32_t filter[kFilterWords];
    uint32_t qkey[kBuildWarps][kQueue];
    unsigned acc2[kEdgeSub][32];
    unsigned acc3[kEdgeSub][32];
    long long krs[32];
    long long jrs[kEdgeSub];
    uint32_t rrs[kRowBatch];       // row start (offset into col) of the frontier rows of the current batch
    int rd[kRowBatch];             // their lengths
    int rcp[kRowBatch + 1];        // exclusive prefix of their chunk counts
    int wsum[kBuildWarps];
    unsigned char re[kRowBatch];   // their link slots
    int kp[32];
    int kdeg[32];
    int kpre[33];
    int jdeg[kEdgeSub];
    int jpre[kEdgeSub + 1];
    unsigned m1[kEdgeSub];
    long long unit;
    int chunk, pass;
    int next_item;
    int n_chunks;
};

// queue entries carry the link slot in their top bits: node ids must stay below 2^kTagShift
constexpr int kTagShift = 27;
static_assert(kEdgeSub <= (1 << (32 - kTagShift)), "link slot does not fit the queue tag");

// explicit shared-window accesses (32-bit shared addresses): keeps the hot loops free of the
// generic-to-shared base recomputation the compiler otherwise emits per access
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct SAddr {  // shared-window byte addresses of the hot structures + the geometry used by this unit
    uint32_t table, filter, queue;
    uint32_t tsize;   // table slots in use (prime); a small key set uses (and clears) a small table
    uint32_t fshift;  // filter: byte offset of the word = (hash >> fshift) & fmask
    uint32_t fmask;
};

// table / filter geometry for n keys: smallest prime size with load <= 0.7, >= 16 filter bits per key
__device__ __forceinline__ void pick_geometry(int nkeys, SAddr& A) {
    const uint32_t sizes[5] = {1301u, 2609u, 5227u, 10487u, (uint32_t)kSlots};
    uint32_t ts = (uint32_t)kSlots;
#pragma unroll
    for (int k = 4; k >= 0; --k)
        if ((long long)sizes[k] * 7 >= (long long)nkeys * 10) ts = sizes[k];
    A.tsize = ts;
    int bits = 10;
    while (bits < kFilterBits && (1 << bits) < nkeys * 16) ++bits;
    A.fshift = (uint32_t)(32 - (bits - 5) - 2);
    A.fmask = (uint32_t)(((1 << (bits - 5)) - 1) << 2);
}

// double hashing on a prime-size table: home slot and a key-dependent step in [1, kSlots-1]
__device__ __forceinline__ uint32_t ht_home(uint32_t key, uint32_t tsize) { return __umulhi(key * 2654435769u, tsize); }
__device__ __forceinline__ uint32_t ht_step(uint32_t key, uint32_t tsize) { return 1u + __umulhi(key * 0xc2b2ae35u, tsize - 1u); }

// one-hash bit filter: word from the top bits of the hash, bit from its low 5 bits
__device__ __forceinline__ uint32_t flt_hash(uint32_t key) { return key * 0x85ebca6bu; }
__device__ __forceinline__ uint32_t flt_word(uint32_t h, const SAddr& A) { return (h >> A.fshift) & A.fmask; }  // byte offset
__device__ __forceinline__ uint32_t flt_need(uint32_t h) { return 1u << (h & 31); }

__device__ __forceinline__ void ht_insert(BuildSmem& S, const SAddr& A, uint32_t key, uint32_t bit) {
    const uint32_t h = flt_hash(key);
    atomicOr(&S.filter[flt_word(h, A) >> 2], flt_need(h));
    uint32_t slot = ht_home(key, A.tsize);
    const uint32_t step = ht_step(key, A.tsize);
    while (true) {
        const uint32_t prev = atomicCAS(&S.table[slot].x, kEmpty, key);
        if (prev == kEmpty || prev == key) {
            atomicOr(&S.table[slot].y, bit);
            return;
        }
        slot += step;
        if (slot >= A.tsize) slot -= A.tsize;
    }
}

__device__ __forceinline__ uint32_t ht_lookup(const SAddr& A, uint32_t key) {
    uint32_t slot = ht_home(key, A.tsize);
    const uint32_t step = ht_step(key, A.tsize);
    while (true) {
        const uint2 e = lds64(A.table + slot * 8u);
        if (e.x == key) return e.y;
        if (e.x == kEmpty) return 0u;
        slot += step;
        if (slot >= A.tsize) slot -= A.tsize;
    }
}

// word of the filter with the key's bit shifted down to bit 0 (bits above it are garbage)
__device__ __forceinline__ uint32_t flt_probe(const SAddr& A, uint32_t key) {
    const uint32_t h = flt_hash(key);
    return lds32(A.filter + flt_word(h, A)) >> (h & 31);
}
__device__ __forceinline__ bool flt_test(const SAddr& A, uint32_t key) { return (flt_probe(A, key) & 1u) != 0u; }

__device__ __forceinline__ void add_bits(unsigned* acc, uint32_t mask) {
    while (mask) {
        const int b = __ffs(mask) - 1;
        atomicAdd(&acc[b], 1u);
        mask &= mask - 1;
    }
}

// look up the first `count` (<= 32) queue entries of this warp, one per lane
__device__ __forceinline__ void q_lookup(BuildSmem& S, const SAddr& A, uint32_t sq, int lane, int count) {
    if (lane < count) {
        const uint32_t w = lds32(sq + lane * 4u);
        add_bits(S.acc3[w >> kTagShift], ht_lookup(A, w & ((1u << kTagShift) - 1u)));
    }
}

// after a push: while at least 32 entries are queued, look 32 of them up and shift the rest down
__device__ __forceinline__ void q_service(BuildSmem& S, const SAddr& A, uint32_t sq, int lane, int& tail) {
    while (tail >= 32) {
        __syncwarp();
        q_lookup(S, A, sq, lane, 32);
        __syncwarp();
        const int rem = tail - 32;
        for (int base = 0; base < rem; base += 32) {  // move entries [32, tail) to the front, 32 at a time
            uint32_t k = 0;
            if (base + lane < rem) k = lds32(sq + (32 + base + lane) * 4u);
            __syncwarp();
            if (base + lane < rem) sts32(sq + (base + lane) * 4u, k);
        }
        tail = rem;
        __syncwarp();
    }
}

__device__ __forceinline__ void q_drain(BuildSmem& S, const SAddr& A, uint32_t sq, int lane, int& tail) {
    q_service(S, A, sq, lane, tail);
    __syncwarp();
    q_lookup(S, A, sq, lane, tail);
    __syncwarp();
    tail = 0;
}

// One step of the frontier walk: K x 32 consecutive columns starting at b2 (K independent 128-byte
// loads in flight per warp), filter test, survivors of link slot `tag` compacted into the warp's queue.
template <int K>
__device__ __forceinline__ void walk_step(BuildSmem& S, const SAddr& A, uint32_t sq, const int32_t* __restrict__ rowp,
                                          int d, uint32_t tag, int b2, int lane, int& tail) {
    int32_t l[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int idx = b2 + k * 32 + lane;
        l[k] = idx < d ? ldg_i32(rowp + idx) : -1;
    }
    if (K == 1) {
        const bool pass = (l[0] >= 0) && flt_test(A, (uint32_t)l[0]);
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (bal == 0u) return;
        if (pass) sts32(sq + (uint32_t)(tail + __popc(bal & ((1u << lane) - 1u))) * 4u, (uint32_t)l[0] | tag);
        tail += __popc(bal);
        q_service(S, A, sq, lane, tail);
        return;
    }
    // funnel-shift accumulate: after K steps the K filter bits sit in the top K bits, first column lowest
    unsigned acc = 0u;
#pragma unroll
    for (int k = 0; k < K; ++k) acc = __funnelshift_r(acc, flt_probe(A, (uint32_t)l[k]), 1);
    int nvalid = (d - b2 - lane + 31) >> 5;  // columns of this lane inside the row (the invalid ones trail)
    nvalid = nvalid < 0 ? 0 : (nvalid > K ? K : nvalid);
    const unsigned flags = (acc >> (32 - K)) & ((1u << nvalid) - 1u);
    if (!__any_sync(0xffffffffu, flags != 0u)) return;
    // exclusive warp scan of the per-lane survivor counts
    const int cnt = __popc(flags);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (tail + total <= kQueue) {
        uint32_t off = sq + (uint32_t)(tail + incl - cnt) * 4u;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (flags & (1u << k)) {
                sts32(off, (uint32_t)l[k] | tag);
                off += 4u;
            }
        }
        tail += total;
        q_service(S, A, sq, lane, tail);
    } else {
        // rare: more survivors than queue space; push one column slot at a time
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const bool p = (flags >> k) & 1u;
            const unsigned bal = __ballot_sync(0xffffffffu, p);
            if (p) sts32(sq + (uint32_t)(tail + __popc(bal & ((1u << lane) - 1u))) * 4u, (uint32_t)l[k] | tag);
            tail += __popc(bal);
            q_service(S, A, sq, lane, tail);
        }
    }
}

// stream columns [start, end) of one frontier row, 256 at a time; the last partial piece of the row
// takes a narrower step
__device__ __forceinline__ void walk_row(BuildSmem& S, const SAddr& A, uint32_t sq, const int32_t* __restrict__ rowp,
                                         int d, int e, int start, int end, int lane, int& tail) {
    const uint32_t tag = (uint32_t)e << kTagShift;
    for (int b2 = start; b2 < end; b2 += 256) {
        const int rem = d - b2;
        if (rem > 128) walk_step<8>(S, A, sq, rowp, d, tag, b2, lane, tail);
        else if (rem > 32) walk_step<4>(S, A, sq, rowp, d, tag, b2, lane, tail);
        else walk_step<1>(S, A, sq, rowp, d, tag, b2, lane, tail);
    }
}

__global__ void __launch_bounds__(kBuildThreads, kBuildCtasPerSm)
k_cn_build(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
           const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int order,
           const int64_t* __restrict__ rec_off, const int32_t* __restrict__ run_start,
           const int64_t* __restrict__ run_unit_off, const int64_t* __restrict__ cost_pre,
           int64_t* __restrict__ plan, Record* __restrict__ records) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BuildSmem& S = *reinterpret_cast<BuildSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    SAddr A;
    A.tsize = (uint32_t)kSlots;
    A.fshift = 0u;
    A.fmask = 0u;
    A.table = sbase + (uint32_t)offsetof(BuildSmem, table);
    A.filter = sbase + (uint32_t)offsetof(BuildSmem, filter);
    A.queue = sbase + (uint32_t)offsetof(BuildSmem, qkey);
    const uint32_t sq = A.queue + (uint32_t)warp * (kQueue * 4u);
    if (order <= 2 && plan[OCN_PLAN_USE_DIRECT] != 0) return;  // the table-free kernel handles this stream
    const int64_t n_units = plan[OCN_PLAN_NUM_UNITS];
    const int64_t n_runs = plan[OCN_PLAN_NUM_RUNS];
    const int64_t W = plan[OCN_PLAN_BUDGET];
    int tail = 0;  // this warp's queue fill (warp-uniform)

    while (true) {
        __syncthreads();
        if (tid == 0) S.unit = (long long)atomicAdd((unsigned long long*)&plan[OCN_PLAN_UNIT_COUNTER], 1ull);
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
        // cost windows of the run: window s covers run-local cost [s*W, (s+1)*W)
        const int64_t cbase = cost_pre[t0];
        const int64_t run_cost = cost_pre[t0 + len] - cbase;
        const int64_t n_win = (run_cost + W - 1) / W;
        const int64_t q = local / n_win, win = local - q * n_win;  // q = global table-pass index of this run
        const int64_t c_lo = win * W, c_hi = (c_lo + W < run_cost) ? c_lo + W : run_cost;
        // links whose cost interval [P_t, P_t + c_t) meets the window
        int64_t t_first, t_last;
        {
            int64_t a = t0, b = t0 + len;  // first t with cost_pre[t+1] - cbase > c_lo
            while (a < b) {
                const int64_t mid = (a + b) >> 1;
                if (cost_pre[mid + 1] - cbase > c_lo) b = mid; else a = mid + 1;
            }
            t_first = a;
            a = t0; b = t0 + len;          // first t with cost_pre[t] - cbase >= c_hi
            while (a < b) {
                const int64_t mid = (a + b) >> 1;
                if (cost_pre[mid] - cbase >= c_hi) b = mid; else a = mid + 1;
            }
            t_last = a - 1;
        }

        // warp 0: locate (chunk, pass) of q by walking the chunks of N(i); load the chunk's rows
        if (warp == 0) {
            int64_t rem = q;
            int chunk = 0;
            while (true) {
                const int64_t p = (int64_t)chunk * kPChunk + lane;
                int kd = 0, k = 0;
                long long krs = 0;
                if (p < d) {
                    k = col[rs_i + p];
                    krs = rowptr[k];
                    kd = (int)(rowptr[k + 1] - krs);
                }
                int incl = kd;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const int tot = __shfl_sync(0xffffffffu, incl, 31);
                const int np_pass = (tot + kCap - 1) / kCap;
                if (rem < np_pass || (int64_t)(chunk + 1) * kPChunk >= d) {
                    S.kp[lane] = k;
                    S.krs[lane] = krs;
                    S.kdeg[lane] = kd;
                    S.kpre[lane] = incl - kd;
                    if (lane == 31) S.kpre[32] = incl;
                    if (lane == 0) { S.chunk = chunk; S.pass = (int)rem; }
                    break;
                }
                rem -= np_pass;
                ++chunk;
            }
        }
        __syncthreads();

        const int64_t p0 = (int64_t)S.chunk * kPChunk;
        const int np = (int)((d - p0) < kPChunk ? (d - p0) : kPChunk);
        const int total_keys = S.kpre[32];
        const int k_lo = S.pass * kCap, k_hi = (k_lo + kCap < total_keys) ? k_lo + kCap : total_keys;
        pick_geometry(k_hi - k_lo, A);
        for (int s = tid; s < (int)(A.tsize + 1) / 2; s += kBuildThreads)
            reinterpret_cast<uint4*>(S.table)[s] = make_uint4(kEmpty, 0u, kEmpty, 0u);
        for (int s = tid; s <= (int)(A.fmask >> 2); s += kBuildThreads) S.filter[s] = 0u;
        __syncthreads();
        // keys [k_lo, k_hi) of the flattened list, an equal share per warp; inside a share the warp walks
        // the rows it spans with coalesced loads (no per-key search, no barrier imbalance)
        {
            const int per = (k_hi - k_lo + kBuildWarps - 1) / kBuildWarps;
            int pos = k_lo + warp * per;
            const int end = (pos + per < k_hi) ? pos + per : k_hi;
            if (pos < end) {
                int a = 0, b = np;  // last row a with kpre[a] <= pos
                while (b - a > 1) {
                    const int mid = (a + b) >> 1;
                    if (S.kpre[mid] <= pos) a = mid; else b = mid;
                }
                while (pos < end) {
                    const int seg_end = S.kpre[a + 1] < end ? S.kpre[a + 1] : end;
                    const int32_t* rowp = col + S.krs[a] - S.kpre[a];
                    for (int idx = pos + lane; idx < seg_end; idx += 32) ht_insert(S, A, (uint32_t)ldg_i32(rowp + idx), 1u << a);
                    pos = seg_end;
                    ++a;
                }
            }
        }

        // the links of the window, kEdgeSub slots at a time, all against the same table
        for (int64_t g0 = t_first; g0 <= t_last; g0 += kEdgeSub) {
            const int ne = (int)((t_last - g0 + 1) < kEdgeSub ? (t_last - g0 + 1) : kEdgeSub);
            __syncthreads();  // table complete / previous group flushed
            if (warp == 1) {
                int dj = 0;
                unsigned first_piece = 0u;
                if (lane < ne) {
                    const int64_t t = g0 + lane;
                    const int64_t j = dst[t];
                    const int64_t rs_j = rowptr[j];
                    const int64_t dfull = (order >= 2) ? (rowptr[j + 1] - rs_j) : 0;
                    const int64_t P = cost_pre[t] - cbase, c = cost_pre[t + 1] - cost_pre[t];
                    const int64_t a = P > c_lo ? P : c_lo, b = (P + c) < c_hi ? (P + c) : c_hi;
                    const int64_t row_lo = (a - P) * dfull / c;
                    const int64_t row_hi = (b == P + c) ? dfull : (b - P) * dfull / c;
                    first_piece = (a == P) ? 1u : 0u;
                    S.jrs[lane] = rs_j + row_lo;
                    dj = (int)(row_hi - row_lo);
                    S.jdeg[lane] = dj;
                }
                S.m1[lane] = first_piece;  // bit 0 = "this unit owns the order-1 lookup of the link" (replaced below)
                int incl = dj;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                S.jpre[lane] = incl - dj;
                if (lane == 31) S.jpre[kEdgeSub] = incl;
            }
            for (int s = tid; s < kEdgeSub * 32; s += kBuildThreads) {
                (&S.acc2[0][0])[s] = 0u;
                (&S.acc3[0][0])[s] = 0u;
            }
            __syncthreads();

            // order 1: is j itself a key?  (one lookup per link, by the unit holding the link's first piece)
            if (tid < ne) S.m1[tid] = S.m1[tid] ? ht_lookup(A, (uint32_t)dst[g0 + tid]) : 0u;
            if (order >= 2) {
                // the rows N(m), m in the N(j) slices of the group's links, kRowBatch at a time
                const int n_rows = S.jpre[kEdgeSub];
                for (int r0 = 0; r0 < n_rows; r0 += kRowBatch) {
                    const int nb = (n_rows - r0) < kRowBatch ? (n_rows - r0) : kRowBatch;
                    // describe the rows: m feeds C2; tiny rows are walked right here, one thread each;
                    // the others are cut into chunks of kChunk columns
                    int chunks = 0;
                    if (tid < nb) {
                        const int item = r0 + tid;
                        int e = 0;  // last link slot with jpre[e] <= item
#pragma unroll
                        for (int s2 = kEdgeSub / 2; s2 > 0; s2 >>= 1)
                            if (S.jpre[e + s2] <= item) e += s2;
                        const int32_t m = ldg_i32(col + S.jrs[e] + (item - S.jpre[e]));
                        if (flt_test(A, (uint32_t)m)) add_bits(S.acc2[e], ht_lookup(A, (uint32_t)m));
                        int dm = 0;
                        uint32_t rs32 = 0u;
                        if (order >= 3) {
                            const int64_t rs_m = ldg_i64(rowptr + m);
                            dm = (int)(ldg_i64(rowptr + m + 1) - rs_m);
                            rs32 = (uint32_t)rs_m;
                            if (dm <= kTinyRow) {
                                int32_t l[kTinyRow];  // all loads first: one memory round trip per thread
#pragma unroll
                                for (int u = 0; u < kTinyRow; ++u) l[u] = u < dm ? ldg_i32(col + rs_m + u) : -1;
#pragma unroll
                                for (int u = 0; u < kTinyRow; ++u)
                                    if (l[u] >= 0 && flt_test(A, (uint32_t)l[u])) add_bits(S.acc3[e], ht_lookup(A, (uint32_t)l[u]));
                            } else {
                                chunks = (dm - 1) / kChunk;  // EXTRA chunks beyond the first kChunk columns
                            }
                        }
                        S.rrs[tid] = rs32;
                        S.rd[tid] = dm;
                        S.re[tid] = (unsigned char)e;
                    }
                    if (order < 3) continue;
                    // CTA-wide exclusive scan of the chunk counts
                    int incl = chunks;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    if (lane == 31) S.wsum[warp] = incl;
                    __syncthreads();
                    if (warp == 0) {
                        const int wv = S.wsum[lane];
                        int wi = wv;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_up_sync(0xffffffffu, wi, o);
                            if (lane >= o) wi += v;
                        }
                        S.wsum[lane] = wi - wv;
                        if (lane == 31) { S.n_chunks = wi; S.next_item = 0; }
                    }
                    __syncthreads();
                    if (tid < nb) S.rcp[tid] = incl - chunks + S.wsum[warp];
                    if (tid == 0) S.rcp[nb] = S.n_chunks;
                    __syncthreads();
                    // warps pull items from a shared counter until the batch is done: item c < nb is the first
                    // kChunk columns of row c (no search), the items after that are the extra chunks of long rows
                    const int n_items = nb + S.n_chunks;
                    while (true) {
                        int c = 0;
                        if (lane == 0) c = atomicAdd(&S.next_item, 1);
                        c = __shfl_sync(0xffffffffu, c, 0);
                        if (c >= n_items) break;
                        int a = c, start = 0;
                        if (c >= nb) {
                            const int x = c - nb;
                            int lo2 = 0, hi2 = nb;  // last row with rcp[row] <= x
                            while (hi2 - lo2 > 1) {
                                const int mid = (lo2 + hi2) >> 1;
                                if (S.rcp[mid] <= x) lo2 = mid; else hi2 = mid;
                            }
                            a = lo2;
                            start = (x - S.rcp[a] + 1) * kChunk;
                        }
                        const int d = S.rd[a];
                        if (d <= kTinyRow) continue;  // walked by its describing thread
                        const int end = (start + kChunk < d) ? start + kChunk : d;
                        walk_row(S, A, sq, col + S.rrs[a], d, (int)S.re[a], start, end, lane, tail);
                    }
                    q_drain(S, A, sq, lane, tail);
                    __syncthreads();  // descriptors are reused by the next batch
                }
            }
            __syncthreads();
            // add this group's partial counts into the records
            for (int s = tid; s < ne * 32; s += kBuildThreads) {
                const int e = s >> 5, p = s & 31;
                if (p < np) {
                    const unsigned c1 = (S.m1[e] >> p) & 1u;
                    const unsigned c2 = S.acc2[e][p], c3 = S.acc3[e][p];
                    unsigned* rec = reinterpret_cast<unsigned*>(records + rec_off[g0 + e] + p0 + p);
                    const unsigned x = c2 | (c1 << 31);
                    if (x) atomicAdd(rec, x);
                    if (c3) atomicAdd(rec + 1, c3);
                }
            }
        }
    }
}

// Orders 1 and 2 need no table: C1[p] = [N(i)[p] in N(j)], C2[p] = |N(j) (cap) N(N(i)[p])| are plain sorted-list
// intersections.  The work is flattened over the RECORDS: a warp takes 32 consecutive (link, position) pairs of
// the stream -- one per lane, whichever links they belong to -- so that a link with a 5 000-neighbour source is
// spread over 157 warps instead of being walked by one (training batches draw their links from the edges, i.e.
// degree-biased: one warp per link left a tail of tens of milliseconds per 2048-link batch).  Pass 1: the lane
// walks the shorter of N(j) and N(k_p) and binary-searches the longer one (resuming where the previous search
// ended); pairs whose shorter row exceeds 32 columns are deferred.  Pass 2: the deferred pairs are intersected by
// the whole warp (lanes stride over the shorter row), so a hub x hub pair costs len/32 * log steps instead of
// len * log on one lane.  Records are written directly (no atomics): each (link, p) has one owner.
// (row lengths and positions inside a row are 32-bit here: the search steps are most of this kernel's 4.9 G warp
// instructions on a 16 384-link training batch, and 64-bit index arithmetic was a third of every step)
__device__ __forceinline__ bool row_has(const int32_t* __restrict__ row, int len, int32_t key) {
    int lo = 0, hi = len;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(row + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo < len && __ldg(row + lo) == key;
}

constexpr int kMergeRatio = 6;  // deferred pairs up to this length ratio are merged, longer ones searched
__global__ void __launch_bounds__(256)
k_cn_build_direct(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int64_t* __restrict__ src,
                  const int64_t* __restrict__ dst, int64_t T, int order, const int64_t* __restrict__ rec_off,
                  const int64_t* __restrict__ plan, Record* __restrict__ records) {
    if (plan[OCN_PLAN_USE_DIRECT] == 0) return;  // the table kernel handles this stream
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    const int64_t total = rec_off[T];
    for (int64_t g0 = warp * 32; g0 < total; g0 += nwarps * 32) {
        const int64_t g = g0 + lane;
        bool defer = false;
        int64_t rs_k = 0, rs_j = 0;
        int dk = 0, dj = 0;
        unsigned c1 = 0u;
        if (g < total) {
            int64_t lo_t = 0, hi_t = T;  // link of record g: the last t with rec_off[t] <= g (links without records share offsets)
            while (hi_t - lo_t > 1) {
                const int64_t mid = (lo_t + hi_t) >> 1;
                if (ldg_i64(rec_off + mid) <= g) lo_t = mid; else hi_t = mid;
            }
            const int64_t t = lo_t, p = g - ldg_i64(rec_off + t);
            const int64_t i = src[t], j = dst[t];
            rs_j = ldg_i64(rowptr + j);
            dj = (int)(ldg_i64(rowptr + j + 1) - rs_j);
            const int32_t* nj = col + rs_j;
            const int32_t k = ldg_i32(col + ldg_i64(rowptr + i) + p);
            c1 = row_has(nj, dj, k) ? 1u : 0u;
            unsigned c2 = 0u;
            if (order >= 2) {
                rs_k = ldg_i64(rowptr + k);
                dk = (int)(ldg_i64(rowptr + k + 1) - rs_k);
                const int32_t* a = nj;          // shorter row
                const int32_t* b = col + rs_k;  // longer row
                int la = dj, lb = dk;
                if (la > lb) {
                    const int32_t* tp = a; a = b; b = tp;
                    const int tl = la; la = lb; lb = tl;
                }
                if (la > 32) {
                    defer = true;
                } else {
                    int lo = 0;  // both rows ascend: every search resumes where the previous one ended
                    for (int u = 0; u < la && lo < lb; ++u) {
                        const int32_t v = ldg_i32(a + u);
                        int hi = lb;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (ldg_i32(b + mid) < v) lo = mid + 1; else hi = mid;
                        }
                        if (lo < lb && ldg_i32(b + lo) == v) { ++c2; ++lo; }
                    }
                }
            }
            if (!defer) records[g] = make_uint2(c2 | (c1 << 31), 0u);
        }
        unsigned pending = __ballot_sync(0xffffffffu, defer);
        while (pending) {
            const int sl = __ffs(pending) - 1;
            pending &= pending - 1;
            const int64_t rk = __shfl_sync(0xffffffffu, rs_k, sl), rj = __shfl_sync(0xffffffffu, rs_j, sl);
            const int dkk = __shfl_sync(0xffffffffu, dk, sl), djj = __shfl_sync(0xffffffffu, dj, sl);
            const unsigned cc1 = __shfl_sync(0xffffffffu, c1, sl);
            const int32_t* a = col + rj;
            const int32_t* b = col + rk;
            int la = djj, lb = dkk;
            if (la > lb) {
                const int32_t* tp = a; a = b; b = tp;
                const int tl = la; la = lb; lb = tl;
            }
            unsigned cnt = 0u;
            if (lb <= kMergeRatio * la) {
                // rows of similar length (two thirds of the deferred search steps on a training batch are pairs within a
                // factor of four): a lane takes a contiguous share of the shorter row, finds where it starts in the
                // longer one and merges forward -- (la + lb) / 32 steps a lane instead of la / 32 * log2 lb
                const int share = (la + 31) >> 5, ua = lane * share, ue = ua + share < la ? ua + share : la;
                if (ua < ue) {
                    int u = ua;
                    int32_t va = ldg_i32(a + u);
                    int lo = 0, hi = lb;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (ldg_i32(b + mid) < va) lo = mid + 1; else hi = mid;
                    }
                    int32_t vb = lo < lb ? ldg_i32(b + lo) : 0;
                    while (u < ue && lo < lb) {
                        const bool step_a = va <= vb, step_b = vb <= va;
                        cnt += (step_a && step_b) ? 1u : 0u;
                        if (step_a && ++u < ue) va = ldg_i32(a + u);
                        if (step_b && ++lo < lb) vb = ldg_i32(b + lo);
                    }
                }
            } else {
                for (int u = lane; u < la; u += 32) cnt += row_has(b, lb, ldg_i32(a + u)) ? 1u : 0u;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0) records[g0 + sl] = make_uint2(cnt | (cc1 << 31), 0u);
        }
    }
}

// The link a record belongs to (the last t with rec_off[t] <= g; links without records share offsets), for a lane that
// walks ascending records: the current link's bounds stay in registers, so a step inside one link costs a compare, a step
// into one of the next few links a load each, and only the first call (or a jump over many links) a binary search.  (A
// search from scratch per record was 590 warp instructions per 32 records in k_cn_build_from_a2: the ALU pipe was its limit.)
struct RecordLink {
    int64_t t = -1, off = 0, next = 0;  // rec_off[t] = off <= g < next = rec_off[t + 1]
    __device__ __forceinline__ bool seek(const int64_t* __restrict__ rec_off, int64_t T, int64_t g) {  // true: another link
        if (t >= 0 && g < next) return false;
        int64_t lo = t < 0 ? 0 : t;
        bool found = false;
        if (t >= 0)
            for (int s = 0; s < 4 && !found; ++s) {
                if (ldg_i64(rec_off + lo + 1) <= g) ++lo; else found = true;
            }
        if (!found) {
            int64_t hi = T;
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (ldg_i64(rec_off + mid) <= g) lo = mid; else hi = mid;
            }
        }
        t = lo;
        off = ldg_i64(rec_off + lo);
        next = ldg_i64(rec_off + lo + 1);
        return true;
    }
};

// Orders 1-2 on a dense graph: a warp takes 32 consecutive records (link, position p), finds their links with one
// search per lane, then computes four records at a time with eight lanes over the words of the two bit-vector
// rows: C2 = popc(row(dst) & row(k_p)) summed over the words, C1 = bit dst of row(k_p).
__global__ void __launch_bounds__(256)
k_cn_build_dense(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int64_t* __restrict__ src,
                 const int64_t* __restrict__ dst, int64_t T, int order, const int64_t* __restrict__ rec_off,
                 const uint32_t* __restrict__ bits, int W, Record* __restrict__ records) {
    const int lane = lane_id();
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5, warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t total = rec_off[T];
    const int64_t span = ((total + nwarps * 32 - 1) / (nwarps * 32)) * 32;  // a warp's records are consecutive
    const int64_t end = (warp + 1) * span < total ? (warp + 1) * span : total;
    RecordLink L;
    int64_t j = -1, rs = 0;
    for (int64_t g0 = warp * span; g0 < end; g0 += 32) {
        const int64_t g = g0 + lane;
        int32_t k = 0;
        if (g < total) {
            if (L.seek(rec_off, T, g)) {
                j = dst[L.t];
                rs = ldg_i64(rowptr + src[L.t]);
            }
            k = ldg_i32(col + rs + (g - L.off));
        }
        const int cnt = (int)((total - g0) < 32 ? (total - g0) : 32);
        unsigned mine = 0u;
        // four records at a time, eight lanes a record, 128-bit loads (W is a multiple of 4 words, rows are 16-byte
        // aligned): 4 x 128 B per request.  One record at a time with the warp across 32-bit words cost 205 warp
        // instructions a record at ddi (W = 136), most of them the per-record shuffle reduction and loop control.
        const int sub = lane & 7, grp = lane >> 3, W4 = W >> 2;
        for (int q0 = 0; q0 < cnt; q0 += 4) {
            const int q = q0 + grp;
            const int64_t jq = __shfl_sync(0xffffffffu, j, q & 31);
            const int32_t kq = __shfl_sync(0xffffffffu, k, q & 31);
            const bool valid = q < cnt;
            unsigned c2 = 0u;
            if (valid && order >= 2) {
                const uint4* rj = reinterpret_cast<const uint4*>(bits + jq * W);
                const uint4* rk = reinterpret_cast<const uint4*>(bits + (int64_t)kq * W);
#pragma unroll 2
                for (int w = sub; w < W4; w += 8) {
                    const uint4 a = __ldg(rj + w), b = __ldg(rk + w);
                    c2 += __popc(a.x & b.x) + __popc(a.y & b.y) + __popc(a.z & b.z) + __popc(a.w & b.w);
                }
            }
            c2 += __shfl_xor_sync(0xffffffffu, c2, 1);
            c2 += __shfl_xor_sync(0xffffffffu, c2, 2);
            c2 += __shfl_xor_sync(0xffffffffu, c2, 4);
            unsigned val = 0u;
            if (valid && sub == 0) val = c2 | (((__ldg(bits + (int64_t)kq * W + (jq >> 5)) >> (jq & 31)) & 1u) << 31);
            const unsigned got = __shfl_sync(0xffffffffu, val, ((lane - q0) & 3) * 8);
            if (lane >= q0 && lane < q0 + 4) mine = got;
        }
        if (g < total) records[g] = make_uint2(mine, 0u);
    }
}

// The whole 2-walk count matrix of a dense symmetric graph, a2[j][k] = popc(row(j) & row(k)): a 64 x 64 tile per CTA,
// 4 x 4 entries per thread (rows ty + 16 a, columns tx + 16 b: consecutive lanes read consecutive shared-memory rows of
// 9 x 16 bytes, conflict free), the bit rows staged 32 words at a time.  Population counts are a quarter-rate
// instruction, so the four words of a 128-bit step go through a carry-save adder tree first (ones / twos carried along
// the row, one count of the fours per step): 13 instructions per entry and step.
constexpr int kA2Tile = 64, kA2Chunk4 = 8, kA2Stride = kA2Chunk4 + 1;
__global__ void __launch_bounds__(256)
k_dense_a2(const uint32_t* __restrict__ bits, int64_t n, int W, uint32_t* __restrict__ a2,
           const long long* __restrict__ skip_stamp, long long e0, long long e1, long long e2, long long e3) {
    // (spgemm.cu: the matrix of this graph is already in a2 when the scratch carries its stamp)
    if (skip_stamp != nullptr && skip_stamp[0] == e0 && skip_stamp[1] == e1 && skip_stamp[2] == e2 && skip_stamp[3] == e3) return;
    __shared__ uint4 sa[kA2Tile * kA2Stride], sb[kA2Tile * kA2Stride];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    if (blockIdx.x < blockIdx.y) return;  // written by the CTA of the mirror tile
    const int64_t j0 = (int64_t)blockIdx.y * kA2Tile, k0 = (int64_t)blockIdx.x * kA2Tile;
    unsigned ones[4][4], twos[4][4], fours[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) ones[a][b] = twos[a][b] = fours[a][b] = 0u;
    const int W4 = W >> 2;
    for (int c0 = 0; c0 < W4; c0 += kA2Chunk4) {
        for (int e = threadIdx.x; e < kA2Tile * kA2Chunk4; e += 256) {
            const int r = e >> 3, c = e & 7;
            uint4 va = make_uint4(0u, 0u, 0u, 0u), vb = va;
            if (c0 + c < W4) {
                if (j0 + r < n) va = __ldg(reinterpret_cast<const uint4*>(bits + (j0 + r) * W) + c0 + c);
                if (k0 + r < n) vb = __ldg(reinterpret_cast<const uint4*>(bits + (k0 + r) * W) + c0 + c);
            }
            sa[r * kA2Stride + c] = va;
            sb[r * kA2Stride + c] = vb;
        }
        __syncthreads();
#pragma unroll 2
        for (int c = 0; c < kA2Chunk4; ++c) {
            uint4 va[4], vb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                va[a] = sa[(ty + 16 * a) * kA2Stride + c];
                vb[a] = sb[(tx + 16 * a) * kA2Stride + c];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const unsigned v0 = va[a].x & vb[b].x, v1 = va[a].y & vb[b].y, v2 = va[a].z & vb[b].z, v3 = va[a].w & vb[b].w;
                    const unsigned o = ones[a][b], t = twos[a][b];
                    const unsigned ta = (o & v0) | ((o ^ v0) & v1), o1 = o ^ v0 ^ v1;
                    const unsigned tb = (o1 & v2) | ((o1 ^ v2) & v3);
                    ones[a][b] = o1 ^ v2 ^ v3;
                    fours[a][b] += __popc((t & ta) | ((t ^ ta) & tb));
                    twos[a][b] = t ^ ta ^ tb;
                }
        }
        __syncthreads();
    }
    unsigned out[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) out[a][b] = 4u * fours[a][b] + 2u * __popc(twos[a][b]) + __popc(ones[a][b]);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t j = j0 + ty + 16 * a;
        if (j >= n) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t k = k0 + tx + 16 * b;
            if (k < n) a2[j * n + k] = out[a][b];
        }
    }
    if (blockIdx.x == blockIdx.y) return;
    // the mirror tile (the graph is symmetric): transposed through shared memory so that its rows are written whole
    __shared__ uint32_t st[kA2Tile][kA2Tile + 1];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) st[tx + 16 * b][ty + 16 * a] = out[a][b];
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int64_t k = k0 + ty + 16 * a;
        if (k >= n) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int64_t j = j0 + tx + 16 * b;
            if (j < n) a2[k * n + j] = st[ty + 16 * a][tx + 16 * b];
        }
    }
}

// records of a dense graph from the whole matrix: one lane per record, C2 = a2[dst][k_p], C1 = bit k_p of row(dst)
__global__ void __launch_bounds__(256)
k_cn_build_from_a2(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int64_t* __restrict__ src,
                   const int64_t* __restrict__ dst, int64_t T, int order, const int64_t* __restrict__ rec_off,
                   const uint32_t* __restrict__ bits, int W, const uint32_t* __restrict__ a2, int64_t n,
                   Record* __restrict__ records) {
    const int64_t total = rec_off[T];
    const int lane = lane_id();
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5, warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t span = ((total + nwarps * 32 - 1) / (nwarps * 32)) * 32;  // a warp's records are consecutive
    const int64_t end = (warp + 1) * span < total ? (warp + 1) * span : total;
    RecordLink L;
    int64_t j = 0, rs = 0;
    for (int64_t g = warp * span + lane; g < end; g += 32) {
        if (L.seek(rec_off, T, g)) {
            j = dst[L.t];
            rs = ldg_i64(rowptr + src[L.t]);
        }
        const int32_t k = ldg_i32(col + rs + (g - L.off));
        const unsigned c2 = order >= 2 ? __ldg(a2 + j * n + k) : 0u;
        const unsigned c1 = (__ldg(bits + j * W + (k >> 5)) >> (k & 31)) & 1u;
        records[g] = make_uint2(c2 | (c1 << 31), 0u);
    }
}

// per-batch column statistics from the finished records: one warp per link
__global__ void k_cn_colstat(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                             const int64_t* __restrict__ src, int64_t T, int64_t batch_size, int weighted,
                             const int64_t* __restrict__ rec_off, const Record* __restrict__ records,
                             ColStat* __restrict__ colstat, int64_t min_deg) {
    const int wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    int64_t warp = (int64_t)blockIdx.x * wpb + wib;
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int lane = lane_id();
    // positions p0, p0 + stride, ... of link t
    auto walk = [&](int64_t t, int64_t rs, int64_t d, int64_t p0, int64_t stride) {
        const int64_t ro = rec_off[t];
        ColStat* cs = colstat + (t / batch_size) * n;
        for (int64_t p = p0; p < d; p += stride) {
            const Record rec = records[ro + p];
            if (rec.x | rec.y) {
                ColStat* c = cs + ldg_i32(col + rs + p);
                const unsigned c2 = rec.x & 0x7fffffffu;
                if (rec.x >> 31) atomicAdd(&c->c1, 1u);
                const unsigned long long v2 = weighted ? c2 : (c2 ? 1u : 0u);
                const unsigned long long v3 = weighted ? rec.y : (rec.y ? 1u : 0u);
                if (v2) atomicAdd(&c->s2, v2);
                if (v3) atomicAdd(&c->s3, v3);
            }
        }
    };
    for (int64_t t = warp; t < T; t += nwarps) {  // one warp per link ...
        const int64_t i = src[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        if (d <= kHeavyLink && d > min_deg) walk(t, rs, d, lane, 32);
    }
    for_each_heavy_link(rowptr, src, T, rec_off, [&](int64_t t) {  // ... a whole CTA per link with a heavy source
        const int64_t i = src[t];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        walk(t, rs, d, threadIdx.x, blockDim.x);
    });
}

// Column statistics of a dense graph (orders <= 2): n is small and every batch touches every column hundreds of times, so
// the global atomics of k_cn_colstat pile up on a few thousand addresses (ddi: 25 M updates on 16 x 4267 entries, 262 us).
// Here a CTA takes a share of ONE batch's links, counts into a table in shared memory (c1 and s2 as 32-bit: the host checks
// that a CTA's share cannot overflow them) and adds its non-zero entries to the batch's statistics at the end.
__global__ void __launch_bounds__(256)
k_cn_colstat_table(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                   const int64_t* __restrict__ src, int64_t T, int64_t batch_size, int weighted, int ctas_per_batch,
                   const int64_t* __restrict__ rec_off, const Record* __restrict__ records, ColStat* __restrict__ colstat) {
    extern __shared__ uint32_t cs_tab[];  // [n] c1 | [n] s2
    const int lane = lane_id(), warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t b = blockIdx.x / ctas_per_batch, part = blockIdx.x % ctas_per_batch;
    const int64_t t_begin = b * batch_size, t_end = t_begin + batch_size < T ? t_begin + batch_size : T;
    for (int64_t k = threadIdx.x; k < 2 * n; k += blockDim.x) cs_tab[k] = 0u;
    __syncthreads();
    for (int64_t t = t_begin + part * wpb + warp; t < t_end; t += (int64_t)wpb * ctas_per_batch) {
        const int64_t i = src[t], rs = rowptr[i], ro = rec_off[t];
        const int d = (int)(rowptr[i + 1] - rs);
        for (int p0 = 0; p0 < d; p0 += 128) {  // four positions a lane in flight
            uint32_t rx[4];
            int32_t kk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int p = p0 + q * 32 + lane;
                rx[q] = p < d ? records[ro + p].x : 0u;
                kk[q] = p < d ? ldg_i32(col + rs + p) : 0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (rx[q] == 0u) continue;
                const uint32_t c2 = rx[q] & 0x7fffffffu;
                if (rx[q] >> 31) atomicAdd(&cs_tab[kk[q]], 1u);
                const uint32_t v2 = weighted ? c2 : (c2 ? 1u : 0u);
                if (v2) atomicAdd(&cs_tab[n + kk[q]], v2);
            }
        }
    }
    __syncthreads();
    ColStat* cs = colstat + b * n;
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
        const uint32_t c1 = cs_tab[k], s2 = cs_tab[n + k];
        if (c1) atomicAdd(&cs[k].c1, c1);
        if (s2) atomicAdd(&cs[k].s2, (unsigned long long)s2);
    }
}

}  // namespace ocn

using namespace ocn;

namespace ocn {
__global__ void k_records_spd(Record* __restrict__ records, int64_t num_records) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < num_records; r += stride) {
        const uint32_t x = records[r].x;
        if ((x >> 31) && (x & 0x7fffffffu)) records[r].x = 0x80000000u;
    }
}
}  // namespace ocn

extern "C" int ocn_cn_build(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src,
                            const int64_t* dst, int64_t num_edges, int64_t batch_size, int order, int weighted,
                            const void* plan_scratch, const int64_t* plan, void* records, int64_t records_capacity,
                            void* colstat, int64_t nnz, const int64_t* plan_host, void* hub_scratch,
                            size_t hub_scratch_bytes, void* node_scratch, void* stream) {
    OCN_RANGE("ocn_cn_build");
    OCN_CHECK_ARG(rowptr && col && src && dst && plan_scratch && plan, "ocn_cn_build: null pointer");
    OCN_CHECK_ARG(order >= 1 && order <= 3, "ocn_cn_build: order must be 1, 2 or 3 (got %d)", order);
    OCN_CHECK_ARG(n > 0 && num_edges > 0 && batch_size > 0, "ocn_cn_build: sizes must be positive");
    OCN_CHECK_ARG(records || records_capacity == 0, "ocn_cn_build: records is null");
    OCN_CHECK_ARG(plan_host == nullptr || plan_host[OCN_PLAN_BAD_LINKS] == 0,
                  "ocn_cn_build: the plan found %lld links with an endpoint outside [0, n)",
                  (long long)(plan_host ? plan_host[OCN_PLAN_BAD_LINKS] : 0));
    OCN_CHECK_ARG(n < (int64_t(1) << kTagShift), "ocn_cn_build: at most 2^27 nodes (queue tag width)");
    // row starts are kept as 32-bit offsets in shared memory; rowptr[n] < 2^32 is the caller's contract (checked by ocn_graph_validate users)
    static_assert(kEdgeSub == 32, "one lane per link slot");
    cudaStream_t st = (cudaStream_t)stream;
    PlanLayout L = plan_layout(num_edges);
    const char* base = (const char*)plan_scratch;
    const int64_t* rec_off = (const int64_t*)(base + L.rec_off);
    OCN_CUDA(cudaFuncSetAttribute(k_cn_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BuildSmem)));
    OCN_CUDA(cudaMemsetAsync((void*)(plan + OCN_PLAN_UNIT_COUNTER), 0, sizeof(int64_t), st));  // restart the dynamic unit counter
    if (records_capacity > 0)  // the table path accumulates with atomics
        OCN_CUDA(cudaMemsetAsync(records, 0, sizeof(Record) * (size_t)records_capacity, st));
    // order 3 on a stream with few runs: inverted-index path (cn_hub.cu) instead of the per-run tables
    bool indexed = false;
    if (order >= 3 && plan_host != nullptr && hub_scratch != nullptr && plan_host[OCN_PLAN_HUB_DEGREE] > 0) {
        OCN_CHECK_ARG(node_scratch != nullptr, "ocn_cn_build: hub stage needs node_scratch");
        OCN_CHECK_ARG(nnz > 0, "ocn_cn_build: hub stage needs nnz");
        int rc = run_hub_stage(rowptr, col, n, src, dst, num_edges, plan_scratch, plan, plan_host, hub_scratch,
                               hub_scratch_bytes, node_scratch, (Record*)records, nnz, st);
        if (rc != OCN_OK) return rc;
        indexed = true;
    }
    const int direct_known = plan_host != nullptr ? (plan_host[OCN_PLAN_USE_DIRECT] != 0 ? 1 : 0) : -1;  // (a launch saved on small batches)
    const bool dense = order <= 2 && plan_host != nullptr && plan_host[OCN_PLAN_DENSE] != 0 && hub_scratch != nullptr;
    if (dense) {
        const int W = dense_words(n);
        const size_t need = dense_whole_a2(n, plan_host) ? dense_bits_bytes(n) + sizeof(uint32_t) * (size_t)n * (size_t)n
                                                         : sizeof(uint32_t) * (size_t)n * (size_t)W;
        if (hub_scratch_bytes < need) return fail(OCN_ENOSPACE, "ocn_cn_build: dense scratch %zu < %zu bytes", hub_scratch_bytes, need);
        uint32_t* bits = (uint32_t*)hub_scratch;
        k_dense_bits<<<(int)((n * 32 + 255) / 256), 256, 0, st>>>(rowptr, col, n, W, bits);
        OCN_LAUNCH_CHECK();
        int64_t want = (records_capacity / 32 + 7) / 8 + 1;
        int64_t cap = (int64_t)sm_count() * 16;
        if (dense_whole_a2(n, plan_host)) {
            uint32_t* a2 = (uint32_t*)((char*)hub_scratch + dense_bits_bytes(n));
            const unsigned tiles = (unsigned)((n + kA2Tile - 1) / kA2Tile);
            if (order >= 2) k_dense_a2<<<dim3(tiles, tiles), 256, 0, st>>>(bits, n, W, a2, nullptr, 0, 0, 0, 0);
            OCN_LAUNCH_CHECK();
            const int64_t want1 = (records_capacity + 255) / 256 + 1;  // a lane per record
            k_cn_build_from_a2<<<(int)(want1 < cap ? want1 : cap), 256, 0, st>>>(rowptr, col, src, dst, num_edges, order, rec_off,
                                                                                bits, W, a2, n, (Record*)records);
        } else {
            k_cn_build_dense<<<(int)(want < cap ? want : cap), 256, 0, st>>>(rowptr, col, src, dst, num_edges, order, rec_off,
                                                                            bits, W, (Record*)records);
        }
        OCN_LAUNCH_CHECK();
        indexed = true;  // (skips the table kernel below)
    } else if (order <= 2 && direct_known != 0) {  // the plan picked one of the two on the device (plan[OCN_PLAN_USE_DIRECT]);
        // without the host copy of the plan both are launched and the other returns at once
        int64_t want = (records_capacity / 32 + 7) / 8 + 1;  // a warp per 32 records
        int64_t cap = (int64_t)sm_count() * 16;
        k_cn_build_direct<<<(int)(want < cap ? want : cap), 256, 0, st>>>(rowptr, col, src, dst, num_edges, order, rec_off,
                                                                         plan, (Record*)records);
        OCN_LAUNCH_CHECK();
    }
    if (!indexed && !(order <= 2 && direct_known == 1)) {
        const int blocks = sm_count() * kBuildCtasPerSm;
        k_cn_build<<<blocks, kBuildThreads, sizeof(BuildSmem), st>>>(
            rowptr, col, n, src, dst, order, rec_off, (const int32_t*)(base + L.run_start),
            (const int64_t*)(base + L.run_unit_off), (const int64_t*)(base + L.cost_pre), (int64_t*)plan,
            (Record*)records);
        OCN_LAUNCH_CHECK();
    }
    if (weighted == 2 && order >= 2 && records_capacity > 0) {
        // shortest-path variant (SPD.py:65-126): a 2-walk count only stands for nodes at distance exactly 2 from the
        // destination, i.e. C2[p] = 0 wherever k_p is itself a neighbour of it -- which is the C1 bit of the record
        const int64_t want = (records_capacity + 255) / 256, cap = (int64_t)sm_count() * 16;
        k_records_spd<<<(int)(want < cap ? want : cap), 256, 0, st>>>((Record*)records, records_capacity);
        OCN_LAUNCH_CHECK();
    }
    const bool grouped = colstat != nullptr && use_grouped(num_edges, plan_host);
    if (grouped)
        if (int rc = grouped_colstat(rowptr, col, n, src, num_edges, batch_size, weighted, plan_scratch, (const Record*)records,
                                     (ColStat*)colstat, st))
            return rc;
    // dense graphs: a shared-memory table per CTA (orders <= 2: no C3 sums; 8 B per column must fit, and a CTA's share of a
    // batch times the largest 2-walk count must stay below 2^32)
    const int64_t num_batches = (num_edges + batch_size - 1) / batch_size;
    int64_t ctas_per_batch = ((int64_t)sm_count() * 4 + num_batches - 1) / num_batches;
    if (ctas_per_batch > (batch_size + 7) / 8) ctas_per_batch = (batch_size + 7) / 8;
    if (ctas_per_batch < 1) ctas_per_batch = 1;
    const bool table = dense && colstat != nullptr && !grouped && 8 * n <= 96 * 1024 && num_batches * ctas_per_batch <= 65535 * 16 &&
                       (batch_size / ctas_per_batch + 8) * n < (int64_t)1 << 31;
    if (table) {
        OCN_CUDA(cudaFuncSetAttribute(k_cn_colstat_table, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * n)));
        k_cn_colstat_table<<<(int)(num_batches * ctas_per_batch), 256, (size_t)(8 * n), st>>>(
            rowptr, col, n, src, num_edges, batch_size, weighted, (int)ctas_per_batch, rec_off, (const Record*)records,
            (ColStat*)colstat);
        OCN_LAUNCH_CHECK();
        return OCN_OK;
    }
    if (colstat != nullptr && (!grouped || plan_host[OCN_PLAN_WIDE_LINKS] > 0)) {
        int64_t want = (num_edges + 7) / 8;
        int64_t cap = (int64_t)sm_count() * 8;
        k_cn_colstat<<<(int)(want < cap ? want : cap), 256, 0, st>>>(rowptr, col, n, src, num_edges, batch_size, weighted,
                                                                    rec_off, (const Record*)records, (ColStat*)colstat,
                                                                    grouped ? (int64_t)kGroupedMaxDeg : (int64_t)-1);
        OCN_LAUNCH_CHECK();
    }
    return OCN_OK;
}
