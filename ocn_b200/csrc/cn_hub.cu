// Indexed path of ocn_cn_build for order 3: CN_1..3 of a stream of links through an inverted index
// of the source side, with the rows shared by many links streamed once.
//
//     C1[t][p] = [ dst in N(k_p) ]
//     C2[t][p] = #{ m in N(dst)              : m in N(k_p) }
//     C3[t][p] = #{ m in N(dst), l in N(m)   : l in N(k_p) },          k_p = N(src[t])[p]
//
// cn_build.cu answers "l in N(k_p)?" from a shared-memory table per run of links with one source
// and walks N(m) once per (link, m).  In a stream of T links with spread-out destinations a row
// N(m) is walked by about T*d(m)/n links: for the citation2 evaluation stream
// (NeighborOverlapCitation2.py:241-254, every source against 1000 uniform destinations) 87 % of
// the walked columns sit in rows that several links of the same call walk, and a link's own part is
// ~150 columns.  This path therefore
//
//   * builds ONE inverted index for the whole stream:  entries(l) = { position pi = (run r, p) :
//     l in N(k_{r,p}) }, a list sorted by (l, pi); node_index[l] = (first, last+1, 64-bit signature
//     of the runs in the list).  The positions of a run are contiguous (run_pos_off[r] + p);
//   * k_cn_hub_count: every row N(m) with d(m) >= plan[OCN_PLAN_HUB_DEGREE] that some destination
//     touches is streamed ONCE; the entries of its columns are counted per position in warp-private
//     shared memory (U[pi] += 1: a shared-memory atomic costs about a load on sm_100a, a global RED
//     15x more -- scripts/micro/atoms_bench.cu), then every link t next to m (the pair list L_m)
//     adds U[positions of its run] into its records: one RED per non-zero counter, not per walk;
//   * k_cn_link: per link, C1, C2 and the C3 part through the short rows, by looking (l, run of t)
//     up in the same index (the run's entries are a contiguous piece of l's list).
//
// All sums are integer RED.ADDs into the zeroed records, so the result is exact and independent of
// the order of execution.  Sizes come from the plan the caller read back (pairs, entries, positions).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <map>
#include <mutex>
#include <set>
#include <string>
#include <utility>

#include "common.cuh"

namespace ocn {

constexpr int kHubSeg = 1024;       // columns of a shared row per work item (16-bit counters: < 65536); A/B: 1024 beats 2048 by 6 % on the kernel alone (tail balance), 512 and 256 lose to item overhead
constexpr int kHubThreads = 256;
constexpr int kHubWarps = kHubThreads / 32;
constexpr int kShortList = 4;       // entry lists up to this length are walked by their own lane
constexpr int kMidList = 32;        // up to this length flattened over the lanes, longer ones by the whole warp
constexpr int kHubQueue = 64;         // queued entry lists per warp (8 bytes each)
constexpr int kHubWindow = 4096;      // positions a warp-private counter array holds (8 KB of packed counters per warp)
constexpr int kHubCtaWindow = 32768;  // positions per pass of the CTA-per-item variant (64 KB of packed counters per CTA)
static_assert(kHubSeg < 65536, "16-bit walk counters per item");

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

static int key_bits(int64_t n) {  // bits of the keys 0 .. n (n itself is the "no pair" sentinel of the pair sort)
    int b = 1;
    while (b < 32 && (int64_t(1) << b) <= n) ++b;
    return b;
}

HubLayout hub_layout(int64_t n, int64_t nnz, int64_t pairs, int64_t entries, int64_t positions, int64_t chunks) {
    HubLayout L;
    size_t off = 0;
    const size_t P = (size_t)(pairs > 0 ? pairs : 1), E = (size_t)(entries > 0 ? entries : 1);
    for (int k = 0; k < 2; ++k) { L.pkey[k] = off; off += align256(sizeof(uint32_t) * P); }
    for (int k = 0; k < 2; ++k) { L.pval[k] = off; off += align256(sizeof(uint32_t) * P); }
    for (int k = 0; k < 2; ++k) { L.ekey[k] = off; off += align256(sizeof(uint32_t) * E); }
    for (int k = 0; k < 2; ++k) { L.eval[k] = off; off += align256(sizeof(uint32_t) * E); }
    L.ent_off = off;   off += align256(sizeof(int64_t) * (size_t)(positions + 2));
    L.max_items = (int64_t)P + nnz / kHubSeg + 2;
    L.items = off;     off += align256(sizeof(uint2) * (size_t)L.max_items);
    L.counters = off;  off += align256(sizeof(unsigned long long) * 4);
    L.prun = off;      off += align256(sizeof(int32_t) * P);
    L.prec = off;      off += align256(sizeof(unsigned long long) * P);
    L.key_bits = off;  off += align256(sizeof(uint32_t) * (size_t)((n + 32) / 32));
    L.item_link = off; off += align256(sizeof(int32_t) * (size_t)((chunks > 0 ? chunks : 0) + 1));
    L.sidx = off;      off += align256(sizeof(uint16_t) * E);
    size_t b1 = 0, b2 = 0, b3 = 0;
    {
        cub::DoubleBuffer<uint32_t> k(nullptr, nullptr), v(nullptr, nullptr);
        cub::DeviceRadixSort::SortPairs(nullptr, b1, k, v, (int)P, 0, key_bits(n));
        cub::DeviceRadixSort::SortPairs(nullptr, b2, k, v, (int)E, 0, key_bits(n));
        cub::DeviceScan::ExclusiveSum(nullptr, b3, (int64_t*)nullptr, (int64_t*)nullptr, (int)(positions + 2));
    }
    size_t b = b2 > b3 ? b2 : b3;
    L.cub_temp = off;                      // entries: scan + sort
    L.cub_temp_bytes = align256(b + 256);
    off += L.cub_temp_bytes;
    L.cub_temp2 = off;                     // pairs: sort (runs concurrently on the auxiliary stream)
    L.cub_temp2_bytes = align256(b1 + 256);
    off += L.cub_temp2_bytes;
    L.total = off;
    return L;
}

// ---- the inverted index ------------------------------------------------------------------------
// position pi of the stream = (run r, position p of N(src of r)); run_pos_off is the exclusive
// prefix of deg(src) over the runs
__device__ __forceinline__ void locate_position(int64_t pi, const int64_t* __restrict__ run_pos_off, int64_t n_runs,
                                                int64_t& r, int64_t& p) {
    int64_t lo = 0, hi = n_runs;  // last r with run_pos_off[r] <= pi
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (run_pos_off[mid] <= pi) lo = mid; else hi = mid;
    }
    r = lo;
    p = pi - run_pos_off[lo];
}

__global__ void k_hub_pos_count(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                const int64_t* __restrict__ src, const int32_t* __restrict__ run_start,
                                const int64_t* __restrict__ run_pos_off, int64_t n_runs, int64_t n_pos,
                                int64_t* __restrict__ ent_off) {
    const int64_t pi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi > n_pos) return;
    if (pi == n_pos) {
        ent_off[pi] = 0;
        return;
    }
    int64_t r, p;
    locate_position(pi, run_pos_off, n_runs, r, p);
    const int64_t i = src[run_start[r]];
    const int32_t k = ldg_i32(col + rowptr[i] + p);
    ent_off[pi] = rowptr[k + 1] - rowptr[k];
}

// one warp per position: the row N(k) of the position's node becomes entries (l, pi), emitted in
// position order (the stable sort by l then leaves every list ascending in pi)
__global__ void k_hub_emit_entries(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                   const int64_t* __restrict__ src, const int32_t* __restrict__ run_start,
                                   const int64_t* __restrict__ run_pos_off, int64_t n_runs, int64_t n_pos,
                                   const int64_t* __restrict__ ent_off, uint32_t* __restrict__ ekey,
                                   uint32_t* __restrict__ eval) {
    const int64_t pi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pi >= n_pos) return;
    int64_t r, p;
    locate_position(pi, run_pos_off, n_runs, r, p);
    const int64_t i = src[run_start[r]];
    const int32_t k = ldg_i32(col + rowptr[i] + p);
    const int64_t rs = rowptr[k], d = rowptr[k + 1] - rs;
    const int64_t off = ent_off[pi];
    for (int64_t idx = lane; idx < d; idx += 32) {
        ekey[off + idx] = (uint32_t)ldg_i32(col + rs + idx);
        eval[off + idx] = (uint32_t)pi;
    }
}

// run of a position: last r with run_pos_off[r] <= pi (runs are few: the search stays in L1)
__device__ __forceinline__ uint32_t run_of_position(uint32_t pi, const int64_t* __restrict__ run_pos_off, int n_runs) {
    int lo = 0, hi = n_runs;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((uint32_t)run_pos_off[mid] <= pi) lo = mid; else hi = mid;
    }
    return (uint32_t)lo;
}

// node_index[l] = (first entry of l, one past its last entry, signature of the runs of its entries:
// bit (run mod 64)); all zero = l is no key.  set == 0 restores the zeros.
__global__ void k_hub_entry_heads(const uint32_t* __restrict__ ekey, const uint32_t* __restrict__ eval,
                                  const int64_t* __restrict__ run_pos_off, int n_runs, int64_t E, int set,
                                  uint4* __restrict__ node_index, uint32_t* __restrict__ key_bits) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint32_t k = ekey[e];
    unsigned* ni = reinterpret_cast<unsigned*>(node_index + k);
    const bool first = (e == 0 || ekey[e - 1] != k);
    if (!set) {
        if (first) node_index[k] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    if (first) {
        ni[0] = (unsigned)e;
        atomicOr(key_bits + (k >> 5), 1u << (k & 31u));  // 1 bit per node: 8 MB of index entries are filtered by 366 KB
    }
    if (e == E - 1 || ekey[e + 1] != k) ni[1] = (unsigned)(e + 1);
    const uint32_t r = run_of_position(eval[e], run_pos_off, n_runs);
    atomicOr(ni + 2 + ((r >> 5) & 1u), 1u << (r & 31u));
}

// ---- (shared row, link) pairs, in link order ---------------------------------------------------
__device__ __forceinline__ bool is_hub_row(const int64_t* __restrict__ rowptr, int32_t m, int64_t hub_d) {
    return (ldg_i64(rowptr + m + 1) - ldg_i64(rowptr + m)) >= hub_d;
}

// one warp per link over the first kLongRow neighbours of dst (positions by ballot rank) ...
// A link whose run has no positions in this pass (it belongs to the other class of runs) still owns its slots of
// the pair arrays; they are filled with the sentinel row n, which sorts last and yields no work item.
__device__ __forceinline__ bool link_in_pass(const int32_t* __restrict__ run_id, const int64_t* __restrict__ run_pos_off,
                                             int64_t t) {
    const int r = run_id[t] - 1;
    return run_pos_off[r + 1] > run_pos_off[r];
}

__global__ void k_hub_emit_pairs(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                                 const int64_t* __restrict__ dst, int64_t T, int64_t hub_d,
                                 const int32_t* __restrict__ hub_off, const int32_t* __restrict__ run_id,
                                 const int64_t* __restrict__ run_pos_off, uint32_t* __restrict__ pkey,
                                 uint32_t* __restrict__ pval) {
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    const bool in_pass = link_in_pass(run_id, run_pos_off, t);
    const int64_t j = dst[t];
    const int64_t rs = rowptr[j];
    int64_t d = rowptr[j + 1] - rs;
    if (d > kLongRow) d = kLongRow;
    int64_t base = hub_off[t];
    for (int64_t o = 0; o < d; o += 32) {
        const int64_t idx = o + lane;
        int32_t m = -1;
        bool hub = false;
        if (idx < d) {
            m = ldg_i32(col + rs + idx);
            hub = is_hub_row(rowptr, m, hub_d);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hub);
        if (hub) {
            const int64_t pos = base + __popc(bal & ((1u << lane) - 1u));
            pkey[pos] = in_pass ? (uint32_t)m : (uint32_t)n;
            pval[pos] = (uint32_t)t;
        }
        base += __popc(bal);
    }
}

// ... and one CTA per link of the long-destination list for the rest (the order of a link's pairs is free:
// its rows are distinct, and the sort by row keeps links in stream order)
__global__ void __launch_bounds__(1024)
k_hub_emit_pairs_long(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                      const int64_t* __restrict__ dst, int64_t hub_d, const int32_t* __restrict__ hub_off,
                      const int32_t* __restrict__ long_list, const int64_t* __restrict__ plan,
                      const int32_t* __restrict__ run_id, const int64_t* __restrict__ run_pos_off,
                      uint32_t* __restrict__ pkey, uint32_t* __restrict__ pval) {
    __shared__ int s_pos;
    const int64_t n_long = plan[OCN_PLAN_LONG_COUNT];
    for (int64_t i = blockIdx.x; i < n_long; i += gridDim.x) {
        const int64_t t = long_list[i];
        const bool in_pass = link_in_pass(run_id, run_pos_off, t);
        const int64_t j = dst[t];
        const int64_t rs = rowptr[j], d = rowptr[j + 1] - rs;
        if (threadIdx.x == 0) s_pos = 0;
        __syncthreads();
        if (threadIdx.x < kLongRow && is_hub_row(rowptr, ldg_i32(col + rs + threadIdx.x), hub_d)) atomicAdd(&s_pos, 1);
        __syncthreads();
        const int64_t base = hub_off[t];
        for (int64_t o = kLongRow + threadIdx.x; o < d; o += blockDim.x) {
            const int32_t m = ldg_i32(col + rs + o);
            if (is_hub_row(rowptr, m, hub_d)) {
                const int64_t pos = base + atomicAdd(&s_pos, 1);
                pkey[pos] = in_pass ? (uint32_t)m : (uint32_t)n;
                pval[pos] = (uint32_t)t;
            }
        }
        __syncthreads();
    }
}

// work items (row, segment of kHubSeg columns), per-pair run and record offset
__global__ void k_hub_items(const int64_t* __restrict__ rowptr, int64_t n, const int32_t* __restrict__ run_id,
                            const int64_t* __restrict__ rec_off, const uint32_t* __restrict__ pkey,
                            const uint32_t* __restrict__ pval, int64_t P, int32_t* __restrict__ prun,
                            unsigned long long* __restrict__ prec, uint2* __restrict__ items, int64_t max_items,
                            unsigned long long* __restrict__ counters) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int nseg = 0;
    if (q < P) {
        const uint32_t t = pval[q];
        prun[q] = run_id[t] - 1;
        prec[q] = (unsigned long long)rec_off[t];
        const uint32_t m = pkey[q];
        if (m < (uint32_t)n && (q == 0 || pkey[q - 1] != m)) nseg = (int)((rowptr[m + 1] - rowptr[m] + kHubSeg - 1) / kHubSeg);
    }
    // one atomic per warp: lane 0 reserves the warp's items, every head row takes its share
    int incl = nseg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned long long base = 0;
    if (lane == 0 && total > 0) base = atomicAdd(&counters[0], (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 0) + (unsigned long long)(incl - nseg);
    for (int s = 0; s < nseg; ++s)
        if ((int64_t)(base + s) < max_items) items[base + s] = make_uint2((uint32_t)q, (uint32_t)s);
}

// ---- warp helpers --------------------------------------------------------------------------------
// lane's owner in a flattened expansion: largest s with excl[s] <= j (excl = exclusive scan of the
// per-lane counts, one value per lane)
__device__ __forceinline__ int owner_lane(int excl, int j) {
    int s = 0;
#pragma unroll
    for (int st = 16; st > 0; st >>= 1) {
        const int ev = __shfl_sync(0xffffffffu, excl, s + st);
        if (ev <= j) s += st;
    }
    return s;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// U holds 16-bit counters packed in pairs, for the positions [win_lo, win_lo + win_n) of this pass
template <bool kWin>
__device__ __forceinline__ void count_walk(uint32_t* U, uint32_t pos, uint32_t win_lo, uint32_t win_n) {
    const uint32_t rel = kWin ? pos - win_lo : pos;  // kWin == false: one window holds every position, no test
    if (!kWin || rel < win_n) atomicAdd(U + (rel >> 1), 1u << ((rel & 1u) << 4));
}

// ---- the walk through the shared rows ------------------------------------------------------------
// One warp (kCta == false: streams with up to kHubWindow positions, counters private to the warp) or one
// CTA (kCta == true: more positions -- a hub source -- counters shared by the CTA's warps) per item
// (row m, segment of its columns).  Entries of runs without a link next to m only touch counters
// nobody reads; a column whose run signature misses every run of L_m is skipped.  Entry lists are
// walked by their own lane (<= kShortList entries), flattened over the lanes (<= kMidList) or by the
// whole warp, so that neither the many short lists nor the few long ones (54 % of the visits are in
// lists of > 32 entries at citation2 shape) leave lanes idle.  Positions beyond the counter window are
// handled by further passes over the items ([win_lo, win_lo + win_n) per launch).
template <bool kCta, bool kWin, bool kExact>
__global__ void __launch_bounds__(kHubThreads, kCta ? 3 : 6)
k_cn_hub_count(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const uint32_t* __restrict__ pkey,
               const int32_t* __restrict__ prun, const unsigned long long* __restrict__ prec, int64_t P,
               const uint32_t* __restrict__ eval, const uint4* __restrict__ node_index,
               const uint32_t* __restrict__ key_bits,
               const int64_t* __restrict__ run_pos_off, int n_runs, uint32_t win_lo, uint32_t win_n,
               const uint2* __restrict__ items, int64_t max_items, unsigned long long* __restrict__ counters,
               Record* __restrict__ records) {
    extern __shared__ uint32_t hub_smem[];
    __shared__ unsigned long long s_item;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = kCta ? kHubWarps : 1;   // warps sharing an item
    const int wi = kCta ? warp : 0;        // this warp's index among them
    const int rpad = (n_runs + 1 + 3) & ~3;
    const int uwords = (int)((((win_n + 1) >> 1) + 3) & ~3u);   // packed 16-bit counters, a multiple of 4 words
    uint32_t* s_pos = hub_smem;                                  // run -> first position (n_runs + 1 entries)
    uint32_t* U = hub_smem + rpad + (kCta ? (size_t)0 : (size_t)warp * uwords);
    // warp-private queue of the entry lists that passed the signature test (head, length): lists are walked 32 at a
    // time, one per lane, instead of by the few lanes of a 32-column step that happen to hold one
    uint2* Q = reinterpret_cast<uint2*>(hub_smem + rpad + (size_t)uwords * (kCta ? 1 : kHubWarps)) + (size_t)warp * kHubQueue;
    for (int r = threadIdx.x; r <= n_runs; r += blockDim.x) s_pos[r] = (uint32_t)run_pos_off[r];
    __syncthreads();
    unsigned long long n_items = counters[0];
    if (n_items > (unsigned long long)max_items) n_items = (unsigned long long)max_items;  // cannot happen with ocn_cn_hub_bytes
    unsigned* rec32 = reinterpret_cast<unsigned*>(records);
    auto item_sync = [&]() { if (kCta) __syncthreads(); else __syncwarp(); };
    while (true) {
        unsigned long long it = 0;
        if (kCta) {
            if (threadIdx.x == 0) s_item = atomicAdd(&counters[1], 1ull);
            __syncthreads();
            it = s_item;
        } else {
            if (lane == 0) it = atomicAdd(&counters[1], 1ull);
            it = __shfl_sync(0xffffffffu, it, 0);
        }
        if (it >= n_items) break;
        const uint2 item = items[it];
        const int64_t q0 = item.x;
        const uint32_t m = pkey[q0];
        int64_t c = 0;  // |L_m|
        // the runs that have a link next to m: exact set (kExact: up to 128 runs, 32-byte node entries whose second
        // half is the exact run set of the list) or folded into 64 bits
        unsigned act_lo = 0u, act_hi = 0u, act_2 = 0u, act_3 = 0u;
        while (true) {
            const int64_t qi = q0 + c + lane;
            const bool ok = qi < P && pkey[qi] == m;
            if (ok) {
                const int r = prun[qi];
                if (kExact) {
                    const unsigned bit = 1u << (r & 31), w = (unsigned)r >> 5;
                    act_lo |= w == 0 ? bit : 0u;
                    act_hi |= w == 1 ? bit : 0u;
                    act_2 |= w == 2 ? bit : 0u;
                    act_3 |= w == 3 ? bit : 0u;
                } else if (r & 32) act_hi |= 1u << (r & 31); else act_lo |= 1u << (r & 31);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            c += __popc(bal);
            if (bal != 0xffffffffu) break;
        }
        act_lo = __reduce_or_sync(0xffffffffu, act_lo);
        act_hi = __reduce_or_sync(0xffffffffu, act_hi);
        if (kExact) {
            act_2 = __reduce_or_sync(0xffffffffu, act_2);
            act_3 = __reduce_or_sync(0xffffffffu, act_3);
        }
        for (int s = (kCta ? threadIdx.x : lane) * 4; s < uwords; s += (kCta ? kHubThreads : 32) * 4)
            *reinterpret_cast<uint4*>(U + s) = make_uint4(0u, 0u, 0u, 0u);
        item_sync();
        const int64_t rs = rowptr[m];
        const int64_t d = rowptr[m + 1] - rs;
        const int64_t c0 = (int64_t)item.y * kHubSeg;
        const int64_t c1 = (c0 + kHubSeg < d) ? c0 + kHubSeg : d;
        // software pipeline: the column of step b+3, the key bit of step b+2 and the index entry of step b+1 are in
        // flight while the lists of step b are walked
        const int64_t stride = 32 * nw;
        auto load_col = [&](int64_t bb) -> int32_t { return (bb + lane < c1) ? ldg_i32(col + rs + bb + lane) : -1; };
        // a column is a key of the index only if its bit is set: the 16-byte entry is fetched for those alone
        auto load_bit = [&](int32_t l) -> bool { return l >= 0 && ((__ldg(key_bits + (l >> 5)) >> (l & 31)) & 1u) != 0u; };
        // node entry -> (first entry, length of the list if it holds a run of an active link, else 0); the loads of step
        // b+1 are issued here, the test runs when the values are used
        uint4 he_raw = make_uint4(0u, 0u, 0u, 0u);
        uint2 he_fe = make_uint2(0u, 0u);
        auto load_entry = [&](int32_t l, bool key) {
            if (kExact) {
                if (key) {
                    he_fe = __ldg(reinterpret_cast<const uint2*>(node_index + 2 * (size_t)l));
                    he_raw = __ldg(node_index + 2 * (size_t)l + 1);
                } else {
                    he_raw = make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
                he_raw = key ? __ldg(node_index + l) : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        auto entry_first = [&]() -> uint32_t { return kExact ? he_fe.x : he_raw.x; };
        auto entry_count = [&]() -> int {
            if (kExact) return ((he_raw.x & act_lo) | (he_raw.y & act_hi) | (he_raw.z & act_2) | (he_raw.w & act_3)) ? (int)(he_fe.y - he_fe.x) : 0;
            return ((he_raw.z & act_lo) | (he_raw.w & act_hi)) ? (int)(he_raw.y - he_raw.x) : 0;
        };
        // one queued list per lane (length 0 = none): the first kShortList entries by the lane itself, the rest of
        // the middle lists flattened over the lanes, 32 entries at a time
        auto drain = [&](uint2 ql) {
            const int cnt = (int)ql.y;
            {
                const uint32_t* list = eval + ql.x;
                uint32_t pos[kShortList];
#pragma unroll
                for (int k = 0; k < kShortList; ++k) pos[k] = k < cnt ? __ldg(list + k) : 0u;
#pragma unroll
                for (int k = 0; k < kShortList; ++k)
                    if (k < cnt) count_walk<kWin>(U, pos[k], win_lo, win_n);
            }
            const int cm = cnt > kShortList ? cnt - kShortList : 0;
            if (__any_sync(0xffffffffu, cm != 0)) {
                const int incl = warp_incl_scan(cm, lane);
                const int total = __shfl_sync(0xffffffffu, incl, 31);
                const int excl = incl - cm;
                for (int j0 = 0; j0 < total; j0 += 32) {
                    const int j = j0 + lane;
                    const int s = owner_lane(excl, j);
                    const uint32_t ehead = __shfl_sync(0xffffffffu, ql.x, s) + kShortList;
                    const int eexcl = __shfl_sync(0xffffffffu, excl, s);
                    if (j < total) count_walk<kWin>(U, __ldg(eval + ehead + (uint32_t)(j - eexcl)), win_lo, win_n);
                }
            }
        };
        int qn = 0;  // queued lists (warp-uniform)
        const int64_t b0 = c0 + 32 * wi;
        int32_t col_n1 = load_col(b0 + stride), col_n2 = load_col(b0 + 2 * stride);
        {
            const int32_t l0 = load_col(b0);
            load_entry(l0, load_bit(l0));
        }
        bool bit_n1 = load_bit(col_n1);
        for (int64_t b = b0; b < c1; b += stride) {
            const int cnt = entry_count();
            const uint32_t he_x = entry_first(), he_y = he_x + (uint32_t)cnt;
            load_entry(col_n1, bit_n1);
            col_n1 = col_n2;
            bit_n1 = load_bit(col_n1);
            col_n2 = load_col(b + 3 * stride);
            if (!__any_sync(0xffffffffu, cnt != 0)) continue;
            // short and middle lists go to the queue; a full batch of 32 is walked as soon as there is one
            {
                const bool push = cnt > 0 && cnt <= kMidList;
                const unsigned pm = __ballot_sync(0xffffffffu, push);
                if (push) Q[qn + __popc(pm & ((1u << lane) - 1u))] = make_uint2(he_x, (uint32_t)cnt);
                qn += __popc(pm);
                __syncwarp();
                if (qn >= 32) {
                    qn -= 32;
                    drain(Q[qn + lane]);
                    __syncwarp();
                }
            }
            // long lists: the whole warp per list (ascending positions: only the window's piece is read when the
            // stream needs several passes)
            unsigned lg = __ballot_sync(0xffffffffu, cnt > kMidList);
            while (lg) {
                const int sl = __ffs(lg) - 1;
                lg &= lg - 1;
                uint32_t e0 = __shfl_sync(0xffffffffu, he_x, sl);
                uint32_t e1 = __shfl_sync(0xffffffffu, he_y, sl);
                if (kWin && win_lo > 0u) {  // first entry with position >= win_lo
                    uint32_t lo = e0, hi = e1;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (__ldg(eval + mid) < win_lo) lo = mid + 1; else hi = mid;
                    }
                    e0 = lo;
                }
                if (kWin) {
                    for (uint32_t e = e0 + lane; e0 < e1; e0 += 32, e += 32) {
                        const uint32_t pos = e < e1 ? __ldg(eval + e) : 0xffffffffu;
                        count_walk<true>(U, pos, win_lo, win_n);
                        if (__any_sync(0xffffffffu, pos - win_lo >= win_n)) break;  // past the window (or the list)
                    }
                } else {
                    // the lists are 33 .. ~230 entries long: 128 entries per round, all loads issued before the first
                    // counter is touched
                    for (uint32_t e = e0 + lane; e < e1; e += 128) {
                        uint32_t pv[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) pv[k] = (e + 32u * k < e1) ? __ldg(eval + e + 32u * k) : 0xffffffffu;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (pv[k] != 0xffffffffu) count_walk<false>(U, pv[k], 0u, 0u);
                    }
                }
            }
        }
        if (qn > 0) drain(lane < qn ? Q[lane] : make_uint2(0u, 0u));
        item_sync();
        // hand the counts to the links next to m
        // (the run and the record offset of 32 links at a time are fetched by the lanes up front: no dependent
        // global load between two links)
        for (int64_t qb = 0; qb < c; qb += 32) {
            int r_l = 0;
            unsigned long long prec_l = 0ull;
            if (qb + lane < c) {
                r_l = prun[q0 + qb + lane];
                prec_l = prec[q0 + qb + lane];
            }
            const int cnt = (int)((c - qb) < 32 ? (c - qb) : 32);
            for (int qq = wi; qq < cnt; qq += nw) {
                const int r = __shfl_sync(0xffffffffu, r_l, qq);
                const unsigned long long pr = __shfl_sync(0xffffffffu, prec_l, qq);
                const uint32_t pb = s_pos[r], pe = s_pos[r + 1];
                const uint32_t lo = (kWin && pb < win_lo) ? win_lo : pb, hi = (kWin && pe > win_lo + win_n) ? win_lo + win_n : pe;
                unsigned* rec = rec32 + 2 * pr + 1;
                for (uint32_t pos = lo + lane; pos < hi; pos += 32) {
                    const uint32_t rel = kWin ? pos - win_lo : pos;
                    const uint32_t u = (U[rel >> 1] >> ((rel & 1u) << 4)) & 0xffffu;
                    if (u) atomicAdd(rec + 2 * (pos - pb), u);
                }
            }
        }
        item_sync();
    }
}

// ---- run-segment index (streams of up to kSegRuns runs, positions in one warp-private window) -------------
// The census of the walk above (scripts/hub_visit_census2.py, bench slice 0) says: of the 104 M entries it visits
// per launch only 19.6 M belong to a run that has a link next to the streamed row -- a list that holds one active
// run is walked whole.  With at most 128 runs the node entry can carry the EXACT set of runs of its list
// (128 bits), and because a list is sorted by position, i.e. by run, the entries of run r are the segment number
// rank(r) = popc(runs(l) & below(r)) of the list.  Node entry (32 bytes, one sector):
//     a = (first entry, one past the last, -, -)      b = run set (bit r = the list holds an entry of run r)
// Lists with one entry per run ("simple": popc(b) == length; 97 % of the keys) need nothing else -- segment s is
// entry first + s.  The others store the start of segment s (relative to first) in sidx[first + s]: a list has no
// more segments than entries, so the array parallel to the entries has room.  A streamed column then costs its
// hit segments only (13.7 M per launch, 19.6 M entries) instead of its whole list, and the per-link look-up
// (k_cn_link) finds a run's piece of a list without a search.
constexpr int kSegRuns = 128;

__device__ __forceinline__ int popc4(uint4 s) { return __popc(s.x) + __popc(s.y) + __popc(s.z) + __popc(s.w); }

// number of runs of the set `s` below run r, and whether r itself is in the set
__device__ __forceinline__ int seg_rank(uint4 s, uint32_t r, bool& present) {
    const uint32_t w = r >> 5, b = r & 31u;
    const uint32_t sw = w == 0 ? s.x : (w == 1 ? s.y : (w == 2 ? s.z : s.w));
    present = ((sw >> b) & 1u) != 0u;
    int rank = __popc(sw & ((1u << b) - 1u));
    if (w > 0) rank += __popc(s.x);
    if (w > 1) rank += __popc(s.y);
    if (w > 2) rank += __popc(s.z);
    return rank;
}

// position (0..31) of the j-th (0-based) set bit of w; j < popc(w)
__device__ __forceinline__ int nth_set_bit(uint32_t w, int j) {
    int pos = 0, t;
    t = __popc(w & 0xffffu); if (j >= t) { j -= t; w >>= 16; pos = 16; }
    t = __popc(w & 0xffu);   if (j >= t) { j -= t; w >>= 8;  pos += 8; }
    t = __popc(w & 0xfu);    if (j >= t) { j -= t; w >>= 4;  pos += 4; }
    t = __popc(w & 0x3u);    if (j >= t) { j -= t; w >>= 2;  pos += 2; }
    if (j >= (int)(w & 1u)) pos += 1;
    return pos;
}

__global__ void k_hub_entry_heads_seg(const uint32_t* __restrict__ ekey, const uint32_t* __restrict__ eval,
                                      const int64_t* __restrict__ run_pos_off, int n_runs, int64_t E, int set,
                                      uint4* __restrict__ node_index, uint32_t* __restrict__ key_bits) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint32_t k = ekey[e];
    unsigned* ni = reinterpret_cast<unsigned*>(node_index + 2 * (size_t)k);
    const bool first = (e == 0 || ekey[e - 1] != k);
    if (!set) {
        if (first) {
            node_index[2 * (size_t)k] = make_uint4(0u, 0u, 0u, 0u);
            node_index[2 * (size_t)k + 1] = make_uint4(0u, 0u, 0u, 0u);
        }
        return;
    }
    if (first) {
        ni[0] = (unsigned)e;
        atomicOr(key_bits + (k >> 5), 1u << (k & 31u));
    }
    if (e == E - 1 || ekey[e + 1] != k) ni[1] = (unsigned)(e + 1);
    const uint32_t r = run_of_position(eval[e], run_pos_off, n_runs);
    atomicOr(ni + 4 + (r >> 5), 1u << (r & 31u));
}

// segment starts of the lists that hold several entries of one run (after the run sets are complete)
__global__ void k_hub_segments(const uint32_t* __restrict__ ekey, const uint32_t* __restrict__ eval,
                               const int64_t* __restrict__ run_pos_off, int n_runs, int64_t E,
                               const uint4* __restrict__ node_index, uint16_t* __restrict__ sidx) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint32_t k = ekey[e];
    const uint4 a = node_index[2 * (size_t)k], s = node_index[2 * (size_t)k + 1];
    if ((uint32_t)popc4(s) == a.y - a.x) return;  // one entry per run: segment s is entry first + s
    const uint32_t r = run_of_position(eval[e], run_pos_off, n_runs);
    if (e != (int64_t)a.x && run_of_position(eval[e - 1], run_pos_off, n_runs) == r) return;  // inside a segment
    bool present;
    const int j = seg_rank(s, r, present);
    sidx[a.x + (uint32_t)j] = (uint16_t)(e - a.x);
}

// the entries [e0, e1) of run r in the list of node entry (a, s); e0 == e1 when the list holds none
__device__ __forceinline__ void seg_bounds(uint4 s, uint32_t first, uint32_t end, int rank,
                                           const uint16_t* __restrict__ sidx, uint32_t& e0, uint32_t& e1) {
    const int k = popc4(s);
    const uint32_t cnt = end - first;
    if ((uint32_t)k == cnt) {
        e0 = first + (uint32_t)rank;
        e1 = e0 + 1u;
    } else {
        e0 = first + (uint32_t)__ldg(sidx + first + rank);
        e1 = (rank + 1 < k) ? first + (uint32_t)__ldg(sidx + first + rank + 1) : end;
    }
}

// k_cn_hub_count with the run-segment index (opt-in: ocn_set_option(OCN_OPT_HUB_WALKER, 1)): one warp per item (row m,
// segment of its columns), counters private to the warp.  Columns whose run set meets the active runs of m are queued
// (node entry, 24 bytes) in warp-private shared memory; every 32 queued columns -- and at the end of the item --
// their (column, active run) hits are flattened over the lanes: slot j finds its column (owner) and the j-th set bit
// of that column's hit set, then counts the entries of that run segment.
// MEASURED SLOWER than the whole-list walk above and therefore not the default (bench slices, kernel alone, B200):
// whole lists 0.58 ms (104 M entry visits, 359 M warp instructions, 73 % issue slots) against 0.74 - 0.89 ms for three
// versions of this kernel (22 M visits, 330 - 395 M warp instructions, 57 - 65 % issue slots, 5 CTAs/SM at 48 registers):
// finding a hit's segment costs ~12 warp instructions per useful visit where walking a list costs 2.6 per visit, useful
// or not, and the per-segment entry loops run with 3 - 4 lanes active (profiles/r02_seg_v1_*, r02_seg_v2_*).
// Kept as a tested alternative walker for streams whose lists are much longer than the bench's.
constexpr int kSegQueue = 64;        // queued node entries per warp
constexpr int kSegLanePairs = 8;     // rows with at least this many links hand their counters out one lane per link
#ifndef OCN_SEG_MINB
#define OCN_SEG_MINB 5
#endif

__global__ void __launch_bounds__(kHubThreads, OCN_SEG_MINB)
k_cn_hub_count_seg(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const uint32_t* __restrict__ pkey,
                   const int32_t* __restrict__ prun, const unsigned long long* __restrict__ prec, int64_t P,
                   const uint32_t* __restrict__ eval, const uint16_t* __restrict__ sidx,
                   const uint4* __restrict__ node_index, const uint32_t* __restrict__ key_bits,
                   const int64_t* __restrict__ run_pos_off, int n_runs, uint32_t n_pos,
                   const uint2* __restrict__ items, int64_t max_items, unsigned long long* __restrict__ counters,
                   Record* __restrict__ records) {
    extern __shared__ uint32_t hub_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rpad = (n_runs + 1 + 3) & ~3;
    const int uwords = (int)((((n_pos + 1) >> 1) + 3) & ~3u);
    uint32_t* s_pos = hub_smem;
    uint32_t* U = hub_smem + rpad + (size_t)warp * uwords;
    uint4* Qs = reinterpret_cast<uint4*>(hub_smem + rpad + (size_t)kHubWarps * uwords) + (size_t)warp * kSegQueue;          // run sets
    uint2* Qa = reinterpret_cast<uint2*>(hub_smem + rpad + (size_t)kHubWarps * uwords + 4 * kSegQueue * kHubWarps) +
                (size_t)warp * kSegQueue;                                                                                    // (first, end)
    for (int r = threadIdx.x; r <= n_runs; r += blockDim.x) s_pos[r] = (uint32_t)run_pos_off[r];
    __syncthreads();
    unsigned long long n_items = counters[0];
    if (n_items > (unsigned long long)max_items) n_items = (unsigned long long)max_items;
    unsigned* rec32 = reinterpret_cast<unsigned*>(records);
    while (true) {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(&counters[1], 1ull);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= n_items) break;
        const uint2 item = items[it];
        const int64_t q0 = item.x;
        const uint32_t m = pkey[q0];
        int64_t c = 0;  // |L_m|
        uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;  // the runs that have a link next to m
        while (true) {
            const int64_t qi = q0 + c + lane;
            const bool ok = qi < P && pkey[qi] == m;
            if (ok) {
                const uint32_t r = (uint32_t)prun[qi];
                const uint32_t bit = 1u << (r & 31u), w = r >> 5;
                a0 |= w == 0 ? bit : 0u;
                a1 |= w == 1 ? bit : 0u;
                a2 |= w == 2 ? bit : 0u;
                a3 |= w == 3 ? bit : 0u;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            c += __popc(bal);
            if (bal != 0xffffffffu) break;
        }
        a0 = __reduce_or_sync(0xffffffffu, a0);
        a1 = __reduce_or_sync(0xffffffffu, a1);
        a2 = __reduce_or_sync(0xffffffffu, a2);
        a3 = __reduce_or_sync(0xffffffffu, a3);
        for (int s = lane * 4; s < uwords; s += 128) *reinterpret_cast<uint4*>(U + s) = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        const int64_t rs = rowptr[m];
        const int64_t d = rowptr[m + 1] - rs;
        const int64_t c0 = (int64_t)item.y * kHubSeg;
        const int64_t c1 = (c0 + kHubSeg < d) ? c0 + kHubSeg : d;
        auto load_col = [&](int64_t bb) -> int32_t { return (bb + lane < c1) ? ldg_i32(col + rs + bb + lane) : -1; };
        auto load_bit = [&](int32_t l) -> bool { return l >= 0 && ((__ldg(key_bits + (l >> 5)) >> (l & 31)) & 1u) != 0u; };
        // the hits of the queued columns [base, base + count), flattened over the lanes
        auto drain = [&](int base, int count) {
            int h = 0;
            if (lane < count) {
                const uint4 g = Qs[base + lane];
                h = __popc(g.x & a0) + __popc(g.y & a1) + __popc(g.z & a2) + __popc(g.w & a3);
            }
            const int incl = warp_incl_scan(h, lane);
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - h;
            for (int j0 = 0; j0 < total; j0 += 32) {
                const int j = j0 + lane;
                const int sl = owner_lane(excl, j);
                const int jl = j - __shfl_sync(0xffffffffu, excl, sl);
                if (j < total) {
                    const uint4 g = Qs[base + sl];
                    const uint2 fe = Qa[base + sl];
                    // the jl-th active run of the column: word, then bit
                    const uint32_t x0 = g.x & a0, x1 = g.y & a1, x2 = g.z & a2, x3 = g.w & a3;
                    const int p0 = __popc(x0), p1 = p0 + __popc(x1), p2 = p1 + __popc(x2);
                    const int s0 = __popc(g.x), s1 = s0 + __popc(g.y), s2 = s1 + __popc(g.z), k = s2 + __popc(g.w);
                    uint32_t xw, sw;
                    int jj, below;
                    if (jl < p0)      { xw = x0; sw = g.x; jj = jl;      below = 0; }
                    else if (jl < p1) { xw = x1; sw = g.y; jj = jl - p0; below = s0; }
                    else if (jl < p2) { xw = x2; sw = g.z; jj = jl - p1; below = s1; }
                    else              { xw = x3; sw = g.w; jj = jl - p2; below = s2; }
                    const int bit = nth_set_bit(xw, jj);
                    const int rank = below + __popc(sw & ((1u << bit) - 1u));
                    if ((uint32_t)k == fe.y - fe.x) {   // one entry per run
                        count_walk<false>(U, __ldg(eval + fe.x + rank), 0u, 0u);
                    } else {
                        uint32_t e = fe.x + (uint32_t)__ldg(sidx + fe.x + rank);
                        const uint32_t e1 = (rank + 1 < k) ? fe.x + (uint32_t)__ldg(sidx + fe.x + rank + 1) : fe.y;
                        for (; e < e1; ++e) count_walk<false>(U, __ldg(eval + e), 0u, 0u);
                    }
                }
            }
        };
        // software pipeline as in k_cn_hub_count: column of step b+3, key bit of step b+2, node entry of step b+1
        uint2 ha_next = make_uint2(0u, 0u);
        uint4 hs_next = make_uint4(0u, 0u, 0u, 0u);
        auto load_entry = [&](int32_t l, bool key) {
            if (key) {
                ha_next = __ldg(reinterpret_cast<const uint2*>(node_index + 2 * (size_t)l));
                hs_next = __ldg(node_index + 2 * (size_t)l + 1);
            } else {
                hs_next = make_uint4(0u, 0u, 0u, 0u);
            }
        };
        int32_t col_n1 = load_col(c0 + 32), col_n2 = load_col(c0 + 64);
        {
            const int32_t l0 = load_col(c0);
            load_entry(l0, load_bit(l0));
        }
        bool bit_n1 = load_bit(col_n1);
        int qn = 0;  // queued columns (warp-uniform)
        for (int64_t b = c0; b < c1; b += 32) {
            const uint2 ha = ha_next;
            const uint4 hs = hs_next;
            load_entry(col_n1, bit_n1);
            col_n1 = col_n2;
            bit_n1 = load_bit(col_n1);
            col_n2 = load_col(b + 96);
            const bool push = ((hs.x & a0) | (hs.y & a1) | (hs.z & a2) | (hs.w & a3)) != 0u;
            const unsigned pm = __ballot_sync(0xffffffffu, push);
            if (pm == 0u) continue;
            if (push) {
                const int qi = qn + __popc(pm & ((1u << lane) - 1u));
                Qs[qi] = hs;
                Qa[qi] = ha;
            }
            qn += __popc(pm);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                drain(qn, 32);
                __syncwarp();
            }
        }
        if (qn > 0) drain(0, qn);
        __syncwarp();
        // hand the counts to the links next to m
        if (c >= kSegLanePairs) {
            // many links: one lane per link walks the positions of its run (links of one run read the same counters)
            for (int64_t qb = 0; qb < c; qb += 32) {
                uint32_t pb = 0u, pe = 0u;
                unsigned* rec = rec32;
                if (qb + lane < c) {
                    const int r = prun[q0 + qb + lane];
                    pb = s_pos[r];
                    pe = s_pos[r + 1];
                    rec = rec32 + 2 * prec[q0 + qb + lane] + 1;
                }
                for (uint32_t pos = pb; pos < pe; ++pos, rec += 2) {
                    const uint32_t u = (U[pos >> 1] >> ((pos & 1u) << 4)) & 0xffffu;
                    if (u) atomicAdd(rec, u);
                }
            }
        } else {
            int r_l = 0;
            unsigned long long prec_l = 0ull;
            if (lane < c) {
                r_l = prun[q0 + lane];
                prec_l = prec[q0 + lane];
            }
            for (int qq = 0; qq < (int)c; ++qq) {
                const int r = __shfl_sync(0xffffffffu, r_l, qq);
                const unsigned long long pr = __shfl_sync(0xffffffffu, prec_l, qq);
                const uint32_t pb = s_pos[r], pe = s_pos[r + 1];
                unsigned* rec = rec32 + 2 * pr + 1;
                for (uint32_t pos = pb + lane; pos < pe; pos += 32) {
                    const uint32_t u = (U[pos >> 1] >> ((pos & 1u) << 4)) & 0xffffu;
                    if (u) atomicAdd(rec + 2 * (pos - pb), u);
                }
            }
        }
        __syncwarp();
    }
}

// per-link look-up in the run-segment index: the entries of run r in l's list, without a search
__device__ __forceinline__ void link_lookup_seg(const uint4* __restrict__ node_index, const uint32_t* __restrict__ key_bits,
                                                const uint32_t* __restrict__ eval, const uint16_t* __restrict__ sidx,
                                                uint32_t l, uint32_t r, uint32_t pos_lo, unsigned* __restrict__ rec, unsigned inc) {
    if (key_bits != nullptr && ((__ldg(key_bits + (l >> 5)) >> (l & 31u)) & 1u) == 0u) return;
    const uint4 s = __ldg(node_index + 2 * (size_t)l + 1);
    bool present;
    const int rank = seg_rank(s, r, present);
    if (!present) return;
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(node_index + 2 * (size_t)l));
    uint32_t e0, e1;
    seg_bounds(s, a.x, a.y, rank, sidx, e0, e1);
    for (uint32_t e = e0; e < e1; ++e) atomicAdd(rec + 2 * (__ldg(eval + e) - pos_lo), inc);
}

// ---- everything a link does not share ------------------------------------------------------------
// One warp per item (link t, 32 consecutive rows of N(dst[t])): C1 (first item of the link), C2 for
// its rows, C3 through those of its rows that have fewer than hub_d columns, their columns
// flattened over the lanes.  lookup: the entries of l that belong to the run of t are the piece of
// l's list with positions in [pos_lo, pos_hi).
__device__ __forceinline__ void link_lookup(const uint4* __restrict__ node_index, const uint32_t* __restrict__ key_bits,
                                            const uint32_t* __restrict__ eval,
                                            uint32_t l, uint32_t r, uint32_t pos_lo, uint32_t pos_hi,
                                            unsigned* __restrict__ rec, unsigned inc) {
    if (key_bits != nullptr && ((__ldg(key_bits + (l >> 5)) >> (l & 31u)) & 1u) == 0u) return;  // l has no entry list (most nodes)
    const uint4 he = __ldg(node_index + l);
    if ((((r & 32u) ? he.w : he.z) >> (r & 31u) & 1u) == 0u) return;  // no entry of l in a run of this signature bit
    uint32_t lo = he.x, hi = he.y;
    while (lo < hi) {  // first entry of l with position >= pos_lo
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(eval + mid) < pos_lo) lo = mid + 1; else hi = mid;
    }
    for (; lo < he.y; ++lo) {
        const uint32_t pos = __ldg(eval + lo);
        if (pos >= pos_hi) break;
        atomicAdd(rec + 2 * (pos - pos_lo), inc);
    }
}

// link of every work item (chunk_off is the exclusive prefix of the per-link item counts): one warp per link
__global__ void k_hub_item_links(const int32_t* __restrict__ chunk_off, int64_t T, int32_t* __restrict__ item_link) {
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= T) return;
    const int32_t a = chunk_off[t], b = chunk_off[t + 1];
    for (int32_t it = a + (threadIdx.x & 31); it < b; it += 32) item_link[it] = (int32_t)t;
}

template <bool kSeg>
__global__ void __launch_bounds__(256)
k_cn_link(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int64_t* __restrict__ dst,
          int64_t T, const int32_t* __restrict__ run_id, const int64_t* __restrict__ rec_off,
          const int32_t* __restrict__ chunk_off, const int64_t* __restrict__ run_pos_off,
          const uint32_t* __restrict__ eval, const uint16_t* __restrict__ sidx, const uint4* __restrict__ node_index,
          const uint32_t* __restrict__ key_bits, const int32_t* __restrict__ item_link, int64_t hub_d,
          Record* __restrict__ records) {
    // (l, run r of the link) -> the entries of r in l's list: by rank in the run-segment index, by search otherwise
    auto lookup = [&](const uint32_t* kb, uint32_t l, uint32_t r, uint32_t pos_lo, uint32_t pos_hi, unsigned* rec, unsigned inc) {
        if (kSeg) link_lookup_seg(node_index, kb, eval, sidx, l, r, pos_lo, rec, inc);
        else link_lookup(node_index, kb, eval, l, r, pos_lo, pos_hi, rec, inc);
    };
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_items = chunk_off[T];
    unsigned* rec32 = reinterpret_cast<unsigned*>(records);
    for (int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items; item += nwarps) {
        const int64_t t = item_link[item];
        const int64_t ch = item - chunk_off[t];
        const int64_t j = dst[t];
        const int64_t rs_j = rowptr[j], d_j = rowptr[j + 1] - rs_j;
        const uint32_t r = (uint32_t)(run_id[t] - 1);
        const uint32_t pos_lo = (uint32_t)run_pos_off[r], pos_hi = (uint32_t)run_pos_off[r + 1];
        if (pos_lo == pos_hi) continue;  // the run has no positions in this pass (other class of runs, or an isolated source)
        unsigned* rec = rec32 + 2 * rec_off[t];
        if (ch == 0 && lane == 0) lookup(key_bits, (uint32_t)j, r, pos_lo, pos_hi, rec, 0x80000000u);  // C1
        const int64_t oi = ch * 32 + lane;
        int64_t rs_m = 0;
        int cnt = 0;
        if (oi < d_j) {
            const int32_t m = ldg_i32(col + rs_j + oi);
            lookup(key_bits, (uint32_t)m, r, pos_lo, pos_hi, rec, 1u);  // C2
            rs_m = ldg_i64(rowptr + m);
            const int64_t d_m = ldg_i64(rowptr + m + 1) - rs_m;
            if (d_m < hub_d) cnt = (int)d_m;
        }
        const int incl = warp_incl_scan(cnt, lane);
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - cnt;
        // C3 through the rows this link does not share: 128 columns per round -- the four column loads, then the four
        // key-map words are in flight together; the few columns that are keys go on to the index
        for (int j0 = 0; j0 < total; j0 += 128) {
            int32_t lc[4];
            uint32_t bw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int jj = j0 + 32 * k + lane;
                const int s = owner_lane(excl, jj);
                const int64_t rs_o = __shfl_sync(0xffffffffu, rs_m, s);
                const int eexcl = __shfl_sync(0xffffffffu, excl, s);
                lc[k] = (jj < total) ? ldg_i32(col + rs_o + (jj - eexcl)) : -1;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) bw[k] = lc[k] >= 0 ? __ldg(key_bits + (lc[k] >> 5)) : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((bw[k] >> (lc[k] & 31)) & 1u) lookup(nullptr, (uint32_t)lc[k], r, pos_lo, pos_hi, rec + 1, 1u);
        }
    }
}


static int grid_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// optional CUDA events around the dominant kernel (bench.py's roofline leg)
static cudaEvent_t g_hub_events[2] = {nullptr, nullptr};
static void hub_timing_record(int which, cudaStream_t st) {
    if (g_hub_events[which] != nullptr) cudaEventRecord(g_hub_events[which], st);
}

// Auxiliary stream per device: the pair pipeline runs beside the entry pipeline and the per-link
// kernel beside the shared-row walk (small latency-bound kernels; both only add into the records).
// Every call forks from and joins back into the caller's stream, so the caller sees one ordered call.
struct HubAux {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t heavy_stream = nullptr;  // the pass over the runs of heavy sources runs beside the light pass
    cudaEvent_t hev[2] = {nullptr, nullptr};
};
static std::mutex g_aux_mutex;
static std::map<std::pair<int, cudaStream_t>, HubAux> g_aux;  // one auxiliary stream per (device, caller stream)

static int hub_aux(cudaStream_t caller, HubAux** out) {
    int dev = 0;
    OCN_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_aux_mutex);
    HubAux& a = g_aux[std::make_pair(dev, caller)];  // calls on different caller streams (two sessions in flight) do not share one
    if (a.stream == nullptr) {
        OCN_CUDA(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
        for (int k = 0; k < 4; ++k) OCN_CUDA(cudaEventCreateWithFlags(&a.ev[k], cudaEventDisableTiming));
        OCN_CUDA(cudaStreamCreateWithFlags(&a.heavy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) OCN_CUDA(cudaEventCreateWithFlags(&a.hev[k], cudaEventDisableTiming));
    }
    *out = &a;
    return OCN_OK;
}

// One pass of the indexed stage over the runs that have positions in `run_pos_off` (an exclusive prefix over ALL
// runs in which the runs of the other class contribute nothing): inverted index, pairs, shared-row walk, per-link
// kernel.  Records of links outside the pass are not touched.
struct PassState {            // what a failing pass has to undo (hub_pass_cleanup)
    bool forked = false;      // the auxiliary stream holds work of this call
    bool index_set = false;   // node_index / key_bitmap hold the index of this call
    bool seg = false;
    const uint32_t* ekey = nullptr;
    const uint32_t* eval = nullptr;
};

static int hub_pass_body(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, const int64_t* dst, int64_t T,
                         const void* plan_scratch, const int64_t* plan_dev, const int64_t* run_pos_off, int64_t hub_d, int64_t P,
                         int64_t E, int64_t NP, int64_t R, const HubLayout& H, void* hub_scratch, uint4* node_index,
                         Record* records, HubAux* aux, bool first_pass, cudaStream_t st, PassState& ps) {
    cudaStream_t sa = aux->stream;
    PlanLayout L = plan_layout(T);
    const char* pb = (const char*)plan_scratch;
    const int64_t* rec_off = (const int64_t*)(pb + L.rec_off);
    const int32_t* run_id = (const int32_t*)(pb + L.run_id);
    const int32_t* run_start = (const int32_t*)(pb + L.run_start);
    const int32_t* hub_off = (const int32_t*)(pb + L.hub_off);
    const int32_t* chunk_off = (const int32_t*)(pb + L.chunk_off);
    const int32_t* long_list = (const int32_t*)(pb + L.long_list);
    char* hb = (char*)hub_scratch;
    uint32_t* pkey[2] = {(uint32_t*)(hb + H.pkey[0]), (uint32_t*)(hb + H.pkey[1])};
    uint32_t* pval[2] = {(uint32_t*)(hb + H.pval[0]), (uint32_t*)(hb + H.pval[1])};
    uint32_t* ekey[2] = {(uint32_t*)(hb + H.ekey[0]), (uint32_t*)(hb + H.ekey[1])};
    uint32_t* eval[2] = {(uint32_t*)(hb + H.eval[0]), (uint32_t*)(hb + H.eval[1])};
    int64_t* ent_off = (int64_t*)(hb + H.ent_off);
    uint2* items = (uint2*)(hb + H.items);
    int32_t* prun = (int32_t*)(hb + H.prun);
    unsigned long long* prec = (unsigned long long*)(hb + H.prec);
    unsigned long long* counters = (unsigned long long*)(hb + H.counters);
    uint32_t* key_bitmap = (uint32_t*)(hb + H.key_bits);
    int32_t* item_link = (int32_t*)(hb + H.item_link);
    uint16_t* sidx = (uint16_t*)(hb + H.sidx);
    const int bits = key_bits(n);
    const int th = 256;

    // positions are counted a window at a time: one pass of warp-per-item for the usual stream; a pass with
    // more positions (hub sources) uses the CTA-per-item variant with a CTA-wide window
    const int64_t warp_win = option(OCN_OPT_HUB_WINDOW, kHubWindow), cta_win = option(OCN_OPT_HUB_CTA_WINDOW, kHubCtaWindow);
    const bool cta = NP > warp_win;
    const int64_t win = cta ? (NP < cta_win ? NP : cta_win) : NP;
    const bool windowed = NP > win;  // several passes: every visit is tested against the window
    // exact index layout: 32-byte node entries with the exact run set of every list (<= 128 runs) and the run-segment
    // starts -- the shared pass skips the lists without an active run, the per-link kernel finds a run's entries by rank.
    // `seg`: the opt-in segment walker of the shared pass (measured slower than the whole-list walk, see above)
    // Both are opt-in (ocn_set_option): on the bench workload the 32-byte entries cost more than the 10 % of entry visits
    // the exact sets save (kernel 0.587 ms against 0.571 ms with the folded sets).
    const bool want_seg = option(OCN_OPT_HUB_WALKER, 0) == 1;
    const bool exact = !cta && !windowed && R <= kSegRuns && NP < 65536 && (want_seg || option(OCN_OPT_HUB_EXACT, 0) == 1);
    const bool seg = exact && want_seg;
    ps.seg = exact;

    OCN_CUDA(cudaMemsetAsync(counters, 0, sizeof(unsigned long long) * 4, st));
    OCN_CUDA(cudaMemsetAsync(key_bitmap, 0, sizeof(uint32_t) * (size_t)((n + 32) / 32), st));
    OCN_CUDA(cudaEventRecord(aux->ev[0], st));  // fork: everything before this call is visible to the auxiliary stream
    OCN_CUDA(cudaStreamWaitEvent(sa, aux->ev[0], 0));
    ps.forked = true;

    // auxiliary stream: (shared row, link) pairs, sorted by row, and the work items
    k_hub_item_links<<<grid_for(T * 32, th), th, 0, sa>>>(chunk_off, T, item_link);
    OCN_LAUNCH_CHECK();
    cub::DoubleBuffer<uint32_t> dk(pkey[0], pkey[1]), dv(pval[0], pval[1]);
    if (P > 0) {
        k_hub_emit_pairs<<<grid_for(T * 32, th), th, 0, sa>>>(rowptr, col, n, dst, T, hub_d, hub_off, run_id, run_pos_off,
                                                             pkey[0], pval[0]);
        OCN_LAUNCH_CHECK();
        k_hub_emit_pairs_long<<<sm_count() * 2, 1024, 0, sa>>>(rowptr, col, n, dst, hub_d, hub_off, long_list, plan_dev,
                                                              run_id, run_pos_off, pkey[0], pval[0]);
        OCN_LAUNCH_CHECK();
        size_t tb2 = H.cub_temp2_bytes;
        OCN_CUDA(cub::DeviceRadixSort::SortPairs(hb + H.cub_temp2, tb2, dk, dv, (int)P, 0, bits, sa));
        k_hub_items<<<grid_for(P, th), th, 0, sa>>>(rowptr, n, run_id, rec_off, dk.Current(), dv.Current(), P, prun, prec,
                                                    items, H.max_items, counters);
        OCN_LAUNCH_CHECK();
    }
    OCN_CUDA(cudaEventRecord(aux->ev[1], sa));

    // caller's stream: the inverted index of the source side
    void* tmp = hb + H.cub_temp;
    size_t tmp_bytes = H.cub_temp_bytes;
    cub::DoubleBuffer<uint32_t> ek(ekey[0], ekey[1]), ev(eval[0], eval[1]);
    k_hub_pos_count<<<grid_for(NP + 1, th), th, 0, st>>>(rowptr, col, src, run_start, run_pos_off, R, NP, ent_off);
    OCN_LAUNCH_CHECK();
    OCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, ent_off, ent_off, (int)(NP + 1), st));
    k_hub_emit_entries<<<grid_for(NP * 32, th), th, 0, st>>>(rowptr, col, src, run_start, run_pos_off, R, NP, ent_off,
                                                            ekey[0], eval[0]);
    OCN_LAUNCH_CHECK();
    OCN_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, ek, ev, (int)E, 0, bits, st));
    ps.ekey = ek.Current();
    ps.eval = ev.Current();
    ps.index_set = true;
    if (exact) {
        k_hub_entry_heads_seg<<<grid_for(E, th), th, 0, st>>>(ek.Current(), ev.Current(), run_pos_off, (int)R, E, 1, node_index, key_bitmap);
        OCN_LAUNCH_CHECK();
        k_hub_segments<<<grid_for(E, th), th, 0, st>>>(ek.Current(), ev.Current(), run_pos_off, (int)R, E, node_index, sidx);
    } else {
        k_hub_entry_heads<<<grid_for(E, th), th, 0, st>>>(ek.Current(), ev.Current(), run_pos_off, (int)R, E, 1, node_index, key_bitmap);
    }
    OCN_LAUNCH_CHECK();
    OCN_CUDA(cudaEventRecord(aux->ev[2], st));  // index complete

    // auxiliary stream: per link, C1, C2 and C3 through the rows that are not shared (needs the index only).
    // With the measurement hook set the two kernels run one after the other, so that k_cn_hub_count is timed alone.
    auto launch_link = [&](cudaStream_t s_) {
        if (exact)
            k_cn_link<true><<<sm_count() * 8, 256, 0, s_>>>(rowptr, col, dst, T, run_id, rec_off, chunk_off, run_pos_off, ev.Current(),
                                                            sidx, node_index, key_bitmap, item_link, hub_d, records);
        else
            k_cn_link<false><<<sm_count() * 8, 256, 0, s_>>>(rowptr, col, dst, T, run_id, rec_off, chunk_off, run_pos_off, ev.Current(),
                                                             sidx, node_index, key_bitmap, item_link, hub_d, records);
    };
    const bool timed_alone = g_hub_events[0] != nullptr;
    if (timed_alone) OCN_CUDA(cudaStreamWaitEvent(sa, aux->ev[1], 0));
    OCN_CUDA(cudaStreamWaitEvent(sa, aux->ev[2], 0));
    if (!timed_alone) launch_link(sa);
    OCN_LAUNCH_CHECK();
    OCN_CUDA(cudaEventRecord(aux->ev[3], sa));

    // caller's stream: the walk through the shared rows (needs the pairs and the items)
    OCN_CUDA(cudaStreamWaitEvent(st, aux->ev[1], 0));
    if (P > 0) {
        const int rpad = (int)((R + 1 + 3) & ~int64_t(3));
        const int uwords = (int)((((win + 1) >> 1) + 3) & ~int64_t(3));
        if (seg) {
            const size_t smem = sizeof(uint32_t) * ((size_t)rpad + (size_t)uwords * kHubWarps) + (size_t)24 * kSegQueue * kHubWarps;
            OCN_CUDA(cudaFuncSetAttribute(k_cn_hub_count_seg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = (int)((200u * 1024u) / (smem + 1024));
            const int cap = (int)option(OCN_OPT_HUB_SEG_CTAS, OCN_SEG_MINB);
            if (per_sm > cap) per_sm = cap;
            if (per_sm < 1) per_sm = 1;
            if (first_pass) hub_timing_record(0, st);
            k_cn_hub_count_seg<<<sm_count() * per_sm, kHubThreads, smem, st>>>(
                rowptr, col, dk.Current(), prun, prec, P, ev.Current(), sidx, node_index, key_bitmap, run_pos_off, (int)R,
                (uint32_t)NP, items, H.max_items, counters, records);
            OCN_LAUNCH_CHECK();
            if (first_pass) hub_timing_record(1, st);
        } else {
            const size_t smem = sizeof(uint32_t) * ((size_t)rpad + (size_t)uwords * (cta ? 1 : kHubWarps)) +
                                sizeof(uint2) * (size_t)kHubQueue * kHubWarps;
            auto kern = cta ? (windowed ? k_cn_hub_count<true, true, false> : k_cn_hub_count<true, false, false>)
                            : (windowed ? k_cn_hub_count<false, true, false>
                                        : (exact ? k_cn_hub_count<false, false, true> : k_cn_hub_count<false, false, false>));
            OCN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = (int)((200u * 1024u) / smem);
            if (per_sm > (cta ? 3 : 8)) per_sm = cta ? 3 : 8;
            if (per_sm < 1) per_sm = 1;
            if (first_pass) hub_timing_record(0, st);  // (the heavy pass, if any, runs on another stream: not part of the timing)
            for (int64_t w0 = 0; w0 < NP; w0 += win) {
                if (w0 > 0) OCN_CUDA(cudaMemsetAsync(counters + 1, 0, sizeof(unsigned long long), st));  // restart the item counter
                const int64_t wn = (NP - w0) < win ? (NP - w0) : win;
                kern<<<sm_count() * per_sm, kHubThreads, smem, st>>>(
                    rowptr, col, dk.Current(), prun, prec, P, ev.Current(), node_index, key_bitmap, run_pos_off, (int)R, (uint32_t)w0,
                    (uint32_t)wn, items, H.max_items, counters, records);
                OCN_LAUNCH_CHECK();
            }
            if (first_pass) hub_timing_record(1, st);
        }
    }
    if (timed_alone) {
        launch_link(st);
        OCN_LAUNCH_CHECK();
    }
    OCN_CUDA(cudaStreamWaitEvent(st, aux->ev[3], 0));  // join
    ps.forked = false;
    if (exact)
        k_hub_entry_heads_seg<<<grid_for(E, th), th, 0, st>>>(ek.Current(), ev.Current(), run_pos_off, (int)R, E, 0, node_index, key_bitmap);
    else
        k_hub_entry_heads<<<grid_for(E, th), th, 0, st>>>(ek.Current(), ev.Current(), run_pos_off, (int)R, E, 0, node_index, key_bitmap);
    OCN_LAUNCH_CHECK();
    ps.index_set = false;
    return OCN_OK;
}

// node_scratch buffers a failed call could not restore to zero: the next build on one of them is refused (the caller
// has to hand in a zeroed buffer) instead of computing from a dirty index
static std::mutex g_poison_mutex;
static std::set<const void*> g_poisoned;

static int hub_pass(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, const int64_t* dst, int64_t T,
                    const void* plan_scratch, const int64_t* plan_dev, const int64_t* run_pos_off, int64_t hub_d, int64_t P,
                    int64_t E, int64_t NP, int64_t R, const HubLayout& H, void* hub_scratch, void* node_scratch,
                    Record* records, HubAux* aux, bool first_pass, cudaStream_t st, const void* poison_key) {
    if (NP <= 0 || E <= 0) return OCN_OK;  // no source of this pass has a neighbour: its record sets are empty
    PassState ps;
    uint4* node_index = (uint4*)node_scratch;
    const int rc = hub_pass_body(rowptr, col, n, src, dst, T, plan_scratch, plan_dev, run_pos_off, hub_d, P, E, NP, R, H, hub_scratch,
                                 node_index, records, aux, first_pass, st, ps);
    if (rc == OCN_OK) return rc;
    // a launch or a runtime call failed half way: join the auxiliary stream and put the index back to zero, so that the
    // caller's stream sees one ordered (failed) call and the next call starts from the contract's all-zero state
    const std::string msg = last_error();
    bool clean = true;
    if (ps.forked) {
        clean = cudaEventRecord(aux->ev[3], aux->stream) == cudaSuccess && cudaStreamWaitEvent(st, aux->ev[3], 0) == cudaSuccess;
    }
    if (ps.index_set && clean) {
        const int th = 256;
        PlanLayout L = plan_layout(T);
        (void)L;
        if (ps.seg)
            k_hub_entry_heads_seg<<<grid_for(E, th), th, 0, st>>>(ps.ekey, ps.eval, run_pos_off, (int)R, E, 0, node_index, nullptr);
        else
            k_hub_entry_heads<<<grid_for(E, th), th, 0, st>>>(ps.ekey, ps.eval, run_pos_off, (int)R, E, 0, node_index, nullptr);
        clean = cudaGetLastError() == cudaSuccess;
    }
    if (!clean) {
        std::lock_guard<std::mutex> lock(g_poison_mutex);
        g_poisoned.insert(poison_key);
    }
    last_error() = msg + (clean ? " [index restored]" : " [node_scratch left dirty: hand in a zeroed buffer]");
    return rc;
}

// called by ocn_cn_build between the zeroing of the records and the column statistics.
// Runs of heavy sources (more than kHeavyRun neighbours; ocn_cn_plan) are indexed in a pass of their own: such a
// source alone contributes thousands of positions and millions of index entries, and in a common index every
// shared row of the stream would walk them (measured: a 65 536-link slice with one source of 5 637 neighbours took
// 9.2 ms instead of 1.2 ms).  In its own pass only the rows next to ITS links are walked.
int run_hub_stage(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, const int64_t* dst,
                  int64_t T, const void* plan_scratch, const int64_t* plan_dev, const int64_t* plan_host, void* hub_scratch,
                  size_t hub_scratch_bytes, void* node_scratch, Record* records, int64_t nnz, cudaStream_t st) {
    const int64_t hub_d = plan_host[OCN_PLAN_HUB_DEGREE];
    const int64_t P = plan_host[OCN_PLAN_HUB_PAIRS], E = plan_host[OCN_PLAN_HUB_ENTRIES];
    const int64_t NP = plan_host[OCN_PLAN_HUB_POSITIONS], R = plan_host[OCN_PLAN_NUM_RUNS];
    const int64_t E_heavy = plan_host[OCN_PLAN_HUB_ENTRIES_HEAVY], NP_heavy = plan_host[OCN_PLAN_HUB_POSITIONS_HEAVY];
    OCN_CHECK_ARG(hub_d > 0, "ocn_cn_build: the indexed path is off in this plan");
    OCN_CHECK_ARG(P < (int64_t(1) << 31) && E < (int64_t(1) << 31), "ocn_cn_build: indexed path limited to 2^31 pairs / entries");
    OCN_CHECK_ARG(R <= kHubMaxRuns && NP < (int64_t(1) << 31), "ocn_cn_build: indexed path with %lld runs, %lld positions",
                  (long long)R, (long long)NP);
    OCN_CHECK_ARG(E_heavy >= 0 && E_heavy <= E && NP_heavy >= 0 && NP_heavy <= NP, "ocn_cn_build: inconsistent plan");
    if (NP <= 0 || E <= 0) return OCN_OK;  // no source has a neighbour: every record set is empty
    {
        std::lock_guard<std::mutex> lock(g_poison_mutex);
        if (g_poisoned.count(node_scratch))
            return fail(OCN_EINVAL, "ocn_cn_build: node_scratch was left dirty by a failed call; pass a zeroed buffer "
                                    "(ocn_cn_hub_scratch_reset after zeroing this one)");
    }
    HubLayout H = hub_layout(n, nnz, P, E, NP, plan_host[OCN_PLAN_NUM_CHUNKS]);
    const size_t need = NP_heavy > 0 ? 2 * H.total : H.total;  // a heavy pass works in a second copy of the layout
    if (hub_scratch_bytes < need)
        return fail(OCN_ENOSPACE, "ocn_cn_build: hub scratch %zu < %zu bytes", hub_scratch_bytes, need);
    HubAux* aux = nullptr;
    if (int rc = hub_aux(st, &aux)) return rc;
    HubAux* aux_h = nullptr;  // created with the first call on this stream, not with the first hub source
    if (int rc = hub_aux(aux->heavy_stream, &aux_h)) return rc;
    PlanLayout L = plan_layout(T);
    const char* pb = (const char*)plan_scratch;
    if (NP_heavy == 0)  // the usual stream: one pass over the plain prefix
        return hub_pass(rowptr, col, n, src, dst, T, plan_scratch, plan_dev, (const int64_t*)(pb + L.run_pos_off), hub_d, P, E,
                        NP, R, H, hub_scratch, node_scratch, records, aux, true, st, node_scratch);
    // two passes side by side (disjoint records): the heavy one on its own stream with the second half of the
    // scratch and of the node index -- its kernels are short of parallelism (few links, long rows) and hide behind
    // the light pass
    OCN_CUDA(cudaEventRecord(aux->hev[0], st));
    OCN_CUDA(cudaStreamWaitEvent(aux->heavy_stream, aux->hev[0], 0));
    if (int rc = hub_pass(rowptr, col, n, src, dst, T, plan_scratch, plan_dev, (const int64_t*)(pb + L.run_pos_heavy), hub_d, P,
                          E_heavy, NP_heavy, R, H, (char*)hub_scratch + H.total, (uint4*)node_scratch + 2 * n, records, aux_h, false,
                          aux->heavy_stream, node_scratch))
        return rc;
    OCN_CUDA(cudaEventRecord(aux->hev[1], aux->heavy_stream));
    if (int rc = hub_pass(rowptr, col, n, src, dst, T, plan_scratch, plan_dev, (const int64_t*)(pb + L.pos_scanN), hub_d, P,
                          E - E_heavy, NP - NP_heavy, R, H, hub_scratch, node_scratch, records, aux, true, st, node_scratch))
        return rc;
    OCN_CUDA(cudaStreamWaitEvent(st, aux->hev[1], 0));
    return OCN_OK;
}

}  // namespace ocn

using namespace ocn;

extern "C" size_t ocn_cn_hub_bytes(int64_t n, int64_t nnz, const int64_t* plan_host) {
    if (n <= 0 || nnz < 0 || plan_host == nullptr) return 0;
    if (plan_host[OCN_PLAN_DENSE] != 0)
        return dense_bits_bytes(n) + (dense_whole_a2(n, plan_host) ? sizeof(uint32_t) * (size_t)n * (size_t)n : 0) + 256;
    if (plan_host[OCN_PLAN_HUB_DEGREE] <= 0) return 0;
    const size_t one = hub_layout(n, nnz, plan_host[OCN_PLAN_HUB_PAIRS], plan_host[OCN_PLAN_HUB_ENTRIES],
                                  plan_host[OCN_PLAN_HUB_POSITIONS], plan_host[OCN_PLAN_NUM_CHUNKS]).total;
    return plan_host[OCN_PLAN_HUB_POSITIONS_HEAVY] > 0 ? 2 * one : one;
}

extern "C" int ocn_cn_hub_scratch_reset(const void* node_scratch) {
    std::lock_guard<std::mutex> lock(g_poison_mutex);
    g_poisoned.erase(node_scratch);
    return OCN_OK;
}

extern "C" int ocn_cn_hub_timing_events(void* start_event, void* stop_event) {
    g_hub_events[0] = (cudaEvent_t)start_event;
    g_hub_events[1] = (cudaEvent_t)stop_event;
    return OCN_OK;
}
