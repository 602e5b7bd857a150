// ocn_cn_plan: device-side schedule for a stream of target links.
//
// The stream is cut into batches of batch_size links (utils.py:8-36 PermIterator).  Maximal runs of
// consecutive links with the same source node share everything that depends on N(src) -- in the
// citation2 evaluation stream every source is repeated against 1000 destinations
// (NeighborOverlapCitation2.py:248-252).  A run may cross a batch boundary: the records of a link do not
// depend on its batch, only the column statistics do (k_cn_colstat takes the batch from the link's
// position in the stream).  Round 1 cut the runs at the batch boundaries: 97 runs for the 66 sources of a
// 65 536-link session, i.e. a third of the index entries (and of the entry visits of the shared pass)
// were duplicates of another run of the same source.  A work unit of ocn_cn_build is
// (run, chunk of kPChunk positions of N(src), sub-list of kEdgeSub links of the run).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace ocn {

static size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

PlanLayout plan_layout(int64_t T) {
    PlanLayout L;
    size_t off = 0;
    L.rec_off = off;       off += align16(sizeof(int64_t) * (T + 2));  // [T + 1]: number of links with a heavy source
    L.run_id = off;        off += align16(sizeof(int32_t) * (T + 1));
    L.run_start = off;     off += align16(sizeof(int32_t) * (T + 2));
    L.run_unit_off = off;  off += align16(sizeof(int64_t) * (T + 2));
    L.cost_pre = off;      off += align16(sizeof(int64_t) * (T + 2));
    L.partial = off;       off += align16(sizeof(float) * 3 * (T + 1));
    L.hub_off = off;       off += align16(sizeof(int32_t) * (T + 2));
    L.run_pos_off = off;   off += align16(sizeof(int64_t) * (T + 2));
    L.run_pos_heavy = off; off += align16(sizeof(int64_t) * (T + 2));
    L.pos_start = off;     off += align16(sizeof(int64_t) * (T + 2));
    L.pos_scanN = off;     off += align16(sizeof(int64_t) * (T + 2));
    L.chunk_off = off;     off += align16(sizeof(int32_t) * (T + 2));
    L.long_list = off;     off += align16(sizeof(int32_t) * (T + 2));
    size_t b1 = 0, b2 = 0, b3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b1, (int64_t*)nullptr, (int64_t*)nullptr, (int)(T + 2));
    cub::DeviceScan::InclusiveSum(nullptr, b2, (int32_t*)nullptr, (int32_t*)nullptr, (int)(T + 2));
    cub::DeviceScan::ExclusiveSum(nullptr, b3, (int32_t*)nullptr, (int32_t*)nullptr, (int)(T + 2));
    if (b3 > b2) b2 = b3;
    L.cub_temp = off;
    L.cub_temp_bytes = align16((b1 > b2 ? b1 : b2) + 256);
    off += L.cub_temp_bytes;
    L.total = off;
    return L;
}

__global__ void k_plan_edges(const int64_t* __restrict__ rowptr, int64_t n, const int64_t* __restrict__ src,
                             const int64_t* __restrict__ dst, int64_t T,
                             int64_t batch_size, int64_t* __restrict__ rec_off, int32_t* __restrict__ flag,
                             int64_t* __restrict__ plan) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > T) return;
    if (t == T) {
        rec_off[t] = 0;
        flag[t] = 0;
        return;
    }
    int64_t i = src[t];
    // the reference raises IndexError for a link outside the graph; here such links are counted (the count rides on the
    // plan's read-back) and every later kernel of the plan returns at once, so nothing indexes rowptr with them
    const bool bad = (uint64_t)i >= (uint64_t)n || (uint64_t)dst[t] >= (uint64_t)n;
    if (bad) atomicAdd(reinterpret_cast<unsigned long long*>(plan + OCN_PLAN_BAD_LINKS), 1ull);
    if ((uint64_t)i >= (uint64_t)n) i = 0;
    const int64_t d = rowptr[i + 1] - rowptr[i];
    rec_off[t] = d;
    flag[t] = (t == 0 || src[t - 1] != i) ? 1 : 0;  // runs ignore batch boundaries (see the header)
    // the per-link kernels walk such a link with a whole CTA; they skip that phase when the count is zero
    if (d > kHeavyLink) atomicAdd(reinterpret_cast<unsigned long long*>(rec_off + T + 1), 1ull);
    // ... and the run-grouped kernels leave links with a source of more than kGroupedMaxDeg neighbours to them (rare in
    // a stream of uniformly drawn sources: one atomic per such link)
    if (d > kGroupedMaxDeg) atomicAdd(reinterpret_cast<unsigned long long*>(plan + OCN_PLAN_WIDE_LINKS), 1ull);
}

// after the inclusive scan flag[t] = run index + 1
__global__ void k_plan_runs(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ src, int64_t T,
                            int64_t batch_size, const int32_t* __restrict__ run_incl, int32_t* __restrict__ run_start,
                            int64_t* __restrict__ plan) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || plan[OCN_PLAN_BAD_LINKS] != 0) return;
    bool first = (t == 0) || src[t - 1] != src[t];
    int32_t r = run_incl[t] - 1;
    if (first) {
        run_start[r] = (int32_t)t;
        // positions of the stream = sum over runs of deg(src) (k_plan_hub_decide needs it before the run-level scan)
        atomicAdd((unsigned long long*)&plan[OCN_PLAN_HUB_POSITIONS], (unsigned long long)(rowptr[src[t] + 1] - rowptr[src[t]]));
    }
    if (t == T - 1) {
        run_start[r + 1] = (int32_t)T;
        plan[OCN_PLAN_NUM_RUNS] = r + 1;
    }
}

// one warp per link: walk cost = number of columns the j-side walk probes (+ a constant)
// Rows N(m) of at least plan[OCN_PLAN_HUB_DEGREE] columns are not walked per link: the hub stage
// (cn_hub.cu) streams each of them once for all links of the stream; they are counted into hub_cnt.
// The warp covers the first kLongRow neighbours of dst; a destination with more is queued for
// k_plan_cost_long (one CTA per such link), so that a hub destination is not a serial tail.
__device__ __forceinline__ void cost_of_row(const int64_t* __restrict__ rowptr, int32_t m, int64_t hub_d, long long& w, int& hubs) {
    const long long dm = ldg_i64(rowptr + m + 1) - ldg_i64(rowptr + m);
    if (hub_d > 0 && dm >= hub_d) ++hubs; else w += dm;
}

__global__ void k_plan_cost(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                            const int64_t* __restrict__ dst, int64_t T, int order, int64_t* __restrict__ cost,
                            int32_t* __restrict__ hub_cnt, int32_t* __restrict__ chunk_cnt, int32_t* __restrict__ long_list,
                            int64_t* __restrict__ plan) {
    __shared__ unsigned long long s_cost;  // one global reduction per CTA, not one per link on a single address
    if (plan[OCN_PLAN_BAD_LINKS] != 0) return;
    if (threadIdx.x == 0) s_cost = 0ull;
    __syncthreads();
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t == T && lane == 0) { cost[t] = 0; hub_cnt[t] = 0; chunk_cnt[t] = 0; }
    if (t < T) {
        const int64_t hub_d = plan[OCN_PLAN_HUB_DEGREE];
        const int64_t j = dst[t];
        const int64_t rs = rowptr[j], d = rowptr[j + 1] - rs;
        long long w = 0;
        int hubs = 0;
        if (order >= 3) {
            const int64_t dcap = d < kLongRow ? d : kLongRow;
            for (int64_t o = lane; o < dcap; o += 32) cost_of_row(rowptr, ldg_i32(col + rs + o), hub_d, w, hubs);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                w += __shfl_xor_sync(0xffffffffu, w, o);
                hubs += __shfl_xor_sync(0xffffffffu, hubs, o);
            }
        }
        if (order >= 2) w += d;
        w += kLinkCost;
        if (lane == 0) {
            cost[t] = w;
            hub_cnt[t] = hubs;
            chunk_cnt[t] = (int32_t)((d + 31) >> 5);
            atomicAdd(&s_cost, (unsigned long long)w);
            if (order >= 3 && d > kLongRow)
                long_list[atomicAdd((unsigned long long*)&plan[OCN_PLAN_LONG_COUNT], 1ull)] = (int32_t)t;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cost) atomicAdd((unsigned long long*)&plan[OCN_PLAN_TOTAL_COST], s_cost);
}

__global__ void __launch_bounds__(1024)
k_plan_cost_long(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int64_t* __restrict__ dst,
                 int64_t* __restrict__ cost, int32_t* __restrict__ hub_cnt, const int32_t* __restrict__ long_list,
                 int64_t* __restrict__ plan) {
    __shared__ unsigned long long s_w;
    __shared__ int s_h;
    if (plan[OCN_PLAN_BAD_LINKS] != 0) return;
    const int64_t n_long = plan[OCN_PLAN_LONG_COUNT];
    const int64_t hub_d = plan[OCN_PLAN_HUB_DEGREE];
    for (int64_t i = blockIdx.x; i < n_long; i += gridDim.x) {
        const int64_t t = long_list[i];
        const int64_t j = dst[t];
        const int64_t rs = rowptr[j], d = rowptr[j + 1] - rs;
        if (threadIdx.x == 0) { s_w = 0ull; s_h = 0; }
        __syncthreads();
        long long w = 0;
        int hubs = 0;
        for (int64_t o = kLongRow + threadIdx.x; o < d; o += blockDim.x) cost_of_row(rowptr, ldg_i32(col + rs + o), hub_d, w, hubs);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            w += __shfl_xor_sync(0xffffffffu, w, o);
            hubs += __shfl_xor_sync(0xffffffffu, hubs, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&s_w, (unsigned long long)w);
            atomicAdd(&s_h, hubs);
        }
        __syncthreads();
        if (threadIdx.x == 0) {  // each long link is owned by one CTA: plain read-modify-write
            cost[t] += (int64_t)s_w;
            hub_cnt[t] += s_h;
            atomicAdd((unsigned long long*)&plan[OCN_PLAN_TOTAL_COST], s_w);
        }
        __syncthreads();
    }
}

// the indexed path (cn_hub.cu) is used for order 3 when the stream has few runs (its run -> position table
// lives in shared memory, and its index holds one entry per (position, neighbour of the position's node))
// In automatic mode the stream must also HAVE runs: with one link per run (ungrouped links, the training shape) the
// index holds a private neighbourhood per link and nothing is shared -- measured at citation2 shape, 2048 uniform links:
// 6.6 ms indexed against 1.1 ms with the per-run tables (scripts/probe_uniform_order3.py).
__global__ void k_plan_hub_decide(int order, int64_t hub_degree, int automatic, int64_t T, int64_t* __restrict__ plan) {
    const int64_t runs = plan[OCN_PLAN_NUM_RUNS];
    const bool fits = runs <= kHubMaxRuns && (!automatic || T >= 4 * runs);
    plan[OCN_PLAN_HUB_DEGREE] = (order >= 3 && hub_degree > 0 && fits) ? hub_degree : 0;
}

// one warp per run: (table passes over all 32-position chunks of N(src)) x (cost windows of the run)
__global__ void k_plan_units(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                             const int64_t* __restrict__ src, int64_t T, const int32_t* __restrict__ run_start,
                             const int64_t* __restrict__ cost_pre, int resident_ctas,
                             int64_t heavy_run, int order, int64_t* __restrict__ plan, int64_t* __restrict__ run_units,
                             int64_t* __restrict__ run_pos_light, int64_t* __restrict__ run_pos_heavy) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r > T + 1 || plan[OCN_PLAN_BAD_LINKS] != 0) return;
    const int64_t n_runs = plan[OCN_PLAN_NUM_RUNS];
    int64_t u = 0, npos = 0;
    // orders <= 2 on short runs go to the table-free kernel (k_plan_finish: OCN_PLAN_USE_DIRECT): no table units to plan
    const bool direct = order <= 2 && T <= 3 * n_runs;
    if (r < n_runs && !direct) {
        const int64_t t0 = run_start[r], len = run_start[r + 1] - t0;
        const int64_t i = src[t0];
        const int64_t rs = rowptr[i], d = rowptr[i + 1] - rs;
        int64_t passes = 0;
        long long keys = 0;
        for (int64_t c0 = 0; c0 < d; c0 += kPChunk) {
            long long f = 0;
            if (c0 + lane < d) {
                const int32_t k = col[rs + c0 + lane];
                f = rowptr[k + 1] - rowptr[k];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
            passes += (f + kCap - 1) / kCap;
            keys += f;
        }
        if (plan[OCN_PLAN_HUB_DEGREE] > 0) {  // the hub stage lists every (key, run, position) entry
            npos = d;
            if (lane == 0) {
                atomicAdd((unsigned long long*)&plan[OCN_PLAN_HUB_ENTRIES], (unsigned long long)keys);
                if (d > heavy_run) atomicAdd((unsigned long long*)&plan[OCN_PLAN_HUB_ENTRIES_HEAVY], (unsigned long long)keys);
            }
        }
        const long long W = unit_budget(plan[OCN_PLAN_TOTAL_COST], resident_ctas);
        const long long run_cost = cost_pre[t0 + len] - cost_pre[t0];
        u = passes * ((run_cost + W - 1) / W);
    }
    if (lane == 0) {
        run_units[r] = u;
        // the runs of heavy sources (long rows N(src)) get an index of their own (cn_hub.cu): two position
        // numberings, one over the light runs and one over the heavy runs (a run of the other class has no positions)
        const bool heavy = npos > heavy_run;
        run_pos_light[r] = heavy ? 0 : npos;
        run_pos_heavy[r] = heavy ? npos : 0;
    }
}

// after the two run-level scans: origin[r] = first position of run r in run order (ascending in r),
// start[r] = first position of run r in the stream's position numbering (light runs first, heavy runs last)
__global__ void k_plan_positions(const int64_t* __restrict__ plan, const int64_t* __restrict__ scan_light,
                                 const int64_t* __restrict__ scan_heavy, int64_t* __restrict__ origin,
                                 int64_t* __restrict__ start) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_runs = plan[OCN_PLAN_NUM_RUNS];
    if (r > n_runs) return;
    const int64_t total_light = scan_light[n_runs];
    origin[r] = scan_light[r] + scan_heavy[r];
    if (r < n_runs) {
        const bool heavy = scan_heavy[r + 1] != scan_heavy[r];
        start[r] = heavy ? total_light + scan_heavy[r] : scan_light[r];
    } else {
        start[r] = total_light + scan_heavy[r];
    }
}

__global__ void k_plan_finish(int64_t T, int64_t batch_size, int resident_ctas, int order, int64_t n, const int64_t* __restrict__ rowptr,
                              const int64_t* __restrict__ rec_off,
                              const int64_t* __restrict__ run_unit_off, const int32_t* __restrict__ hub_off,
                              const int64_t* __restrict__ run_pos_off, const int64_t* __restrict__ run_pos_heavy_off,
                              const int32_t* __restrict__ chunk_off, int64_t* __restrict__ plan) {
    plan[OCN_PLAN_NUM_RECORDS] = rec_off[T];
    plan[OCN_PLAN_NUM_UNITS] = run_unit_off[plan[OCN_PLAN_NUM_RUNS]];
    plan[OCN_PLAN_NUM_BATCHES] = (T + batch_size - 1) / batch_size;
    plan[OCN_PLAN_UNIT_COUNTER] = 0;  // dynamic unit counter of ocn_cn_build
    plan[OCN_PLAN_BUDGET] = unit_budget(plan[OCN_PLAN_TOTAL_COST], resident_ctas);
    // orders <= 2: the table pays off only when several links share it (runs of >= 3 links on average)
    plan[OCN_PLAN_USE_DIRECT] = (T <= 3 * plan[OCN_PLAN_NUM_RUNS]) ? 1 : 0;
    // dense graphs (ddi: density 0.147, the 2-hop frontier of a link is the whole graph several times over): C2 of a record
    // is popc(row(dst) & row(k_p)) on bit-vector rows -- 0.3 M links/s through the walk kernels, ~30 M/s this way
    const int64_t nnz = rowptr[n];
    plan[OCN_PLAN_DENSE] = (order <= 2 && n <= 32768 && nnz >= n * (n / 64 + 1) && plan[OCN_PLAN_BAD_LINKS] == 0) ? 1 : 0;
    plan[OCN_PLAN_HUB_PAIRS] = hub_off[T];
    const int64_t n_runs = plan[OCN_PLAN_NUM_RUNS];
    plan[OCN_PLAN_HUB_POSITIONS] = run_pos_off[n_runs] + run_pos_heavy_off[n_runs];  // 0 when the indexed path is off
    plan[OCN_PLAN_HUB_POSITIONS_HEAVY] = run_pos_heavy_off[n_runs];
    plan[OCN_PLAN_NUM_CHUNKS] = chunk_off[T];
}

// Small streams (a Cora / Pubmed batch: 1 - 2 thousand links) are launch-bound: the plan's eight prefix sums were sixteen
// launches of cub::DeviceScan (init + scan).  Up to kSmallScan elements the sums of a group run as ONE launch, a CTA per
// array (block scans of 1024 elements with a running carry, in place).
constexpr int kSmallScan = 8192;
struct ScanJob { void* data; int count; int is64; int inclusive; };
struct ScanJobs { ScanJob job[3]; };
__global__ void __launch_bounds__(1024) k_scan_small(ScanJobs jobs) {
    using Scan = cub::BlockScan<long long, 1024>;
    __shared__ typename Scan::TempStorage tmp;
    const ScanJob J = jobs.job[blockIdx.x];
    long long running = 0;
    for (int base = 0; base < J.count; base += 1024) {
        const int i = base + (int)threadIdx.x;
        long long v = 0;
        if (i < J.count) v = J.is64 ? reinterpret_cast<const long long*>(J.data)[i] : (long long)reinterpret_cast<const int*>(J.data)[i];
        long long ex = 0, total = 0;
        Scan(tmp).ExclusiveSum(v, ex, total);
        const long long out = running + ex + (J.inclusive ? v : 0);
        if (i < J.count) {
            if (J.is64) reinterpret_cast<long long*>(J.data)[i] = out; else reinterpret_cast<int*>(J.data)[i] = (int)out;
        }
        running += total;
        __syncthreads();
    }
}

}  // namespace ocn

using namespace ocn;

extern "C" {

size_t ocn_cn_plan_bytes(int64_t num_edges) {
    if (num_edges < 0) return 0;
    return plan_layout(num_edges).total;
}
size_t ocn_cn_colstat_bytes(int64_t n) { return n < 0 ? 0 : sizeof(ColStat) * (size_t)n; }
size_t ocn_cn_record_bytes(void) { return sizeof(Record); }

int ocn_cn_plan(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, const int64_t* dst, int64_t num_edges,
                int64_t batch_size, int order, int64_t hub_degree, void* plan_scratch, size_t plan_scratch_bytes,
                int64_t* out_plan, void* stream) {
    OCN_RANGE("ocn_cn_plan");
    OCN_CHECK_ARG(rowptr && col && out_plan && plan_scratch, "ocn_cn_plan: null pointer");
    OCN_CHECK_ARG(n > 0 && num_edges > 0 && batch_size > 0, "ocn_cn_plan: n, num_edges and batch_size must be positive");
    OCN_CHECK_ARG(num_edges < (int64_t(1) << 30), "ocn_cn_plan: at most 2^30 links per call");
    OCN_CHECK_ARG(src && dst, "ocn_cn_plan: null edge pointer");
    OCN_CHECK_ARG(order >= 1 && order <= 3, "ocn_cn_plan: order must be 1..3");
    OCN_CHECK_ARG(hub_degree >= -1, "ocn_cn_plan: hub_degree must be -1 (off), 0 (auto) or a degree");
    PlanLayout L = plan_layout(num_edges);
    if (plan_scratch_bytes < L.total)
        return fail(OCN_ENOSPACE, "ocn_cn_plan: plan scratch %zu < %zu bytes", plan_scratch_bytes, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)plan_scratch;
    int64_t* rec_off = (int64_t*)(base + L.rec_off);
    int32_t* run_id = (int32_t*)(base + L.run_id);
    int32_t* run_start = (int32_t*)(base + L.run_start);
    int64_t* run_unit_off = (int64_t*)(base + L.run_unit_off);
    void* tmp = base + L.cub_temp;
    size_t tmp_bytes = L.cub_temp_bytes;
    int64_t T = num_edges;
    int64_t* cost_pre = (int64_t*)(base + L.cost_pre);
    const int resident_ctas = sm_count() * kBuildCtasPerSm;
    int threads = 256;
    int blocks = (int)((T + 1 + threads - 1) / threads);
    OCN_CUDA(cudaMemsetAsync(out_plan, 0, sizeof(int64_t) * OCN_PLAN_WORDS, st));
    int32_t* hub_off = (int32_t*)(base + L.hub_off);
    int64_t* run_pos_off = (int64_t*)(base + L.run_pos_off);
    int64_t* run_pos_heavy = (int64_t*)(base + L.run_pos_heavy);
    int64_t* pos_start = (int64_t*)(base + L.pos_start);
    int64_t* pos_scanN = (int64_t*)(base + L.pos_scanN);
    const int automatic = hub_degree == 0;
    if (hub_degree == 0) {  // auto: share a row once about one link of the stream is expected to walk it
        hub_degree = (n + num_edges - 1) / num_edges;  // (citation2 shape, n/T = 45: flat optimum between 24 and 64)
        if (hub_degree < 32) hub_degree = 32;
    }
    OCN_CUDA(cudaMemsetAsync(rec_off + T + 1, 0, sizeof(int64_t), st));
    k_plan_edges<<<blocks, threads, 0, st>>>(rowptr, n, src, dst, T, batch_size, rec_off, run_id, out_plan);
    OCN_LAUNCH_CHECK();
    const bool small_scans = T + 2 <= kSmallScan;
    if (small_scans) {
        const ScanJobs jobs = {{{rec_off, (int)(T + 1), 1, 0}, {run_id, (int)(T + 1), 0, 1}, {nullptr, 0, 0, 0}}};
        k_scan_small<<<2, 1024, 0, st>>>(jobs);
        OCN_LAUNCH_CHECK();
    } else {
        OCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, rec_off, rec_off, (int)(T + 1), st));
        OCN_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, run_id, run_id, (int)(T + 1), st));
    }
    k_plan_runs<<<blocks, threads, 0, st>>>(rowptr, src, T, batch_size, run_id, run_start, out_plan);
    OCN_LAUNCH_CHECK();
    k_plan_hub_decide<<<1, 1, 0, st>>>(order, hub_degree, automatic, T, out_plan);
    OCN_LAUNCH_CHECK();
    int blocks_w = (int)(((T + 1) * 32 + threads - 1) / threads);
    int32_t* chunk_off = (int32_t*)(base + L.chunk_off);
    int32_t* long_list = (int32_t*)(base + L.long_list);
    k_plan_cost<<<blocks_w, threads, 0, st>>>(rowptr, col, dst, T, order, cost_pre, hub_off, chunk_off, long_list, out_plan);
    OCN_LAUNCH_CHECK();
    if (order >= 3) {
        k_plan_cost_long<<<sm_count() * 2, 1024, 0, st>>>(rowptr, col, dst, cost_pre, hub_off, long_list, out_plan);
        OCN_LAUNCH_CHECK();
    }
    if (small_scans) {
        const ScanJobs jobs = {{{cost_pre, (int)(T + 1), 1, 0}, {hub_off, (int)(T + 1), 0, 0}, {chunk_off, (int)(T + 1), 0, 0}}};
        k_scan_small<<<3, 1024, 0, st>>>(jobs);
        OCN_LAUNCH_CHECK();
    } else {
        OCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cost_pre, cost_pre, (int)(T + 1), st));
        OCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, hub_off, hub_off, (int)(T + 1), st));
        OCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, chunk_off, chunk_off, (int)(T + 1), st));
    }
    int blocks2 = (int)(((T + 2) * 32 + threads - 1) / threads);
    const int64_t heavy_run = option(OCN_OPT_HUB_HEAVY_RUN, kHeavyRun);
    k_plan_units<<<blocks2, threads, 0, st>>>(rowptr, col, src, T, run_start, cost_pre, resident_ctas, heavy_run, order,
                                              out_plan, run_unit_off, pos_scanN, run_pos_heavy);
    OCN_LAUNCH_CHECK();
    if (small_scans) {
        const ScanJobs jobs = {{{run_unit_off, (int)(T + 2), 1, 0}, {pos_scanN, (int)(T + 2), 1, 0}, {run_pos_heavy, (int)(T + 2), 1, 0}}};
        k_scan_small<<<3, 1024, 0, st>>>(jobs);
        OCN_LAUNCH_CHECK();
    } else {
        OCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, run_unit_off, run_unit_off, (int)(T + 2), st));
        OCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, pos_scanN, pos_scanN, (int)(T + 2), st));
        OCN_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, run_pos_heavy, run_pos_heavy, (int)(T + 2), st));
    }
    k_plan_positions<<<(int)((T + 2 + threads - 1) / threads), threads, 0, st>>>(out_plan, pos_scanN, run_pos_heavy, run_pos_off,
                                                                                pos_start);
    OCN_LAUNCH_CHECK();
    k_plan_finish<<<1, 1, 0, st>>>(T, batch_size, resident_ctas, order, n, rowptr, rec_off, run_unit_off, hub_off, pos_scanN, run_pos_heavy, chunk_off, out_plan);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // extern "C"
