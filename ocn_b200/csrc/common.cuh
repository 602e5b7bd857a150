// Shared device/host helpers of libocn_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/ocn_b200.h"

namespace ocn {

// ---- error plumbing (no exceptions cross the C ABI) --------------------------------------
std::string& last_error();
int fail(int code, const char* fmt, ...);

#define OCN_CHECK_ARG(cond, ...)                         \
    do {                                                 \
        if (!(cond)) return ::ocn::fail(OCN_EINVAL, __VA_ARGS__); \
    } while (0)

#define OCN_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::ocn::fail(OCN_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                         \
    } while (0)

// every launch of one of the library's own kernels is followed by this: the error check, and the count bench.py reports
// as gpu_launches (ocn_launch_count)
void count_launch();
long long launch_count();
#define OCN_LAUNCH_CHECK()            \
    do {                              \
        ::ocn::count_launch();        \
        OCN_CUDA(cudaGetLastError()); \
    } while (0)

int sm_count();

// NVTX range around an entry point (tracing): visible in Nsight Systems / ncu --nvtx, a no-op (one branch on a
// cached flag inside the header-only nvtx3 shim) when no tool is attached
struct NvtxRange {
    explicit NvtxRange(const char* name);
    ~NvtxRange();
};
#define OCN_RANGE(name) ::ocn::NvtxRange _ocn_range_(name)
int64_t option(int key, int64_t dflt);  // ocn_set_option value, or dflt when unset (0)
int set_option(int key, int64_t value);

// ---- constants shared by the plan and the build kernel -----------------------------------
constexpr int kPChunk = 32;   // positions of N(src) handled by one work unit (one mask word)
constexpr int kEdgeSub = 32;  // target links of a run handled by one work unit
constexpr int kSlots = 20399;  // (key, mask) slots of the shared-memory table of ocn_cn_build (prime, 159 KB)
constexpr int kCap = (kSlots * 7) / 10;  // keys inserted per table pass (load factor 0.7)

// per-batch, per-node column statistics (32 B = one L2 sector)
struct __align__(32) ColStat {
    uint32_t c1;   // number of links of the batch that have this node in CN1
    uint32_t pad0;
    unsigned long long s2;  // sum over links of C2 (weighted) or [C2>0]
    unsigned long long s3;  // same for C3
    unsigned long long pad1;
};
static_assert(sizeof(ColStat) == 32, "ColStat must be one sector");

// plan scratch layout (offsets in bytes, all 16-B aligned)
struct PlanLayout {
    size_t rec_off;        // int64[T+2]: exclusive prefix of deg(src) over the links; [T+1] = number of links with a heavy source
    size_t run_id;         // int32[T+1]
    size_t run_start;      // int32[T+2]
    size_t run_unit_off;   // int64[T+2]
    size_t cost_pre;       // int64[T+2]  exclusive prefix of the per-link walk cost
    size_t partial;        // float[3*T]
    size_t hub_off;        // int32[T+2]  exclusive prefix of the per-link number of hub rows in N(dst)
    size_t run_pos_off;    // int64[T+2]  exclusive prefix of deg(src) over the runs, in run order (ascending)
    size_t run_pos_heavy;  // int64[T+2]  the same prefix over the heavy runs only
    size_t pos_start;      // int64[T+2]  first position of every run in the index numbering (light runs first, heavy last)
    size_t pos_scanN;      // int64[T+2]  the prefix over the light runs only
    size_t chunk_off;      // int32[T+2]  exclusive prefix of ceil(deg(dst)/32) (items of the per-link kernel)
    size_t long_list;      // int32[T+2]  links whose destination has more than kLongRow neighbours
    size_t cub_temp;       // bytes
    size_t cub_temp_bytes;
    size_t total;
};
PlanLayout plan_layout(int64_t num_edges);

// one record per (target link, position p in N(src)):
//   x = C2 | (C1 << 31)   (C1 in {0,1}; C2 = #2-walks dst->..->N(src)[p], < 2^31)
//   y = C3                (#3-walks)
using Record = uint2;

// hub stage scratch (cn_hub.cu), offsets in bytes
struct HubLayout {
    size_t pkey[2], pval[2];  // uint32[P] x2 each: (hub row, link) pairs, radix-sort double buffers
    size_t ekey[2], eval[2];  // uint32[E] x2 each: (key, position of the stream) entries
    size_t ent_off;           // int64[positions + 2]
    size_t items;             // uint2[max_items]
    int64_t max_items;
    size_t counters;          // uint64[4]: number of items, next item
    size_t prun, prec;        // int32[P] run / uint64[P] record offset of the link of every sorted pair
    size_t key_bits;          // uint32[ceil((n + 1) / 32)]: bit l = node l is a key of the index (has an entry list)
    size_t item_link;         // int32[chunks + 1]: link of every work item of the per-link kernel
    size_t sidx;              // uint16[E]: run-segment starts of the lists with several entries per run (k_hub_segments)
    size_t cub_temp, cub_temp_bytes;    // entry pipeline (caller's stream)
    size_t cub_temp2, cub_temp2_bytes;  // pair pipeline (auxiliary stream)
    size_t total;
};
HubLayout hub_layout(int64_t n, int64_t nnz, int64_t pairs, int64_t entries, int64_t positions, int64_t chunks);
int run_hub_stage(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, const int64_t* dst,
                  int64_t T, const void* plan_scratch, const int64_t* plan_dev, const int64_t* plan_host, void* hub_scratch,
                  size_t hub_scratch_bytes, void* node_scratch, Record* records, int64_t nnz, cudaStream_t st);

// rows of the graph as bit vectors, W words each (spgemm.cu; also the dense order-2 build of cn_build.cu)
__global__ void k_dense_bits(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n, int W,
                             uint32_t* __restrict__ bits);
inline int dense_words(int64_t n) { return (int)(((n + 31) / 32 + 3) & ~int64_t(3)); }
// The dense build takes its 2-walk counts from the whole matrix A^2 (n x n uint32 behind the bit rows in the dense scratch, a
// tiled AND / popcount product) when the stream has at least half as many positions as the matrix has entries and the matrix
// stays within 256 MB; per-position row products otherwise.
inline bool dense_whole_a2(int64_t n, const int64_t* plan_host) {
    return n <= 8192 && plan_host[OCN_PLAN_NUM_RECORDS] * 2 >= n * n;
}
inline size_t dense_bits_bytes(int64_t n) { return (sizeof(uint32_t) * (size_t)n * (size_t)dense_words(n) + 255) & ~size_t(255); }

// run-grouped kernels after the build (cn_grouped.cu): used when the stream averages >= kGroupedMinRun links per run
constexpr int kGroupedMinRun = 16;
constexpr int kGroupedMaxDeg = 64;   // sources with more neighbours (more than two position tiles) stay with the per-link kernels
bool use_grouped(int64_t T, const int64_t* plan_host);
int grouped_colstat(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, int64_t T, int64_t batch_size,
                    int weighted, const void* plan_scratch, const Record* records, ColStat* colstat, cudaStream_t st);
int grouped_stats(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, int64_t T, int64_t batch_size,
                  int order, int weighted, int variant, float fill, const float* ip, int stage, const void* plan_scratch,
                  const Record* records, const ColStat* colstat, float* bscal, float* partial, cudaStream_t st);
size_t grouped_aggregate_smem(int nvec);
int grouped_aggregate(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, const int64_t* dst, int64_t T,
                      int64_t batch_size, int order, int weighted, int variant, float fill, const float* ip,
                      const void* plan_scratch, const Record* records, const ColStat* colstat, const float* bscal,
                      const float* x, int nvec, float* xcn1, float* xcn2, float* xcn3, float* xij, cudaStream_t st);
int grouped_release(const int64_t* rowptr, const int32_t* col, int64_t n, const int64_t* src, int64_t T, int64_t batch_size,
                    const void* plan_scratch, ColStat* colstat, cudaStream_t st);

// cost window (in probed columns) one work unit of ocn_cn_build covers: about 8 units per resident CTA,
// clamped so that a unit amortises its table build but a heavy link is still split over many CTAs
__host__ __device__ inline long long unit_budget(long long total_cost, int resident_ctas) {
    long long w = total_cost / (4ll * (resident_ctas > 0 ? resident_ctas : 1));
    if (w < 65536) w = 65536;
    if (w > 1048576) w = 1048576;
    return w;
}
constexpr int kBuildCtasPerSm = 1;  // one 1024-thread CTA per SM owns (almost) all shared memory
constexpr int kLinkCost = 64;  // fixed cost added to every link so that empty links still advance

// plan[] words beyond the public ones
#define OCN_PLAN_UNIT_COUNTER 4
#define OCN_PLAN_BUDGET 5
#define OCN_PLAN_TOTAL_COST 6
#define OCN_PLAN_USE_DIRECT 7  /* orders <= 2 only: 1 = table-free kernel (short runs), 0 = table kernel */
#define OCN_PLAN_LONG_COUNT 12 /* entries of the long-destination list */
#define OCN_PLAN_HUB_ENTRIES_HEAVY 13   /* the part of OCN_PLAN_HUB_ENTRIES that belongs to runs of heavy sources */
#define OCN_PLAN_NUM_CHUNKS 15           /* work items (link, 32 rows of N(dst)) of the per-link kernel of the indexed path */
#define OCN_PLAN_HUB_POSITIONS_HEAVY 14 /* the part of OCN_PLAN_HUB_POSITIONS that belongs to runs of heavy sources */
constexpr int kHeavyRun = 1024;    // a run whose source has more neighbours than this is "heavy": indexed in a pass of its own
constexpr int kHeavyLink = 1024;   // a link whose source has more neighbours is walked by a whole CTA in the per-link kernels
constexpr int kLongRow = 256;      // neighbours of dst a single warp walks in the plan / pair kernels; the rest goes to a CTA
constexpr int kHubMaxRuns = 2048;       // indexed path: run -> first position table in shared memory

// ---- small device helpers ------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ int32_t ldg_i32(const int32_t* p) { return __ldg(p); }
__device__ __forceinline__ int64_t ldg_i64(const int64_t* p) { return __ldg(reinterpret_cast<const long long*>(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// lower_bound on an ascending int32 row in global memory; returns true if key present
__device__ __forceinline__ bool row_contains(const int32_t* __restrict__ row, int64_t len, int32_t key) {
    int64_t lo = 0, hi = len;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        int32_t v = __ldg(row + mid);
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return lo < len && __ldg(row + lo) == key;
}

// The links with a heavy source (more than kHeavyLink neighbours) among t = blockIdx.x, blockIdx.x + gridDim.x, ...:
// the CTA tests blockDim.x candidates at a time (two dependent loads in all, not two per candidate in turn), collects
// the heavy ones in shared memory and calls body(t) for each with ALL its threads.  rec_off[T + 1] is the plan's count
// of such links in the stream (0: nothing to do).
template <typename Body>
__device__ __forceinline__ void for_each_heavy_link(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ src,
                                                    int64_t T, const int64_t* __restrict__ rec_off, Body&& body) {
    __shared__ int s_n;
    __shared__ long long s_list[1024];
    if (rec_off[T + 1] == 0) return;
    for (int64_t c0 = 0; (int64_t)blockIdx.x + c0 * gridDim.x < T; c0 += blockDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const int64_t t = (int64_t)blockIdx.x + (c0 + threadIdx.x) * gridDim.x;
        if (t < T) {
            const int64_t i = src[t];
            if (rowptr[i + 1] - rowptr[i] > kHeavyLink) s_list[atomicAdd(&s_n, 1)] = t;
        }
        __syncthreads();
        const int n = s_n;
        for (int k = 0; k < n; ++k) body((int64_t)s_list[k]);
    }
}

}  // namespace ocn
