// ABI basics, graph validation and the generic two-matrix row intersection.
#include <stdarg.h>

#include <atomic>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

namespace ocn {

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

std::string& last_error() {
    static thread_local std::string s;
    return s;
}

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

static int64_t g_options[OCN_OPT_COUNT] = {0};
int64_t option(int key, int64_t dflt) {
    const int64_t v = (key >= 0 && key < OCN_OPT_COUNT) ? g_options[key] : 0;
    return v != 0 ? v : dflt;
}
int set_option(int key, int64_t value) {
    if (key < 0 || key >= OCN_OPT_COUNT) return fail(OCN_EINVAL, "ocn_set_option: unknown key %d", key);
    g_options[key] = value;
    return OCN_OK;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

// ---- validation ---------------------------------------------------------------------------
__global__ void k_validate(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                           int64_t nnz, int32_t* __restrict__ flags) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int lane = lane_id();
    int bad = 0;
    for (int64_t r = warp; r < n; r += nwarps) {
        int64_t s = rowptr[r], e = rowptr[r + 1];
        if (s > e || s < 0 || e > nnz) { bad |= 4; continue; }
        for (int64_t o = s + lane; o < e; o += 32) {
            int32_t c = col[o];
            if (c < 0 || c >= n) { bad |= 1; continue; }
            if (o > s && col[o - 1] >= c) bad |= 2;
            // symmetry: r must be in row c
            int64_t cs = rowptr[c], ce = rowptr[c + 1];
            if (cs <= ce && cs >= 0 && ce <= nnz) {
                if (!row_contains(col + cs, ce - cs, (int32_t)r)) bad |= 8;
            }
        }
    }
    if (warp == 0 && lane == 0 && (rowptr[0] != 0 || rowptr[n] != nnz)) bad |= 4;
    if (bad) atomicOr(flags, bad);
}

// ---- generic rows intersect ---------------------------------------------------------------
// One warp per target link.  The shorter row is walked 32 columns at a time, each lane binary
// searches its column in the longer row; matches are emitted in ascending column order with a
// ballot (both rows are ascending, so walking either one keeps the output ascending).
template <bool FILL>
__global__ void k_rows_intersect(const int64_t* __restrict__ rowptr1, const int32_t* __restrict__ col1, int64_t n1,
                                 const int64_t* __restrict__ rowptr2, const int32_t* __restrict__ col2, int64_t n2,
                                 const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                 int64_t num_edges, int64_t* __restrict__ out_counts,
                                 const int64_t* __restrict__ out_rowptr, int64_t* __restrict__ out_col) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int lane = lane_id();
    for (int64_t t = warp; t < num_edges; t += nwarps) {
        int64_t i = src[t], j = dst[t];
        if ((uint64_t)i >= (uint64_t)n1 || (uint64_t)j >= (uint64_t)n2) {  // reference: IndexError; here: counted, row left empty
            if (!FILL && lane == 0) { out_counts[t] = 0; atomicAdd(reinterpret_cast<unsigned long long*>(out_counts + num_edges), 1ull); }
            continue;
        }
        int64_t s1 = rowptr1[i], l1 = rowptr1[i + 1] - s1;
        int64_t s2 = rowptr2[j], l2 = rowptr2[j + 1] - s2;
        const int32_t* a = col1 + s1;
        const int32_t* b = col2 + s2;
        if (l1 > l2) {
            const int32_t* tp = a; a = b; b = tp;
            int64_t tl = l1; l1 = l2; l2 = tl;
        }
        int64_t count = 0;
        int64_t obase = FILL ? out_rowptr[t] : 0;
        for (int64_t base = 0; base < l1; base += 32) {
            int64_t o = base + lane;
            bool hit = false;
            int32_t c = 0;
            if (o < l1) {
                c = __ldg(a + o);
                hit = row_contains(b, l2, c);
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            if (FILL && hit) out_col[obase + count + __popc(m & ((1u << lane) - 1))] = c;
            count += __popc(m);
        }
        if (!FILL && lane == 0) out_counts[t] = count;
    }
}

// Set difference for the completion predictors' calresadj=True branch (utils.py:260-274 ->
// spmoverlap_notoverlap_ :210-244): out row b = adj1[src[b]] \ adj2[dst[b]], columns ascending.  One warp
// per link walks row 1 and binary-searches row 2 (the other difference is the same call with the
// matrices and the link ends swapped).
template <bool FILL>
__global__ void k_rows_difference(const int64_t* __restrict__ rowptr1, const int32_t* __restrict__ col1, int64_t n1,
                                  const int64_t* __restrict__ rowptr2, const int32_t* __restrict__ col2, int64_t n2,
                                  const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                  int64_t num_edges, int64_t* __restrict__ out_counts,
                                  const int64_t* __restrict__ out_rowptr, int64_t* __restrict__ out_col) {
    int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int lane = lane_id();
    for (int64_t t = warp; t < num_edges; t += nwarps) {
        const int64_t i = src[t], j = dst[t];
        if ((uint64_t)i >= (uint64_t)n1 || (uint64_t)j >= (uint64_t)n2) {
            if (!FILL && lane == 0) { out_counts[t] = 0; atomicAdd(reinterpret_cast<unsigned long long*>(out_counts + num_edges), 1ull); }
            continue;
        }
        const int64_t s1 = rowptr1[i], l1 = rowptr1[i + 1] - s1;
        const int64_t s2 = rowptr2[j], l2 = rowptr2[j + 1] - s2;
        const int32_t* a = col1 + s1;
        const int32_t* b = col2 + s2;
        int64_t count = 0;
        const int64_t obase = FILL ? out_rowptr[t] : 0;
        for (int64_t base = 0; base < l1; base += 32) {
            const int64_t o = base + lane;
            bool keep = false;
            int32_t c = 0;
            if (o < l1) {
                c = __ldg(a + o);
                keep = !row_contains(b, l2, c);
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (FILL && keep) out_col[obase + count + __popc(m & ((1u << lane) - 1))] = c;
            count += __popc(m);
        }
        if (!FILL && lane == 0) out_counts[t] = count;
    }
}

// ---- per-entry selection (DropAdj, model.py:219-229: torch_sparse.masked_select_nnz + value * ratio) -------------
// one warp per row: kept entries counted / compacted by ballot, order preserved
__global__ void k_select_count(const int64_t* __restrict__ rowptr, const uint8_t* __restrict__ keep, int64_t n,
                               int64_t* __restrict__ out_counts) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    if (warp >= n) return;
    const int64_t s = rowptr[warp], e = rowptr[warp + 1];
    int64_t c = 0;
    for (int64_t o = s; o < e; o += 32) {
        const bool k = (o + lane < e) && keep[o + lane] != 0;
        c += __popc(__ballot_sync(0xffffffffu, k));
    }
    if (lane == 0) out_counts[warp] = c;
}

__global__ void k_select_fill(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                              const float* __restrict__ val, const uint8_t* __restrict__ keep, int64_t n, float scale,
                              const int64_t* __restrict__ out_rowptr, int32_t* __restrict__ out_col,
                              float* __restrict__ out_val) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = lane_id();
    if (warp >= n) return;
    const int64_t s = rowptr[warp], e = rowptr[warp + 1];
    int64_t w = out_rowptr[warp];
    for (int64_t o = s; o < e; o += 32) {
        const bool k = (o + lane < e) && keep[o + lane] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (k) {
            const int64_t q = w + __popc(m & ((1u << lane) - 1u));
            out_col[q] = col[o + lane];
            if (out_val) out_val[q] = (val ? val[o + lane] : 1.0f) * scale;
        }
        w += __popc(m);
    }
}

}  // namespace ocn

using namespace ocn;

extern "C" {

int ocn_abi_version(void) { return OCN_ABI_VERSION; }
const char* ocn_last_error(void) { return last_error().c_str(); }
int ocn_device_sm_count(void) { return sm_count(); }
int ocn_set_option(int key, int64_t value) { return set_option(key, value); }
int64_t ocn_get_option(int key) { return option(key, 0); }
long long ocn_launch_count(void) { return launch_count(); }

int ocn_graph_select_count(const int64_t* rowptr, const uint8_t* keep, int64_t n, int64_t* out_counts, void* stream) {
    OCN_CHECK_ARG(rowptr && keep && out_counts, "ocn_graph_select_count: null pointer");
    OCN_CHECK_ARG(n > 0, "ocn_graph_select_count: n must be positive");
    k_select_count<<<(int)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rowptr, keep, n, out_counts);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_graph_select_fill(const int64_t* rowptr, const int32_t* col, const float* val, const uint8_t* keep, int64_t n,
                          float scale, const int64_t* out_rowptr, int32_t* out_col, float* out_val, void* stream) {
    OCN_CHECK_ARG(rowptr && col && keep && out_rowptr && out_col, "ocn_graph_select_fill: null pointer");
    OCN_CHECK_ARG(n > 0, "ocn_graph_select_fill: n must be positive");
    k_select_fill<<<(int)((n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rowptr, col, val, keep, n, scale, out_rowptr,
                                                                                  out_col, out_val);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_graph_validate(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t nnz, int32_t* out_flags,
                       void* stream) {
    OCN_CHECK_ARG(rowptr && out_flags && n >= 0 && nnz >= 0, "ocn_graph_validate: null pointer or negative size");
    OCN_CHECK_ARG(col || nnz == 0, "ocn_graph_validate: col is null");
    cudaStream_t st = (cudaStream_t)stream;
    OCN_CUDA(cudaMemsetAsync(out_flags, 0, sizeof(int32_t), st));
    if (n == 0) return OCN_OK;
    int blocks = sm_count() * 8;
    k_validate<<<blocks, 256, 0, st>>>(rowptr, col, n, nnz, out_flags);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_rows_intersect_count(const int64_t* rowptr1, const int32_t* col1, int64_t n1, const int64_t* rowptr2,
                             const int32_t* col2, int64_t n2, const int64_t* src, const int64_t* dst, int64_t num_edges,
                             int64_t* out_counts, void* stream) {
    OCN_CHECK_ARG(rowptr1 && rowptr2 && num_edges >= 0, "ocn_rows_intersect_count: bad arguments");
    if (num_edges == 0) return OCN_OK;
    OCN_CHECK_ARG(src && dst && out_counts, "ocn_rows_intersect_count: null edge/out pointer");
    int64_t want = (num_edges + 7) / 8;
    int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    k_rows_intersect<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(rowptr1, col1, n1, rowptr2, col2, n2, src, dst,
                                                                       num_edges, out_counts, nullptr, nullptr);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_rows_intersect_fill(const int64_t* rowptr1, const int32_t* col1, int64_t n1, const int64_t* rowptr2,
                            const int32_t* col2, int64_t n2, const int64_t* src, const int64_t* dst, int64_t num_edges,
                            const int64_t* out_rowptr, int64_t* out_col, void* stream) {
    OCN_RANGE("ocn_rows_intersect_fill");
    OCN_CHECK_ARG(rowptr1 && rowptr2 && num_edges >= 0, "ocn_rows_intersect_fill: bad arguments");
    if (num_edges == 0) return OCN_OK;
    OCN_CHECK_ARG(src && dst && out_rowptr, "ocn_rows_intersect_fill: null edge/out pointer");
    int64_t want = (num_edges + 7) / 8;
    int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    k_rows_intersect<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(rowptr1, col1, n1, rowptr2, col2, n2, src, dst,
                                                                      num_edges, nullptr, out_rowptr, out_col);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_rows_difference_count(const int64_t* rowptr1, const int32_t* col1, int64_t n1, const int64_t* rowptr2,
                              const int32_t* col2, int64_t n2, const int64_t* src, const int64_t* dst, int64_t num_edges,
                              int64_t* out_counts, void* stream) {
    OCN_CHECK_ARG(rowptr1 && rowptr2 && num_edges >= 0, "ocn_rows_difference_count: bad arguments");
    if (num_edges == 0) return OCN_OK;
    OCN_CHECK_ARG(src && dst && out_counts, "ocn_rows_difference_count: null edge/out pointer");
    int64_t want = (num_edges + 7) / 8;
    int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    k_rows_difference<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(rowptr1, col1, n1, rowptr2, col2, n2, src, dst,
                                                                        num_edges, out_counts, nullptr, nullptr);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_rows_difference_fill(const int64_t* rowptr1, const int32_t* col1, int64_t n1, const int64_t* rowptr2,
                             const int32_t* col2, int64_t n2, const int64_t* src, const int64_t* dst, int64_t num_edges,
                             const int64_t* out_rowptr, int64_t* out_col, void* stream) {
    OCN_RANGE("ocn_rows_difference_fill");
    OCN_CHECK_ARG(rowptr1 && rowptr2 && num_edges >= 0, "ocn_rows_difference_fill: bad arguments");
    if (num_edges == 0) return OCN_OK;
    OCN_CHECK_ARG(src && dst && out_rowptr, "ocn_rows_difference_fill: null edge/out pointer");
    int64_t want = (num_edges + 7) / 8;
    int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    k_rows_difference<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(rowptr1, col1, n1, rowptr2, col2, n2, src, dst,
                                                                       num_edges, nullptr, out_rowptr, out_col);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // extern "C"
