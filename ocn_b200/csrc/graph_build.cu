// The step BEFORE the hot path: building the CSR adjacency on the device.
//
//   SparseTensor.from_edge_index(tei, sparse_sizes=(N, N)).to_symmetric()
//       ogbdataset.py:44-45, NeighborOverlap_large.py:56-63, NeighborOverlapCitation2.py:135-143
//
// Under --maskinput the reference re-sorts and re-symmetrises the WHOLE edge list every training batch,
// only to drop the <= batch_size target links of that batch.  Two entry-point families replace it:
//
// * ocn_graph_build_*: edge list (+ optional keep mask) -> CSR.  64-bit keys row * n + col of both
//   directions, ONE radix sort over the significant bits only (CUB onesweep), run-length encode -> unique
//   entries and their multiplicity (how many list edges map onto an entry), row pointers by binary search
//   of row * n in the unique keys.  Done once per graph.
// * ocn_graph_mask_*: the per-batch masked adjacency WITHOUT a sort.  Every masked link decrements the
//   multiplicity of its (two) entries -- a binary search in one row each -- an entry survives while some
//   unmasked list edge still maps onto it (exactly what rebuilding from the remaining list yields, also
//   when the list holds duplicates or both directions of a link); the few dead positions are sorted and
//   the survivors move in one flat coalesced copy.  Traffic: one streaming pass over col/mult/dec (read)
//   and the new col/mult (write) instead of ~6 sort passes over 16-byte pairs.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace ocn {

namespace {

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct BuildLayout {
    size_t keys_a, keys_b, uniq, cnt, nruns, cub_temp, cub_bytes, total;
};

BuildLayout build_layout(int64_t E, int symmetric) {
    const int64_t K = (E > 0 ? E : 1) * (symmetric ? 2 : 1);
    BuildLayout L;
    size_t off = 0;
    L.keys_a = off; off += align256(sizeof(unsigned long long) * K);
    L.keys_b = off; off += align256(sizeof(unsigned long long) * K);
    L.uniq = off;   off += align256(sizeof(unsigned long long) * K);
    L.cnt = off;    off += align256(sizeof(int32_t) * K);
    L.nruns = off;  off += 256;
    size_t b1 = 0, b2 = 0;
    cub::DoubleBuffer<unsigned long long> db(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, b1, db, (int)K, 0, 64);
    cub::DeviceRunLengthEncode::Encode(nullptr, b2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                       (int32_t*)nullptr, (int64_t*)nullptr, (int)K);
    L.cub_bytes = (b1 > b2 ? b1 : b2) + 256;
    L.cub_temp = off; off += align256(L.cub_bytes);
    L.total = off;
    return L;
}

int key_bits(int64_t n) {  // bits of the sentinel n * n (the largest key that is ever sorted)
    unsigned long long s = (unsigned long long)n * (unsigned long long)n;
    int b = 1;
    while (b < 64 && (s >> b) != 0ull) ++b;
    return b;
}

}  // namespace

__global__ void k_graph_emit_keys(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                  const uint8_t* __restrict__ keep, int64_t E, int64_t n, int symmetric,
                                  unsigned long long* __restrict__ keys, int64_t* __restrict__ info) {
    const unsigned long long sentinel = (unsigned long long)n * (unsigned long long)n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < E; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = src[t], v = dst[t];
        const bool kept = keep == nullptr || keep[t] != 0;
        const bool in_range = u >= 0 && u < n && v >= 0 && v < n;
        if (kept && !in_range) atomicAdd(reinterpret_cast<unsigned long long*>(info + 1), 1ull);
        const bool ok = kept && in_range;
        if (symmetric) {
            keys[2 * t] = ok ? (unsigned long long)u * n + v : sentinel;
            keys[2 * t + 1] = ok ? (unsigned long long)v * n + u : sentinel;
        } else {
            keys[t] = ok ? (unsigned long long)u * n + v : sentinel;
        }
    }
}

// rowptr[r] = first unique key >= r * n; the sentinel run (dropped edges) sorts last, so rowptr[n] is the
// number of entries
__global__ void k_graph_rowptr(const unsigned long long* __restrict__ uniq, const int64_t* __restrict__ nruns_p,
                               int64_t n, int64_t* __restrict__ rowptr, int64_t* __restrict__ info) {
    const int64_t nruns = *nruns_p;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = (unsigned long long)r * (unsigned long long)n;
        int64_t lo = 0, hi = nruns;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (uniq[mid] < key) lo = mid + 1; else hi = mid;
        }
        rowptr[r] = lo;
        if (r == n) info[0] = lo;
    }
}

__global__ void k_graph_fill(const unsigned long long* __restrict__ uniq, const int32_t* __restrict__ cnt, int64_t n,
                             int64_t nnz, int32_t* __restrict__ col, int32_t* __restrict__ mult) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        col[i] = (int32_t)(uniq[i] % (unsigned long long)n);
        if (mult) mult[i] = cnt[i];
    }
}

// ---- per-batch masking ---------------------------------------------------------------------------
// The masked adjacency is the full one minus a FEW dead entries (at most two per masked link), so it is built by
// a flat, fully coalesced copy: the dead entry positions are collected while the multiplicities are decremented,
// sorted (a few thousand keys), and every surviving entry p moves to p - #{dead positions < p}; the new row
// pointers are the old ones minus the same count.  No per-row work, no scan over the entries.
struct MaskLayout {
    size_t rem_a, rem_b, counter, cub_temp, cub_bytes, total;
    int64_t cap;
};

static MaskLayout mask_layout(int64_t num_masked) {
    MaskLayout L;
    L.cap = 2 * (num_masked > 0 ? num_masked : 0) + 1;
    size_t off = 0;
    L.rem_a = off;   off += align256(sizeof(int64_t) * (size_t)L.cap);
    L.rem_b = off;   off += align256(sizeof(int64_t) * (size_t)L.cap);
    L.counter = off; off += 256;
    size_t b = 0;
    cub::DoubleBuffer<int64_t> db(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, b, db, (int)L.cap, 0, 64);
    L.cub_bytes = b + 256;
    L.cub_temp = off; off += align256(L.cub_bytes);
    L.total = off;
    return L;
}

__device__ __forceinline__ int64_t entry_of(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                            int64_t u, int64_t v) {
    int64_t lo = rowptr[u], hi = rowptr[u + 1];
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(col + mid) < (int32_t)v) lo = mid + 1; else hi = mid;
    }
    return (lo < end && __ldg(col + lo) == (int32_t)v) ? lo : -1;
}

__global__ void k_mask_init(int64_t* __restrict__ rem, int64_t cap, int64_t nnz, unsigned long long* __restrict__ counter) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (int64_t)gridDim.x * blockDim.x) rem[i] = nnz;
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0ull;
}

// kReset == false: dec[p]++ for the (two) entries of every masked link; the decrement that uses up an entry's
// multiplicity records its position as dead.  kReset == true: dec back to zero.
template <bool kReset>
__global__ void k_mask_mark(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                            const int32_t* __restrict__ mult, int64_t n, const int64_t* __restrict__ src,
                            const int64_t* __restrict__ dst, int64_t M, int symmetric, int32_t* __restrict__ dec,
                            int64_t* __restrict__ rem, unsigned long long* __restrict__ counter, int64_t* __restrict__ info) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < M; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = src[t], v = dst[t];
        if (u < 0 || u >= n || v < 0 || v >= n) {
            if (!kReset) atomicAdd(reinterpret_cast<unsigned long long*>(info + 1), 1ull);
            continue;
        }
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (side == 1 && !symmetric) break;
            const int64_t p = side == 0 ? entry_of(rowptr, col, u, v) : entry_of(rowptr, col, v, u);
            if (p < 0) {
                if (!kReset && side == 0) atomicAdd(reinterpret_cast<unsigned long long*>(info + 1), 1ull);
                continue;
            }
            if (kReset) {
                dec[p] = 0;
            } else {
                const int32_t old = atomicAdd(dec + p, 1);
                if (old + 1 == (mult ? __ldg(mult + p) : 1)) rem[atomicAdd(counter, 1ull)] = p;
            }
        }
    }
}

// number of dead positions < key (the list is ascending and padded with the sentinel nnz)
__device__ __forceinline__ int64_t dead_before(const int64_t* __restrict__ rem, int64_t cap, int64_t key) {
    int64_t lo = 0, hi = cap;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(rem + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void k_mask_rowptr(const int64_t* __restrict__ rowptr, int64_t n, const int64_t* __restrict__ rem, int64_t cap,
                              int64_t* __restrict__ out_rowptr, int64_t* __restrict__ info) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = rowptr[r];
        const int64_t v = s - dead_before(rem, cap, s);
        out_rowptr[r] = v;
        if (r == n) info[0] = v;
    }
}

constexpr int kMaskChunk = 4096;  // entries per CTA of the flat copy

__global__ void __launch_bounds__(256)
k_mask_copy(const int32_t* __restrict__ col, const int32_t* __restrict__ mult, const int32_t* __restrict__ dec, int64_t nnz,
            const int64_t* __restrict__ rem, int64_t cap, int32_t* __restrict__ out_col, int32_t* __restrict__ out_mult) {
    __shared__ int64_t s_base, s_cnt;
    __shared__ int64_t s_dead[64];
    for (int64_t p0 = (int64_t)blockIdx.x * kMaskChunk; p0 < nnz; p0 += (int64_t)gridDim.x * kMaskChunk) {
        const int64_t p1 = (p0 + kMaskChunk < nnz) ? p0 + kMaskChunk : nnz;
        __syncthreads();
        if (threadIdx.x == 0) {
            s_base = dead_before(rem, cap, p0);
            s_cnt = dead_before(rem, cap, p1) - s_base;
        }
        __syncthreads();
        const int64_t base = s_base, cnt = s_cnt;
        if (cnt > 0 && cnt <= 64 && (int64_t)threadIdx.x < cnt) s_dead[threadIdx.x] = rem[base + threadIdx.x];
        __syncthreads();
        for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
            int64_t shift = base;
            bool dead = false;
            if (cnt > 64) {  // many dead entries in one chunk (a masked hub): search the list
                shift = dead_before(rem, cap, p);
                dead = shift < cap && __ldg(rem + shift) == p;
            } else {
                for (int k = 0; k < (int)cnt; ++k) {
                    const int64_t q = s_dead[k];
                    shift += q < p;
                    dead |= q == p;
                }
            }
            if (!dead) {
                out_col[p - shift] = ldg_i32(col + p);
                if (out_mult) out_mult[p - shift] = (mult ? __ldg(mult + p) : 1) - __ldg(dec + p);
            }
        }
    }
}

static int nnz_bits(int64_t nnz) {  // bits of the keys 0 .. nnz (nnz itself is the padding sentinel)
    int b = 1;
    while (b < 63 && (int64_t(1) << b) <= nnz) ++b;
    return b;
}

static int grid_for(int64_t items, int per_block) {
    int64_t want = (items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace ocn

using namespace ocn;

extern "C" {

size_t ocn_graph_build_bytes(int64_t num_edges, int symmetric) { return build_layout(num_edges, symmetric).total; }

int ocn_graph_build_count(const int64_t* src, const int64_t* dst, const uint8_t* keep, int64_t num_edges, int64_t n,
                          int symmetric, void* scratch, size_t scratch_bytes, int64_t* out_rowptr, int64_t* out_info,
                          void* stream) {
    OCN_CHECK_ARG(num_edges >= 0 && n > 0 && n < (int64_t(1) << 31), "ocn_graph_build_count: bad sizes");
    OCN_CHECK_ARG((num_edges == 0 || (src && dst)) && scratch && out_rowptr && out_info, "ocn_graph_build_count: null pointer");
    const int64_t K = num_edges * (symmetric ? 2 : 1);
    OCN_CHECK_ARG(K < (int64_t(1) << 31), "ocn_graph_build_count: more than 2^31 directed entries");
    const BuildLayout L = build_layout(num_edges, symmetric);
    OCN_CHECK_ARG(scratch_bytes >= L.total, "ocn_graph_build_count: scratch too small (%zu < %zu)", scratch_bytes, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)scratch;
    unsigned long long* ka = (unsigned long long*)(base + L.keys_a);
    unsigned long long* kb = (unsigned long long*)(base + L.keys_b);
    unsigned long long* uniq = (unsigned long long*)(base + L.uniq);
    int32_t* cnt = (int32_t*)(base + L.cnt);
    int64_t* nruns = (int64_t*)(base + L.nruns);
    OCN_CUDA(cudaMemsetAsync(out_info, 0, 2 * sizeof(int64_t), st));
    OCN_CUDA(cudaMemsetAsync(nruns, 0, sizeof(int64_t), st));
    if (K > 0) {
        k_graph_emit_keys<<<grid_for(num_edges, 256), 256, 0, st>>>(src, dst, keep, num_edges, n, symmetric, ka, out_info);
        OCN_LAUNCH_CHECK();
        cub::DoubleBuffer<unsigned long long> db(ka, kb);
        size_t tb = L.cub_bytes;
        OCN_CUDA(cub::DeviceRadixSort::SortKeys(base + L.cub_temp, tb, db, (int)K, 0, key_bits(n), st));
        tb = L.cub_bytes;
        OCN_CUDA(cub::DeviceRunLengthEncode::Encode(base + L.cub_temp, tb, (const unsigned long long*)db.Current(), uniq,
                                                    cnt, nruns, (int)K, st));
    }
    k_graph_rowptr<<<grid_for(n + 1, 256), 256, 0, st>>>(uniq, nruns, n, out_rowptr, out_info);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_graph_build_fill(const void* scratch, int64_t num_edges, int symmetric, int64_t n, int64_t nnz, int32_t* out_col,
                         int32_t* out_mult, void* stream) {
    OCN_RANGE("ocn_graph_build_fill");
    OCN_CHECK_ARG(scratch && nnz >= 0 && n > 0, "ocn_graph_build_fill: bad arguments");
    if (nnz == 0) return OCN_OK;
    OCN_CHECK_ARG(out_col, "ocn_graph_build_fill: null output");
    const BuildLayout L = build_layout(num_edges, symmetric);
    const char* base = (const char*)scratch;
    k_graph_fill<<<grid_for(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long*)(base + L.uniq), (const int32_t*)(base + L.cnt), n, nnz, out_col, out_mult);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

size_t ocn_graph_mask_bytes(int64_t n, int64_t num_masked) {
    (void)n;
    return mask_layout(num_masked).total;
}

int ocn_graph_mask_count(const int64_t* rowptr, const int32_t* col, const int32_t* mult, int64_t n, int64_t nnz,
                         const int64_t* src, const int64_t* dst, int64_t num_masked, int symmetric, int32_t* dec,
                         void* scratch, size_t scratch_bytes, int64_t* out_rowptr, int64_t* out_info, void* stream) {
    OCN_CHECK_ARG(rowptr && dec && scratch && out_rowptr && out_info && (col || nnz == 0), "ocn_graph_mask_count: null pointer");
    OCN_CHECK_ARG(n > 0 && nnz >= 0 && num_masked >= 0 && (num_masked == 0 || (src && dst)), "ocn_graph_mask_count: bad arguments");
    OCN_CHECK_ARG(num_masked < (int64_t(1) << 30), "ocn_graph_mask_count: too many masked links");
    const MaskLayout L = mask_layout(num_masked);
    OCN_CHECK_ARG(scratch_bytes >= L.total, "ocn_graph_mask_count: scratch too small (%zu < %zu)", scratch_bytes, L.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)scratch;
    int64_t* rem_a = (int64_t*)(base + L.rem_a);
    int64_t* rem_b = (int64_t*)(base + L.rem_b);
    unsigned long long* counter = (unsigned long long*)(base + L.counter);
    OCN_CUDA(cudaMemsetAsync(out_info, 0, 2 * sizeof(int64_t), st));
    k_mask_init<<<grid_for(L.cap, 256), 256, 0, st>>>(rem_a, L.cap, nnz, counter);
    OCN_LAUNCH_CHECK();
    const int64_t* rem = rem_a;
    if (num_masked > 0) {
        k_mask_mark<false><<<grid_for(num_masked, 256), 256, 0, st>>>(rowptr, col, mult, n, src, dst, num_masked, symmetric, dec,
                                                                     rem_a, counter, out_info);
        OCN_LAUNCH_CHECK();
        cub::DoubleBuffer<int64_t> db(rem_a, rem_b);
        size_t tb = L.cub_bytes;
        OCN_CUDA(cub::DeviceRadixSort::SortKeys(base + L.cub_temp, tb, db, (int)L.cap, 0, nnz_bits(nnz), st));
        if (db.Current() != rem_a)  // fill() reads the sorted list from the first buffer
            OCN_CUDA(cudaMemcpyAsync(rem_a, db.Current(), sizeof(int64_t) * (size_t)L.cap, cudaMemcpyDeviceToDevice, st));
    }
    k_mask_rowptr<<<grid_for(n + 1, 256), 256, 0, st>>>(rowptr, n, rem, L.cap, out_rowptr, out_info);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_graph_mask_fill(const int64_t* rowptr, const int32_t* col, const int32_t* mult, int64_t n, int64_t nnz,
                        const int64_t* src, const int64_t* dst, int64_t num_masked, int symmetric, int32_t* dec,
                        const void* scratch, int32_t* out_col, int32_t* out_mult, void* stream) {
    OCN_RANGE("ocn_graph_mask_fill");
    OCN_CHECK_ARG(rowptr && dec && scratch && (col || nnz == 0), "ocn_graph_mask_fill: null pointer");
    OCN_CHECK_ARG(n > 0 && nnz >= 0 && num_masked >= 0, "ocn_graph_mask_fill: bad arguments");
    const MaskLayout L = mask_layout(num_masked);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t* rem = (const int64_t*)((const char*)scratch + L.rem_a);
    if (out_col && nnz > 0) {
        k_mask_copy<<<grid_for(nnz, kMaskChunk), 256, 0, st>>>(col, mult, dec, nnz, rem, L.cap, out_col, out_mult);
        OCN_LAUNCH_CHECK();
    }
    if (num_masked > 0) {  // hand the decrement array back all zero
        k_mask_mark<true><<<grid_for(num_masked, 256), 256, 0, st>>>(rowptr, col, mult, n, src, dst, num_masked, symmetric, dec,
                                                                    nullptr, nullptr, nullptr);
        OCN_LAUNCH_CHECK();
    }
    return OCN_OK;
}

}  // extern "C"
