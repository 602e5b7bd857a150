// Container-level sparse-row kernels behind the import shims (ocn_b200/shim): the torch_sparse / pygho methods the
// reference's model.py and drivers call on [B x N] and [N x N] matrices, on the library's CSR layout
// (rowptr int64, col int32, optional fp32 values):
//   adj[idx], SparseTensor.index_select             -> ocn_rows_gather_*      (utils.py:256-257, NeighborOverlapCitation2.py:79-81)
//   pygho spsphadamard(A, B) on two explicit matrices -> ocn_rows_hadamard_*  (model.py:2243, innerprod1)
//   SparseTensor.sum(dim=0)                          -> ocn_csr_colsum        (model.py:2261)
#include "common.cuh"

namespace ocn {

// one warp per output row: copy row idx[b] of the source (columns, and values when both pointers are given)
template <bool FILL>
__global__ void k_rows_gather(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                              const float* __restrict__ val, int64_t n, const int64_t* __restrict__ idx, int64_t num_rows,
                              int64_t* __restrict__ out_counts, const int64_t* __restrict__ out_rowptr,
                              int32_t* __restrict__ out_col, float* __restrict__ out_val) {
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (int64_t b = warp; b < num_rows; b += nwarps) {
        const int64_t r = idx[b];
        if ((uint64_t)r >= (uint64_t)n) {  // torch raises IndexError: counted in the spare word, row left empty
            if (!FILL && lane == 0) {
                out_counts[b] = 0;
                atomicAdd(reinterpret_cast<unsigned long long*>(out_counts + num_rows), 1ull);
            }
            continue;
        }
        const int64_t s = rowptr[r], len = rowptr[r + 1] - s;
        if (!FILL) {
            if (lane == 0) out_counts[b] = len;
            continue;
        }
        const int64_t o = out_rowptr[b];
        for (int64_t k = lane; k < len; k += 32) {
            out_col[o + k] = __ldg(col + s + k);
            if (out_val) out_val[o + k] = val ? __ldg(val + s + k) : 1.0f;
        }
    }
}

// one warp per row b: entries of A's row b that also sit in B's row b, value va * vb.  A's row is walked (ascending
// output for free), B's row is binary-searched for the POSITION of the column (the values are paired by it).
__device__ __forceinline__ int64_t row_find(const int32_t* __restrict__ row, int64_t len, int32_t key) {
    int64_t lo = 0, hi = len;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(row + mid) < key) lo = mid + 1; else hi = mid;
    }
    return (lo < len && __ldg(row + lo) == key) ? lo : -1;
}

template <bool FILL>
__global__ void k_rows_hadamard(const int64_t* __restrict__ rowptr_a, const int32_t* __restrict__ col_a,
                                const float* __restrict__ val_a, const int64_t* __restrict__ rowptr_b,
                                const int32_t* __restrict__ col_b, const float* __restrict__ val_b, int64_t num_rows,
                                int64_t* __restrict__ out_counts, const int64_t* __restrict__ out_rowptr,
                                int32_t* __restrict__ out_col, float* __restrict__ out_val) {
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (int64_t b = warp; b < num_rows; b += nwarps) {
        const int64_t sa = rowptr_a[b], la = rowptr_a[b + 1] - sa;
        const int64_t sb = rowptr_b[b], lb = rowptr_b[b + 1] - sb;
        int64_t count = 0;
        const int64_t obase = FILL ? out_rowptr[b] : 0;
        for (int64_t base = 0; base < la; base += 32) {
            const int64_t o = base + lane;
            int64_t pos = -1;
            int32_t c = 0;
            if (o < la) {
                c = __ldg(col_a + sa + o);
                pos = row_find(col_b + sb, lb, c);
            }
            const unsigned m = __ballot_sync(0xffffffffu, pos >= 0);
            if (FILL && pos >= 0) {
                const int64_t q = obase + count + __popc(m & ((1u << lane) - 1u));
                out_col[q] = c;
                out_val[q] = (val_a ? __ldg(val_a + sa + o) : 1.0f) * (val_b ? __ldg(val_b + sb + pos) : 1.0f);
            }
            count += __popc(m);
        }
        if (!FILL && lane == 0) out_counts[b] = count;
    }
}

// flat over the entries: out[col[e]] += val[e] (or 1).  Counts of ones are exact in fp32 below 2^24; weighted sums are
// order-dependent, as the reference's scatter_add is.
__global__ void k_csr_colsum(const int32_t* __restrict__ col, const float* __restrict__ val, int64_t nnz, int64_t n_cols,
                             float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
        const int32_t c = __ldg(col + e);
        if ((uint32_t)c < (uint64_t)n_cols) atomicAdd(out + c, val ? __ldg(val + e) : 1.0f);
    }
}

static int warp_grid(int64_t rows) {
    const int64_t want = (rows + 7) / 8, cap = (int64_t)sm_count() * 16;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace ocn

using namespace ocn;

extern "C" {

int ocn_rows_gather_count(const int64_t* rowptr, int64_t n, const int64_t* idx, int64_t num_rows, int64_t* out_counts,
                          void* stream) {
    OCN_CHECK_ARG(rowptr && n >= 0 && num_rows >= 0, "ocn_rows_gather_count: bad arguments");
    if (num_rows == 0) return OCN_OK;
    OCN_CHECK_ARG(idx && out_counts, "ocn_rows_gather_count: null index/out pointer");
    k_rows_gather<false><<<warp_grid(num_rows), 256, 0, (cudaStream_t)stream>>>(rowptr, nullptr, nullptr, n, idx, num_rows,
                                                                                out_counts, nullptr, nullptr, nullptr);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_rows_gather_fill(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n, const int64_t* idx,
                         int64_t num_rows, const int64_t* out_rowptr, int32_t* out_col, float* out_val, void* stream) {
    OCN_CHECK_ARG(rowptr && n >= 0 && num_rows >= 0, "ocn_rows_gather_fill: bad arguments");
    if (num_rows == 0) return OCN_OK;
    OCN_CHECK_ARG(col && idx && out_rowptr && out_col, "ocn_rows_gather_fill: null pointer");
    k_rows_gather<true><<<warp_grid(num_rows), 256, 0, (cudaStream_t)stream>>>(rowptr, col, val, n, idx, num_rows, nullptr,
                                                                               out_rowptr, out_col, out_val);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_rows_hadamard_count(const int64_t* rowptr_a, const int32_t* col_a, const int64_t* rowptr_b, const int32_t* col_b,
                            int64_t num_rows, int64_t* out_counts, void* stream) {
    OCN_CHECK_ARG(rowptr_a && rowptr_b && num_rows >= 0, "ocn_rows_hadamard_count: bad arguments");
    if (num_rows == 0) return OCN_OK;
    OCN_CHECK_ARG(out_counts, "ocn_rows_hadamard_count: null out pointer");
    k_rows_hadamard<false><<<warp_grid(num_rows), 256, 0, (cudaStream_t)stream>>>(
        rowptr_a, col_a, nullptr, rowptr_b, col_b, nullptr, num_rows, out_counts, nullptr, nullptr, nullptr);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_rows_hadamard_fill(const int64_t* rowptr_a, const int32_t* col_a, const float* val_a, const int64_t* rowptr_b,
                           const int32_t* col_b, const float* val_b, int64_t num_rows, const int64_t* out_rowptr,
                           int32_t* out_col, float* out_val, void* stream) {
    OCN_CHECK_ARG(rowptr_a && rowptr_b && num_rows >= 0, "ocn_rows_hadamard_fill: bad arguments");
    if (num_rows == 0) return OCN_OK;
    OCN_CHECK_ARG(out_rowptr && out_col && out_val, "ocn_rows_hadamard_fill: null out pointer");
    k_rows_hadamard<true><<<warp_grid(num_rows), 256, 0, (cudaStream_t)stream>>>(
        rowptr_a, col_a, val_a, rowptr_b, col_b, val_b, num_rows, nullptr, out_rowptr, out_col, out_val);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_csr_colsum(const int32_t* col, const float* val, int64_t nnz, int64_t n_cols, float* out, void* stream) {
    OCN_CHECK_ARG(nnz >= 0 && n_cols >= 0, "ocn_csr_colsum: bad sizes");
    if (nnz == 0) return OCN_OK;
    OCN_CHECK_ARG(col && out, "ocn_csr_colsum: null pointer");
    const int64_t want = (nnz + 255) / 256, cap = (int64_t)sm_count() * 16;
    k_csr_colsum<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(col, val, nnz, n_cols, out);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // extern "C"
