// A^2 SpGEMM: `spadj @ spadj` (NeighborOverlap_large.py:74,119) and the reference's
// `sparse_tensor_multiply` / `block_matrix_multiply` (utils.py:287-329).
//
// Row-wise Gustavson with a dense per-CTA accumulator in global scratch (L2 resident for the
// graphs the reference materialises A^2 for: N <= 2.4e5): a bitmap of touched columns + fp32-exact
// uint32 walk counts.  Columns come out ascending by scanning the touched word range of the
// bitmap, which is what torch_sparse's CSR needs.  fold = bs > 0 reproduces the reference's
// adj2byblock result: block (I,J) of A^2 is added at block-local coordinates, i.e.
//   folded[r, c] = sum over i = r mod bs, j = c mod bs of A^2[i, j]            (SURVEY Q6).
#include <cub/block/block_scan.cuh>

#include "common.cuh"

namespace ocn {

constexpr int kGemmThreads = 256;
constexpr int kGemmSlotsPerSm = 2;

struct GemmScratch {
    size_t words;      // bitmap words per slot
    size_t slot_bytes; // bitmap + counts
    int slots;
};

static GemmScratch gemm_scratch(int64_t n) {
    GemmScratch g;
    g.words = (size_t)((n + 31) / 32 + 1);
    size_t bytes = g.words * 4 + (size_t)n * 4;
    g.slot_bytes = (bytes + 255) & ~size_t(255);
    g.slots = sm_count() * kGemmSlotsPerSm;
    return g;
}

template <bool NUMERIC>
__global__ void __launch_bounds__(kGemmThreads)
k_spgemm_a2(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n, int64_t fold,
            unsigned char* __restrict__ scratch, size_t words, size_t slot_bytes, int64_t* __restrict__ out_row_nnz,
            const int64_t* __restrict__ out_rowptr, int32_t* __restrict__ out_col, float* __restrict__ out_val) {
    using Scan = cub::BlockScan<int, kGemmThreads>;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ unsigned s_wmin, s_wmax;
    __shared__ int s_running;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(scratch + (size_t)blockIdx.x * slot_bytes);
    uint32_t* counts = bitmap + words;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t out_rows = fold > 0 ? (fold < n ? fold : n) : n;
    const bool want_val = NUMERIC && out_val != nullptr;
    for (int64_t r = blockIdx.x; r < out_rows; r += gridDim.x) {
        if (tid == 0) { s_wmin = 0xffffffffu; s_wmax = 0u; s_running = 0; }
        __syncthreads();
        unsigned wmin = 0xffffffffu, wmax = 0u;
        const int64_t step = fold > 0 ? fold : n;
        for (int64_t i = r; i < n; i += step) {
            const int64_t s = rowptr[i], e = rowptr[i + 1];
            for (int64_t o = s + warp; o < e; o += kGemmThreads / 32) {
                const int32_t m = ldg_i32(col + o);
                const int64_t ms = rowptr[m], me = rowptr[m + 1];
                for (int64_t oo = ms + lane; oo < me; oo += 32) {
                    int64_t l = ldg_i32(col + oo);
                    if (fold > 0) l %= fold;
                    const unsigned w = (unsigned)(l >> 5);
                    atomicOr(&bitmap[w], 1u << (l & 31));
                    if (want_val) atomicAdd(&counts[l], 1u);
                    wmin = w < wmin ? w : wmin;
                    wmax = w > wmax ? w : wmax;
                }
            }
        }
        if (wmin != 0xffffffffu) { atomicMin(&s_wmin, wmin); atomicMax(&s_wmax, wmax); }
        __syncthreads();
        const unsigned lo = s_wmin, hi = s_wmax;
        int64_t total = 0;
        if (lo != 0xffffffffu) {
            const int64_t obase = NUMERIC ? out_rowptr[r] : 0;
            for (unsigned w0 = lo; w0 <= hi; w0 += kGemmThreads) {
                const unsigned w = w0 + tid;
                uint32_t bits = (w <= hi) ? __ldcg(&bitmap[w]) : 0u;
                int pc = __popc(bits), off = 0, tile_total = 0;
                Scan(scan_tmp).ExclusiveSum(pc, off, tile_total);
                const int running = s_running;
                __syncthreads();
                if (NUMERIC && bits) {
                    int64_t o = obase + running + off;
                    uint32_t b = bits;
                    while (b) {
                        const int bit = __ffs(b) - 1;
                        b &= b - 1;
                        const int64_t l = (int64_t)w * 32 + bit;
                        out_col[o] = (int32_t)l;
                        if (want_val) { out_val[o] = (float)__ldcg(&counts[l]); __stcg(&counts[l], 0u); }
                        ++o;
                    }
                }
                if (bits) __stcg(&bitmap[w], 0u);
                if (tid == 0) s_running = running + tile_total;
                __syncthreads();
            }
            total = s_running;
        }
        if (!NUMERIC && tid == 0) out_row_nnz[r] = total;
        __syncthreads();
    }
    if (!NUMERIC && fold > 0) {
        for (int64_t r = out_rows + (int64_t)blockIdx.x * kGemmThreads + tid; r < n; r += (int64_t)gridDim.x * kGemmThreads)
            out_row_nnz[r] = 0;
    }
}

}  // namespace ocn

using namespace ocn;

extern "C" {

size_t ocn_spgemm_scratch_bytes(int64_t n) {
    if (n <= 0) return 0;
    GemmScratch g = gemm_scratch(n);
    return g.slot_bytes * (size_t)g.slots;
}

int ocn_spgemm_a2_symbolic(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t fold, void* scratch,
                           int64_t* out_row_nnz, void* stream) {
    OCN_CHECK_ARG(rowptr && col && scratch && out_row_nnz, "ocn_spgemm_a2_symbolic: null pointer");
    OCN_CHECK_ARG(n > 0 && fold >= 0, "ocn_spgemm_a2_symbolic: bad sizes");
    GemmScratch g = gemm_scratch(n);
    k_spgemm_a2<false><<<g.slots, kGemmThreads, 0, (cudaStream_t)stream>>>(
        rowptr, col, n, fold, (unsigned char*)scratch, g.words, g.slot_bytes, out_row_nnz, nullptr, nullptr, nullptr);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_spgemm_a2_numeric(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t fold, void* scratch,
                          const int64_t* out_rowptr, int32_t* out_col, float* out_val, void* stream) {
    OCN_CHECK_ARG(rowptr && col && scratch && out_rowptr && out_col, "ocn_spgemm_a2_numeric: null pointer");
    OCN_CHECK_ARG(n > 0 && fold >= 0, "ocn_spgemm_a2_numeric: bad sizes");
    GemmScratch g = gemm_scratch(n);
    k_spgemm_a2<true><<<g.slots, kGemmThreads, 0, (cudaStream_t)stream>>>(
        rowptr, col, n, fold, (unsigned char*)scratch, g.words, g.slot_bytes, nullptr, out_rowptr, out_col, out_val);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // extern "C"
