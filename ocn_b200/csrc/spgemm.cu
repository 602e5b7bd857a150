// A^2 SpGEMM: `spadj @ spadj` (NeighborOverlap_large.py:74,119) and the reference's
// `sparse_tensor_multiply` / `block_matrix_multiply` (utils.py:287-329).
//
// Three row-wise Gustavson variants, picked from (n, nnz, fold) by gemm_mode():
//   SMEM   (k_spgemm_row_smem; the graphs the reference materialises A^2 for, n <= ~7e5): the accumulator of a row
//          lives in shared memory -- a two-level bitmap of the touched columns, the rank of every touched bitmap word,
//          and the walk counts / columns of the row staged by rank -- so a product is one shared-memory atomic, the
//          columns come out ascending from the touched-word list (no scan over the column range) and leave in
//          coalesced stores.  Round 1 accumulated in per-CTA global scratch with two global atomics per product and
//          scanned the words wmin..wmax of a 7 372-word bitmap per row: collab 48.7 ms (21 GB/s).
//   DENSE  (k_dense_*; ddi: n = 4 267, density 0.147, A^2 full): rows as bit vectors; the structure of a row is the OR
//          of its neighbours' vectors, a walk count is popc(row_i & row_l) -- no atomics at all.  The reference does
//          this with 25 dense SGEMMs (utils.py:287-323).
//   GLOBAL (k_spgemm_a2, round 1's kernel): a dense per-CTA accumulator in global scratch, for column spaces that do
//          not fit shared memory.
// Columns come out ascending, which is what torch_sparse's CSR needs.
// Tried for the 97 % of collab's rows with <= 2048 products and removed: one warp per row, the products gathered into a
// warp-private buffer, bitonic-sorted and run-length encoded (no barriers, 16 rows per SM in flight) -- 10.8 ms against
// 7.6 ms for the CTA-per-row bitmap below (collab), 1.06 against 0.29 ms (pubmed): ~5 K warp instructions per row for
// the sorting network, twice (symbolic + numeric).  fold = bs > 0 reproduces the reference's
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


// ---- SMEM: per-row accumulator in shared memory ---------------------------------------------------------------
constexpr int kRowThreads = 256;
constexpr int kRowStage = 1024;     // entries of a row staged in shared memory (columns + counts); longer rows go to global

struct RowSmem {                     // dynamic shared memory layout (words)
    size_t w0, w1;                   // level-0 / level-1 bitmap words
    size_t off_l1, off_pref, off_cnt, off_colbuf, total_bytes;
};
static RowSmem row_smem(int64_t ncols, bool numeric) {
    RowSmem r;
    r.w0 = (size_t)((ncols + 31) / 32);
    r.w1 = (r.w0 + 31) / 32;
    size_t off = r.w0;
    r.off_l1 = off;   off += r.w1;
    r.off_pref = off; off += numeric ? r.w1 + (r.w0 + 1) / 2 : 0;   // ranks per level-1 word (32 bit) and per word (16 bit)
    r.off_cnt = off;  off += numeric ? kRowStage : 0;
    r.off_colbuf = off; off += numeric ? 2 * kRowStage : 0;   // staged columns by rank + the row's products
    r.total_bytes = off * 4;
    return r;
}

template <bool NUMERIC>
__global__ void __launch_bounds__(kRowThreads)
k_spgemm_row_smem(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n, int64_t fold,
                  int64_t w0, int64_t w1, int64_t* __restrict__ out_row_nnz, const int64_t* __restrict__ out_rowptr,
                  int32_t* __restrict__ out_col, float* __restrict__ out_val, unsigned long long* __restrict__ row_counter) {
    extern __shared__ uint32_t rs_smem[];
    using Scan = cub::BlockScan<int, kRowThreads>;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ int s_total;
    __shared__ long long s_row;
    uint32_t* L0 = rs_smem;
    uint32_t* L1 = rs_smem + w0;
    uint32_t* pref = L1 + w1;                         // NUMERIC: rank of the first column under every level-1 word ...
    uint16_t* pref16 = reinterpret_cast<uint16_t*>(pref + (NUMERIC ? w1 : 0));   // ... and of every touched word (rows of < 65536 columns)
    uint32_t* cnt = pref + (NUMERIC ? w1 + (w0 + 1) / 2 : 0);   // NUMERIC: walk counts by rank
    uint32_t* colbuf = cnt + (NUMERIC ? kRowStage : 0);
    uint32_t* prodbuf = colbuf + (NUMERIC ? kRowStage : 0);   // NUMERIC: the products of the row (second pass without global loads)
    __shared__ int s_np;
    constexpr int kHubNeighbour = 256, kHubSlots = 64;
    __shared__ long long s_hms[kHubSlots], s_hme[kHubSlots];
    __shared__ int s_nh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_nh = 0;
    const int64_t out_rows = fold > 0 ? (fold < n ? fold : n) : n;
    const int64_t step = fold > 0 ? fold : n;
    const bool want_val = NUMERIC && out_val != nullptr;
    for (int64_t w = tid; w < w0 + w1; w += kRowThreads) rs_smem[w] = 0u;  // both bitmaps: zero between rows
    __syncthreads();
    // every product (i, m, l) of output row r, lanes over the columns of N(m), warps over the neighbours m
    auto for_each_product = [&](int64_t r, auto&& body) {
        for (int64_t i = r; i < n; i += step) {
            const int64_t s = rowptr[i], e = rowptr[i + 1];
            // the warp's neighbours (every 8th of the row), 32 at a time: their ranges are fetched by the lanes side by side,
            // so the walk below waits for one load per neighbour instead of a chain of three
            constexpr int kW = kRowThreads / 32;
            for (int64_t o0 = s + warp; o0 < e; o0 += 32 * kW) {
                int64_t ms_l = 0, me_l = 0;
                if (o0 + (int64_t)lane * kW < e) {
                    const int32_t m = ldg_i32(col + o0 + (int64_t)lane * kW);
                    ms_l = ldg_i64(rowptr + m);
                    me_l = ldg_i64(rowptr + m + 1);
                }
                const int64_t left = (e - o0 + kW - 1) / kW;
                const int cnt = (int)(left < 32 ? left : 32);
                for (int q = 0; q < cnt; ++q) {
                    const int64_t ms = __shfl_sync(0xffffffffu, ms_l, q), me = __shfl_sync(0xffffffffu, me_l, q);
                    if (me - ms > kHubNeighbour) {  // a hub among the neighbours: walked by the whole CTA below
                        int slot = 0;
                        if (lane == 0) slot = atomicAdd(&s_nh, 1);
                        slot = __shfl_sync(0xffffffffu, slot, 0);
                        if (slot < kHubSlots) {
                            if (lane == 0) { s_hms[slot] = ms; s_hme[slot] = me; }
                            continue;
                        }
                    }
                    for (int64_t oo = ms + lane; oo < me; oo += 32) {
                        int64_t l = ldg_i32(col + oo);
                        if (fold > 0) l %= fold;
                        body((uint32_t)l);
                    }
                }
            }
        }
        // the hubs among the neighbours (ncu: a quarter of the stall samples sat at the barrier after this walk -- the warp
        // that drew a 5 000-column neighbour kept the other seven waiting)
        __syncthreads();
        const int nh = s_nh < kHubSlots ? s_nh : kHubSlots;
        for (int h = 0; h < nh; ++h) {
            const int64_t ms = s_hms[h], me = s_hme[h];
            for (int64_t oo = ms + tid; oo < me; oo += kRowThreads) {
                int64_t l = ldg_i32(col + oo);
                if (fold > 0) l %= fold;
                body((uint32_t)l);
            }
        }
        __syncthreads();
        if (tid == 0) s_nh = 0;
    };
    while (true) {
        // rows are handed out by a counter: their cost spans four orders of magnitude
        if (tid == 0) { s_row = (long long)atomicAdd(row_counter, 1ull); s_np = 0; }
        __syncthreads();
        const int64_t r = s_row;
        if (r >= out_rows) break;
        for_each_product(r, [&](uint32_t l) {
            const uint32_t w = l >> 5;
            if (atomicOr(&L0[w], 1u << (l & 31u)) == 0u) atomicOr(&L1[w >> 5], 1u << (w & 31u));  // first touch of the word
            if (NUMERIC) {  // keep the product for the counting pass (one shared-memory counter update per warp step)
                const unsigned am = __activemask();
                const int leader = __ffs(am) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&s_np, __popc(am));
                base = __shfl_sync(am, base, leader) + __popc(am & ((1u << lane) - 1u));
                if (base < kRowStage) prodbuf[base] = l;
            }
        });
        __syncthreads();
        // distinct columns per level-1 word -> exclusive ranks (block scan, kRowThreads level-1 words at a time)
        int running = 0;
        for (int64_t j0 = 0; j0 < w1; j0 += kRowThreads) {
            const int64_t j = j0 + tid;
            uint32_t b1 = j < w1 ? L1[j] : 0u;
            int mine = 0;
            for (uint32_t b = b1; b; b &= b - 1) mine += __popc(L0[j * 32 + (__ffs(b) - 1)]);
            int off = 0, tile = 0;
            Scan(scan_tmp).ExclusiveSum(mine, off, tile);
            if (NUMERIC && j < w1) {
                pref[j] = (uint32_t)(running + off);
                uint32_t rank = (uint32_t)(running + off);
                for (uint32_t b = b1; b; b &= b - 1) {
                    const int64_t w = j * 32 + (__ffs(b) - 1);
                    pref16[w] = (uint16_t)rank;   // (wraps for rows of >= 65536 columns: those take the level-1 walk below)
                    rank += __popc(L0[w]);
                }
            }
            running += tile;
            __syncthreads();
        }
        const int total = running;
        if (!NUMERIC) {
            if (tid == 0) out_row_nnz[r] = total;
        } else {
            const int64_t obase = out_rowptr[r];
            const int np = s_np;
            const bool staged = total <= kRowStage;
            uint32_t* gcnt = reinterpret_cast<uint32_t*>(out_val) + obase;  // long rows count in place (converted below)
            const bool small = total < 65536;
            auto rank_of = [&](uint32_t l) -> uint32_t {
                // rank of column l: the rank of its bitmap word + the columns below it in the word.  (First version: the word's
                // rank summed on the fly over the earlier touched words of its level-1 word -- 25 % of the kernel's
                // instructions, the heavy rows touch most of the 32; kept for rows whose ranks do not fit 16 bits.)
                const uint32_t w = l >> 5, j = w >> 5;
                const uint32_t below = __popc(L0[w] & ((1u << (l & 31u)) - 1u));
                if (small) return (uint32_t)pref16[w] + below;
                uint32_t rank = pref[j] + below;
                for (uint32_t b = L1[j] & ((1u << (w & 31u)) - 1u); b; b &= b - 1) rank += __popc(L0[j * 32 + (__ffs(b) - 1)]);
                return rank;
            };
            if (np <= kRowStage) {
                // the usual row (97 % at collab shape): its products sit in shared memory -- count and place them by rank
                for (int k = tid; k < total; k += kRowThreads) cnt[k] = 0u;
                __syncthreads();
                for (int k = tid; k < np; k += kRowThreads) {
                    const uint32_t l = prodbuf[k];
                    const uint32_t rank = rank_of(l);
                    atomicAdd(&cnt[rank], 1u);
                    colbuf[rank] = l;   // the products of one column all write the same value
                }
                __syncthreads();
                for (int k = tid; k < total; k += kRowThreads) {
                    out_col[obase + k] = (int32_t)colbuf[k];
                    if (want_val) out_val[obase + k] = (float)cnt[k];
                }
            } else {
            if (want_val) {
                if (staged) for (int k = tid; k < total; k += kRowThreads) cnt[k] = 0u;
                else for (int k = tid; k < total; k += kRowThreads) gcnt[k] = 0u;
                __syncthreads();
                for_each_product(r, [&](uint32_t l) {
                    const uint32_t rank = rank_of(l);
                    if (staged) atomicAdd(&cnt[rank], 1u); else atomicAdd(&gcnt[rank], 1u);
                });
                __syncthreads();
            }
            // columns by rank.  Rows whose ranks fit 16 bits: the threads stride over the bitmap words, each word knows its rank
            // (a thread per level-1 word expanded up to 1024 columns one after the other while the others of the 231 waited:
            // the 12 % of collab's rows with more than 1024 distinct columns took most of the numeric pass)
            if (small) {
                for (int64_t w = tid; w < w0; w += kRowThreads) {
                    uint32_t bits = L0[w];
                    if (bits == 0u) continue;
                    uint32_t rank = pref16[w];
                    for (; bits; bits &= bits - 1) {
                        const uint32_t c = (uint32_t)(w * 32 + (__ffs(bits) - 1));
                        if (staged) colbuf[rank] = c; else out_col[obase + rank] = (int32_t)c;
                        ++rank;
                    }
                }
            } else
            for (int64_t j = tid; j < w1; j += kRowThreads) {
                uint32_t rank = pref[j];
                for (uint32_t b = L1[j]; b; b &= b - 1) {
                    const int64_t w = j * 32 + (__ffs(b) - 1);
                    for (uint32_t bits = L0[w]; bits; bits &= bits - 1) {
                        const uint32_t c = (uint32_t)(w * 32 + (__ffs(bits) - 1));
                        if (staged) colbuf[rank] = c; else out_col[obase + rank] = (int32_t)c;
                        ++rank;
                    }
                }
            }
            __syncthreads();
            if (staged) {
                for (int k = tid; k < total; k += kRowThreads) {
                    out_col[obase + k] = (int32_t)colbuf[k];
                    if (want_val) out_val[obase + k] = (float)cnt[k];
                }
            } else if (want_val) {
                for (int k = tid; k < total; k += kRowThreads) out_val[obase + k] = (float)gcnt[k];
            }
            }
        }
        // the bitmaps go back to zero through the touched-word list
        for (int64_t j = tid; j < w1; j += kRowThreads) {
            for (uint32_t b = L1[j]; b; b &= b - 1) L0[j * 32 + (__ffs(b) - 1)] = 0u;
            L1[j] = 0u;
        }
        __syncthreads();
    }
    if (!NUMERIC && fold > 0) {
        for (int64_t r = out_rows + (int64_t)blockIdx.x * kRowThreads + tid; r < n; r += (int64_t)gridDim.x * kRowThreads)
            out_row_nnz[r] = 0;
    }
}

// ---- DENSE: rows as bit vectors ------------------------------------------------------------------------------------
// Bits[i][w]: bit l of row i = A[i, l]; W words per row (a multiple of 4)
__global__ void k_dense_bits(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n, int W,
                             uint32_t* __restrict__ bits) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    uint32_t* row = bits + i * W;
    for (int w = lane; w < W; w += 32) row[w] = 0u;
    __syncwarp();
    for (int64_t o = rowptr[i] + lane; o < rowptr[i + 1]; o += 32) {
        const int32_t l = ldg_i32(col + o);
        atomicOr(&row[l >> 5], 1u << (l & 31));
    }
}

// structure of output row r (one CTA per row): S[r] = OR over i = r (mod fold), m in N(i) of Bits[m], folded to Wf words;
// Sp[r][w] = number of set bits before word w; out_row_nnz[r] = popc(S[r])
__global__ void __launch_bounds__(256)
k_dense_structure(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n, int64_t fold, int W, int Wf,
                  const uint32_t* __restrict__ bits, uint32_t* __restrict__ S, uint32_t* __restrict__ Sp,
                  int64_t* __restrict__ out_row_nnz) {
    using Scan = cub::BlockScan<int, 256>;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ uint32_t s_acc[1024];   // Wf <= 1024 (n <= 32768)
    const int64_t r = blockIdx.x;
    const int64_t step = fold > 0 ? fold : n;
    for (int w = threadIdx.x; w < Wf; w += 256) s_acc[w] = 0u;
    __syncthreads();
    for (int64_t i = r; i < n; i += step) {
        for (int w = threadIdx.x; w < W; w += 256) {
            uint32_t acc = 0u;
            for (int64_t o = rowptr[i]; o < rowptr[i + 1]; ++o) acc |= __ldg(bits + (int64_t)ldg_i32(col + o) * W + w);
            if (acc) atomicOr(&s_acc[w % Wf], acc);   // fold is a multiple of 32: word w folds onto word w mod Wf
        }
    }
    __syncthreads();
    int running = 0;
    for (int w0 = 0; w0 < Wf; w0 += 256) {
        const int w = w0 + threadIdx.x;
        const uint32_t b = w < Wf ? s_acc[w] : 0u;
        int off = 0, tile = 0;
        Scan(scan_tmp).ExclusiveSum(__popc(b), off, tile);
        if (w < Wf) {
            S[r * Wf + w] = b;
            Sp[r * Wf + w] = (uint32_t)(running + off);
        }
        running += tile;
        __syncthreads();
    }
    if (out_row_nnz != nullptr && threadIdx.x == 0) out_row_nnz[r] = running;
}

// walk counts of a 64 x 64 tile of the (folded) output: count[r][c] = sum over the fold classes a, b of
// popc(Bits[r + a fold] & Bits[c + b fold]); a thread owns 4 x 4 outputs, operands word-major in shared memory
constexpr int kDT = 64;
__global__ void __launch_bounds__(256)
k_dense_counts(int64_t n, int64_t fold, int64_t out_rows, int W, int Wf, const uint32_t* __restrict__ bits,
               const uint32_t* __restrict__ S, const uint32_t* __restrict__ Sp, const int64_t* __restrict__ out_rowptr,
               int32_t* __restrict__ out_col, float* __restrict__ out_val) {
    extern __shared__ uint32_t dt_smem[];
    const int WC = 32;                                   // words per chunk
    uint32_t* As = dt_smem;                              // [WC][kDT + 4] word-major
    uint32_t* Bs = dt_smem + WC * (kDT + 4);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t r0 = (int64_t)blockIdx.y * kDT, c0 = (int64_t)blockIdx.x * kDT;
    const int64_t step = fold > 0 ? fold : n;
    unsigned acc[4][4], ones[4][4], twos[4][4];  // acc counts the fours
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = ones[a][b] = twos[a][b] = 0u;
    for (int64_t ra = 0; r0 + ra < n; ra += step) {
        for (int64_t cb = 0; c0 + cb < n; cb += step) {
            for (int wc = 0; wc < W; wc += WC) {
                __syncthreads();
                // 64 rows x WC words of each operand, read row-major (coalesced along w), stored word-major
                for (int idx = threadIdx.x; idx < kDT * WC; idx += 256) {
                    const int rr = idx / WC, w = idx - rr * WC;
                    const int64_t gi = r0 + ra + rr, gj = c0 + cb + rr;
                    const bool okw = wc + w < W;
                    // rows of the tile beyond out_rows (the last tile) or beyond n contribute nothing
                    As[w * (kDT + 4) + rr] = (okw && r0 + rr < out_rows && gi < n) ? __ldg(bits + gi * W + wc + w) : 0u;
                    Bs[w * (kDT + 4) + rr] = (okw && c0 + rr < out_rows && gj < n) ? __ldg(bits + gj * W + wc + w) : 0u;
                }
                __syncthreads();
                // four words a step: the population count is a quarter-rate instruction, so the four ANDs of an output go
                // through a carry-save adder tree first (ones / twos carried along, one count of the fours per step)
#pragma unroll 2
                for (int w = 0; w < WC; w += 4) {
                    uint32_t av[4][4], bv[4][4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint4 a = *reinterpret_cast<const uint4*>(As + (w + k) * (kDT + 4) + 4 * ty);
                        const uint4 b = *reinterpret_cast<const uint4*>(Bs + (w + k) * (kDT + 4) + 4 * tx);
                        av[k][0] = a.x; av[k][1] = a.y; av[k][2] = a.z; av[k][3] = a.w;
                        bv[k][0] = b.x; bv[k][1] = b.y; bv[k][2] = b.z; bv[k][3] = b.w;
                    }
#pragma unroll
                    for (int p = 0; p < 4; ++p)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const unsigned v0 = av[0][p] & bv[0][q], v1 = av[1][p] & bv[1][q];
                            const unsigned v2 = av[2][p] & bv[2][q], v3 = av[3][p] & bv[3][q];
                            const unsigned o = ones[p][q], t = twos[p][q];
                            const unsigned ta = (o & v0) | ((o ^ v0) & v1), o1 = o ^ v0 ^ v1;
                            const unsigned tb = (o1 & v2) | ((o1 ^ v2) & v3);
                            ones[p][q] = o1 ^ v2 ^ v3;
                            acc[p][q] += __popc((t & ta) | ((t ^ ta) & tb));
                            twos[p][q] = t ^ ta ^ tb;
                        }
                }
            }
        }
    }
    // entries of the structure get their rank from the row's prefix counts
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int64_t r = r0 + 4 * ty + p;
        if (r >= out_rows) continue;
        const int64_t obase = out_rowptr[r];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t c = c0 + 4 * tx + q;
            if (c >= out_rows) continue;
            const uint32_t word = S[r * Wf + (c >> 5)];
            if ((word >> (c & 31)) & 1u) {
                const int64_t o = obase + Sp[r * Wf + (c >> 5)] + __popc(word & ((1u << (c & 31)) - 1u));
                out_col[o] = (int32_t)c;
                if (out_val != nullptr) out_val[o] = (float)(4u * acc[p][q] + 2u * __popc(twos[p][q]) + __popc(ones[p][q]));
            }
        }
    }
}

// the whole count matrix of a dense graph (cn_build.cu: 64 x 64 AND / popcount tiles through carry-save adders, mirror
// tiles written by the transposed CTA)
__global__ void k_dense_a2(const uint32_t* __restrict__ bits, int64_t n, int W, uint32_t* __restrict__ a2,
                           const long long* __restrict__ skip_stamp, long long e0, long long e1, long long e2, long long e3);

// rows of the count matrix -> number of non-zero entries (a warp per row)
__global__ void k_a2_row_nnz(const uint32_t* __restrict__ a2, int64_t n, int64_t* __restrict__ out_row_nnz) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t* row = a2 + r * n;
    int cnt = 0;
    for (int64_t c = lane; c < n; c += 32) cnt += __ldg(row + c) != 0u ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) out_row_nnz[r] = cnt;
}

// rows of the count matrix -> CSR entries in ascending column order (a warp per row, ballot ranks)
__global__ void k_a2_fill(const uint32_t* __restrict__ a2, int64_t n, const int64_t* __restrict__ out_rowptr,
                          int32_t* __restrict__ out_col, float* __restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint32_t* row = a2 + r * n;
    int64_t base = out_rowptr[r];
    for (int64_t c0 = 0; c0 < n; c0 += 32) {
        const int64_t c = c0 + lane;
        const uint32_t v = c < n ? __ldg(row + c) : 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, v != 0u);
        if (v != 0u) {
            const int64_t o = base + __popc(bal & ((1u << lane) - 1u));
            out_col[o] = (int32_t)c;
            if (out_val != nullptr) out_val[o] = (float)v;
        }
        base += __popc(bal);
    }
}

enum GemmMode { kModeGlobal = 0, kModeSmem = 1, kModeDense = 2, kModeDenseWhole = 3 };

// out-of-line so that the symbolic and the numeric call (and ocn_spgemm_scratch_bytes) agree
static GemmMode gemm_mode(int64_t n, int64_t nnz, int64_t fold) {
    const int64_t forced = option(OCN_OPT_SPGEMM_MODE, 0);  // 1 global, 2 smem, 3 dense (tests / A-B runs)
    const int64_t ncols = fold > 0 ? (fold < n ? fold : n) : n;
    const bool dense_ok = n <= 32768 && (fold == 0 || fold % 32 == 0);
    const bool smem_ok = row_smem(ncols, true).total_bytes <= 200 * 1024;
    // the true A^2 of a dense graph whose count matrix stays within 256 MB: tiles for the whole matrix, then a compaction
    const bool whole_ok = dense_ok && fold == 0 && n <= 8192;
    if (forced == 3 && dense_ok) return kModeDense;
    if (forced == 4 && whole_ok) return kModeDenseWhole;
    if (forced == 2 && smem_ok) return kModeSmem;
    if (forced == 1) return kModeGlobal;
    if (dense_ok && nnz >= n * (n / 64 + 1)) return whole_ok ? kModeDenseWhole : kModeDense;   // mean degree >= n / 64: A^2 is (close to) full
    return smem_ok ? kModeSmem : kModeGlobal;
}

struct DenseLayout { int W, Wf; size_t bits, S, Sp, total; };
static DenseLayout dense_layout(int64_t n, int64_t fold) {
    DenseLayout d;
    d.W = dense_words(n);
    const int64_t out = fold > 0 ? (fold < n ? fold : n) : n;
    d.Wf = fold > 0 ? (int)((out + 31) / 32) : d.W;
    size_t off = 0;
    d.bits = off; off += (size_t)n * d.W * 4;
    d.S = off;    off += (size_t)out * d.Wf * 4;
    d.Sp = off;   off += (size_t)out * d.Wf * 4;
    d.total = off + 256;
    return d;
}

}  // namespace ocn

using namespace ocn;

extern "C" {

size_t ocn_spgemm_scratch_bytes(int64_t n, int64_t nnz, int64_t fold) {
    if (n <= 0) return 0;
    switch (gemm_mode(n, nnz, fold)) {
        case kModeDense: return dense_layout(n, fold).total;
        case kModeDenseWhole: return 256 + dense_bits_bytes(n) + sizeof(uint32_t) * (size_t)n * (size_t)n;
        case kModeSmem: return 256;  // the row counter
        default: break;
    }
    GemmScratch g = gemm_scratch(n);
    return g.slot_bytes * (size_t)g.slots;
}

static int dense_structure(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t fold, void* scratch,
                           int64_t* out_row_nnz, cudaStream_t st) {
    const DenseLayout d = dense_layout(n, fold);
    const int64_t out_rows = fold > 0 ? (fold < n ? fold : n) : n;
    OCN_CHECK_ARG(d.Wf <= 1024, "ocn_spgemm_a2: dense mode holds up to 32768 columns");
    uint32_t* bits = (uint32_t*)((char*)scratch + d.bits);
    k_dense_bits<<<(int)((n * 32 + 255) / 256), 256, 0, st>>>(rowptr, col, n, d.W, bits);
    OCN_LAUNCH_CHECK();
    k_dense_structure<<<(int)out_rows, 256, 0, st>>>(rowptr, col, n, fold, d.W, d.Wf, bits, (uint32_t*)((char*)scratch + d.S),
                                                      (uint32_t*)((char*)scratch + d.Sp), out_row_nnz);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

// kModeDenseWhole scratch: [stamp: 4 x int64 | bits | a2].  The symbolic call leaves the count matrix and a stamp of the
// graph it was computed from (pointers, n, nnz); the numeric call on the same scratch and graph compacts it without
// recomputing (the tile kernel's CTAs return at once when they find the stamp: no host read-back), then clears the stamp;
// without the stamp (a caller that did not run symbolic first) the matrix is computed again.
struct WholeStamp { long long rowptr, col, n, nnz; };
__global__ void k_whole_stamp(WholeStamp* s, WholeStamp v) { *s = v; }

static int dense_whole(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t nnz, void* scratch, bool skip_if_stamped,
                       cudaStream_t st) {
    const int W = dense_words(n);
    uint32_t* bits = (uint32_t*)((char*)scratch + 256);
    uint32_t* a2 = (uint32_t*)((char*)scratch + 256 + dense_bits_bytes(n));
    k_dense_bits<<<(int)((n * 32 + 255) / 256), 256, 0, st>>>(rowptr, col, n, W, bits);
    OCN_LAUNCH_CHECK();
    const unsigned tiles = (unsigned)((n + 63) / 64);
    // (skip_if_stamped: every CTA returns at once when the scratch carries this graph's stamp -- no host read-back)
    k_dense_a2<<<dim3(tiles, tiles), 256, 0, st>>>(bits, n, W, a2, skip_if_stamped ? (const long long*)scratch : nullptr,
                                                   (long long)(uintptr_t)rowptr, (long long)(uintptr_t)col, (long long)n, (long long)nnz);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_spgemm_a2_symbolic(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t nnz, int64_t fold, void* scratch,
                           int64_t* out_row_nnz, void* stream) {
    OCN_RANGE("ocn_spgemm_a2_symbolic");
    OCN_CHECK_ARG(rowptr && col && scratch && out_row_nnz, "ocn_spgemm_a2_symbolic: null pointer");
    OCN_CHECK_ARG(n > 0 && nnz >= 0 && fold >= 0, "ocn_spgemm_a2_symbolic: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const GemmMode mode = gemm_mode(n, nnz, fold);
    const int64_t out_rows = fold > 0 ? (fold < n ? fold : n) : n;
    if (mode == kModeDenseWhole) {
        if (int rc = dense_whole(rowptr, col, n, nnz, scratch, false, st)) return rc;
        const uint32_t* a2 = (const uint32_t*)((char*)scratch + 256 + dense_bits_bytes(n));
        k_a2_row_nnz<<<(int)((n * 32 + 255) / 256), 256, 0, st>>>(a2, n, out_row_nnz);
        OCN_LAUNCH_CHECK();
        const WholeStamp stamp = {(long long)(uintptr_t)rowptr, (long long)(uintptr_t)col, (long long)n, (long long)nnz};
        k_whole_stamp<<<1, 1, 0, st>>>((WholeStamp*)scratch, stamp);
        OCN_LAUNCH_CHECK();
        return OCN_OK;
    }
    if (mode == kModeDense) {
        if (fold > 0 && out_rows < n) OCN_CUDA(cudaMemsetAsync(out_row_nnz + out_rows, 0, sizeof(int64_t) * (size_t)(n - out_rows), st));
        return dense_structure(rowptr, col, n, fold, scratch, out_row_nnz, st);
    }
    if (mode == kModeSmem) {
        const RowSmem rsm = row_smem(out_rows, false);
        OCN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned long long), st));
        unsigned long long* counters = (unsigned long long*)scratch;
        OCN_CUDA(cudaFuncSetAttribute(k_spgemm_row_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm.total_bytes));
        int per_sm = (int)((200u * 1024u) / (rsm.total_bytes + 2048));
        per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
        k_spgemm_row_smem<false><<<sm_count() * per_sm, kRowThreads, rsm.total_bytes, st>>>(
            rowptr, col, n, fold, (int64_t)rsm.w0, (int64_t)rsm.w1, out_row_nnz, nullptr, nullptr, nullptr, counters);
        OCN_LAUNCH_CHECK();
        return OCN_OK;
    }
    GemmScratch g = gemm_scratch(n);
    k_spgemm_a2<false><<<g.slots, kGemmThreads, 0, st>>>(
        rowptr, col, n, fold, (unsigned char*)scratch, g.words, g.slot_bytes, out_row_nnz, nullptr, nullptr, nullptr);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

int ocn_spgemm_a2_numeric(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t nnz, int64_t fold, void* scratch,
                          const int64_t* out_rowptr, int32_t* out_col, float* out_val, void* stream) {
    OCN_RANGE("ocn_spgemm_a2_numeric");
    OCN_CHECK_ARG(rowptr && col && scratch && out_rowptr && out_col, "ocn_spgemm_a2_numeric: null pointer");
    OCN_CHECK_ARG(n > 0 && nnz >= 0 && fold >= 0, "ocn_spgemm_a2_numeric: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const GemmMode mode = gemm_mode(n, nnz, fold);
    const int64_t out_rows = fold > 0 ? (fold < n ? fold : n) : n;
    if (mode == kModeDenseWhole) {
        if (int rc = dense_whole(rowptr, col, n, nnz, scratch, true, st)) return rc;
        const uint32_t* a2 = (const uint32_t*)((char*)scratch + 256 + dense_bits_bytes(n));
        k_a2_fill<<<(int)((n * 32 + 255) / 256), 256, 0, st>>>(a2, n, out_rowptr, out_col, out_val);
        OCN_LAUNCH_CHECK();
        OCN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(WholeStamp), st));
        return OCN_OK;
    }
    if (mode == kModeDense) {
        if (int rc = dense_structure(rowptr, col, n, fold, scratch, nullptr, st)) return rc;
        const DenseLayout d = dense_layout(n, fold);
        const size_t smem = sizeof(uint32_t) * 2 * 32 * (kDT + 4);
        dim3 grid((unsigned)((out_rows + kDT - 1) / kDT), (unsigned)((out_rows + kDT - 1) / kDT));
        k_dense_counts<<<grid, 256, smem, st>>>(n, fold, out_rows, d.W, d.Wf, (const uint32_t*)((char*)scratch + d.bits),
                                                (const uint32_t*)((char*)scratch + d.S), (const uint32_t*)((char*)scratch + d.Sp),
                                                out_rowptr, out_col, out_val);
        OCN_LAUNCH_CHECK();
        return OCN_OK;
    }
    if (mode == kModeSmem) {
        const RowSmem rsm = row_smem(out_rows, true);
        OCN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned long long), st));
        unsigned long long* counters = (unsigned long long*)scratch;
        OCN_CUDA(cudaFuncSetAttribute(k_spgemm_row_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsm.total_bytes));
        int per_sm = (int)((200u * 1024u) / (rsm.total_bytes + 2048));
        per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
        k_spgemm_row_smem<true><<<sm_count() * per_sm, kRowThreads, rsm.total_bytes, st>>>(
            rowptr, col, n, fold, (int64_t)rsm.w0, (int64_t)rsm.w1, nullptr, out_rowptr, out_col, out_val, counters);
        OCN_LAUNCH_CHECK();
        return OCN_OK;
    }
    GemmScratch g = gemm_scratch(n);
    k_spgemm_a2<true><<<g.slots, kGemmThreads, 0, st>>>(
        rowptr, col, n, fold, (unsigned char*)scratch, g.words, g.slot_bytes, nullptr, out_rowptr, out_col, out_val);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // extern "C"
