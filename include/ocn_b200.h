/*
 * ocn_b200.h -- C ABI of the B200-native OCN common-neighbour hot path.
 *
 * The reference (qingpingmo/OCN) has no FFI of its own: it is pure Python over torch_sparse /
 * pygho / PyG.  Each entry point below therefore names the reference *Python call site* whose
 * arithmetic it replaces (file:line under /root/reference) -- that is the interface a
 * maintainer binds (see INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - every function returns 0 on success, a negative OCN_E* code on failure, never throws;
 *     ocn_last_error() returns a thread-local description of the last failure;
 *   - all pointers are DEVICE pointers owned by the caller (torch), nothing is allocated or
 *     freed inside the library; variable-size outputs use count -> caller allocates -> fill;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) of the current
 *     device; no call synchronises;
 *   - graphs are CSR: rowptr int64[n+1], col int32[nnz], columns ascending and duplicate free
 *     inside a row (what torch_sparse.SparseTensor.csr() holds, with col narrowed to int32);
 *   - features / weights are fp32, row-major contiguous; target edges are int64.
 *   - the fused CN path (ocn_cn_*) requires a symmetric graph (the reference builds every adj
 *     with to_symmetric(): ogbdataset.py:44-45, NeighborOverlap_large.py:63).
 */
#ifndef OCN_B200_H
#define OCN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCN_OK 0
#define OCN_EINVAL (-1)   /* bad argument */
#define OCN_ECUDA (-2)    /* CUDA runtime / launch error */
#define OCN_ENOSPACE (-3) /* caller-provided buffer too small */

#define OCN_ABI_VERSION 3

int ocn_abi_version(void);
const char* ocn_last_error(void);
/* number of SMs of the current device (grid sizing is done inside; exposed for bench.py) */
int ocn_device_sm_count(void);

/* Tuning / test options of the library (process-wide; not part of the data path's contract).  The parity tests use
 * them to force, on small graphs, the code paths that production sizes reach on their own (position windows,
 * heavy-source passes), and bench.py / scripts/ab.py use them for one-process A/B runs of kernel variants.
 * value 0 always means "the built-in default".  Returns OCN_EINVAL for an unknown key. */
#define OCN_OPT_HUB_WINDOW 0      /* positions a warp-private counter window of k_cn_hub_count holds (default 4096) */
#define OCN_OPT_HUB_CTA_WINDOW 1  /* positions of the CTA-wide counter window (default 32768) */
#define OCN_OPT_HUB_HEAVY_RUN 2   /* a run whose source has more neighbours is indexed in its own pass (default 1024) */
#define OCN_OPT_HUB_WALKER 3      /* 1: the shared pass walks run segments (k_cn_hub_count_seg; measured slower), 0: whole lists */
#define OCN_OPT_HUB_SEG_CTAS 4    /* resident CTAs per SM of k_cn_hub_count_seg (default 5) */
#define OCN_OPT_HUB_EXACT 5       /* 1: 32-byte node entries with exact run sets + run-segment starts for streams of <= 128 runs
                                     (measured slower than the folded 64-bit sets on the bench workload; implied by WALKER = 1) */
#define OCN_OPT_GROUPED_OFF 6     /* 1: never use the run-grouped statistics / aggregation kernels (cn_grouped.cu) */
#define OCN_OPT_SPGEMM_MODE 7     /* A^2 kernel: 0 automatic, 1 global scratch, 2 shared-memory rows, 3 dense bit matrix (structure + tiled counts), 4 dense whole count matrix + compaction (fold == 0, n <= 8192) */
#define OCN_OPT_SPMM_TMA 8        /* k_spmm: 0 automatic, 1 gather neighbour rows with cp.async.bulk + mbarrier, 2 register gather, 3 lane-per-feature gather, 4 shared-memory ring fed by per-thread cp.async */
#define OCN_OPT_HEAD_TC 9         /* ocn_cn_head at in = hidden = 32: 0 / 3 tcgen05 kernel (tf32 x 3 split, fp32 accuracy; A operand in tensor memory, 4 pipelines per SM), 1 the same with A through shared memory (2 pipelines), 2 CUDA-core kernel */
#define OCN_OPT_COUNT 16
/* launches of the library's own kernels since the process started (CUB scans / sorts it calls are not counted) */
long long ocn_launch_count(void);
int ocn_set_option(int key, int64_t value);
int64_t ocn_get_option(int key);

/* ---- graph ------------------------------------------------------------------------------ */

/* Checks, on device, what torch_sparse guarantees for an adj built by
 * SparseTensor.from_edge_index(...).to_symmetric() (NeighborOverlap_large.py:59-63):
 * out_flags[0] bit0 = a column out of [0,n), bit1 = a row not strictly ascending,
 * bit2 = rowptr not monotone / rowptr[n] != nnz, bit3 = not symmetric. */
int ocn_graph_validate(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t nnz,
                       int32_t* out_flags, void* stream);

/* ---- the step before the path: building / masking the adjacency on the device ----
 * SparseTensor.from_edge_index(tei, sparse_sizes=(n, n))[.to_symmetric()] (ogbdataset.py:44-45,
 * NeighborOverlap_large.py:56-63, NeighborOverlapCitation2.py:135-143): edge list (int64 src/dst, optional
 * keep mask uint8[num_edges] = the reference's `adjmask`) -> CSR with ascending duplicate-free rows.
 * count: fills out_rowptr[n+1]; out_info[0] = nnz, out_info[1] = kept edges with an endpoint outside [0, n)
 * (they are dropped; the host mirror raises).  The caller reads nnz, allocates col int32[nnz] (and
 * optionally mult int32[nnz]: how many list edges map onto each entry) and calls fill with the same scratch. */
size_t ocn_graph_build_bytes(int64_t num_edges, int symmetric);
int ocn_graph_build_count(const int64_t* src, const int64_t* dst, const uint8_t* keep, int64_t num_edges,
                          int64_t n, int symmetric, void* scratch, size_t scratch_bytes,
                          int64_t* out_rowptr, int64_t* out_info, void* stream);
int ocn_graph_build_fill(const void* scratch, int64_t num_edges, int symmetric, int64_t n, int64_t nnz,
                         int32_t* out_col, int32_t* out_mult, void* stream);

/* Per-batch target-link masking under --maskinput (NeighborOverlap_large.py:56-63: adjmask[perm] = 0, rebuild,
 * to_symmetric) WITHOUT re-sorting the edge list: (rowptr, col, mult) is the full graph from ocn_graph_build_*
 * (mult NULL = every entry comes from exactly one list edge), (src, dst)[num_masked] the masked links.  An
 * entry survives while an unmasked list edge still maps onto it -- the result equals a rebuild from the
 * remaining list.  dec is an int32[nnz] work array, all zero on entry and on return of fill.
 * count: out_rowptr[n+1] of the masked graph, out_info[0] = its nnz, out_info[1] = masked links not found in
 * the graph; fill: out_col int32[nnz'] (and out_mult).  fill must follow count with the same scratch (count
 * leaves the sorted positions of the dead entries there; fill restores dec).  The survivors move in one flat
 * coalesced copy: entry p goes to p - #{dead positions < p}. */
size_t ocn_graph_mask_bytes(int64_t n, int64_t num_masked);
int ocn_graph_mask_count(const int64_t* rowptr, const int32_t* col, const int32_t* mult, int64_t n, int64_t nnz,
                         const int64_t* src, const int64_t* dst, int64_t num_masked, int symmetric,
                         int32_t* dec, void* scratch, size_t scratch_bytes,
                         int64_t* out_rowptr, int64_t* out_info, void* stream);
int ocn_graph_mask_fill(const int64_t* rowptr, const int32_t* col, const int32_t* mult, int64_t n, int64_t nnz,
                        const int64_t* src, const int64_t* dst, int64_t num_masked, int symmetric,
                        int32_t* dec, const void* scratch /* as left by count */,
                        int32_t* out_col, int32_t* out_mult, void* stream);

/* DropAdj (model.py:211-229: mask = rand_like(col) > dp; torch_sparse.masked_select_nnz(adj, mask); value * 1/(1-dp)),
 * applied to the GNN's adjacency in every layer of a training forward (model.py:312): the entries with keep[e] != 0
 * move, in order, into a new CSR whose values are (val or 1) * scale.  The mask comes from the caller's RNG (torch), so
 * that a seeded run drops the same entries as the reference.  count: kept entries per row; the caller scans them into
 * out_rowptr and calls fill. */
int ocn_graph_select_count(const int64_t* rowptr, const uint8_t* keep, int64_t n, int64_t* out_counts, void* stream);
int ocn_graph_select_fill(const int64_t* rowptr, const int32_t* col, const float* val /* NULL = ones */,
                          const uint8_t* keep, int64_t n, float scale, const int64_t* out_rowptr, int32_t* out_col,
                          float* out_val /* NULL = structure only */, void* stream);

/* ---- piece 1: generic per-target-edge row intersection ----------------------------------
 * adjoverlap(adj1, adj2, tarei) with calresadj=False (utils.py:248-285 -> spmoverlap_
 * utils.py:163-183): out row b = adj1[src[b]] (cap) adj2[dst[b]], columns ascending, value 1.
 * adj1/adj2 may be different matrices with the same column space (e.g. A and an explicit A^2,
 * NeighborOverlap_large.py:78-79). */
int ocn_rows_intersect_count(const int64_t* rowptr1, const int32_t* col1, int64_t n1 /* rows of matrix 1 */,
                             const int64_t* rowptr2, const int32_t* col2, int64_t n2 /* rows of matrix 2 */,
                             const int64_t* src, const int64_t* dst, int64_t num_edges,
                             int64_t* out_counts /* [num_edges + 1], the last word ZERO on entry: it receives the number
                                                    of links with src outside [0, n1) or dst outside [0, n2) -- the
                                                    reference raises IndexError there; their rows stay empty */,
                             void* stream);
/* out_rowptr = exclusive scan of out_counts (int64[num_edges+1]), computed by the caller. */
int ocn_rows_intersect_fill(const int64_t* rowptr1, const int32_t* col1, int64_t n1,
                            const int64_t* rowptr2, const int32_t* col2, int64_t n2,
                            const int64_t* src, const int64_t* dst, int64_t num_edges,
                            const int64_t* out_rowptr, int64_t* out_col, void* stream);

/* calresadj=True branch of adjoverlap (utils.py:260-274 -> spmoverlap_notoverlap_ utils.py:210-244), used by the
 * completion predictors: out row b = adj1[src[b]] \ adj2[dst[b]] (columns ascending).  The second residual,
 * adj2[dst] \ adj1[src], is the same call with the matrices and link ends swapped. */
int ocn_rows_difference_count(const int64_t* rowptr1, const int32_t* col1, int64_t n1,
                              const int64_t* rowptr2, const int32_t* col2, int64_t n2,
                              const int64_t* src, const int64_t* dst, int64_t num_edges,
                              int64_t* out_counts, void* stream);
int ocn_rows_difference_fill(const int64_t* rowptr1, const int32_t* col1, int64_t n1,
                             const int64_t* rowptr2, const int32_t* col2, int64_t n2,
                             const int64_t* src, const int64_t* dst, int64_t num_edges,
                             const int64_t* out_rowptr, int64_t* out_col, void* stream);

/* ---- container methods behind the import shims (ocn_b200/shim: torch_sparse / pygho stand-ins on CUDA) -----------
 * adj[idx] / SparseTensor.index_select(0, idx) / pygho index_select([0], idx) (utils.py:256-257,
 * NeighborOverlapCitation2.py:79-81): out row b = row idx[b] of the matrix, columns (and values) in order.
 * out_counts has num_rows + 1 words, the last ZERO on entry: it receives the number of idx outside [0, n). */
int ocn_rows_gather_count(const int64_t* rowptr, int64_t n, const int64_t* idx, int64_t num_rows, int64_t* out_counts,
                          void* stream);
int ocn_rows_gather_fill(const int64_t* rowptr, const int32_t* col, const float* val /* NULL = ones */, int64_t n,
                         const int64_t* idx, int64_t num_rows, const int64_t* out_rowptr, int32_t* out_col,
                         float* out_val /* NULL = structure only */, void* stream);
/* pygho.backend.Spspmm.spsphadamard(A, B) on two explicit [num_rows x N] matrices (model.py:2243 innerprod1; the
 * general case of NeighborOverlapCitation2.py:82-85): entries present in both, value va * vb (NULL values = ones),
 * columns ascending. */
int ocn_rows_hadamard_count(const int64_t* rowptr_a, const int32_t* col_a, const int64_t* rowptr_b, const int32_t* col_b,
                            int64_t num_rows, int64_t* out_counts /* [num_rows] */, void* stream);
int ocn_rows_hadamard_fill(const int64_t* rowptr_a, const int32_t* col_a, const float* val_a, const int64_t* rowptr_b,
                           const int32_t* col_b, const float* val_b, int64_t num_rows, const int64_t* out_rowptr,
                           int32_t* out_col, float* out_val, void* stream);
/* SparseTensor.sum(dim=0) (model.py:2261, 3114): out[col[e]] += val[e] (NULL = 1) into the caller's ZEROED out[n_cols]. */
int ocn_csr_colsum(const int32_t* col, const float* val, int64_t nnz, int64_t n_cols, float* out, void* stream);

/* ---- pieces 1 + 1b + 2, fused: higher-order CN sets over A, A^2, A^3 ------------------------
 * Replaces get_cn1_cn2 (NeighborOverlapCitation2.py:78-104, NeighborOverlap_large_ppa.py:147-173)
 * and adjoverlap(adj, adj, e) / adjoverlap(adj, adj2, e) (NeighborOverlap_large.py:78-79)
 * without materialising A^2/A^3, then the cn5 / cn7 / cn6 combination
 * (model.py:2261-2429, 3114-3216, 2546-2940).
 *
 * A call processes a STREAM of num_edges target links cut into consecutive batches of
 * batch_size links (the reference's PermIterator batches, utils.py:8-36); every batch is
 * normalised on its own, exactly as one multidomainforward call would.
 *
 * Step 1  ocn_cn_plan      sizes + schedule              (sync-free; caller reads 4 int64 back)
 * Step 2  ocn_cn_build     CN_k(e) = A[i] (*) A^k[j], k=1..order, as dense records over N(i)
 *                          + per-batch column statistics
 * Step 3  ocn_cn_stats     per-batch scalars (scale factor, inner products)
 * Step 4  ocn_cn_aggregate xcn_k = C-hat_k @ x, x_i * x_j
 *         ocn_cn_aggregate_bwd  grad_x += C-hat_k^T @ grad_xcn_k
 *         ocn_cn_extract_count / _fill   the sparse [B x N] CN matrices themselves
 * Step 5  ocn_cn_release   resets the column statistics touched by this stream
 */

/* plan[] int64 layout written by ocn_cn_plan (device): */
#define OCN_PLAN_NUM_RECORDS 0 /* sum over edges of deg(src) */
#define OCN_PLAN_NUM_RUNS 1    /* maximal runs of consecutive links with one source */
#define OCN_PLAN_NUM_UNITS 2   /* work units of ocn_cn_build */
#define OCN_PLAN_NUM_BATCHES 3
/* words 4..7 are internal to the library */
#define OCN_PLAN_HUB_DEGREE 8    /* rows of >= this many columns go through the hub stage (0: stage off) */
#define OCN_PLAN_HUB_PAIRS 9     /* (hub row, link) pairs of the stream */
#define OCN_PLAN_HUB_ENTRIES 10  /* (key, run, position) entries of the stream */
#define OCN_PLAN_HUB_POSITIONS 11 /* sum over runs of deg(src) */
#define OCN_PLAN_BAD_LINKS 17    /* links with an endpoint outside [0, n): the plan stops, the caller must not build */
#define OCN_PLAN_WIDE_LINKS 16   /* links whose source has more than 64 neighbours (they stay with the per-link kernels) */
#define OCN_PLAN_DENSE 18        /* 1: orders <= 2 on a dense graph (mean degree >= n / 64, n <= 32768) are built from bit-vector
                                    rows (AND + popcount per record); ocn_cn_hub_bytes then sizes the bit matrix */
#define OCN_PLAN_WORDS 24

/* bytes of plan scratch the caller must provide for a stream of num_edges links */
size_t ocn_cn_plan_bytes(int64_t num_edges);
/* bytes of per-batch column statistics for one batch over n nodes */
size_t ocn_cn_colstat_bytes(int64_t n);
/* bytes per record */
size_t ocn_cn_record_bytes(void);

int ocn_cn_plan(const int64_t* rowptr, const int32_t* col, int64_t n,
                const int64_t* src, const int64_t* dst, int64_t num_edges, int64_t batch_size,
                int order /* highest CN order that will be built: sizes the work units */,
                int64_t hub_degree /* 0: automatic, -1: hub stage off, > 0: rows of at least this many columns */,
                void* plan_scratch, size_t plan_scratch_bytes,
                int64_t* out_plan /* device int64[OCN_PLAN_WORDS] */, void* stream);

/* records: device buffer of >= plan[NUM_RECORDS] * ocn_cn_record_bytes() bytes.
 * colstat: device buffer of num_batches * ocn_cn_colstat_bytes(n) bytes, ZERO on entry
 *          (ocn_cn_release restores that), or NULL when only the sets are wanted.
 * order in {1,2,3}; weighted != 0 keeps walk counts as values (pygho path), == 0 uses the 0/1
 * structure (torch_sparse path).  weighted == 2 (ocn_cn_build only; the later stages take 1): the
 * shortest-path variant of SPD.py:65-126 -- the 2-walk count of a node that is itself a neighbour of the destination
 * is zeroed (compute_adj2_with_shortest_paths masks the entries of A out of Ej . A). */
int ocn_cn_build(const int64_t* rowptr, const int32_t* col, int64_t n,
                 const int64_t* src, const int64_t* dst, int64_t num_edges, int64_t batch_size,
                 int order, int weighted,
                 const void* plan_scratch, const int64_t* plan,
                 void* records, int64_t records_capacity, void* colstat,
                 int64_t nnz /* rowptr[n]; only read by the hub stage */,
                 const int64_t* plan_host /* HOST copy of plan[OCN_PLAN_WORDS] (the caller read it back to size
                                             the buffers); NULL = hub stage off */,
                 void* hub_scratch /* ocn_cn_hub_bytes(...) bytes, NULL = hub stage off */, size_t hub_scratch_bytes,
                 void* node_scratch /* 2 x 32 bytes per node (uint4[4n]), ZERO on entry, zero again when the call has run */,
                 void* stream);

/* Hub stage of the order-3 walk (cn_hub.cu): a row N(m) that many links of the stream would walk
 * is streamed once for all of them.  Scratch bytes for the sizes ocn_cn_plan reported.  When the plan chose the
 * dense build for orders <= 2 (plan[OCN_PLAN_DENSE]) the same scratch holds the graph's rows as bit vectors. */
size_t ocn_cn_hub_bytes(int64_t n, int64_t nnz, const int64_t* plan_host);
/* A build that fails half way restores node_scratch to zero itself; if even that fails (a sticky CUDA error) the
 * buffer is remembered as dirty and every later ocn_cn_build on it is refused until the caller has zeroed it and
 * called this. */
int ocn_cn_hub_scratch_reset(const void* node_scratch);
/* Measurement hook: two cudaEvent_t recorded on the build's stream right before / after the
 * dominant kernel of the indexed path (k_cn_hub_count).  NULL, NULL switches it off. */
int ocn_cn_hub_timing_events(void* start_event, void* stop_event);

/* variant: 5 = cn5/OCN (order 2) or its order-3 form (cn6 template), 7 = cn7/OCNP.
 * fill   : weight of a node that is CN1 of exactly one edge of the batch (cn5: 0, cn7: args.sum).
 * ip     : device fp32[3] = inner-product buffer value used for (C2|C1-hat), (C3|C1-hat), (C3|C2-hat);
 *          eval mode passes the stored buffer three times (model.py:2241-2250, Q9).  NULL (aggregate / aggregate_bwd /
 *          extract / stats stage 1): every batch reads its own three coefficients from batch_scalars[b*8 + 5..7], which
 *          the caller wrote after stage 0 -- the sub-batches of one optimiser step as ONE session, each with the running
 *          mean it would have seen in sequence.
 * out_batch_scalars: device fp32[num_batches * 8]:
 *          [0] scale = max|C1-hat|, [1] s12 = sum(C2 * C1-hat), [2] s13, [3] s23 (training only). */
int ocn_cn_stats(const int64_t* rowptr, const int32_t* col, int64_t n,
                 const int64_t* src, int64_t num_edges, int64_t batch_size,
                 int order, int weighted, int variant, float fill, const float* ip,
                 int stage /* 0: scale + s12, 1: s13 + s23 (needs stage 0 and final ip[0]) */,
                 const void* plan_scratch, const void* records, const void* colstat,
                 float* batch_scalars,
                 const int64_t* plan_host /* HOST copy of plan[OCN_PLAN_WORDS] as the caller read it back, or NULL.
                                             Streams of long runs (>= 16 links per run on average) use the run-grouped
                                             kernels of cn_grouped.cu; NULL keeps the per-link kernels */,
                 void* stream);

int ocn_cn_aggregate(const int64_t* rowptr, const int32_t* col, int64_t n,
                     const int64_t* src, const int64_t* dst, int64_t num_edges, int64_t batch_size,
                     int order, int weighted, int variant, float fill, const float* ip,
                     const void* plan_scratch, const void* records, const void* colstat,
                     const float* batch_scalars,
                     const float* x, int64_t feat,
                     float* xcn1, float* xcn2, float* xcn3 /* NULL unless order 3 */,
                     float* xij /* x[src]*x[dst], may be NULL */, const int64_t* plan_host /* as for ocn_cn_stats */, void* stream);

int ocn_cn_aggregate_bwd(const int64_t* rowptr, const int32_t* col, int64_t n,
                         const int64_t* src, const int64_t* dst, int64_t num_edges, int64_t batch_size,
                         int order, int weighted, int variant, float fill, const float* ip,
                         const void* plan_scratch, const void* records, const void* colstat,
                         const float* batch_scalars,
                         const float* x, int64_t feat,
                         const float* g_xcn1, const float* g_xcn2, const float* g_xcn3, const float* g_xij,
                         float* grad_x /* [n, feat], accumulated into */, void* stream);

/* which: 1,2,3 = raw CN_k (value 1 / walk count); 11,12,13 = normalised C-hat_k as the predictor
 * builds them (pattern of C-hat_2 = CN1 u CN2 etc., explicit zeros kept, model.py:2352-2391). */
int ocn_cn_extract_count(const int64_t* rowptr, int64_t n, const int64_t* src, int64_t num_edges,
                         int which, int weighted, const void* plan_scratch, const void* records,
                         int64_t* out_counts, void* stream);
int ocn_cn_extract_fill(const int64_t* rowptr, const int32_t* col, int64_t n,
                        const int64_t* src, int64_t num_edges, int64_t batch_size,
                        int which, int weighted, int variant, float fill, const float* ip,
                        const void* plan_scratch, const void* records, const void* colstat,
                        const float* batch_scalars,
                        const int64_t* out_rowptr, int64_t* out_col, float* out_val, void* stream);

int ocn_cn_release(const int64_t* rowptr, const int32_t* col, int64_t n,
                   const int64_t* src, int64_t num_edges, int64_t batch_size,
                   const void* plan_scratch, const void* records, void* colstat, const int64_t* plan_host /* as for ocn_cn_stats */,
                   void* stream);

/* ---- piece 2 / 3a: CSR SpMM ---------------------------------------------------------------
 * spmm_add / spmm_mean / spmm_max(src, other) (torch_sparse.matmul; model.py:6,45-53,2426-2427).
 * val may be NULL (all ones). reduce: 0 sum, 1 mean, 2 max (empty rows -> 0). */
int ocn_spmm_csr(const int64_t* rowptr, const int32_t* col, const float* val, int64_t num_rows,
                 const float* x, int64_t feat, int reduce, float* out, void* stream);
/* grad_x[col] += val * grad_out[row]  (transpose SpMM by scatter; reduce sum/mean only) */
int ocn_spmm_csr_bwd(const int64_t* rowptr, const int32_t* col, const float* val, int64_t num_rows,
                     const float* grad_out, int64_t feat, int reduce, float* grad_x, void* stream);
/* spmm_max backward (PureConv aggr="max", model.py:47): grad_x[argmax col of (row, f)] += val * grad_out[row, f];
 * the first maximum in row order wins, as in torch_sparse's running strict comparison; x is the forward input */
int ocn_spmm_csr_max_bwd(const int64_t* rowptr, const int32_t* col, const float* val, int64_t num_rows,
                         const float* x, const float* grad_out, int64_t feat, float* grad_x, void* stream);

/* GCN-style aggregation fused with its degree normalisation.
 * mode 3: PureConv "gcn" (model.py:51-54)   out = nrm * (A (nrm*x) + nrm*x), nrm = rsqrt(1 + rowsum)
 *         (== PyG GCNConv with self loops on a loop-free unit-weight graph, model.py:58-60)
 * mode 4: PureConv3 "gcn" (model.py:136-140) out = (nrm_r * a_rc * nrm_c) x, no self term
 * edge_w may be NULL (unit weights; DropAdj supplies rescaled values, model.py:219-229). */
int ocn_gcn_norm(const int64_t* rowptr, const float* edge_w, int64_t n, float* out_norm, void* stream);
int ocn_gcn_spmm(const int64_t* rowptr, const int32_t* col, const float* edge_w, int64_t n,
                 const float* norm, int mode, const float* x, int64_t feat, float* out, void* stream);

/* ---- piece 3b: A^2 SpGEMM -----------------------------------------------------------------
 * spadj @ spadj (NeighborOverlap_large.py:74,119) and sparse_tensor_multiply (utils.py:287-329).
 * fold == 0: the true A^2.  fold == bs > 0: the reference's adj2byblock result, every bs x bs
 * block summed into the top-left corner.  Two phase: symbolic writes the nnz of
 * every output row, numeric fills ascending columns (+ fp32 2-walk counts if out_val != NULL).
 * scratch: ocn_spgemm_scratch_bytes(n, nnz, fold) bytes, zero on first use (left zero on return).  The library
 * picks one of four kernels from (n, nnz, fold) -- dense bit-matrix rows (ddi; fold == 0 and n <= 8192: the whole count
 * matrix by AND / popcount tiles, compacted by the numeric call, which reuses the symbolic call's matrix when it runs on
 * the same scratch and graph; both dense forms count popc(row r & row c), i.e. they assume A symmetric as every adj of
 * the reference is), a shared-memory row accumulator
 * (collab, Planetoid), a global-scratch accumulator (column spaces beyond shared memory); OCN_OPT_SPGEMM_MODE forces
 * one for tests. */
size_t ocn_spgemm_scratch_bytes(int64_t n, int64_t nnz, int64_t fold);
int ocn_spgemm_a2_symbolic(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t nnz, int64_t fold,
                           void* scratch, int64_t* out_row_nnz, void* stream);
int ocn_spgemm_a2_numeric(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t nnz, int64_t fold,
                          void* scratch, const int64_t* out_rowptr, int32_t* out_col, float* out_val,
                          void* stream);

/* ---- the step after the aggregation: fused inference-mode MLP head --------------------
 * out[b, :] = lin(mix[0]*xcn1lin(xcn1[b]) + mix[1]*xcn2lin(xcn2[b]) [+ mix[2]*xcn3lin(xcn3[b])] + mix[3]*xijlin(xij[b]))
 * (model.py:2192-2235 modules, :2429-2437 combination; mix = sigmoid(alpha).cumprod and beta).  Served for
 * in_ch, hid in {32, 64} with <= 200 KB of parameters (they live in shared memory); ocn_cn_head_params returns
 * the length of the packed fp32 parameter buffer (layout documented in csrc/head.cu) or -1 when the head is
 * not served (the caller keeps its GEMMs).  flags: bit0 LayerNorm, bit1 tailact, bit2 twolayerlin.
 * xcn3 NULL = two CN branches (cn5 / cn7). */
int64_t ocn_cn_head_params(int in_ch, int hid, int out_ch, int flags, int branches);
int ocn_cn_head(const float* xcn1, const float* xcn2, const float* xcn3, const float* xij, int64_t num_links,
                int in_ch, int hid, int out_ch, int flags, const float* params, int64_t params_len,
                const float* mix, float* out, void* stream);

/* ---- wider heads: one fp32-accurate Linear layer with its element-wise tail on the tensor cores (csrc/linear_tc.cu) ----
 * v = a[rows, k] . w[n, k]^T + bias; optional LayerNorm(gamma, beta, eps 1e-5); optional ReLU; then any of
 *   out[rows, n] = v;   z[rows, n] = (z_accumulate ? z : 0) + z_scale * v;   out_final[rows, out_ch] = v . wo[out_ch, n]^T + bo
 * (the layers of the nn.Sequential heads, model.py:2192-2235, 2429-2437, with the branch mix and the last Linear fused).
 * n in {32, 64, 128, 256}, k a multiple of 32 (<= 1024).  `prepped` is w split hi + lo and laid out per 32-wide K chunk
 * by ocn_linear_tc_prep (ocn_linear_tc_prep_floats(n, k) floats, -1 = shape not served); redo it when w changes.
 * tcgen05.mma kind::tf32 with three products per K step: results agree with an fp32 GEMM to ~1e-6 relative. */
int64_t ocn_linear_tc_prep_floats(int n, int k);
int ocn_linear_tc_prep(const float* w, int n, int k, float* prepped, void* stream);
int ocn_linear_tc(const float* a, int64_t rows, int k, int n, const float* prepped, const float* bias,
                  const float* ln_gamma /* NULL = no LayerNorm */, const float* ln_beta, int relu,
                  float* out /* or NULL */, float* z /* or NULL */, float z_scale, int z_accumulate,
                  const float* wo /* or NULL */, const float* bo, int out_ch, float* out_final, void* stream);

/* ---- the step after the path: ranking metrics on the device --------------------------
 * ogb Evaluator.eval as the drivers call it (NeighborOverlap_large.py:162-179: Hits@K over all positive /
 * negative scores of a split; NeighborOverlapCitation2.py:256-259: MRR of every source against its own
 * 1000 negatives), so that scores stay on the device.
 * ocn_mrr: out_mrr[b] = 1 / (0.5 * (#{neg[b,:] > pos[b]} + #{neg[b,:] >= pos[b]}) + 1)   (ogb >= 1.3.3)
 * ocn_hits_at_k: out_hits[0] = mean(pos > k-th largest neg), 1.0 when num_neg < k. */
int ocn_mrr(const float* pos, const float* neg, int64_t num_sources, int64_t negs_per_source,
            float* out_mrr, void* stream);
size_t ocn_hits_bytes(int64_t num_neg);
int ocn_hits_at_k(const float* pos, int64_t num_pos, const float* neg, int64_t num_neg, int64_t k,
                  void* scratch, size_t scratch_bytes, float* out_hits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCN_B200_H */
