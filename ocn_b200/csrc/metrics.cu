// The step AFTER the path: the ranking metrics of the drivers, on the device, so that
// the scores never take the per-batch `.cpu()` hop of NeighborOverlap_large.py:122-160 /
// NeighborOverlapCitation2.py:241-259.  Definitions follow ogb 1.3.6 `Evaluator` (absent here; restated in
// oracle/ref_ops.py): Hits@K = mean(pos > K-th largest negative) (1.0 when there are fewer than K
// negatives); MRR per source = 1 / (0.5 * (#neg > pos + #neg >= pos) + 1) over its own row of negatives.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace ocn {

// one warp per source row: exact integer rank counts, one division
__global__ void __launch_bounds__(256)
k_mrr_rows(const float* __restrict__ pos, const float* __restrict__ neg, int64_t B, int64_t K, float* __restrict__ out) {
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int lane = lane_id();
    for (int64_t b = warp; b < B; b += nwarps) {
        const float p = pos[b];
        int gt = 0, ge = 0;
        for (int64_t k = lane; k < K; k += 32) {
            const float v = __ldg(neg + b * K + k);
            gt += v > p;
            ge += v >= p;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gt += __shfl_xor_sync(0xffffffffu, gt, o);
            ge += __shfl_xor_sync(0xffffffffu, ge, o);
        }
        if (lane == 0) out[b] = __fdiv_rn(1.0f, 0.5f * (float)(gt + ge) + 1.0f);
    }
}

__global__ void k_hits_count(const float* __restrict__ pos, int64_t P, const float* __restrict__ neg_desc, int64_t M,
                             int64_t K, unsigned long long* __restrict__ counter) {
    if (M < K) return;  // fewer than K negatives: every positive is a hit (finalised below)
    const float kth = neg_desc[K - 1];
    unsigned long long c = 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < P; t += (int64_t)gridDim.x * blockDim.x)
        c += pos[t] > kth;
    c = __reduce_add_sync(0xffffffffu, (unsigned)c);
    if (lane_id() == 0 && c) atomicAdd(counter, c);
}

__global__ void k_hits_finish(const unsigned long long* __restrict__ counter, int64_t P, int64_t M, int64_t K,
                              float* __restrict__ out) {
    out[0] = (M < K) ? 1.0f : (float)((double)(*counter) / (double)(P > 0 ? P : 1));
}

static size_t sort_bytes(int64_t M) {
    size_t b = 0;
    cub::DeviceRadixSort::SortKeysDescending(nullptr, b, (const float*)nullptr, (float*)nullptr, (int)(M > 0 ? M : 1));
    return b;
}

}  // namespace ocn

using namespace ocn;

extern "C" {

int ocn_mrr(const float* pos, const float* neg, int64_t num_sources, int64_t negs_per_source, float* out_mrr, void* stream) {
    OCN_CHECK_ARG(num_sources >= 0 && negs_per_source >= 0, "ocn_mrr: bad sizes");
    if (num_sources == 0) return OCN_OK;
    OCN_CHECK_ARG(pos && out_mrr && (neg || negs_per_source == 0), "ocn_mrr: null pointer");
    int64_t want = (num_sources + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 16;
    k_mrr_rows<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(pos, neg, num_sources, negs_per_source, out_mrr);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

size_t ocn_hits_bytes(int64_t num_neg) {
    const size_t a = (sizeof(float) * (size_t)(num_neg > 0 ? num_neg : 1) + 255) & ~size_t(255);
    return a + 256 + ((sort_bytes(num_neg) + 255) & ~size_t(255));
}

int ocn_hits_at_k(const float* pos, int64_t num_pos, const float* neg, int64_t num_neg, int64_t k, void* scratch,
                  size_t scratch_bytes, float* out_hits, void* stream) {
    OCN_CHECK_ARG(num_pos >= 0 && num_neg >= 0 && k > 0 && num_neg < (int64_t(1) << 31), "ocn_hits_at_k: bad sizes");
    OCN_CHECK_ARG(out_hits && scratch && (pos || num_pos == 0) && (neg || num_neg == 0), "ocn_hits_at_k: null pointer");
    OCN_CHECK_ARG(scratch_bytes >= ocn_hits_bytes(num_neg), "ocn_hits_at_k: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)scratch;
    const size_t a = (sizeof(float) * (size_t)(num_neg > 0 ? num_neg : 1) + 255) & ~size_t(255);
    float* sorted = (float*)base;
    unsigned long long* counter = (unsigned long long*)(base + a);
    OCN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    if (num_neg >= k) {
        size_t tb = scratch_bytes - a - 256;
        OCN_CUDA(cub::DeviceRadixSort::SortKeysDescending(base + a + 256, tb, neg, sorted, (int)num_neg, 0, 32, st));
        if (num_pos > 0) {
            int64_t want = (num_pos + 255) / 256;
            const int64_t cap = (int64_t)sm_count() * 8;
            k_hits_count<<<(int)(want < cap ? want : cap), 256, 0, st>>>(pos, num_pos, sorted, num_neg, k, counter);
            OCN_LAUNCH_CHECK();
        }
    }
    k_hits_finish<<<1, 1, 0, st>>>(counter, num_pos, num_neg, k, out_hits);
    OCN_LAUNCH_CHECK();
    return OCN_OK;
}

}  // extern "C"
