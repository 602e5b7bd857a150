// Weight algebra of the cn5 / cn6 / cn7 combination, shared by the per-link kernels (cn_aggregate.cu) and the
// run-grouped kernels (cn_grouped.cu).  See the header of cn_aggregate.cu for the formulas and the reference lines.
#pragma once

#include "common.cuh"

namespace ocn {

struct WeightParams {
    int order, weighted, variant;
    float fill, ipn_a, ipn_b, ipn_c;
};

struct EntryWeights {
    float w1, w2, w3;   // C1h, C2h (cn7: raw C2), C3h values of this (link, node)
    bool in1, in2, in3; // membership in the three patterns
};

// The part of the weights that depends on the node (column statistics of the batch) only: the three divisions.  The
// run-grouped kernels evaluate it once per (batch, position) instead of once per record.
struct NodeWeights {
    float w1k;    // 1/c1 (c1 >= 2) or fill
    float inv2;   // 1 / c2(k)
    float inv3;   // 1 / c3(k)
};

__device__ __forceinline__ NodeWeights node_weights(uint32_t c1cnt, unsigned long long s2, unsigned long long s3,
                                                    const WeightParams& P) {
    NodeWeights N;
    N.w1k = (c1cnt >= 2u) ? __fdiv_rn(1.0f, (float)c1cnt) : P.fill;
    N.inv2 = 0.0f;
    N.inv3 = 0.0f;
    if (P.variant == 7 || P.order < 2) return N;
    const float c1f = (float)c1cnt;
    const float corr1a = (c1cnt >= 2u) ? __fmul_rn(__fmul_rn(P.ipn_a, N.w1k), c1f) : 0.0f;
    const float c2raw = __fsub_rn((float)s2, corr1a);
    const float c2sum = (c2raw == 0.0f) ? 1.0f : c2raw;
    N.inv2 = __fdiv_rn(1.0f, c2sum);
    if (P.order < 3) return N;
    const float corr1b = (c1cnt >= 2u) ? __fmul_rn(__fmul_rn(P.ipn_b, N.w1k), c1f) : 0.0f;
    const float t2 = (c2raw == 0.0f) ? 0.0f : __fmul_rn(c2raw, N.inv2);
    const float c3raw = __fsub_rn(__fsub_rn((float)s3, corr1b), __fmul_rn(P.ipn_c, t2));
    const float c3sum = (c3raw == 0.0f) ? 1.0f : c3raw;
    N.inv3 = __fdiv_rn(1.0f, c3sum);
    return N;
}

// ... and the part that depends on the record
__device__ __forceinline__ EntryWeights record_weights(Record rec, const NodeWeights& N, const WeightParams& P) {
    EntryWeights W;
    const bool has1 = (rec.x >> 31) != 0u;
    const uint32_t C2 = rec.x & 0x7fffffffu, C3 = rec.y;
    const float c2v = P.weighted ? (float)C2 : (C2 ? 1.0f : 0.0f);
    const float c3v = P.weighted ? (float)C3 : (C3 ? 1.0f : 0.0f);
    const float h1 = has1 ? N.w1k : 0.0f;
    W.in1 = has1;
    W.w1 = h1;
    W.in2 = has1 || (P.order >= 2 && C2 != 0u);
    W.in3 = W.in2 || (P.order >= 3 && C3 != 0u);
    W.w2 = 0.0f;
    W.w3 = 0.0f;
    if (P.variant == 7) {
        W.in2 = (P.order >= 2 && C2 != 0u);
        W.w2 = c2v;
        W.in3 = false;
        return W;
    }
    if (P.order < 2) return W;
    const float v2 = __fsub_rn(c2v, __fmul_rn(P.ipn_a, h1));
    W.w2 = W.in2 ? __fmul_rn(v2, N.inv2) : 0.0f;
    if (P.order < 3) return W;
    const float v3 = __fsub_rn(__fsub_rn(c3v, __fmul_rn(P.ipn_b, h1)), __fmul_rn(P.ipn_c, W.w2));
    W.w3 = W.in3 ? __fmul_rn(v3, N.inv3) : 0.0f;
    return W;
}

__device__ __forceinline__ EntryWeights entry_weights(Record rec, uint32_t c1cnt, unsigned long long s2,
                                                      unsigned long long s3, const WeightParams& P) {
    return record_weights(rec, node_weights(c1cnt, s2, s3, P), P);
}

__device__ __forceinline__ WeightParams make_params(int order, int weighted, int variant, float fill,
                                                    const float* __restrict__ ip, const float* __restrict__ bscal) {
    WeightParams P;
    P.order = order;
    P.weighted = weighted;
    P.variant = variant;
    P.fill = fill;
    const float scale = bscal ? bscal[0] : 0.0f;
    // ip == NULL: every batch carries its own coefficients in its scalar slots 5..7 (a training step whose sub-batches run
    // as one session: sub-batch u sees the running mean after u updates, model.py:2245-2248)
    const float a = ip ? ip[0] : (bscal ? bscal[5] : 0.0f), b = ip ? ip[1] : (bscal ? bscal[6] : 0.0f),
                c = ip ? ip[2] : (bscal ? bscal[7] : 0.0f);
    P.ipn_a = scale > 0.0f ? __fdiv_rn(a, scale) : a;
    P.ipn_b = scale > 0.0f ? __fdiv_rn(b, scale) : b;
    P.ipn_c = scale > 0.0f ? __fdiv_rn(c, scale) : c;
    return P;
}

__device__ __forceinline__ void load_colstat(const ColStat* cs, uint32_t& c1, unsigned long long& s2,
                                             unsigned long long& s3) {
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(cs));
    c1 = a.x;
    s2 = ((unsigned long long)a.w << 32) | a.z;
    const uint2 b = __ldcg(reinterpret_cast<const uint2*>(cs) + 2);
    s3 = ((unsigned long long)b.y << 32) | b.x;
}

}  // namespace ocn
