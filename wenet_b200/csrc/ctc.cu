// CTC head tail: fp32 log-softmax over the vocabulary fused with the per-frame top-k that both
// searches consume, and the greedy collapse.
// Replaces wenet/models/transformer/ctc.py:73-81 (log_softmax), asr_model.py:254-265 (blank penalty),
// search.py:158 (logp.topk(beam_size) per frame) and search.py:109-124 + ctc_utils.py:23-33 (greedy).
//
// HBM-bound: one read (+ one write when the normalised matrix is requested) of the [frames, V] fp32 matrix
// (12.7 MB per 30 s utterance at V = 4233).
#include "common.cuh"
#include "kernels.h"

namespace wb {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// Log-softmax + per-frame top-k, one warp per row, the row is read twice (HBM, then L2).  The normalised matrix is
// written back only when out_logp != nullptr (ctc_logprobs API parity); decode() passes nullptr because the searches
// consume nothing but the per-frame top-k (search.py:158).
//   pass 1: online log-sum-exp (5 MUFU per 4 elements) + the two largest elements of every lane;
//           tau = k-th largest of those 64 values  =>  every true top-k element is >= tau;
//   pass 2: elements >= tau are compacted into a per-warp candidate list (ballot / popc), from which k rounds of warp
//           arg-max pick the winners in (value desc, index asc) order - torch.topk's order on distinct values.
// Rows with more than LT_CAP candidates (flat posteriors) take k ordered re-scans of the row instead.
constexpr int LT_WARPS = 4;
constexpr int LT_CAP = 192;

// (v, i) precedes (pv, pi) in the output order
__device__ __forceinline__ bool lt_before(float v, int i, float pv, int pi) { return v > pv || (v == pv && i < pi); }

__device__ __forceinline__ void lt_warp_best(float& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (lt_before(ov, oi, v, i)) {
            v = ov;
            i = oi;
        }
    }
}

__global__ void __launch_bounds__(LT_WARPS * 32)
lse_topk_kernel(const float* logits, long long ldl, int M, int V_all, int blank_id, float blank_penalty, int topk,
                float* __restrict__ topk_val, int* __restrict__ topk_idx, float* out_logp /* may alias logits */, int slices,
                int slice_len, float2* __restrict__ part_ml) {
    __shared__ float s_cv[LT_WARPS][LT_CAP];
    __shared__ int s_ci[LT_WARPS][LT_CAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // sliced mode (slices > 1; few rows over a huge vocabulary: autoregressive decoding): a warp takes ONE slice of a row and
    // leaves the raw (un-normalised) top-k of its slice plus the slice's (max, sum of exp) - lse_topk_merge_kernel
    // finishes the row.  Row id = row * slices + slice; indices are global.
    const long long row_id = (long long)blockIdx.x * LT_WARPS + warp;
    if (row_id >= (long long)M * slices) return;
    const long long row = (slices > 1) ? row_id / slices : row_id;
    const int col0 = (slices > 1) ? (int)(row_id - row * slices) * slice_len : 0;
    const int V = (slices > 1) ? min(slice_len, V_all - col0) : V_all;
    blank_id -= col0;
    const float* g = logits + row * ldl + col0;
    const int n4 = (V + 3) >> 2;
    constexpr float kLog2e = 1.4426950408889634f;
    auto load4_from = [&](const float* src, float pen, int i4, float (&v)[4]) {
        const float4 x = *reinterpret_cast<const float4*>(src + 4 * i4);   // ldl >= round_up(V, 4): in bounds
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = 4 * i4 + e;
            if (idx == blank_id) v[e] -= pen;
            if (idx >= V) v[e] = -INFINITY;
        }
    };
    auto load4 = [&](int i4, float (&v)[4]) { load4_from(g, blank_penalty, i4, v); };

    // ---- pass 1 ----
    float m = -INFINITY, ssum = 0.f;
    float v1 = -INFINITY, v2 = -INFINITY;
    for (int i4 = lane; i4 < n4; i4 += 32) {
        float v[4];
        load4(i4, v);
        const float gm = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
        const float mn = fmaxf(m, gm);
        if (mn > -INFINITY) {
            float acc = ssum * fast_exp2((m - mn) * kLog2e);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc += fast_exp2((v[e] - mn) * kLog2e);
            ssum = acc;
            m = mn;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (v[e] > v2) {          // only the VALUES of the two largest elements are needed for tau
                if (v[e] > v1) {
                    v2 = v1;
                    v1 = v[e];
                } else {
                    v2 = v[e];
                }
            }
        }
    }
    float mx = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float part = (m > -INFINITY) ? ssum * fast_exp2((m - mx) * kLog2e) : 0.f;
    part = warp_sum(part);
    const float lse = (slices > 1) ? 0.f : mx + logf(part);   // sliced: raw values leave the kernel
    if (slices > 1 && lane == 0) part_ml[row_id] = make_float2(mx, part);

    // tau: k-th largest of the 64 lane maxima
    float tau = -INFINITY;
    {
        float h1 = v1, h2 = v2;
        for (int r = 0; r < topk; ++r) {
            float bv = h1;
            int bl = lane;
            lt_warp_best(bv, bl);
            tau = bv;
            if (bl == lane) {
                h1 = h2;
                h2 = -INFINITY;
            }
        }
    }

    // ---- pass 2: candidates >= tau ----
    int count = 0;
    for (int i0 = 0; i0 < n4; i0 += 32) {
        const int i4 = i0 + lane;
        float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (i4 < n4) {
            load4(i4, v);
            if (out_logp != nullptr) {
                float* o = out_logp + row * ldl + 4 * i4;
                if (4 * i4 + 3 < V) {
                    *reinterpret_cast<float4*>(o) = make_float4(v[0] - lse, v[1] - lse, v[2] - lse, v[3] - lse);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (4 * i4 + e < V) o[e] = v[e] - lse;
                }
            }
        }
        if (topk == 0) continue;   // warp-uniform
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool hit = (i4 < n4) && (4 * i4 + e < V) && (v[e] >= tau);
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int pos = count + __popc(bal & ((1u << lane) - 1));
                if (pos < LT_CAP) {
                    s_cv[warp][pos] = v[e];
                    s_ci[warp][pos] = 4 * i4 + e;
                }
            }
            count += __popc(bal);
        }
    }
    __syncwarp();

    // ordered re-scans read the row again: after an in-place write-back it already holds normalised values
    const bool inplace = (out_logp != nullptr);
    const float* g2 = inplace ? out_logp + row * ldl : g;
    const float pen2 = inplace ? 0.f : blank_penalty;
    const float sub2 = inplace ? 0.f : lse;
    float pv = INFINITY;
    int pi = -1;
    for (int r = 0; r < topk; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        if (count <= LT_CAP) {
            for (int c = lane; c < count; c += 32) {
                const float v = s_cv[warp][c];
                const int i = s_ci[warp][c];
                if (lt_before(pv, pi, v, i) && lt_before(v, i, bv, bi)) {   // after the previous winner, best so far
                    bv = v;
                    bi = i;
                }
            }
        } else {
            for (int i4 = lane; i4 < n4; i4 += 32) {
                float v[4];
                load4_from(g2, pen2, i4, v);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 4 * i4 + e;
                    if (i < V && lt_before(pv, pi, v[e], i) && lt_before(v[e], i, bv, bi)) {
                        bv = v[e];
                        bi = i;
                    }
                }
            }
        }
        lt_warp_best(bv, bi);
        if (lane == 0) {
            topk_val[row_id * topk + r] = (count <= LT_CAP) ? bv - lse : bv - sub2;
            topk_idx[row_id * topk + r] = (bi == 0x7fffffff) ? 0 : bi + col0;
        }
        pv = bv;
        pi = bi;
    }
}

// one warp per sequence
// sliced mode, second half: warp per row - combine the slices' (max, sum) into the row's log-sum-exp and pick the top-k of the
// slices' candidates in (value desc, index asc) order
__global__ void lse_topk_merge_kernel(const float* __restrict__ pval, const int* __restrict__ pidx, const float2* __restrict__ part_ml,
                                      int M, int slices, int topk, float* __restrict__ topk_val, int* __restrict__ topk_idx) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    float mx = -INFINITY;
    for (int s2 = lane; s2 < slices; s2 += 32) mx = fmaxf(mx, part_ml[(long long)row * slices + s2].x);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s2 = lane; s2 < slices; s2 += 32) {
        const float2 ml = part_ml[(long long)row * slices + s2];
        if (ml.y > 0.f) sum += ml.y * __expf(ml.x - mx);
    }
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    const int nc = slices * topk;
    const float* cv = pval + (long long)row * nc;
    const int* ci = pidx + (long long)row * nc;
    float pv = INFINITY;
    int pi = -1;
    for (int r = 0; r < topk; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < nc; c += 32) {
            const float v = cv[c];
            const int i = ci[c];
            if (lt_before(pv, pi, v, i) && lt_before(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
        }
        lt_warp_best(bv, bi);
        if (lane == 0) {
            topk_val[(long long)row * topk + r] = bv - lse;
            topk_idx[(long long)row * topk + r] = (bi == 0x7fffffff) ? 0 : bi;
        }
        pv = bv;
        pi = bi;
    }
}

__global__ void greedy_kernel(const int* __restrict__ topk_idx, int topk, const int* __restrict__ seq_start,
                              const int* __restrict__ seq_len, int blank_id, int* __restrict__ out_tokens,
                              int out_stride, int* __restrict__ out_len) {
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    const int s = seq_start[b], n = seq_len[b];
    int count = 0;
    int carry = -1;  // previous frame's id
    for (int t0 = 0; t0 < n; t0 += 32) {
        const int t = t0 + lane;
        const int id = (t < n) ? topk_idx[(long long)(s + t) * topk] : blank_id;
        int prev = __shfl_up_sync(0xffffffffu, id, 1);
        if (lane == 0) prev = carry;
        const bool keep = (t < n) && (id != blank_id) && (id != prev);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) out_tokens[(long long)b * out_stride + count + __popc(m & ((1u << lane) - 1))] = id;
        count += __popc(m);
        carry = __shfl_sync(0xffffffffu, id, 31);
    }
    if (lane == 0) out_len[b] = count;
}

}  // namespace

static int launch_lse_topk(const float* logits, long long ldl, int M, int V, int blank_id, float blank_penalty, int topk,
                           float* topk_val, int* topk_idx, float* out_logp, cudaStream_t stream) {
    if (M <= 0) return WB_OK;
    WB_REQUIRE(topk >= 0 && topk <= V, WB_ERR_BAD_ARG, "logsoftmax_topk: bad k=%d (V=%d)", topk, V);
    WB_REQUIRE(ldl % 4 == 0 && ldl >= (V + 3) / 4 * 4 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0, WB_ERR_BAD_ARG,
               "logsoftmax_topk: rows must be 16-byte aligned with pitch >= round_up(V, 4) (pitch %lld, V %d)", ldl, V);
    WB_REQUIRE(topk == 0 || (topk_val && topk_idx), WB_ERR_BAD_ARG, "logsoftmax_topk: null top-k output");
    ProfScope _ps(PT_LOGSOFTMAX_TOPK, stream, (double)M * V * (out_logp ? 8.0 : 4.0));
    lse_topk_kernel<<<ceil_div(M, LT_WARPS), LT_WARPS * 32, 0, stream>>>(logits, ldl, M, V, blank_id, blank_penalty, topk,
                                                                         topk_val, topk_idx, out_logp, 1, 0, nullptr);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int ctc_logsoftmax_topk(float* logits, long long ldl, int M, int V, int blank_id, float blank_penalty, int topk,
                        float* topk_val, int* topk_idx, cudaStream_t stream) {
    return launch_lse_topk(logits, ldl, M, V, blank_id, blank_penalty, topk, topk_val, topk_idx, logits, stream);
}

int ctc_lse_topk(const float* logits, long long ldl, int M, int V, int blank_id, float blank_penalty, int topk,
                 float* topk_val, int* topk_idx, cudaStream_t stream) {
    return launch_lse_topk(logits, ldl, M, V, blank_id, blank_penalty, topk, topk_val, topk_idx, nullptr, stream);
}

// Few rows, huge vocabulary (the output layer of autoregressive decoding: 320 rows x 51 866 columns): every row is cut into
// `slices` pieces handled by different warps, then merged.  scratch: M * slices * topk (float + int) + M * slices float2.
size_t lse_topk_sliced_scratch_bytes(int M, int slices, int topk) {
    return (size_t)M * slices * topk * 8 + (size_t)M * slices * 8 + 256;
}
int lse_topk_sliced(const float* logits, long long ldl, int M, int V, int topk, int slices, float* topk_val, int* topk_idx,
                    void* scratch, cudaStream_t stream) {
    if (M <= 0) return WB_OK;
    WB_REQUIRE(topk >= 1 && slices >= 2 && slices <= 64 && scratch != nullptr, WB_ERR_BAD_ARG, "lse_topk_sliced: bad argument");
    const int slice_len = ceil_div(ceil_div(V, slices), 4) * 4;
    WB_REQUIRE((long long)(slices - 1) * slice_len < V && topk <= slice_len - 3, WB_ERR_BAD_ARG, "lse_topk_sliced: %d slices of %d",
               slices, slice_len);
    WB_REQUIRE(ldl % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0, WB_ERR_BAD_ARG, "lse_topk_sliced: alignment");
    float* pval = reinterpret_cast<float*>(scratch);
    int* pidx = reinterpret_cast<int*>(pval + (size_t)M * slices * topk);
    float2* pml = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(scratch) + (((size_t)M * slices * topk * 8 + 15) / 16) * 16);
    ProfScope _ps(PT_LOGSOFTMAX_TOPK, stream, (double)M * V * 4.0);
    lse_topk_kernel<<<ceil_div(M * slices, LT_WARPS), LT_WARPS * 32, 0, stream>>>(logits, ldl, M, V, -1, 0.f, topk, pval, pidx, nullptr,
                                                                                  slices, slice_len, pml);
    count_launch();
    WB_CHECK_LAUNCH();
    lse_topk_merge_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(pval, pidx, pml, M, slices, topk, topk_val, topk_idx);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int ctc_greedy(const int* topk_idx, int topk, const int* seq_start, const int* seq_len, int batch, int blank_id,
               int* out_tokens, int out_stride, int* out_len, cudaStream_t stream) {
    if (batch <= 0) return WB_OK;
    ProfScope _ps(PT_GREEDY, stream, 0.0);
    greedy_kernel<<<batch, 32, 0, stream>>>(topk_idx, topk, seq_start, seq_len, blank_id, out_tokens, out_stride,
                                            out_len);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
