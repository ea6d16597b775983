// CTC head tail: fp32 log-softmax over the vocabulary fused with the per-frame top-k that both
// searches consume, and the greedy collapse.
// Replaces wenet/models/transformer/ctc.py:73-81 (log_softmax), asr_model.py:254-265 (blank penalty),
// search.py:158 (logp.topk(beam_size) per frame) and search.py:109-124 + ctc_utils.py:23-33 (greedy).
//
// HBM-bound: one read + one write of the [frames, V] fp32 matrix (12.7 MB per 30 s utterance at
// V = 4233); the row lives in shared memory for max / sum-exp / k rounds of block arg-max.
#include "common.cuh"
#include "kernels.h"

namespace wb {

namespace {

constexpr int LS_THREADS = 256;

__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
    }
}

__global__ void __launch_bounds__(LS_THREADS)
logsoftmax_topk_kernel(float* __restrict__ logits, long long ldl, int V, int blank_id, float blank_penalty,
                       int topk, float* __restrict__ topk_val, int* __restrict__ topk_idx) {
    extern __shared__ float s_row[];  // [V]
    __shared__ float s_redv[LS_THREADS / 32];
    __shared__ int s_redi[LS_THREADS / 32];
    __shared__ float s_bcast;
    __shared__ int s_bcasti;
    const long long row = blockIdx.x;
    float* g = logits + row * ldl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    float mx = -INFINITY;
    for (int i = threadIdx.x; i < V; i += LS_THREADS) {
        float v = g[i];
        if (i == blank_id) v -= blank_penalty;
        s_row[i] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if (lane == 0) s_redv[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = s_redv[0];
        for (int w = 1; w < LS_THREADS / 32; ++w) m = fmaxf(m, s_redv[w]);
        s_bcast = m;
    }
    __syncthreads();
    mx = s_bcast;
    float sum = 0.f;
    for (int i = threadIdx.x; i < V; i += LS_THREADS) sum += expf(s_row[i] - mx);
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) s_redv[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < LS_THREADS / 32; ++w) s += s_redv[w];
        s_bcast = mx + logf(s);
    }
    __syncthreads();
    const float lse = s_bcast;
    for (int i = threadIdx.x; i < V; i += LS_THREADS) {
        const float lp = s_row[i] - lse;
        s_row[i] = lp;
        g[i] = lp;
    }
    __syncthreads();
    // k rounds of block arg-max (ties -> lowest index), winner removed each round
    for (int k = 0; k < topk; ++k) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = threadIdx.x; i < V; i += LS_THREADS) argmax_combine(bv, bi, s_row[i], i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            argmax_combine(bv, bi, ov, oi);
        }
        if (lane == 0) {
            s_redv[warp] = bv;
            s_redi[warp] = bi;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = s_redv[0];
            int ix = s_redi[0];
            for (int w = 1; w < LS_THREADS / 32; ++w) argmax_combine(v, ix, s_redv[w], s_redi[w]);
            s_bcast = v;
            s_bcasti = ix;
            topk_val[row * topk + k] = v;
            topk_idx[row * topk + k] = ix;
            if (ix >= 0 && ix < V) s_row[ix] = -INFINITY;
        }
        __syncthreads();
    }
}

// one warp per sequence
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

int ctc_logsoftmax_topk(float* logits, long long ldl, int M, int V, int blank_id, float blank_penalty, int topk,
                        float* topk_val, int* topk_idx, cudaStream_t stream) {
    if (M <= 0) return WB_OK;
    WB_REQUIRE(topk >= 0 && topk <= V, WB_ERR_BAD_ARG, "logsoftmax_topk: bad k=%d (V=%d)", topk, V);
    const size_t smem = (size_t)V * sizeof(float);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        WB_REQUIRE(smem <= 200 * 1024, WB_ERR_UNSUPPORTED, "logsoftmax_topk: V=%d too large", V);
        WB_CHECK_CUDA(cudaFuncSetAttribute(logsoftmax_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    ProfScope _ps(PT_LOGSOFTMAX_TOPK, stream, (double)M * V * 8.0);
    logsoftmax_topk_kernel<<<M, LS_THREADS, smem, stream>>>(logits, ldl, V, blank_id, blank_penalty, topk, topk_val,
                                                            topk_idx);
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
