// Rescoring-decoder helpers: token embedding + sinusoidal PE, fused log-softmax + target gather
// (never materialises (hyps, L, V) log-probs), and the per-utterance score combine.
// Replaces wenet/models/transformer/decoder.py:98-103,186 (embed), embedding.py:61-78 (x*sqrt(d)+pe),
// asr_model.py:541-546 (log_softmax) + search.py:421-452 (token gather, l2r/r2l mix, ctc weight, argmax).
#include "common.cuh"
#include "kernels.h"

namespace wb {

namespace {

__global__ void embed_kernel(const int* __restrict__ tokens, const int* __restrict__ pos, int R, int d,
                             const float* __restrict__ emb, const float* __restrict__ pe, float xscale,
                             float* __restrict__ x) {
    pdl_launch_dependents();
    pdl_wait();
    const int dv = d / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)R * dv;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dv;
        const int c = (int)(i - r * dv);
        const float4 e = reinterpret_cast<const float4*>(emb + (long long)tokens[r] * d)[c];
        const float4 p = reinterpret_cast<const float4*>(pe + (long long)pos[r] * d)[c];
        reinterpret_cast<float4*>(x + r * d)[c] =
            make_float4(fmaf(e.x, xscale, p.x), fmaf(e.y, xscale, p.y), fmaf(e.z, xscale, p.z), fmaf(e.w, xscale, p.w));
    }
}

// warp per row: combine the partial (max, sum) pairs of gemm_lse_partials into the log-sum-exp of the row and subtract it
// from the target logit, recomputed here as a 1 x d dot product (a[r] . W[target] + bias) - the [R, V] logits are never
// materialised (1.5 GB per decoder direction at 64 x 30 s x 10 hypotheses).
__global__ void __launch_bounds__(256)
lse_target_kernel(const float2* __restrict__ part, int n_parts, const __nv_bfloat16* __restrict__ a, long long lda,
                  const __nv_bfloat16* __restrict__ w, int d, const float* __restrict__ bias,
                  const int* __restrict__ target, const int* __restrict__ row_map, int R, int V,
                  float* __restrict__ tok_logp) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const int src = row_map ? row_map[row] : row;   // decoder state row that produced this position (prefix sharing)
    const int tg = target[row];
    if (tg < 0 || tg >= V) {
        if (lane == 0) tok_logp[row] = 0.f;
        return;
    }
    const float2* pr = part + (long long)src * n_parts;
    float m = -INFINITY;
    for (int i = lane; i < n_parts; i += 32) m = fmaxf(m, pr[i].x);
    m = warp_max(m);
    float ssum = 0.f;
    for (int i = lane; i < n_parts; i += 32) {
        const float2 q = pr[i];
        if (q.y > 0.f) ssum += q.y * fast_exp2(q.x - m);
    }
    ssum = warp_sum(ssum);
    const float lse = (m + log2f(ssum)) * 0.6931471805599453f;   // back from the log2 domain
    // target logit: bf16 x bf16 products, fp32 accumulation (the tensor core's contract)
    const __nv_bfloat16* ar = a + (long long)src * lda;
    const __nv_bfloat16* wr = w + (long long)tg * d;
    float dot = 0.f;
    for (int c = lane * 8; c < d; c += 256) {
        const uint4 av = *reinterpret_cast<const uint4*>(ar + c);
        const uint4 wv = *reinterpret_cast<const uint4*>(wr + c);
        dot = fmaf(bf16_lo(av.x), bf16_lo(wv.x), dot);
        dot = fmaf(bf16_hi(av.x), bf16_hi(wv.x), dot);
        dot = fmaf(bf16_lo(av.y), bf16_lo(wv.y), dot);
        dot = fmaf(bf16_hi(av.y), bf16_hi(wv.y), dot);
        dot = fmaf(bf16_lo(av.z), bf16_lo(wv.z), dot);
        dot = fmaf(bf16_hi(av.z), bf16_hi(wv.z), dot);
        dot = fmaf(bf16_lo(av.w), bf16_lo(wv.w), dot);
        dot = fmaf(bf16_hi(av.w), bf16_hi(wv.w), dot);
    }
    dot = warp_sum(dot);
    if (lane == 0) tok_logp[row] = (dot + (bias ? bias[tg] : 0.f)) - lse;
}

struct RsDev {
    const float* l2r;
    const float* r2l;
    const int* hyp_row0;
    const int* hyp_len;
    const int* utt_hyp0;
    const int* utt_nhyp;
    const double* ctc_score;
    float ctc_weight, reverse_weight;
    float* hyp_score;
    int* best;
};

// Score combine of search.py:421-452.  The reference adds the token log-probs of a hypothesis one by one in fp32
// (score += ...), so the ORDER is kept: a warp per hypothesis loads 32 values at a time in parallel and then every
// lane replays the same sequential chain of additions through shuffles (a serial thread per utterance spent 340 us
// in dependent global loads).  One CTA per utterance; the first maximum wins (strict >, search.py:448).
constexpr int RS_WARPS = 8;

__device__ __forceinline__ float seq_sum(const float* __restrict__ src, int n, bool reversed_head, int ln, int lane) {
    // sum_{j < n} src[idx(j)], idx(j) = j, or for the r2l branch (ln - 1 - j) for j < ln and ln for j == ln
    float acc = 0.f;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        float v = 0.f;
        if (j < n) v = src[reversed_head ? ((j < ln) ? (ln - 1 - j) : ln) : j];
        const int m = min(32, n - j0);
        for (int i = 0; i < m; ++i) acc += __shfl_sync(0xffffffffu, v, i);
    }
    return acc;
}

__global__ void __launch_bounds__(RS_WARPS * 32) rescore_kernel(RsDev P, int batch) {
    __shared__ float s_score[64];
    const int b = blockIdx.x;
    if (b >= batch) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h0 = P.utt_hyp0[b], nh = P.utt_nhyp[b];
    for (int base = 0; base < nh; base += 64) {       // 64 hypotheses per round (beam sizes beyond that loop)
        const int cnt = min(64, nh - base);
        for (int i = warp; i < cnt; i += RS_WARPS) {
            const int hy = h0 + base + i;
            const int r0 = P.hyp_row0[hy], ln = P.hyp_len[hy];
            float score = seq_sum(P.l2r + r0, ln + 1, false, ln, lane);   // tokens then <eos>
            if (P.reverse_weight > 0.f && P.r2l != nullptr) {
                // r_decoder_out[i][len-j-1][hyp[j]] for j = 0..len-1 (search.py:438-441), then eos
                const float r_score = seq_sum(P.r2l + r0, ln + 1, true, ln, lane);
                score = score * (1.f - P.reverse_weight) + r_score * P.reverse_weight;
            }
            score += (float)(P.ctc_score[hy] * (double)P.ctc_weight);
            if (lane == 0) {
                P.hyp_score[hy] = score;
                s_score[i] = score;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float best_score = (base == 0) ? -INFINITY : P.hyp_score[h0 + P.best[b]];
            int best_i = (base == 0) ? 0 : P.best[b];
            for (int i = 0; i < cnt; ++i)
                if (s_score[i] > best_score) {
                    best_score = s_score[i];
                    best_i = base + i;
                }
            P.best[b] = best_i;
        }
        __syncthreads();
    }
    if (nh == 0 && threadIdx.x == 0) P.best[b] = 0;
}

}  // namespace

int embed_tokens(const int* tokens, const int* pos, int R, int d, const float* emb, const float* pe, float xscale,
                 float* x, cudaStream_t stream) {
    if (R <= 0) return WB_OK;
    WB_REQUIRE(d % 4 == 0, WB_ERR_BAD_ARG, "embed: d %% 4");
    const long long n = (long long)R * (d / 4);
    const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    ProfScope _ps(PT_EMBED, stream, (double)R * d * 12.0);
    WB_CHECK_CUDA(launch_maybe_pdl(embed_kernel, dim3(grid), dim3(256), 0, stream, tokens, pos, R, d, emb, pe, xscale, x));
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int lse_target_logprob(const float2* part, int n_parts, const void* a_bf16, long long lda, const void* w_bf16, int d,
                       const float* bias, const int* target, const int* row_map, int R, int V, float* tok_logp,
                       cudaStream_t stream) {
    if (R <= 0) return WB_OK;
    WB_REQUIRE(d % 8 == 0 && lda % 8 == 0, WB_ERR_BAD_ARG, "lse_target_logprob: d / lda must be multiples of 8");
    ProfScope _ps(PT_GATHER_LOGPROB, stream, (double)R * (n_parts * 8.0 + d * 4.0));
    lse_target_kernel<<<ceil_div(R, 8), 256, 0, stream>>>(part, n_parts, reinterpret_cast<const __nv_bfloat16*>(a_bf16), lda,
                                                          reinterpret_cast<const __nv_bfloat16*>(w_bf16), d, bias, target,
                                                          row_map, R, V, tok_logp);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int rescore_combine(const RescoreArgs& a, cudaStream_t stream) {
    if (a.batch <= 0) return WB_OK;
    RsDev P;
    P.l2r = a.l2r;
    P.r2l = a.r2l;
    P.hyp_row0 = a.hyp_row0;
    P.hyp_len = a.hyp_len;
    P.utt_hyp0 = a.utt_hyp0;
    P.utt_nhyp = a.utt_nhyp;
    P.ctc_score = a.ctc_score;
    P.ctc_weight = a.ctc_weight;
    P.reverse_weight = a.reverse_weight;
    P.hyp_score = a.hyp_score;
    P.best = a.best;
    ProfScope _ps(PT_RESCORE, stream, 0.0);
    rescore_kernel<<<a.batch, RS_WARPS * 32, 0, stream>>>(P, a.batch);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
