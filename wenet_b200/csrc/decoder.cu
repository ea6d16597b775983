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

constexpr int GL_THREADS = 256;
__global__ void __launch_bounds__(GL_THREADS)
gather_logprob_kernel(const float* __restrict__ logits, long long ldl, int V, const int* __restrict__ target,
                      float* __restrict__ tok_logp) {
    __shared__ float s_red[GL_THREADS / 32];
    __shared__ float s_b;
    const long long row = blockIdx.x;
    const float* g = logits + row * ldl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < V; i += GL_THREADS) mx = fmaxf(mx, g[i]);
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = s_red[0];
        for (int w = 1; w < GL_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
        s_b = m;
    }
    __syncthreads();
    mx = s_b;
    float sum = 0.f;
    for (int i = threadIdx.x; i < V; i += GL_THREADS) sum += expf(g[i] - mx);
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < GL_THREADS / 32; ++w) s += s_red[w];
        const int tg = target[row];
        tok_logp[row] = (tg >= 0 && tg < V) ? (g[tg] - (mx + logf(s))) : 0.f;
    }
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

// one thread per utterance (<= 16 hyps x ~100 tokens of sequential fp32 adds, as the reference does)
__global__ void rescore_kernel(RsDev P, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const int h0 = P.utt_hyp0[b], nh = P.utt_nhyp[b];
    float best_score = -INFINITY;
    int best_i = 0;
    for (int i = 0; i < nh; ++i) {
        const int hy = h0 + i;
        const int r0 = P.hyp_row0[hy], ln = P.hyp_len[hy];
        float score = 0.f;
        for (int j = 0; j <= ln; ++j) score += P.l2r[r0 + j];  // tokens then <eos>
        if (P.reverse_weight > 0.f && P.r2l != nullptr) {
            float r_score = 0.f;
            // r_decoder_out[i][len-j-1][hyp[j]] for j = 0..len-1 (search.py:438-441), then eos
            for (int j = 0; j < ln; ++j) r_score += P.r2l[r0 + (ln - j - 1)];
            r_score += P.r2l[r0 + ln];
            score = score * (1.f - P.reverse_weight) + r_score * P.reverse_weight;
        }
        score += (float)(P.ctc_score[hy] * (double)P.ctc_weight);
        P.hyp_score[hy] = score;
        if (score > best_score) {
            best_score = score;
            best_i = i;
        }
    }
    P.best[b] = best_i;
}

}  // namespace

int embed_tokens(const int* tokens, const int* pos, int R, int d, const float* emb, const float* pe, float xscale,
                 float* x, cudaStream_t stream) {
    if (R <= 0) return WB_OK;
    WB_REQUIRE(d % 4 == 0, WB_ERR_BAD_ARG, "embed: d %% 4");
    const long long n = (long long)R * (d / 4);
    const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    ProfScope _ps(PT_EMBED, stream, (double)R * d * 12.0);
    embed_kernel<<<grid, 256, 0, stream>>>(tokens, pos, R, d, emb, pe, xscale, x);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int gather_logprob(const float* logits, long long ldl, int R, int V, const int* target, float* tok_logp,
                   cudaStream_t stream) {
    if (R <= 0) return WB_OK;
    ProfScope _ps(PT_GATHER_LOGPROB, stream, (double)R * V * 4.0);
    gather_logprob_kernel<<<R, GL_THREADS, 0, stream>>>(logits, ldl, V, target, tok_logp);
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
    rescore_kernel<<<ceil_div(a.batch, 64), 64, 0, stream>>>(P, a.batch);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
