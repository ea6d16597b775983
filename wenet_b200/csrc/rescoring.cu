// Host-side orchestration of attention rescoring: (Bi)TransformerDecoder over ALL hypotheses of ALL
// utterances in one varlen batch.  Mirrors wenet/models/transformer/search.py:374-458
// (attention_rescoring), asr_model.py:453-547 (forward_attention_decoder), decoder.py:146-201,430-463
// and decoder_layer.py:68-153 — with the cross-attention K/V of the encoder memory projected once per
// utterance instead of once per hypothesis (the reference repeats encoder_out N times, asr_model.py:478)
// and the output layer fused with log-softmax + target gather (no (N, L, V) tensor).
#include "model.h"
#include <math.h>
#include <vector>

namespace wb {

namespace {

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct RsPlan {
    int batch = 0, n_hyp = 0;
    long long R = 0;  // decoder rows
    int max_hyp_rows = 0, max_utt_rows = 0, max_enc_len = 0;
    size_t o_int = 0, o_dbl = 0, o_x = 0, o_a = 0, o_qkv = 0, o_ctx = 0, o_q = 0, o_h = 0, o_memkv = 0, o_logits = 0,
           o_qkv_full = 0, o_ctx_full = 0, total = 0;
    long long ldl = 0;
    size_t n_int = 0;
    // prefix sharing (dedup): decoder rows that see the same input prefix (same utterance) are computed once
    bool dedup = false;
    int U[2] = {0, 0};           // unique rows, l2r / r2l
    int max_utt_u[2] = {0, 0};   // largest per-utterance unique-row count
};

void rs_layout(const Model* m, long long enc_rows, long long R, int batch, int n_hyp, RsPlan* P) {
    const int d = m->cfg.d_model, ff = m->cfg.dec_ffn_dim;
    P->ldl = (m->cfg.vocab + 7) / 8 * 8;
    P->n_int = (size_t)13 * R + (size_t)4 * n_hyp + (size_t)10 * batch + 64;
    size_t o = 0;
    P->o_int = o; o += align_up(P->n_int * 4);
    P->o_dbl = o; o += align_up((size_t)n_hyp * 8 + 64);
    P->o_x = o; o += align_up((size_t)R * d * 4);
    P->o_a = o; o += align_up((size_t)R * d * 2);
    P->o_qkv = o; o += align_up((size_t)R * 3 * d * 2);
    P->o_ctx = o; o += align_up((size_t)R * d * 2);
    P->o_q = o; o += align_up((size_t)R * d * 2);
    P->o_h = o; o += align_up((size_t)R * ff * 2);
    P->o_memkv = o; o += align_up((size_t)enc_rows * 2 * d * 2);
    P->o_qkv_full = o; o += align_up((size_t)R * 3 * d * 2);   // self-attention runs on every row of every hypothesis
    P->o_ctx_full = o; o += align_up((size_t)R * d * 2);
    // rescoring never materialises the [R, V] logits: the output GEMM leaves per-tile log-sum-exp partials only
    P->o_logits = o; o += align_up((size_t)R * lse_parts(m->cfg.vocab, m->cfg.d_model) * sizeof(float2));
    P->total = o + 256;
}

#define RC(x)                         \
    do {                              \
        int _rc = (x);                \
        if (_rc != WB_OK) return _rc; \
    } while (0)

struct RsDevPtrs {
    int *tok_l2r, *tok_r2l, *pos, *tgt_l2r, *tgt_r2l, *hyp_row0, *hyp_rows, *hyp_len, *hyp_src, *utt_q0, *utt_qn,
        *utt_hyp0, *utt_nhyp, *enc_start, *enc_len;
    // dedup maps per direction (0 = l2r, 1 = r2l): unique row of every row, a representative row of every unique row,
    // decoder inputs of the unique rows, and the unique-row range of every utterance
    int *uniq[2], *rep[2], *tok_u[2], *pos_u[2], *utt_q0_u[2], *utt_qn_u[2];
    double* ctc;
};

// rows of one decoder direction as the layers see them
struct DirRows {
    int n;                      // rows the row-wise layers run on (unique rows, or all rows without dedup)
    const int* tok;             // [n] input tokens
    const int* pos;             // [n] positions
    const int* utt_q0;          // [batch] row range of every utterance (cross-attention queries)
    const int* utt_qn;
    int max_utt_rows;
    const int* uniq_of_row;     // [R] or null (identity)
    const int* rep_row;         // [n] or null
};

// dst[i] = src[idx[i]] for rows of `vecs` 16-byte vectors; one warp per row
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int* __restrict__ idx, int n, int vecs,
                                   uint4* __restrict__ dst) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint4* s = src + (long long)idx[i] * vecs;
    uint4* d = dst + (long long)i * vecs;
    for (int v = threadIdx.x & 31; v < vecs; v += 32) d[v] = s[v];
}
int gather_rows(const void* src, const int* idx, int n, int row_bytes, void* dst, cudaStream_t st) {
    if (n <= 0) return WB_OK;
    ProfScope _ps(PT_MISC, st, (double)n * row_bytes * 2.0);
    gather_rows_kernel<<<ceil_div(n, 8), 256, 0, st>>>(reinterpret_cast<const uint4*>(src), idx, n, row_bytes / 16,
                                                       reinterpret_cast<uint4*>(dst));
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

// x += A W^T + b, then a = LayerNorm(x): one kernel when the tile holds whole rows (d == 256), else GEMM + LayerNorm
static int dec_resid_then_norm(const void* A, long long lda, const Linear& W, int M, int d, float* x, const Norm& n, float eps,
                               void* a_out, cudaStream_t st) {
    if (W.b != nullptr && gemm_resid_ln_supported(d))
        return gemm_resid_ln(A, lda, &W.tmap, W.w, M, d, W.K, W.b, 1.0f, x, d, nullptr, nullptr, n.g, n.b, eps, a_out, d, st);
    int rc = gemm_bf16(A, lda, &W.tmap, W.w, M, d, W.K, W.b, EPI_RESID_F32, 1.0f, x, d, 0, st);
    if (rc != WB_OK) return rc;
    return layernorm_rows(x, d, M, d, n.g, n.b, eps, a_out, d, 0, nullptr, 0, st);
}

// one direction of the decoder: x = embed(tokens); layers; after_norm; logits = out(x)
// Output: either the raw logits [R][ldl] (logits != null; decoder_logprobs API) or, for rescoring, only
// tok_logp[r] = log_softmax(logits[r])[target[r]] via LSE partials (logits == null).
// With prefix sharing (Q.uniq_of_row != null) every row-wise layer runs on the unique rows only; the causal
// self-attention still sees every row of every hypothesis: Q/K/V are expanded unique -> all rows before it and its
// output is compacted back (two row gathers per layer instead of 2 x the FLOPs of the whole decoder).
int run_decoder(const Model* m, const Decoder& D, const RsPlan& P, const RsDevPtrs& dp, const DirRows& Q,
                const void* enc_bf16, long long enc_rows, uint8_t* ws, float* logits, long long ldl, const int* target,
                float* tok_logp, cudaStream_t st) {
    const wb_model_config& c = m->cfg;
    const int d = c.d_model, ff = c.dec_ffn_dim, H = c.dec_heads;
    const int R = (int)P.R;      // all rows (hypothesis-major)
    const int N = Q.n;           // rows of the row-wise layers
    const bool shared = Q.uniq_of_row != nullptr;
    float* x = reinterpret_cast<float*>(ws + P.o_x);
    void* a = ws + P.o_a;
    void* qkv = ws + P.o_qkv;
    void* ctx = ws + P.o_ctx;
    void* q = ws + P.o_q;
    void* h = ws + P.o_h;
    void* memkv = ws + P.o_memkv;
    void* qkv_full = shared ? ws + P.o_qkv_full : qkv;
    void* ctx_full = shared ? ws + P.o_ctx_full : ctx;
    const float scale = 1.0f / sqrtf(64.0f);
    RC(embed_tokens(Q.tok, Q.pos, N, d, D.emb, D.pe ? D.pe : m->pe, D.xscale, x, st));
    // every LayerNorm after the first rides in the epilogue of the residual GEMM in front of it (decoder_layer.py:101-147:
    // norm2 after the self-attention output projection, norm3 after the cross-attention one, the next layer's norm1 - or
    // after_norm - after the feed-forward w_2)
    if (!D.layers.empty())
        RC(layernorm_rows(x, d, N, d, D.layers[0].n1.g, D.layers[0].n1.b, c.dec_ln_eps, a, d, 0, nullptr, 0, st));
    for (size_t li = 0; li < D.layers.size(); ++li) {
        const DecLayer& L = D.layers[li];
        // masked (causal) self-attention, decoder_layer.py:101-118
        RC(gemm_bf16(a, d, &L.sa_qkv.tmap, L.sa_qkv.w, N, 3 * d, d, L.sa_qkv.b, EPI_BF16, 1.0f, qkv, 3 * d, 0, st));
        if (shared) RC(gather_rows(qkv, Q.uniq_of_row, R, 3 * d * 2, qkv_full, st));
        {
            AttnArgs A;
            A.q = qkv_full; A.ldq = 3 * d; A.q_rows = R; A.q_col0 = 0;
            A.k = qkv_full; A.ldk = 3 * d; A.k_rows = R; A.k_col0 = d;
            A.v = qkv_full; A.ldv = 3 * d; A.v_rows = R; A.v_col0 = 2 * d;
            A.kbias = nullptr; A.ld_kbias = 0;
            A.q_start = dp.hyp_row0; A.q_len = dp.hyp_rows; A.k_start = dp.hyp_row0; A.k_len = dp.hyp_rows;
            A.batch = P.n_hyp; A.heads = H; A.max_q_len = P.max_hyp_rows;
            A.chunk_size = 1; A.num_left_chunks = -1; A.scale = scale;   // causal == chunk size 1
            A.out = ctx_full; A.ldo = d; A.out_col0 = 0; A.split3_out = 0; A.v_mode = 0;
            RC(attention_forward(A, st));
        }
        if (shared) RC(gather_rows(ctx_full, Q.rep_row, N, d * 2, ctx, st));
        RC(dec_resid_then_norm(ctx, d, L.sa_out, N, d, x, L.n2, c.dec_ln_eps, a, st));
        // cross-attention over the utterance's encoder frames, decoder_layer.py:120-139
        RC(gemm_bf16(a, d, &L.ca_q.tmap, L.ca_q.w, N, d, d, L.ca_q.b, EPI_BF16, 1.0f, q, d, 0, st));
        RC(gemm_bf16(enc_bf16, m->cfg.precise ? 3 * d : d /*precise: rows are [hi|lo|hi]; the decoder reads hi*/, &L.ca_kv.tmap, L.ca_kv.w, (int)enc_rows, 2 * d, d, L.ca_kv.b, EPI_BF16, 1.0f, memkv,
                     2 * d, 0, st));
        {
            AttnArgs A;
            A.q = q; A.ldq = d; A.q_rows = N; A.q_col0 = 0;
            A.k = memkv; A.ldk = 2 * d; A.k_rows = enc_rows; A.k_col0 = 0;
            A.v = memkv; A.ldv = 2 * d; A.v_rows = enc_rows; A.v_col0 = d;
            A.kbias = nullptr; A.ld_kbias = 0;
            A.q_start = Q.utt_q0; A.q_len = Q.utt_qn; A.k_start = dp.enc_start; A.k_len = dp.enc_len;
            A.batch = P.batch; A.heads = H; A.max_q_len = Q.max_utt_rows;
            A.chunk_size = 0; A.num_left_chunks = -1; A.scale = scale;
            A.out = ctx; A.ldo = d; A.out_col0 = 0; A.split3_out = 0; A.v_mode = 0;
            RC(attention_forward(A, st));
        }
        RC(dec_resid_then_norm(ctx, d, L.ca_out, N, d, x, L.n3, c.dec_ln_eps, a, st));
        // feed-forward (ReLU / GELU), decoder_layer.py:141-147
        RC(gemm_bf16(a, d, &L.ff1.tmap, L.ff1.w, N, ff, d, L.ff1.b, D.act_epi, 1.0f, h, ff, 0, st));
        const Norm& nn = (li + 1 < D.layers.size()) ? D.layers[li + 1].n1 : D.after;
        RC(dec_resid_then_norm(h, ff, L.ff2, N, d, x, nn, c.dec_ln_eps, a, st));
    }
    if (logits != nullptr) {
        WB_REQUIRE(!shared, WB_ERR_BAD_ARG, "decoder logits are only produced without prefix sharing");
        RC(gemm_bf16(a, d, &D.out.tmap, D.out.w, N, c.vocab, d, D.out.b, EPI_F32, 1.0f, logits, ldl, 0, st));
    } else {
        float2* part = reinterpret_cast<float2*>(ws + P.o_logits);
        RC(gemm_lse_partials(a, d, &D.out.tmap, D.out.w, N, c.vocab, d, D.out.b, part, st));
        RC(lse_target_logprob(part, lse_parts(c.vocab, d), a, d, D.out.w, d, D.out.b, target, Q.uniq_of_row, R, c.vocab,
                              tok_logp, st));
    }
    return WB_OK;
}

// decoder inputs from hypotheses that are still on the device (the prefix beam search output): one warp per hypothesis
//   l2r: in = [sos, y0..y_{n-1}], target = [y0..y_{n-1}, eos];  r2l: the same on the reversed hypothesis
__global__ void build_decoder_inputs_kernel(int n_hyp, const int* __restrict__ hyp_row0, const int* __restrict__ hyp_len,
                                            const int* __restrict__ hyp_src, const int* __restrict__ tokens, int sos,
                                            int eos, int* __restrict__ tok_l2r, int* __restrict__ tok_r2l,
                                            int* __restrict__ tgt_l2r, int* __restrict__ tgt_r2l, int* __restrict__ pos) {
    const int h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= n_hyp) return;
    const int lane = threadIdx.x & 31;
    const int n = hyp_len[h], r = hyp_row0[h];
    const int* y = tokens + hyp_src[h];
    for (int j = lane; j <= n; j += 32) {
        tok_l2r[r + j] = (j == 0) ? sos : y[j - 1];
        tok_r2l[r + j] = (j == 0) ? sos : y[n - j];
        tgt_l2r[r + j] = (j < n) ? y[j] : eos;
        tgt_r2l[r + j] = (j < n) ? y[n - 1 - j] : eos;
        pos[r + j] = j;
    }
}

// Prefix sharing tables of one decoder direction (0 = l2r, 1 = r2l: hypotheses reversed).
// Row (h, j) of a direction sees the input prefix [sos, s_0 .. s_{j-1}] (s = the hypothesis in that direction) and
// the utterance's encoder memory, nothing else: rows of the SAME utterance with equal prefixes are the same
// computation.  Hypothesis h shares rows 0 .. lcp with the earlier hypothesis g* of its utterance that has the
// longest common prefix; those rows point at g*'s unique rows, the rest get new unique rows (hypothesis-major, so
// the unique rows of an utterance are contiguous).  Returns the number of unique rows.
int build_prefix_share(int dir, int batch, const int* utt_hyp0, const int* utt_nhyp, const int* hyp_row0,
                       const int32_t* hyp_len, const int32_t* hyp_tok0, const int32_t* hyp_tokens, int sos, int* uniq,
                       int* rep, int* tok_u, int* pos_u, int* utt_q0_u, int* utt_qn_u, int* max_utt_u) {
    int U = 0;
    auto tokat = [&](int h, int i) {   // i-th token of the sequence hypothesis h presents in this direction
        const int32_t* y = hyp_tokens + hyp_tok0[h];
        return dir == 0 ? y[i] : y[hyp_len[h] - 1 - i];
    };
    *max_utt_u = 0;
    for (int b = 0; b < batch; ++b) {
        const int hb = utt_hyp0[b], he = hb + utt_nhyp[b];
        utt_q0_u[b] = U;
        for (int h = hb; h < he; ++h) {
            const int n = hyp_len[h], r0 = hyp_row0[h];
            int best = -1, best_g = -1;   // rows 0 .. best are shared with hypothesis best_g
            for (int g = hb; g < h; ++g) {
                const int lim = n < hyp_len[g] ? n : hyp_len[g];
                int l = 0;
                while (l < lim && tokat(h, l) == tokat(g, l)) ++l;
                if (l > best) {
                    best = l;
                    best_g = g;
                }
            }
            for (int j = 0; j <= n; ++j) {
                if (best_g >= 0 && j <= best) {
                    uniq[r0 + j] = uniq[hyp_row0[best_g] + j];
                } else {
                    uniq[r0 + j] = U;
                    rep[U] = r0 + j;
                    tok_u[U] = (j == 0) ? sos : tokat(h, j - 1);
                    pos_u[U] = j;
                    ++U;
                }
            }
        }
        utt_qn_u[b] = U - utt_q0_u[b];
        if (utt_qn_u[b] > *max_utt_u) *max_utt_u = utt_qn_u[b];
    }
    return U;
}

// builds the flattened decoder inputs on the host and uploads them (one copy)
int prepare(const Model* m, long long enc_rows, const int32_t* seq_start_host, const int32_t* seq_len_host, int batch,
            int n_hyp, const int32_t* hyp_utt, const int32_t* hyp_len, const int32_t* hyp_tok0,
            const int32_t* hyp_tokens, bool tokens_on_device, bool dedup, const double* ctc_score, int sos, int eos,
            uint8_t* ws, size_t ws_bytes, RsPlan* P, RsDevPtrs* dp, cudaStream_t st) {
    long long R = 0;
    for (int h = 0; h < n_hyp; ++h) R += hyp_len[h] + 1;
    rs_layout(m, enc_rows, R, batch, n_hyp, P);
    P->batch = batch;
    P->n_hyp = n_hyp;
    P->R = R;
    WB_REQUIRE(ws_bytes >= P->total, WB_ERR_WORKSPACE, "rescoring: workspace %zu < required %zu", ws_bytes, P->total);
    WB_REQUIRE(R < 2147483647LL, WB_ERR_UNSUPPORTED, "rescoring: too many decoder rows");
    std::vector<int> buf(P->n_int, 0);
    int* tok_l2r = buf.data();
    int* tok_r2l = tok_l2r + R;
    int* pos = tok_r2l + R;
    int* tgt_l2r = pos + R;
    int* tgt_r2l = tgt_l2r + R;
    int* hyp_row0 = tgt_r2l + R;
    int* hyp_rows = hyp_row0 + n_hyp;
    int* hyp_ln = hyp_rows + n_hyp;
    int* hyp_src = hyp_ln + n_hyp;
    int* utt_q0 = hyp_src + n_hyp;
    int* utt_qn = utt_q0 + batch;
    int* utt_hyp0 = utt_qn + batch;
    int* utt_nhyp = utt_hyp0 + batch;
    int* enc_start = utt_nhyp + batch;
    int* enc_len = enc_start + batch;
    // dedup tables (l2r then r2l): uniq[R], rep[R], tok_u[R], pos_u[R] each, then utt_q0_u[batch], utt_qn_u[batch] each
    int* dd = enc_len + batch;
    int* h_uniq[2] = {dd, dd + 4 * R};
    int* h_rep[2] = {dd + R, dd + 5 * R};
    int* h_tok_u[2] = {dd + 2 * R, dd + 6 * R};
    int* h_pos_u[2] = {dd + 3 * R, dd + 7 * R};
    int* h_q0_u[2] = {dd + 8 * R, dd + 8 * R + 2 * batch};
    int* h_qn_u[2] = {dd + 8 * R + batch, dd + 8 * R + 3 * batch};
    P->dedup = dedup && !tokens_on_device && hyp_tokens != nullptr;
    for (int b = 0; b < batch; ++b) {
        utt_q0[b] = 0;
        utt_qn[b] = 0;
        utt_hyp0[b] = 0;
        utt_nhyp[b] = 0;
        enc_start[b] = seq_start_host[b];
        enc_len[b] = seq_len_host[b];
        if (seq_len_host[b] > P->max_enc_len) P->max_enc_len = seq_len_host[b];
    }
    long long r = 0;
    int prev_utt = -1;
    for (int h = 0; h < n_hyp; ++h) {
        const int b = hyp_utt[h], n = hyp_len[h];
        WB_REQUIRE(b >= 0 && b < batch && b >= prev_utt, WB_ERR_BAD_ARG, "rescoring: hyp_utt must be non-decreasing in [0,batch)");
        if (b != prev_utt) {
            utt_q0[b] = (int)r;
            utt_hyp0[b] = h;
            prev_utt = b;
        }
        utt_nhyp[b] += 1;
        utt_qn[b] += n + 1;
        hyp_row0[h] = (int)r;
        hyp_rows[h] = n + 1;
        hyp_ln[h] = n;
        hyp_src[h] = hyp_tok0[h];
        if (n + 1 > P->max_hyp_rows) P->max_hyp_rows = n + 1;
        const int32_t* y = hyp_tokens + hyp_tok0[h];
        // l2r: in = [sos, y0..y_{n-1}], target = [y0..y_{n-1}, eos]      (common.py:113-156 add_sos_eos)
        // r2l: in = [sos, y_{n-1}..y0], target = [y_{n-1}..y0, eos]      (asr_model.py:487-536)
        for (int j = 0; j <= n && !tokens_on_device; ++j) {
            tok_l2r[r + j] = (j == 0) ? sos : y[j - 1];
            tok_r2l[r + j] = (j == 0) ? sos : y[n - j];
            tgt_l2r[r + j] = (j < n) ? y[j] : eos;
            tgt_r2l[r + j] = (j < n) ? y[n - 1 - j] : eos;
            pos[r + j] = j;
        }
        r += n + 1;
    }
    for (int b = 0; b < batch; ++b)
        if (utt_qn[b] > P->max_utt_rows) P->max_utt_rows = utt_qn[b];
    if (P->dedup) {
        for (int dir = 0; dir < 2; ++dir)
            P->U[dir] = build_prefix_share(dir, batch, utt_hyp0, utt_nhyp, hyp_row0, hyp_len, hyp_tok0, hyp_tokens, sos,
                                           h_uniq[dir], h_rep[dir], h_tok_u[dir], h_pos_u[dir], h_q0_u[dir], h_qn_u[dir],
                                           &P->max_utt_u[dir]);
    }
    WB_REQUIRE(P->max_hyp_rows <= m->cfg.max_pos, WB_ERR_UNSUPPORTED, "hypothesis longer than the positional table");
    WB_CHECK_CUDA(cudaMemcpyAsync(ws + P->o_int, buf.data(), P->n_int * 4, cudaMemcpyHostToDevice, st));
    if (ctc_score)
        WB_CHECK_CUDA(cudaMemcpyAsync(ws + P->o_dbl, ctc_score, (size_t)n_hyp * 8, cudaMemcpyHostToDevice, st));
    WB_CHECK_CUDA(cudaStreamSynchronize(st));  // host staging buffers go out of scope
    int* base = reinterpret_cast<int*>(ws + P->o_int);
    dp->tok_l2r = base;
    dp->tok_r2l = base + R;
    dp->pos = base + 2 * R;
    dp->tgt_l2r = base + 3 * R;
    dp->tgt_r2l = base + 4 * R;
    dp->hyp_row0 = base + 5 * R;
    dp->hyp_rows = dp->hyp_row0 + n_hyp;
    dp->hyp_len = dp->hyp_rows + n_hyp;
    dp->hyp_src = dp->hyp_len + n_hyp;
    dp->utt_q0 = dp->hyp_src + n_hyp;
    dp->utt_qn = dp->utt_q0 + batch;
    dp->utt_hyp0 = dp->utt_qn + batch;
    dp->utt_nhyp = dp->utt_hyp0 + batch;
    dp->enc_start = dp->utt_nhyp + batch;
    dp->enc_len = dp->enc_start + batch;
    {
        int* ddv = dp->enc_len + batch;
        for (int dir = 0; dir < 2; ++dir) {
            dp->uniq[dir] = ddv + (dir ? 4 * R : 0);
            dp->rep[dir] = ddv + (dir ? 5 * R : R);
            dp->tok_u[dir] = ddv + (dir ? 6 * R : 2 * R);
            dp->pos_u[dir] = ddv + (dir ? 7 * R : 3 * R);
            dp->utt_q0_u[dir] = ddv + 8 * R + (dir ? 2 * batch : 0);
            dp->utt_qn_u[dir] = ddv + 8 * R + (dir ? 3 * batch : batch);
        }
    }
    dp->ctc = reinterpret_cast<double*>(ws + P->o_dbl);
    if (tokens_on_device) {
        build_decoder_inputs_kernel<<<ceil_div(n_hyp, 8), 256, 0, st>>>(n_hyp, dp->hyp_row0, dp->hyp_len, dp->hyp_src,
                                                                       hyp_tokens, sos, eos, dp->tok_l2r, dp->tok_r2l,
                                                                       dp->tgt_l2r, dp->tgt_r2l, dp->pos);
        count_launch();
        WB_CHECK_LAUNCH();
    }
    return WB_OK;
}

}  // namespace
}  // namespace wb

using namespace wb;

extern "C" {

size_t wb_rescoring_workspace_bytes(const wb_model* mm, int64_t enc_rows, int64_t total_tokens_plus_hyps) {
    const Model* m = reinterpret_cast<const Model*>(mm);
    if (!m || !m->finalized) return 0;
    RsPlan P;
    // n_hyp and batch only size small integer tables: bound both by the row count
    rs_layout(m, enc_rows, total_tokens_plus_hyps, (int)(total_tokens_plus_hyps < 1 ? 1 : total_tokens_plus_hyps),
              (int)(total_tokens_plus_hyps < 1 ? 1 : total_tokens_plus_hyps), &P);
    return P.total;
}

static int attention_rescoring_impl(const wb_model* mm, const void* enc_out_bf16_dev, int64_t enc_rows,
                                    const int32_t* seq_start_host, const int32_t* seq_len_host, int batch, int n_hyp,
                                    const int32_t* hyp_utt_host, const int32_t* hyp_len_host,
                                    const int32_t* hyp_tok0_host, const int32_t* hyp_tokens, bool tokens_on_device,
                                    const double* ctc_score_host, int sos, int eos, float ctc_weight,
                                    float reverse_weight, float* tok_logp_l2r_dev, float* tok_logp_r2l_dev,
                                    float* hyp_score_dev, int32_t* best_dev, void* workspace_dev,
                                    size_t workspace_bytes, wb_stream_t stream) {
    const Model* m = reinterpret_cast<const Model*>(mm);
    WB_REQUIRE(m && m->finalized, WB_ERR_NOT_LOADED, "attention_rescoring: model not finalized");
    WB_REQUIRE(m->cfg.dec_layers > 0, WB_ERR_UNSUPPORTED, "attention_rescoring: model has no decoder");
    WB_REQUIRE(batch > 0 && n_hyp > 0 && enc_out_bf16_dev && tok_logp_l2r_dev && hyp_score_dev && best_dev &&
                   workspace_dev && ctc_score_host,
               WB_ERR_BAD_ARG, "attention_rescoring: null/empty argument");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace_dev);
    RsPlan P;
    RsDevPtrs dp;
    static const bool dedup_off = getenv("WB_NO_PREFIX_SHARING") != nullptr;
    RC(prepare(m, enc_rows, seq_start_host, seq_len_host, batch, n_hyp, hyp_utt_host, hyp_len_host, hyp_tok0_host,
               hyp_tokens, tokens_on_device, !dedup_off, ctc_score_host, sos, eos, ws, workspace_bytes, &P, &dp, st));
    auto rows_of = [&](int dir) {
        DirRows Q;
        if (P.dedup) {
            Q.n = P.U[dir];
            Q.tok = dp.tok_u[dir];
            Q.pos = dp.pos_u[dir];
            Q.utt_q0 = dp.utt_q0_u[dir];
            Q.utt_qn = dp.utt_qn_u[dir];
            Q.max_utt_rows = P.max_utt_u[dir];
            Q.uniq_of_row = dp.uniq[dir];
            Q.rep_row = dp.rep[dir];
        } else {
            Q.n = (int)P.R;
            Q.tok = dir ? dp.tok_r2l : dp.tok_l2r;
            Q.pos = dp.pos;
            Q.utt_q0 = dp.utt_q0;
            Q.utt_qn = dp.utt_qn;
            Q.max_utt_rows = P.max_utt_rows;
            Q.uniq_of_row = nullptr;
            Q.rep_row = nullptr;
        }
        return Q;
    };
    RC(run_decoder(m, m->left, P, dp, rows_of(0), enc_out_bf16_dev, enc_rows, ws, nullptr, 0, dp.tgt_l2r, tok_logp_l2r_dev,
                   st));
    const bool use_r2l = reverse_weight > 0.f && m->cfg.rdec_layers > 0;
    if (use_r2l) {
        WB_REQUIRE(tok_logp_r2l_dev, WB_ERR_BAD_ARG, "attention_rescoring: r2l output buffer missing");
        RC(run_decoder(m, m->right, P, dp, rows_of(1), enc_out_bf16_dev, enc_rows, ws, nullptr, 0, dp.tgt_r2l,
                       tok_logp_r2l_dev, st));
    }
    RescoreArgs a;
    a.l2r = tok_logp_l2r_dev;
    a.r2l = use_r2l ? tok_logp_r2l_dev : nullptr;
    a.hyp_row0 = dp.hyp_row0;
    a.hyp_len = dp.hyp_len;
    a.utt_hyp0 = dp.utt_hyp0;
    a.utt_nhyp = dp.utt_nhyp;
    a.batch = batch;
    a.ctc_score = dp.ctc;
    a.ctc_weight = ctc_weight;
    a.reverse_weight = use_r2l ? reverse_weight : 0.f;
    a.hyp_score = hyp_score_dev;
    a.best = best_dev;
    return rescore_combine(a, st);
}

int wb_attention_rescoring(const wb_model* mm, const void* enc_out_bf16_dev, int64_t enc_rows,
                           const int32_t* seq_start_host, const int32_t* seq_len_host, int batch, int n_hyp,
                           const int32_t* hyp_utt_host, const int32_t* hyp_len_host, const int32_t* hyp_tok0_host,
                           const int32_t* hyp_tokens_host, const double* ctc_score_host, int sos, int eos,
                           float ctc_weight, float reverse_weight, float* tok_logp_l2r_dev, float* tok_logp_r2l_dev,
                           float* hyp_score_dev, int32_t* best_dev, void* workspace_dev, size_t workspace_bytes,
                           wb_stream_t stream) {
    return attention_rescoring_impl(mm, enc_out_bf16_dev, enc_rows, seq_start_host, seq_len_host, batch, n_hyp,
                                    hyp_utt_host, hyp_len_host, hyp_tok0_host, hyp_tokens_host, false, ctc_score_host,
                                    sos, eos, ctc_weight, reverse_weight, tok_logp_l2r_dev, tok_logp_r2l_dev,
                                    hyp_score_dev, best_dev, workspace_dev, workspace_bytes, stream);
}

int wb_attention_rescoring_dev(const wb_model* mm, const void* enc_out_bf16_dev, int64_t enc_rows,
                               const int32_t* seq_start_host, const int32_t* seq_len_host, int batch, int n_hyp,
                               const int32_t* hyp_utt_host, const int32_t* hyp_len_host, const int32_t* hyp_tok0_host,
                               const int32_t* hyp_tokens_dev, const double* ctc_score_host, int sos, int eos,
                               float ctc_weight, float reverse_weight, float* tok_logp_l2r_dev,
                               float* tok_logp_r2l_dev, float* hyp_score_dev, int32_t* best_dev, void* workspace_dev,
                               size_t workspace_bytes, wb_stream_t stream) {
    WB_REQUIRE(hyp_tokens_dev, WB_ERR_BAD_ARG, "attention_rescoring_dev: null token buffer");
    return attention_rescoring_impl(mm, enc_out_bf16_dev, enc_rows, seq_start_host, seq_len_host, batch, n_hyp,
                                    hyp_utt_host, hyp_len_host, hyp_tok0_host, hyp_tokens_dev, true, ctc_score_host, sos,
                                    eos, ctc_weight, reverse_weight, tok_logp_l2r_dev, tok_logp_r2l_dev, hyp_score_dev,
                                    best_dev, workspace_dev, workspace_bytes, stream);
}

int wb_prefix_share_tables(int dir, int batch, int n_hyp, const int32_t* hyp_utt_host, const int32_t* hyp_len_host,
                           const int32_t* hyp_tok0_host, const int32_t* hyp_tokens_host, int sos, int32_t* uniq_of_row,
                           int32_t* rep_row, int32_t* tok_u, int32_t* pos_u, int32_t* utt_q0_u, int32_t* utt_qn_u) {
    WB_REQUIRE(batch > 0 && n_hyp >= 0 && hyp_utt_host && hyp_len_host && hyp_tok0_host && hyp_tokens_host && uniq_of_row &&
                   rep_row && tok_u && pos_u && utt_q0_u && utt_qn_u && (dir == 0 || dir == 1),
               WB_ERR_BAD_ARG, "prefix_share_tables: bad argument");
    std::vector<int> utt_hyp0(batch, 0), utt_nhyp(batch, 0), hyp_row0(n_hyp > 0 ? n_hyp : 1, 0);
    int prev = -1, r = 0;
    for (int h = 0; h < n_hyp; ++h) {
        const int b = hyp_utt_host[h];
        WB_REQUIRE(b >= 0 && b < batch && b >= prev, WB_ERR_BAD_ARG, "prefix_share_tables: hyp_utt must be non-decreasing");
        if (b != prev) {
            utt_hyp0[b] = h;
            prev = b;
        }
        utt_nhyp[b] += 1;
        hyp_row0[h] = r;
        r += hyp_len_host[h] + 1;
    }
    int max_u = 0;
    return build_prefix_share(dir, batch, utt_hyp0.data(), utt_nhyp.data(), hyp_row0.data(), hyp_len_host, hyp_tok0_host,
                              hyp_tokens_host, sos, uniq_of_row, rep_row, tok_u, pos_u, utt_q0_u, utt_qn_u, &max_u);
}

int wb_decoder_logprobs(const wb_model* mm, const void* enc_out_bf16_dev, int64_t enc_rows,
                        const int32_t* seq_start_host, const int32_t* seq_len_host, int batch, int n_hyp,
                        const int32_t* hyp_utt_host, const int32_t* hyp_len_host, const int32_t* hyp_tok0_host,
                        const int32_t* hyp_tokens_host, int sos, int eos, int use_r2l, float* logp_dev,
                        float* r_logp_dev, int64_t ldl, void* workspace_dev, size_t workspace_bytes,
                        wb_stream_t stream) {
    const Model* m = reinterpret_cast<const Model*>(mm);
    WB_REQUIRE(m && m->finalized, WB_ERR_NOT_LOADED, "decoder_logprobs: model not finalized");
    WB_REQUIRE(m->cfg.dec_layers > 0, WB_ERR_UNSUPPORTED, "decoder_logprobs: model has no decoder");
    WB_REQUIRE(ldl >= m->cfg.vocab && logp_dev, WB_ERR_BAD_ARG, "decoder_logprobs: bad output");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace_dev);
    RsPlan P;
    RsDevPtrs dp;
    RC(prepare(m, enc_rows, seq_start_host, seq_len_host, batch, n_hyp, hyp_utt_host, hyp_len_host, hyp_tok0_host,
               hyp_tokens_host, false, false, nullptr, sos, eos, ws, workspace_bytes, &P, &dp, st));
    DirRows Q;
    Q.n = (int)P.R; Q.tok = dp.tok_l2r; Q.pos = dp.pos; Q.utt_q0 = dp.utt_q0; Q.utt_qn = dp.utt_qn;
    Q.max_utt_rows = P.max_utt_rows; Q.uniq_of_row = nullptr; Q.rep_row = nullptr;
    RC(run_decoder(m, m->left, P, dp, Q, enc_out_bf16_dev, enc_rows, ws, logp_dev, ldl, nullptr, nullptr, st));
    RC(ctc_logsoftmax_topk(logp_dev, ldl, (int)P.R, m->cfg.vocab, -1, 0.f, 0, nullptr, nullptr, st));
    if (use_r2l && m->cfg.rdec_layers > 0 && r_logp_dev) {
        Q.tok = dp.tok_r2l;
        RC(run_decoder(m, m->right, P, dp, Q, enc_out_bf16_dev, enc_rows, ws, r_logp_dev, ldl, nullptr, nullptr, st));
        RC(ctc_logsoftmax_topk(r_logp_dev, ldl, (int)P.R, m->cfg.vocab, -1, 0.f, 0, nullptr, nullptr, st));
    }
    return WB_OK;
}

}  // extern "C"
