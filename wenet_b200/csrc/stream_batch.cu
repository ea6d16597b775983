// Batched streaming: S concurrent forward_chunk sessions advanced in lockstep by ONE pass over the encoder
// ("batched streaming (many concurrent forward_chunk sessions)", the batched-cache design of the
// reference's GPU export, wenet/bin/export_onnx_gpu.py:83-232 StreamingEncoder: caches carry a batch dimension).
// Same arithmetic as wb_encoder_forward_chunk (encoder.py:204-300 per session): every row-wise stage (LayerNorm, the
// tcgen05 GEMMs, FFN, pointwise convs) runs on the S x chunk rows of all sessions at once - a 16-frame chunk alone leaves
// the tensor cores idle, 64 sessions give M = 1024 rows per GEMM - while the sequence-structured stages (conv2d
// subsampling, attention over the session's own K/V history, the causal depthwise conv with its own left context) are
// batched over sessions by their varlen / per-sequence arguments.  All sessions share T and cache_t1 (steady-state
// streaming: fixed window, full cache); each has its own position offset.  bf16 mode only.
//
// Layouts (device): xs [S][T][idim]; att_cache [S][layers][heads][cache_t1][128]; cnn_cache [S][layers][d][K-1];
// y [S][chunk][d]; r_att_cache [S][layers][heads][cache_t1 + chunk - next_cache_start][128]; r_cnn_cache like cnn_cache.
#include "model.h"
#include <math.h>
#include <vector>

namespace wb {

namespace {

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

#define RC(x)                         \
    do {                              \
        int _rc = (x);                \
        if (_rc != WB_OK) return _rc; \
    } while (0)

// per session (blockIdx.y): K/V history + this chunk's k, v -> kcat / vcat bf16 [key_size][d], trimmed fp32 cache back
__global__ void att_cache_concat_batch_kernel(const float* __restrict__ att_cache, long long cache_sess_stride, int cache_t1,
                                              const __nv_bfloat16* __restrict__ qkv, int chunk, int d, int H,
                                              __nv_bfloat16* __restrict__ kcat, __nv_bfloat16* __restrict__ vcat,
                                              float* __restrict__ r_att, long long r_sess_stride, int nxt) {
    const int s = blockIdx.y;
    const int key_size = cache_t1 + chunk;
    const int total = key_size * d;
    const float* cache = att_cache ? att_cache + (size_t)s * cache_sess_stride : nullptr;
    const __nv_bfloat16* q = qkv + (size_t)s * chunk * 3 * d;
    __nv_bfloat16* kc = kcat + (size_t)s * key_size * d;
    __nv_bfloat16* vc = vcat + (size_t)s * key_size * d;
    float* ro = r_att + (size_t)s * r_sess_stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int j = i / d, c = i - j * d;
        const int h = c >> 6, e = c & 63;
        float kv, vv;
        if (j < cache_t1) {
            const float* src = cache + ((size_t)h * cache_t1 + j) * 128;
            kv = src[e];
            vv = src[64 + e];
        } else {
            const __nv_bfloat16* row = q + (size_t)(j - cache_t1) * 3 * d;
            kv = __bfloat162float(row[d + c]);
            vv = __bfloat162float(row[2 * d + c]);
        }
        kc[i] = __float2bfloat16_rn(kv);
        vc[i] = __float2bfloat16_rn(vv);
        if (j >= nxt) {
            float* dst = ro + ((size_t)h * (key_size - nxt) + (j - nxt)) * 128;
            dst[e] = kv;
            dst[64 + e] = vv;
        }
    }
}

// per session: acat[s] = [cnn history (lead rows, bf16) ; LayerNorm rows of this chunk (bf16)], new cache = last `lead`
// rows of [history ; fp32 LayerNorm rows] (convolution.py:122-130)
__global__ void cnn_cache_batch_kernel(const float* __restrict__ cnn_cache, long long sess_stride, const float* __restrict__ a_f32,
                                       const __nv_bfloat16* __restrict__ a_bf16, int chunk, int d, int lead,
                                       __nv_bfloat16* __restrict__ acat, float* __restrict__ r_cnn) {
    const int s = blockIdx.y;
    const float* cc = cnn_cache ? cnn_cache + (size_t)s * sess_stride : nullptr;
    const float* af = a_f32 + (size_t)s * chunk * d;
    const __nv_bfloat16* ab = a_bf16 + (size_t)s * chunk * d;
    __nv_bfloat16* ac = acat + (size_t)s * (lead + chunk) * d;
    float* rc = r_cnn + (size_t)s * sess_stride;
    const int total = (lead + chunk) * d;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / d, c = i - r * d;
        if (r < lead) {
            const float old = cc ? cc[(size_t)c * lead + r] : 0.f;
            ac[i] = __float2bfloat16_rn(old);
            const int src = chunk + r;
            rc[(size_t)c * lead + r] = (src < lead) ? (cc ? cc[(size_t)c * lead + src] : 0.f) : af[(size_t)(src - lead) * d + c];
        } else {
            ac[i] = ab[(size_t)(r - lead) * d + c];
        }
    }
}

// same policy as the single-session path (encoder.cu resid_then_norm): x += alpha (A W^T + b); out = LayerNorm(x) in one
// kernel when the row fits an output tile (d == 256), so that a session's rows are bit-identical to wb_encoder_forward_chunk
int resid_then_norm(const void* A, long long lda_in, const Linear& W, int M, int d, float alpha, float* x, const Norm& n,
                    float eps, void* out, cudaStream_t st) {
    if (W.b != nullptr && gemm_resid_ln_supported(d))
        return gemm_resid_ln(A, lda_in, &W.tmap, W.w, M, d, W.K, W.b, alpha, x, d, nullptr, nullptr, n.g, n.b, eps, out, d, st);
    RC(gemm_bf16(A, lda_in, &W.tmap, W.w, M, d, W.K, W.b, EPI_RESID_F32, alpha, x, d, 0, st));
    return layernorm_rows(x, d, M, d, n.g, n.b, eps, out, d, 0, nullptr, 0, st);
}

struct SbPlan {
    int chunk, key_size, t1n, lead;
    size_t o_meta, o_out1, o_a2, o_out2, o_x, o_a, o_acat, o_af32, o_h, o_qkv, o_kcat, o_vcat, o_kp, o_kbias, o_ctx, o_g, o_g2,
        o_rowpos, total;
};

inline int sub4_len(int T) { return T >= 7 ? ((T - 1) / 2 - 1) / 2 : 0; }

void sb_plan(const Model* m, int T, int cache_t1, int S, SbPlan* P) {
    const wb_model_config& c = m->cfg;
    const size_t d = c.d_model;
    P->chunk = sub4_len(T);
    P->key_size = cache_t1 + P->chunk;
    P->t1n = P->chunk > 0 ? 2 * P->chunk + 1 : 0;
    P->lead = c.cnn_causal ? c.cnn_kernel - 1 : 0;
    const size_t rows = (size_t)S * P->chunk, keys = (size_t)S * P->key_size, cin = (size_t)S * (P->lead + P->chunk);
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o += align_up(bytes + 16);
        return at;
    };
    P->o_meta = take((size_t)S * (8 * 4 + 2 * 8) + 64);
    P->o_out1 = take((size_t)S * P->t1n * m->F1 * d * 2);
    P->o_a2 = take(rows * m->F2 * 9 * d * 2);
    P->o_out2 = take(rows * m->F2 * d * 2);
    P->o_x = take(rows * d * 4);
    P->o_a = take(rows * d * 2);
    P->o_acat = take(cin * d * 2);
    P->o_af32 = take(rows * d * 4);
    P->o_h = take(rows * c.ffn_dim * 2);
    P->o_qkv = take(rows * 3 * d * 2);
    P->o_kcat = take(keys * d * 2);
    P->o_vcat = take(keys * d * 2);
    P->o_kp = take(keys * d * 2);
    P->o_kbias = take(keys * c.heads * 4);
    P->o_ctx = take(rows * d * 2);
    P->o_g = take(cin * d * 2);
    P->o_g2 = take(rows * d * 2);
    P->o_rowpos = take(keys * 4);
    P->total = o + 256;
}

}  // namespace

size_t encoder_chunk_batch_workspace_bytes(const Model* m, int T, int cache_t1, int sessions) {
    SbPlan P;
    sb_plan(m, T, cache_t1, sessions, &P);
    return P.total;
}

// offsets_host != nullptr: regular call (the small tables are uploaded, one stream synchronisation);
// offsets_dev  != nullptr: capture-safe call (no copy, no synchronisation; the workspace must have seen a regular call
//                          with the same T / cache_t1 / sessions; positions are read on the device).
int encoder_forward_chunk_batch(const Model* m, const float* xs, int T, int S, const int32_t* offsets_host,
                                const int32_t* offsets_dev, int required_cache_size, const float* att_cache, int cache_t1,
                                const float* cnn_cache, float* y, float* r_att, float* r_cnn, int* out_chunk,
                                int* out_new_cache_t1, void* ws_v, size_t ws_bytes, cudaStream_t st) {
    const wb_model_config& c = m->cfg;
    WB_REQUIRE(c.arch == 0 && !c.precise, WB_ERR_UNSUPPORTED, "forward_chunk_batch: Conformer handles in bf16 mode only");
    WB_REQUIRE(S >= 1 && (offsets_host != nullptr) != (offsets_dev != nullptr), WB_ERR_BAD_ARG,
               "forward_chunk_batch: sessions >= 1 and exactly one of offsets_host / offsets_dev");
    WB_REQUIRE(cache_t1 == 0 || att_cache, WB_ERR_BAD_ARG, "forward_chunk_batch: att_cache missing");
    const int d = c.d_model, ff = c.ffn_dim, H = c.heads, L = c.enc_layers;
    PdlScope pdl_scope(pdl_stream_allowed());
    SbPlan P;
    sb_plan(m, T, cache_t1, S, &P);
    WB_REQUIRE(P.chunk > 0, WB_ERR_BAD_ARG, "forward_chunk_batch: %d input frames give no output frame", T);
    WB_REQUIRE(ws_bytes >= P.total, WB_ERR_WORKSPACE, "forward_chunk_batch: workspace %zu < required %zu", ws_bytes, P.total);
    const int chunk = P.chunk, key_size = P.key_size, lead = P.lead;
    int nxt;
    if (required_cache_size < 0) nxt = 0;
    else if (required_cache_size == 0) nxt = key_size;
    else nxt = key_size - required_cache_size > 0 ? key_size - required_cache_size : 0;
    if (out_chunk) *out_chunk = chunk;
    if (out_new_cache_t1) *out_new_cache_t1 = key_size - nxt;
    WB_REQUIRE(lead == 0 || r_cnn, WB_ERR_BAD_ARG, "forward_chunk_batch: r_cnn_cache missing");
    uint8_t* ws = reinterpret_cast<uint8_t*>(ws_v);
    // meta: ints t1n[S], len_q[S], q_start[S], k_start[S], k_len[S], cin_start[S], cin_len[S], pos_off[S]; int64 off1[S], off2[S]
    int* d_meta = reinterpret_cast<int*>(ws + P.o_meta);
    long long* d_ll = reinterpret_cast<long long*>(ws + P.o_meta + align_up((size_t)S * 8 * 4, 16));
    if (offsets_host != nullptr) {
        std::vector<int> hi((size_t)8 * S);
        std::vector<long long> hl((size_t)2 * S);
        for (int s = 0; s < S; ++s) {
            WB_REQUIRE(offsets_host[s] - cache_t1 >= 0 && offsets_host[s] + chunk <= c.max_pos, WB_ERR_BAD_ARG,
                       "forward_chunk_batch: session %d positions [%d, %d) outside the positional table", s,
                       offsets_host[s] - cache_t1, offsets_host[s] + chunk);
            hi[s] = P.t1n;
            hi[S + s] = chunk;
            hi[2 * S + s] = s * chunk;
            hi[3 * S + s] = s * key_size;
            hi[4 * S + s] = key_size;
            hi[5 * S + s] = s * (lead + chunk);
            hi[6 * S + s] = lead + chunk;
            hi[7 * S + s] = offsets_host[s];
            hl[s] = (long long)s * P.t1n * m->F1;
            hl[S + s] = (long long)s * chunk * m->F2;
        }
        WB_CHECK_CUDA(cudaMemcpyAsync(d_meta, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice, st));
        WB_CHECK_CUDA(cudaMemcpyAsync(d_ll, hl.data(), hl.size() * 8, cudaMemcpyHostToDevice, st));
        WB_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    const int* d_t1n = d_meta;
    const int* d_qlen = d_meta + S;
    const int* d_qstart = d_meta + 2 * S;
    const int* d_kstart = d_meta + 3 * S;
    const int* d_klen = d_meta + 4 * S;
    const int* d_cstart = d_meta + 5 * S;
    const int* d_clen = d_meta + 6 * S;
    const int* d_posoff = offsets_dev ? offsets_dev : d_meta + 7 * S;
    const long long* d_off1 = d_ll;
    const long long* d_off2 = d_ll + S;

    void* out1 = ws + P.o_out1;
    void* a2 = ws + P.o_a2;
    void* out2 = ws + P.o_out2;
    float* x = reinterpret_cast<float*>(ws + P.o_x);
    __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(ws + P.o_a);
    __nv_bfloat16* acat = reinterpret_cast<__nv_bfloat16*>(ws + P.o_acat);
    float* af32 = reinterpret_cast<float*>(ws + P.o_af32);
    void* h = ws + P.o_h;
    __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(ws + P.o_qkv);
    __nv_bfloat16* kcat = reinterpret_cast<__nv_bfloat16*>(ws + P.o_kcat);
    __nv_bfloat16* vcat = reinterpret_cast<__nv_bfloat16*>(ws + P.o_vcat);
    void* kp = ws + P.o_kp;
    float* kbias = reinterpret_cast<float*>(ws + P.o_kbias);
    void* ctx = ws + P.o_ctx;
    void* g = ws + P.o_g;
    void* g2 = ws + P.o_g2;
    int* d_row_pos = reinterpret_cast<int*>(ws + P.o_rowpos);
    const int rows = S * chunk, keys = S * key_size, cin = S * (lead + chunk);

    RC(subsample_conv1(xs, (long long)T * c.input_dim, c.input_dim, d_t1n, d_off1, S, P.t1n, m->cmvn_mean, m->cmvn_istd, m->conv1_w,
                       m->conv1_b, d, out1, 0, st));
    RC(subsample_im2col(out1, d_off1, d_qlen, d_off2, S, chunk, m->F1, m->F2, d, a2, 0, st));
    RC(gemm_bf16(a2, m->conv2.K, &m->conv2.tmap, m->conv2.w, rows * m->F2, d, m->conv2.K, m->conv2.b, EPI_BF16_RELU, 1.0f, out2, d,
                 0, st));
    RC(gemm_bf16(out2, m->embed_out.K, &m->embed_out.tmap, m->embed_out.w, rows, d, m->embed_out.K, m->embed_out.b, EPI_F32,
                 sqrtf((float)d), x, d, 0, st));
    // key positions of session s: offset_s - cache_t1 + j (clamped to the table)
    RC(fill_row_pos(d_kstart, d_klen, S, -cache_t1, d_row_pos, key_size, st, d_posoff, c.max_pos, /*per_seq_offset=*/1));
    const float att_scale = 1.0f / sqrtf(64.0f);
    const size_t att_l = (size_t)H * cache_t1 * 128, ratt_l = (size_t)H * (key_size - nxt) * 128;
    const size_t cnn_l = (size_t)d * lead;
    for (int li = 0; li < L; ++li) {
        const EncLayer& Ly = m->layers[li];
        if (li == 0) RC(layernorm_rows(x, d, rows, d, Ly.n_ffm.g, Ly.n_ffm.b, c.ln_eps, a, d, 0, nullptr, 0, st));
        RC(gemm_bf16(a, d, &Ly.ffm1.tmap, Ly.ffm1.w, rows, ff, d, Ly.ffm1.b, EPI_BF16_SILU, 1.0f, h, ff, 0, st));
        RC(resid_then_norm(h, ff, Ly.ffm2, rows, d, 0.5f, x, Ly.n_mha, c.ln_eps, a, st));
        RC(gemm_bf16(a, d, &Ly.qkv.tmap, Ly.qkv.w, rows, 3 * d, d, Ly.qkv.b, EPI_BF16, 1.0f, qkv, 3 * d, 0, st));
        {
            dim3 grid(ceil_div(key_size * d, 256), S);
            att_cache_concat_batch_kernel<<<grid, 256, 0, st>>>(cache_t1 > 0 ? att_cache + li * att_l : nullptr, (long long)L * att_l,
                                                                cache_t1, qkv, chunk, d, H, kcat, vcat, r_att + li * ratt_l,
                                                                (long long)L * ratt_l, nxt);
            count_launch();
            WB_CHECK_LAUNCH();
        }
        RC(relpos_kprep(kcat, d, Ly.pos_proj, d_row_pos, Ly.pos_u, Ly.pos_v, keys, H, kp, d, kbias, st, att_scale * 1.4426950408889634f));
        {
            AttnArgs A;
            A.q = qkv; A.ldq = 3 * d; A.q_rows = rows; A.q_col0 = 0;
            A.k = kp; A.ldk = d; A.k_rows = keys; A.k_col0 = 0;
            A.v = vcat; A.ldv = d; A.v_rows = keys; A.v_col0 = 0;
            A.kbias = kbias; A.ld_kbias = H; A.kbias_scaled = 1;
            A.q_start = d_qstart; A.q_len = d_qlen; A.k_start = d_kstart; A.k_len = d_klen;
            A.batch = S; A.heads = H; A.max_q_len = chunk;
            A.chunk_size = 0; A.num_left_chunks = -1; A.scale = att_scale;   // att_mask is all-ones (encoder.py:243-247)
            A.out = ctx; A.ldo = d; A.out_col0 = 0; A.split3_out = 0; A.v_mode = 0;
            RC(attention_forward(A, st));
        }
        RC(gemm_bf16(ctx, d, &Ly.out.tmap, Ly.out.w, rows, d, d, Ly.out.b, EPI_RESID_F32, 1.0f, x, d, 0, st));
        RC(layernorm_rows(x, d, rows, d, Ly.n_conv.g, Ly.n_conv.b, c.ln_eps, a, d, 0, af32, d, st));
        const void* pw1_in = a;
        int pw1_rows = rows;
        if (lead > 0) {
            dim3 grid(ceil_div((lead + chunk) * d, 256), S);
            cnn_cache_batch_kernel<<<grid, 256, 0, st>>>(cnn_cache ? cnn_cache + li * cnn_l : nullptr, (long long)L * cnn_l, af32, a,
                                                         chunk, d, lead, acat, r_cnn + li * cnn_l);
            count_launch();
            WB_CHECK_LAUNCH();
            pw1_in = acat;
            pw1_rows = cin;
        }
        RC(gemm_bf16(pw1_in, d, &Ly.pw1.tmap, Ly.pw1.w, pw1_rows, 2 * d, d, Ly.pw1.b, EPI_GLU_BF16, 1.0f, g, d, 0, st));
        {
            DwConvArgs D;
            D.g = g; D.ldg = d; D.in_split3 = 0; D.seq_start = d_cstart; D.seq_len = d_clen; D.out_start = d_qstart;
            D.batch = S; D.max_len = chunk; D.lead = lead; D.d = d; D.ksize = c.cnn_kernel;
            D.causal = c.cnn_causal; D.w = Ly.dw_w; D.bias = Ly.dw_b; D.norm_type = c.cnn_norm;
            D.gamma = Ly.n_cnn.g; D.beta = Ly.n_cnn.b; D.eps = c.ln_eps; D.pad_vec = Ly.pad_vec; D.pad_until = chunk;
            D.out = g2; D.ldo = d; D.split3 = 0;
            RC(dwconv_norm_silu(D, st));
        }
        RC(resid_then_norm(g2, d, Ly.pw2, rows, d, 1.0f, x, Ly.n_ff, c.ln_eps, a, st));
        RC(gemm_bf16(a, d, &Ly.ff1.tmap, Ly.ff1.w, rows, ff, d, Ly.ff1.b, EPI_BF16_SILU, 1.0f, h, ff, 0, st));
        if (li + 1 < L && Ly.ff2.b != nullptr && gemm_resid_ln_supported(d)) {
            const EncLayer& Ln = m->layers[li + 1];
            RC(gemm_resid_ln(h, ff, &Ly.ff2.tmap, Ly.ff2.w, rows, d, ff, Ly.ff2.b, 0.5f, x, d, Ly.n_final.g, Ly.n_final.b, Ln.n_ffm.g,
                             Ln.n_ffm.b, c.ln_eps, a, d, st));
            continue;
        }
        RC(gemm_bf16(h, ff, &Ly.ff2.tmap, Ly.ff2.w, rows, d, ff, Ly.ff2.b, EPI_RESID_F32, 0.5f, x, d, 0, st));
        if (li + 1 < L) {
            const EncLayer& Ln = m->layers[li + 1];
            RC(layernorm2_rows(x, d, rows, d, Ly.n_final.g, Ly.n_final.b, Ln.n_ffm.g, Ln.n_ffm.b, c.ln_eps, x, d, a, d, 0, nullptr, 0,
                               st));
        } else {
            RC(layernorm2_rows(x, d, rows, d, Ly.n_final.g, Ly.n_final.b, m->after.g, m->after.b, c.ln_eps, nullptr, 0, nullptr, 0,
                               0, y, d, st));
        }
    }
    return WB_OK;
}

}  // namespace wb
