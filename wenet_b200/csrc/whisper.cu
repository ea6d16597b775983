// Whisper front-end and encoder (the `wenet.cli` / BASELINE configs[4] path):
//   log-mel spectrogram  wenet/dataset/processor.py:320-369 (compute_log_mel_spectrogram: hann STFT n_fft 400 / hop 160,
//                        |.|^2, slaney mel filterbank, log10, clamp to max - 8, (x + 4) / 4)
//   Conv1dSubsampling2   wenet/models/transformer/subsampling.py:117-171 (Conv1d k3 p1 + GELU, Conv1d k3 s2 p1 + GELU)
//   abs_pos_whisper      wenet/models/transformer/embedding.py:150-164 (xscale 1, sin|cos table)
//   TransformerEncoder   wenet/models/transformer/encoder.py:365-440 + encoder_layer.py:28-135 (pre-norm MHA + GELU FFN),
//                        after_norm (encoder.py:176-177)
// The convolutions are GEMMs on the tcgen05 kernel over explicit im2col rows ((tap, channel) order); attention is the
// same varlen tcgen05 kernel as the Conformer path with no positional term (key_bias = false is a zero bias slice delivered
// by the packer); every Linear is gemm_tcgen05_kernel (GELU epilogue: EPI_BF16_GELU).  Rows are packed: utterance b owns
// rows [seq_start[b], seq_start[b] + T'_b) with T'_b given by the reference's mask rule x_mask[:, :, (time + 1) % 2::2].
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

// ------------------------------------------------------------------------------------------------------------------
// log-mel
// ------------------------------------------------------------------------------------------------------------------
constexpr int LM_FR = 16;      // frames per CTA
constexpr int LM_THREADS = 256;

// One CTA = LM_FR consecutive frames of one utterance.  Samples (reflect-padded, torch.stft center=True) are staged in
// shared memory; thread k < n_bins accumulates the windowed DFT bin k of all LM_FR frames against the host-built
// tables wcos / wsin [n_fft][n_bins] (window folded in, bin-contiguous so a warp reads 128 B per sample index), then
// the mel projection (dense [n_bins][n_mel], mel-contiguous) and log10.  The per-utterance maximum is merged with an
// atomicMax on the order-preserving integer image of the float.
__device__ __forceinline__ int float_order_key(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : (i ^ 0x7fffffff);
}
__device__ __forceinline__ float float_from_key(int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); }

__global__ void __launch_bounds__(LM_THREADS)
logmel_power_kernel(const float* __restrict__ pcm, long long pcm_stride, const int* __restrict__ num_samples, int n_fft,
                    int hop, int n_bins, int n_mel, const float* __restrict__ wcos, const float* __restrict__ wsin,
                    const float* __restrict__ melT, float* __restrict__ out, long long out_stride_b, int max_frames,
                    int* __restrict__ utt_max_key) {
    extern __shared__ float lm_smem[];
    float* frames = lm_smem;                          // [LM_FR][n_fft]
    float* power = lm_smem + LM_FR * n_fft;           // [LM_FR][n_bins]
    const int b = blockIdx.y;
    const int n = num_samples[b];
    const int n_frames = min(n / hop, max_frames);    // 1 + n / hop frames, the last one dropped (processor.py:357)
    const int f0 = blockIdx.x * LM_FR;
    if (f0 >= n_frames) return;
    const float* x = pcm + (long long)b * pcm_stride;
    const int half = n_fft / 2;
    for (int i = threadIdx.x; i < LM_FR * n_fft; i += LM_THREADS) {
        const int f = i / n_fft, j = i - f * n_fft;
        int s = (f0 + f) * hop + j - half;
        if (s < 0) s = -s;                            // reflect (no edge repeat)
        if (s >= n) s = 2 * (n - 1) - s;
        frames[i] = (f0 + f < n_frames && s >= 0 && s < n) ? x[s] : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_bins; k += LM_THREADS) {
        float re[LM_FR], im[LM_FR];
#pragma unroll
        for (int f = 0; f < LM_FR; ++f) re[f] = im[f] = 0.f;
        for (int j = 0; j < n_fft; ++j) {
            const float c = __ldg(wcos + (long long)j * n_bins + k), s = __ldg(wsin + (long long)j * n_bins + k);
#pragma unroll
            for (int f = 0; f < LM_FR; ++f) {
                const float v = frames[f * n_fft + j];
                re[f] = fmaf(v, c, re[f]);
                im[f] = fmaf(v, s, im[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < LM_FR; ++f) power[f * n_bins + k] = re[f] * re[f] + im[f] * im[f];
    }
    __syncthreads();
    float local_max = -INFINITY;
    for (int i = threadIdx.x; i < LM_FR * n_mel; i += LM_THREADS) {
        const int f = i / n_mel, m = i - f * n_mel;
        if (f0 + f >= n_frames) continue;
        float acc = 0.f;
        const float* pw = power + f * n_bins;
        for (int k = 0; k < n_bins; ++k) acc = fmaf(__ldg(melT + (long long)k * n_mel + m), pw[k], acc);
        const float lg = log10f(fmaxf(acc, 1e-10f));
        out[(long long)b * out_stride_b + (long long)(f0 + f) * n_mel + m] = lg;
        local_max = fmaxf(local_max, lg);
    }
    local_max = warp_max(local_max);
    if ((threadIdx.x & 31) == 0 && local_max > -INFINITY) atomicMax(utt_max_key + b, float_order_key(local_max));
}

// log_spec = (max(log_spec, utterance max - 8) + 4) / 4 on the utterance's own frames, 0 on the batch padding
// (processor.py:365-367, then pad_sequence(..., 0) processor.py:562-566)
__global__ void logmel_finish_kernel(float* __restrict__ out, long long out_stride_b, const int* __restrict__ num_samples,
                                     int hop, int n_mel, int max_frames, const int* __restrict__ utt_max_key) {
    const int b = blockIdx.y;
    const int n_frames = min(num_samples[b] / hop, max_frames);
    const float floor_v = float_from_key(utt_max_key[b]) - 8.0f;
    const long long total = (long long)max_frames * n_mel;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i / n_mel);
        float* p = out + (long long)b * out_stride_b + i;
        *p = (f < n_frames) ? (fmaxf(*p, floor_v) + 4.0f) * 0.25f : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Conv1dSubsampling2 as im2col + GEMM
// ------------------------------------------------------------------------------------------------------------------
// conv1 rows of utterance b: t in [0, rows1_b); row = [x[t-1] | x[t] | x[t+1]] (idim each), x = 0 outside [0, len_b)
// split3 (precise mode): every tap block of idim values is written as [hi | lo | hi] (row = 9 idim wide), matching weights
// packed [hi | hi | lo] per tap block (bf16x3, see gemm.cu)
__global__ void w_im2col1_kernel(const float* __restrict__ feats, long long stride_b, int idim, const int* __restrict__ len,
                                 const int* __restrict__ rows1, const long long* __restrict__ off1, int batch, int split3,
                                 __nv_bfloat16* __restrict__ a1) {
    const int b = blockIdx.y;
    const int n = len[b], nr = rows1[b];
    const int per_row = 3 * idim;
    const int p3 = split3 ? 3 : 1;
    const long long total = (long long)nr * per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / per_row), j = (int)(i - (long long)t * per_row);
        const int tap = j / idim, cch = j - tap * idim;
        const int ts = t + tap - 1;
        const float v = (ts >= 0 && ts < n) ? feats[(long long)b * stride_b + (long long)ts * idim + cch] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        __nv_bfloat16* row = a1 + (off1[b] + t) * (long long)per_row * p3 + (long long)tap * idim * p3;
        row[cch] = hi;
        if (split3) {
            row[idim + cch] = __float2bfloat16_rn(v - __bfloat162float(hi));
            row[2 * idim + cch] = hi;
        }
    }
}
// conv2 rows: t' in [0, T'_b); row = [c1[2t'-1] | c1[2t'] | c1[2t'+1]] (d each), c1 = 0 outside [0, rows1_b).  16-byte copies.
__global__ void w_im2col2_kernel(const uint4* __restrict__ c1, const int* __restrict__ rows1, const long long* __restrict__ off1,
                                 const int* __restrict__ seq_start, const int* __restrict__ seq_len, int d8,
                                 uint4* __restrict__ a2) {
    const int b = blockIdx.y;
    const int nr = rows1[b], tp = seq_len[b];
    const long long total = (long long)tp * 3 * d8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i / (3 * d8)), j = (int)(i - (long long)t * 3 * d8);
        const int tap = j / d8, v8 = j - tap * d8;
        const int ts = 2 * t + tap - 1;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ts >= 0 && ts < nr) v = c1[(off1[b] + ts) * d8 + v8];
        a2[((long long)seq_start[b] + t) * 3 * d8 + j] = v;
    }
}
// x[m] = float(g[m]) + pe[t'] (xscale 1)
__global__ void w_add_pe_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ pe, const int* __restrict__ seq_start,
                                const int* __restrict__ seq_len, int d, int split3, float* __restrict__ x) {
    const int b = blockIdx.y;
    const long long total = (long long)seq_len[b] * d;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = (long long)seq_start[b] * d + i;
        float v;
        if (split3) {   // rows are [hi | lo | hi]
            const long long r = m / d, c = m - r * d;
            v = __bfloat162float(g[r * 3 * d + c]) + __bfloat162float(g[r * 3 * d + d + c]);
        } else {
            v = __bfloat162float(g[m]);
        }
        x[m] = v + pe[i];
    }
}

struct WePlan {
    long long rows1 = 0, M = 0;
    int max_tp = 0;
    size_t o_int = 0, o_a1 = 0, o_c1 = 0, o_a2 = 0, o_g = 0, o_x = 0, o_a = 0, o_qkv = 0, o_ctx = 0, o_h = 0, total = 0;
};

void we_rows(const int32_t* lens, int batch, int time_pad, std::vector<int>* rows1, std::vector<int>* tp) {
    rows1->resize(batch);
    tp->resize(batch);
    for (int b = 0; b < batch; ++b) {
        const int n = lens[b];
        // subsampling.py:171: x_mask[:, :, (time + 1) % 2::2]
        (*tp)[b] = (time_pad % 2 == 0) ? n / 2 : (n + 1) / 2;
        (*rows1)[b] = n + ((n < time_pad) ? 1 : 0);   // the conv1 frame at t = len sees the frame len - 1 (see header)
    }
}

void we_layout(const Model* m, const std::vector<int>& rows1, const std::vector<int>& tp, WePlan* P) {
    const size_t p3 = m->cfg.precise ? 3 : 1;   // precise: bf16 activations are [hi | lo | hi], q / k / v fp32
    const size_t d = (size_t)m->cfg.d_model * p3, ff = (size_t)m->cfg.ffn_dim * p3, idim = (size_t)m->cfg.input_dim * p3;
    const size_t d1 = m->cfg.d_model;
    const int batch = (int)rows1.size();
    P->rows1 = 0;
    P->M = 0;
    P->max_tp = 0;
    for (int b = 0; b < batch; ++b) {
        P->rows1 += rows1[b];
        P->M += tp[b];
        if (tp[b] > P->max_tp) P->max_tp = tp[b];
    }
    size_t o = 0;
    P->o_int = o; o += align_up((size_t)batch * (4 * 4 + 8) + 256);
    P->o_a1 = o; o += align_up((size_t)P->rows1 * 3 * idim * 2);
    P->o_c1 = o; o += align_up((size_t)P->rows1 * d * 2);
    P->o_a2 = o; o += align_up((size_t)P->M * 3 * d * 2);
    P->o_g = o; o += align_up((size_t)P->M * d * 2);
    P->o_x = o; o += align_up((size_t)P->M * d1 * 4);
    P->o_a = o; o += align_up((size_t)P->M * d * 2);
    P->o_qkv = o; o += align_up((size_t)P->M * 3 * d1 * (m->cfg.precise ? 4 : 2));
    P->o_ctx = o; o += align_up((size_t)P->M * d * 2);
    P->o_h = o; o += align_up((size_t)P->M * ff * 2);
    P->total = o + 256;
}

}  // namespace

// ---- log-mel plan ------------------------------------------------------------------------------------------------
struct LogMelPlan {
    int n_fft, hop, n_mel, n_bins;
    float* wcos;   // [n_fft][n_bins]  window[j] * cos(2 pi j k / n_fft)
    float* wsin;   // [n_fft][n_bins]
    float* melT;   // [n_bins][n_mel]
};

int logmel_plan_create(LogMelPlan** out, int n_fft, int hop, int n_mel, const float* window_host, const float* mel_host) {
    WB_REQUIRE(n_fft >= 16 && n_fft <= 1024 && hop >= 1 && n_mel >= 1 && n_mel <= 256, WB_ERR_BAD_ARG, "logmel: bad geometry");
    const int nb = n_fft / 2 + 1;
    std::vector<float> c((size_t)n_fft * nb), s((size_t)n_fft * nb), mt((size_t)nb * n_mel);
    for (int j = 0; j < n_fft; ++j)
        for (int k = 0; k < nb; ++k) {
            const long long jk = ((long long)j * k) % n_fft;   // exact argument reduction
            const double ang = 2.0 * M_PI * (double)jk / (double)n_fft;
            c[(size_t)j * nb + k] = (float)((double)window_host[j] * cos(ang));
            s[(size_t)j * nb + k] = (float)(-(double)window_host[j] * sin(ang));
        }
    for (int mI = 0; mI < n_mel; ++mI)
        for (int k = 0; k < nb; ++k) mt[(size_t)k * n_mel + mI] = mel_host[(size_t)mI * nb + k];
    LogMelPlan* p = new LogMelPlan();
    p->n_fft = n_fft;
    p->hop = hop;
    p->n_mel = n_mel;
    p->n_bins = nb;
    p->wcos = p->wsin = p->melT = nullptr;
    WB_CHECK_CUDA(cudaMalloc((void**)&p->wcos, c.size() * 4));
    WB_CHECK_CUDA(cudaMalloc((void**)&p->wsin, s.size() * 4));
    WB_CHECK_CUDA(cudaMalloc((void**)&p->melT, mt.size() * 4));
    WB_CHECK_CUDA(cudaMemcpy(p->wcos, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
    WB_CHECK_CUDA(cudaMemcpy(p->wsin, s.data(), s.size() * 4, cudaMemcpyHostToDevice));
    WB_CHECK_CUDA(cudaMemcpy(p->melT, mt.data(), mt.size() * 4, cudaMemcpyHostToDevice));
    *out = p;
    return WB_OK;
}

void logmel_plan_destroy(LogMelPlan* p) {
    if (!p) return;
    cudaFree(p->wcos);
    cudaFree(p->wsin);
    cudaFree(p->melT);
    delete p;
}

int logmel_forward(const LogMelPlan* p, const float* pcm, long long pcm_stride, const int* num_samples_dev, int batch,
                   float* out, long long frames_stride, int max_frames, int* scratch_dev /*[batch]*/, cudaStream_t st) {
    if (batch <= 0 || max_frames <= 0) return WB_OK;
    const size_t smem = (size_t)LM_FR * (p->n_fft + p->n_bins) * sizeof(float);
    WB_SET_MAX_DYN_SMEM(logmel_power_kernel, smem);
    // the smallest order key (-inf) as the initial maximum
    WB_CHECK_CUDA(cudaMemsetAsync(scratch_dev, 0x80, (size_t)batch * sizeof(int), st));
    {
        ProfScope _ps(PT_FBANK, st, (double)batch * max_frames * (p->hop * 4.0 + p->n_mel * 4.0));
        dim3 grid(ceil_div(max_frames, LM_FR), batch);
        logmel_power_kernel<<<grid, LM_THREADS, smem, st>>>(pcm, pcm_stride, num_samples_dev, p->n_fft, p->hop, p->n_bins, p->n_mel,
                                                            p->wcos, p->wsin, p->melT, out, frames_stride * p->n_mel, max_frames,
                                                            scratch_dev);
        count_launch();
        WB_CHECK_LAUNCH();
    }
    {
        ProfScope _ps(PT_FBANK, st, (double)batch * max_frames * p->n_mel * 8.0);
        dim3 grid(ceil_div(max_frames * p->n_mel, 256 * 4), batch);
        logmel_finish_kernel<<<grid, 256, 0, st>>>(out, frames_stride * p->n_mel, num_samples_dev, p->hop, p->n_mel, max_frames,
                                                   scratch_dev);
        count_launch();
        WB_CHECK_LAUNCH();
    }
    return WB_OK;
}

// ---- encoder -------------------------------------------------------------------------------------------------------
long long whisper_encoder_out_rows(int batch, const int32_t* lens, int time_pad) {
    std::vector<int> r1, tp;
    we_rows(lens, batch, time_pad, &r1, &tp);
    long long M = 0;
    for (int b = 0; b < batch; ++b) M += tp[b];
    return M;
}

size_t whisper_encoder_workspace_bytes(const Model* m, int batch, const int32_t* lens, int time_pad) {
    std::vector<int> r1, tp;
    we_rows(lens, batch, time_pad, &r1, &tp);
    WePlan P;
    we_layout(m, r1, tp, &P);
    return P.total;
}

int whisper_encoder_forward(const Model* m, const float* feats, long long feats_stride_b, const int32_t* lens_host, int batch,
                            int time_pad, float* enc_out, void* enc_out_bf16, int32_t* seq_start_dev, int32_t* seq_len_dev,
                            void* ws_v, size_t ws_bytes, cudaStream_t st) {
    const wb_model_config& c = m->cfg;
    WB_REQUIRE(c.arch == 1, WB_ERR_BAD_ARG, "whisper_encoder_forward: not a Whisper model handle");
    const WhisperEnc& E = m->wenc;
    const int d = c.d_model, ff = c.ffn_dim, idim = c.input_dim, H = c.heads;
    // precise (parity) mode: bf16x3 GEMMs over [hi | lo | hi] activations, fp32 q / k / v and attention (precise.cu) - the
    // same scheme as the Conformer path (include/wenet_b200.h, wb_model_config.precise)
    const int sp = c.precise ? 1 : 0, p3 = sp ? 3 : 1;
    const long long lda = (long long)d * p3, ldh = (long long)ff * p3;
    std::vector<int> r1, tp;
    we_rows(lens_host, batch, time_pad, &r1, &tp);
    for (int b = 0; b < batch; ++b)
        WB_REQUIRE(lens_host[b] >= 0 && lens_host[b] <= time_pad, WB_ERR_BAD_ARG, "whisper encoder: feature length %d > padded %d",
                   lens_host[b], time_pad);
    WePlan P;
    we_layout(m, r1, tp, &P);
    WB_REQUIRE(ws_bytes >= P.total, WB_ERR_WORKSPACE, "whisper encoder: workspace %zu < required %zu", ws_bytes, P.total);
    WB_REQUIRE(P.max_tp <= c.max_pos, WB_ERR_UNSUPPORTED, "whisper encoder: %d frames exceed the position table (%d)", P.max_tp,
               c.max_pos);
    WB_REQUIRE(P.M < 2147483647LL && P.rows1 < 2147483647LL, WB_ERR_UNSUPPORTED, "whisper encoder: too many rows");
    if (P.M == 0) return WB_OK;
    uint8_t* ws = reinterpret_cast<uint8_t*>(ws_v);
    // small tables: len, rows1, (seq_start, seq_len go to the caller's arrays), off1 (int64)
    std::vector<int> hi((size_t)4 * batch);
    std::vector<long long> ho(batch);
    long long o1 = 0;
    int s0 = 0;
    for (int b = 0; b < batch; ++b) {
        hi[b] = lens_host[b];
        hi[batch + b] = r1[b];
        hi[2 * batch + b] = s0;
        hi[3 * batch + b] = tp[b];
        ho[b] = o1;
        o1 += r1[b];
        s0 += tp[b];
    }
    int* d_int = reinterpret_cast<int*>(ws + P.o_int);
    long long* d_off1 = reinterpret_cast<long long*>(ws + P.o_int + align_up((size_t)4 * batch * 4, 8));
    WB_CHECK_CUDA(cudaMemcpyAsync(d_int, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice, st));
    WB_CHECK_CUDA(cudaMemcpyAsync(d_off1, ho.data(), ho.size() * 8, cudaMemcpyHostToDevice, st));
    WB_CHECK_CUDA(cudaMemcpyAsync(seq_start_dev, hi.data() + 2 * batch, (size_t)batch * 4, cudaMemcpyHostToDevice, st));
    WB_CHECK_CUDA(cudaMemcpyAsync(seq_len_dev, hi.data() + 3 * batch, (size_t)batch * 4, cudaMemcpyHostToDevice, st));
    const int* d_len = d_int;
    const int* d_rows1 = d_int + batch;
    __nv_bfloat16* a1 = reinterpret_cast<__nv_bfloat16*>(ws + P.o_a1);
    __nv_bfloat16* c1 = reinterpret_cast<__nv_bfloat16*>(ws + P.o_c1);
    __nv_bfloat16* a2 = reinterpret_cast<__nv_bfloat16*>(ws + P.o_a2);
    __nv_bfloat16* g = reinterpret_cast<__nv_bfloat16*>(ws + P.o_g);
    float* x = reinterpret_cast<float*>(ws + P.o_x);
    void* a = ws + P.o_a;
    void* qkv = ws + P.o_qkv;
    void* ctx = ws + P.o_ctx;
    void* h = ws + P.o_h;
    int max_r1 = 0;
    for (int b = 0; b < batch; ++b) max_r1 = r1[b] > max_r1 ? r1[b] : max_r1;
    {
        ProfScope _ps(PT_IM2COL, st, (double)P.rows1 * 3 * idim * 6.0);
        dim3 grid(ceil_div(max_r1 * 3 * idim, 256 * 4), batch);
        w_im2col1_kernel<<<grid, 256, 0, st>>>(feats, feats_stride_b, idim, d_len, d_rows1, d_off1, batch, sp, a1);
        count_launch();
        WB_CHECK_LAUNCH();
    }
    RC(gemm_bf16(a1, 3 * idim * p3, &E.conv1.tmap, E.conv1.w, (int)P.rows1, d, 3 * idim * p3, E.conv1.b, EPI_BF16_GELU, 1.0f, c1,
                 lda, sp, st));
    {
        ProfScope _ps(PT_IM2COL, st, (double)P.M * 3 * d * 4.0);
        // (precise: a conv1 row is [hi | lo | hi] = 3 d wide, so a tap block of the im2col row is too)
        const int d8 = d * p3 / 8;
        dim3 grid(ceil_div(P.max_tp * 3 * d8, 256 * 2), batch);
        w_im2col2_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(c1), d_rows1, d_off1, seq_start_dev, seq_len_dev,
                                               d8, reinterpret_cast<uint4*>(a2));
        count_launch();
        WB_CHECK_LAUNCH();
    }
    RC(gemm_bf16(a2, 3 * lda, &E.conv2.tmap, E.conv2.w, (int)P.M, d, 3 * d * p3, E.conv2.b, EPI_BF16_GELU, 1.0f, g, lda, sp, st));
    {
        ProfScope _ps(PT_MISC, st, (double)P.M * d * 10.0);
        dim3 grid(ceil_div(P.max_tp * d, 256 * 4), batch);
        w_add_pe_kernel<<<grid, 256, 0, st>>>(g, E.pe, seq_start_dev, seq_len_dev, d, sp, x);
        count_launch();
        WB_CHECK_LAUNCH();
    }
    const float scale = 1.0f / sqrtf(64.0f);
    const int M = (int)P.M;
    for (size_t li = 0; li < E.layers.size(); ++li) {
        const TrLayer& L = E.layers[li];
        RC(layernorm_rows(x, d, M, d, L.n1.g, L.n1.b, c.ln_eps, a, lda, sp, nullptr, 0, st));
        if (sp) {
            float* qf = reinterpret_cast<float*>(qkv);
            RC(gemm_bf16(a, lda, &L.qkv.tmap, L.qkv.w, M, 3 * d, 3 * d, L.qkv.b, EPI_F32, 1.0f, qf, 3 * d, 0, st));
            AttnF32Args A;
            A.q = qf; A.ldq = 3 * d; A.k = qf + d; A.ldk = 3 * d; A.v = qf + 2 * d; A.ldv = 3 * d;
            A.pos_proj = nullptr; A.row_pos = nullptr; A.pos_u = nullptr; A.pos_v = nullptr;
            A.q_start = seq_start_dev; A.q_len = seq_len_dev; A.k_start = seq_start_dev; A.k_len = seq_len_dev;
            A.batch = batch; A.heads = H; A.max_q_len = P.max_tp;
            A.chunk_size = 0; A.num_left_chunks = -1; A.scale = scale;
            A.out = ctx; A.ldo = lda; A.split3_out = 1;
            RC(attention_f32(A, st));
        } else {
        RC(gemm_bf16(a, d, &L.qkv.tmap, L.qkv.w, M, 3 * d, d, L.qkv.b, EPI_BF16, 1.0f, qkv, 3 * d, 0, st));
        {
            AttnArgs A;
            A.q = qkv; A.ldq = 3 * d; A.q_rows = M; A.q_col0 = 0;
            A.k = qkv; A.ldk = 3 * d; A.k_rows = M; A.k_col0 = d;
            A.v = qkv; A.ldv = 3 * d; A.v_rows = M; A.v_col0 = 2 * d;
            A.kbias = nullptr; A.ld_kbias = 0;
            A.q_start = seq_start_dev; A.q_len = seq_len_dev; A.k_start = seq_start_dev; A.k_len = seq_len_dev;
            A.batch = batch; A.heads = H; A.max_q_len = P.max_tp;
            A.chunk_size = 0; A.num_left_chunks = -1; A.scale = scale;
            A.out = ctx; A.ldo = d; A.out_col0 = 0; A.split3_out = 0; A.v_mode = 0;
            RC(attention_forward(A, st));
        }
        }
        RC(gemm_bf16(ctx, lda, &L.out.tmap, L.out.w, M, d, d * p3, L.out.b, EPI_RESID_F32, 1.0f, x, d, 0, st));
        RC(layernorm_rows(x, d, M, d, L.n2.g, L.n2.b, c.ln_eps, a, lda, sp, nullptr, 0, st));
        RC(gemm_bf16(a, lda, &L.ff1.tmap, L.ff1.w, M, ff, d * p3, L.ff1.b, EPI_BF16_GELU, 1.0f, h, ldh, sp, st));
        RC(gemm_bf16(h, ldh, &L.ff2.tmap, L.ff2.w, M, d, ff * p3, L.ff2.b, EPI_RESID_F32, 1.0f, x, d, 0, st));
    }
    RC(layernorm_rows(x, d, M, d, E.after.g, E.after.b, c.ln_eps, enc_out_bf16, lda, sp, enc_out, d, st));
    return WB_OK;
}

}  // namespace wb
