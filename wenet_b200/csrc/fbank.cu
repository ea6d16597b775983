// Fused Kaldi-compatible log-mel filterbank front-end (one kernel):
//   PCM -> frames (snip_edges) -> DC removal -> pre-emphasis -> povey window -> zero-pad to nfft ->
//   real FFT -> |X|^2 -> mel filterbank (sparse triangles) -> log(max(., eps))
//
// Replaces wenet/dataset/processor.py:226-256 (compute_fbank) -> torchaudio.compliance.kaldi.fbank
// (torchaudio/compliance/kaldi.py:514-645, helpers :44-217, mel banks :436-511).
//
// Layout / schedule:  a CTA (128 threads) handles FR consecutive frames of one utterance.  The
// (FR-1)*shift + frame_len samples they cover are contiguous in HBM and are staged into shared
// memory by ONE bulk-async copy (cp.async.bulk, mbarrier completion) — every sample is read from HBM
// once although adjacent frames overlap by 60 %.  Each warp then runs whole frames: warp-cooperative
// mean, pre-emphasis + window into a packed complex buffer, nfft/2-point radix-2 DIF FFT in shared
// memory, real-FFT split, power spectrum, sparse mel accumulation, log, coalesced store.
// HBM traffic = 4 B/sample in + 4*num_mel B/frame out (2.88 MB per 30 s utterance).
#include "common.cuh"
#include "kernels.h"
#include <math.h>
#include <vector>

namespace wb {

namespace {

constexpr int FR = 8;         // frames per CTA
constexpr int FB_THREADS = 128;
constexpr int MAX_NFFT = 512;

struct FbankDev {
    const float* window;
    const float* twiddle;    // nfft/2 complex: exp(-2*pi*i*k/(nfft/2)), k < nfft/2
    const float* twiddle_r;  // nfft/2+1 complex: exp(-2*pi*i*k/nfft)
    const int* mel_start;
    const int* mel_len;
    const int* mel_off;
    const float* mel_w;
    int num_mel, frame_len, frame_shift, nfft;
    float preemph;
};

__device__ __forceinline__ int bitrev(int x, int bits) { return (int)(__brev((unsigned)x) >> (32 - bits)); }
// base-4 digit reversal of a `bits`-bit index (bits even): bit reversal with the two bits of every digit swapped back
__device__ __forceinline__ int digitrev4(int x, int bits) {
    const unsigned r = __brev((unsigned)x) >> (32 - bits);
    return (int)(((r & 0xAAAAAAAAu) >> 1) | ((r & 0x55555555u) << 1));
}

template <typename T>
__device__ __forceinline__ float load_sample(const T* p, int i);
template <>
__device__ __forceinline__ float load_sample<float>(const float* p, int i) { return p[i]; }
template <>
__device__ __forceinline__ float load_sample<int16_t>(const int16_t* p, int i) { return (float)p[i]; }

template <typename T>
__global__ void __launch_bounds__(FB_THREADS)
fbank_kernel(FbankDev P, const T* __restrict__ pcm, long long pcm_stride, const int* __restrict__ num_samples,
             float scale, float* __restrict__ out, long long out_frames_stride, int max_frames) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * FR;
    const int ns = num_samples[b];
    const int n_frames = (ns >= P.frame_len) ? 1 + (ns - P.frame_len) / P.frame_shift : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = FB_THREADS / 32;
    const int half = P.nfft / 2;          // complex FFT size
    const int span = (FR - 1) * P.frame_shift + P.frame_len;

    // smem carve: [mbarrier 16B][samples: span * sizeof(T), padded][per warp: half complex + (half+1) power]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    T* s_pcm = reinterpret_cast<T*>(smem_raw + 16);
    const int pcm_bytes = (span * (int)sizeof(T) + 15) & ~15;
    float* s_work = reinterpret_cast<float*>(smem_raw + 16 + pcm_bytes);
    float* w_fft = s_work + warp * (2 * half + half + 8);  // re/im interleaved (2*half) + power (half+1)
    float* w_pow = w_fft + 2 * half;

    float* out_b = out + (long long)b * out_frames_stride * P.num_mel;

    if (f0 >= n_frames) {
        // zero-fill padded frames so downstream padded-batch consumers see zeros (processor.padding)
        for (int i = threadIdx.x; i < FR * P.num_mel; i += FB_THREADS) {
            const int f = f0 + i / P.num_mel;
            if (f < max_frames) out_b[(long long)f * P.num_mel + (i % P.num_mel)] = 0.f;
        }
        return;
    }
    const int nf_here = min(FR, n_frames - f0);
    const int need = (nf_here - 1) * P.frame_shift + P.frame_len;  // samples actually needed
    const T* src = pcm + (long long)b * pcm_stride + (long long)f0 * P.frame_shift;
    const uint32_t bytes = (uint32_t)need * (uint32_t)sizeof(T);
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15) == 0);
    if (bulk_ok) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, bytes);
            bulk_load_1d(s_pcm, src, bytes, bar);
        }
        mbar_wait(bar, 0);
    } else {
        for (int i = threadIdx.x; i < need; i += FB_THREADS) s_pcm[i] = src[i];
        __syncthreads();
    }

    const int log2half = 31 - __clz(half);
    const bool radix4 = (log2half & 1) == 0;   // half is a power of 4
    for (int fi = warp; fi < FR; fi += nwarps) {
        const int f = f0 + fi;
        if (f >= max_frames) break;
        float* orow = out_b + (long long)f * P.num_mel;
        if (fi >= nf_here) {
            for (int m = lane; m < P.num_mel; m += 32) orow[m] = 0.f;
            continue;
        }
        const T* x = s_pcm + fi * P.frame_shift;
        // 1. mean (remove_dc_offset, kaldi.py:184-186)
        float acc = 0.f;
        for (int j = lane; j < P.frame_len; j += 32) acc += load_sample<T>(x, j) * scale;
        const float mean = warp_sum(acc) / (float)P.frame_len;
        // 2. pre-emphasis (replicate pad, kaldi.py:194-198) * window, zero pad, pack z[n] = y[2n] + i y[2n+1]
        for (int j = lane; j < P.nfft; j += 32) {
            float y = 0.f;
            if (j < P.frame_len) {
                const float cur = load_sample<T>(x, j) * scale - mean;
                const float prev = load_sample<T>(x, j > 0 ? j - 1 : 0) * scale - mean;
                y = (cur - P.preemph * prev) * P.window[j];
            }
            w_fft[j] = y;  // natural order: (re, im) of z[j/2] interleaved == y itself
        }
        __syncwarp();
        // 3. half-point complex FFT, decimation in frequency (natural in, digit-reversed out).  half = 4^n (the 25 ms /
        //    16 kHz front-end: 512-point real FFT -> 256 complex points) runs radix-4 on float2 elements: a quarter of
        //    the shared-memory instructions of the radix-2 loop, which was bound by exactly those.
        if (radix4) {
            float2* z = reinterpret_cast<float2*>(w_fft);
            const float2* twd = reinterpret_cast<const float2*>(P.twiddle);
            for (int ns = half; ns >= 4; ns >>= 2) {
                const int q = ns >> 2;             // quarter of the current sub-transform
                const int tws = half / ns;         // twiddle stride: W_ns^j = W_half^(j * tws)
                for (int t = lane; t < half / 4; t += 32) {
                    const int grp = t / q, pos = t - grp * q;
                    const int i0 = grp * ns + pos;
                    const float2 x0 = z[i0], x1 = z[i0 + q], x2 = z[i0 + 2 * q], x3 = z[i0 + 3 * q];
                    const float2 a0 = make_float2(x0.x + x2.x, x0.y + x2.y);
                    const float2 a1 = make_float2(x0.x - x2.x, x0.y - x2.y);
                    const float2 a2 = make_float2(x1.x + x3.x, x1.y + x3.y);
                    const float2 a3 = make_float2(x1.y - x3.y, -(x1.x - x3.x));   // -i (x1 - x3)
                    const float2 y0 = make_float2(a0.x + a2.x, a0.y + a2.y);
                    const float2 y1 = make_float2(a1.x + a3.x, a1.y + a3.y);
                    const float2 y2 = make_float2(a0.x - a2.x, a0.y - a2.y);
                    const float2 y3 = make_float2(a1.x - a3.x, a1.y - a3.y);
                    const float2 w1 = twd[pos * tws], w2 = twd[2 * pos * tws], w3 = twd[3 * pos * tws];
                    z[i0] = y0;
                    z[i0 + q] = make_float2(y1.x * w1.x - y1.y * w1.y, y1.x * w1.y + y1.y * w1.x);
                    z[i0 + 2 * q] = make_float2(y2.x * w2.x - y2.y * w2.y, y2.x * w2.y + y2.y * w2.x);
                    z[i0 + 3 * q] = make_float2(y3.x * w3.x - y3.y * w3.y, y3.x * w3.y + y3.y * w3.x);
                }
                __syncwarp();
            }
        } else {
            for (int s = 0; s < log2half; ++s) {
                const int hs = half >> (s + 1);  // butterfly half-span
                for (int t = lane; t < half / 2; t += 32) {
                    const int grp = t / hs, pos = t - grp * hs;
                    const int i0 = grp * 2 * hs + pos, i1 = i0 + hs;
                    const int tw = pos << s;  // twiddle index in units of 2*pi/half
                    const float wr = P.twiddle[2 * tw], wi = P.twiddle[2 * tw + 1];
                    const float ar = w_fft[2 * i0], ai = w_fft[2 * i0 + 1];
                    const float br = w_fft[2 * i1], bi = w_fft[2 * i1 + 1];
                    w_fft[2 * i0] = ar + br;
                    w_fft[2 * i0 + 1] = ai + bi;
                    const float dr = ar - br, di = ai - bi;
                    w_fft[2 * i1] = dr * wr - di * wi;
                    w_fft[2 * i1 + 1] = dr * wi + di * wr;
                }
                __syncwarp();
            }
        }
        // 4. real-FFT split + power:  X[k] = (Z[k] + conj(Z[h-k]))/2 - i*W^k*(Z[k] - conj(Z[h-k]))/2
        for (int k = lane; k <= half; k += 32) {
            const int ka = radix4 ? digitrev4(k & (half - 1), log2half) : bitrev(k & (half - 1), log2half);
            const int kb = radix4 ? digitrev4((half - k) & (half - 1), log2half)
                                  : bitrev((half - k) & (half - 1), log2half);
            const float zr = w_fft[2 * ka], zi = w_fft[2 * ka + 1];
            const float yr = w_fft[2 * kb], yi = -w_fft[2 * kb + 1];  // conj(Z[h-k])
            const float er = 0.5f * (zr + yr), ei = 0.5f * (zi + yi);
            const float or_ = 0.5f * (zr - yr), oi = 0.5f * (zi - yi);
            // -i * (or + i oi) = oi - i or
            const float tr = oi, ti = -or_;
            const float wr = P.twiddle_r[2 * k], wi = P.twiddle_r[2 * k + 1];
            const float xr = er + (tr * wr - ti * wi);
            const float xi = ei + (tr * wi + ti * wr);
            w_pow[k] = xr * xr + xi * xi;
        }
        __syncwarp();
        // 5. mel + log (kaldi.py:620-633)
        for (int m = lane; m < P.num_mel; m += 32) {
            const int st = P.mel_start[m], ln = P.mel_len[m];
            const float* w = P.mel_w + P.mel_off[m];
            float e = 0.f;
            for (int i = 0; i < ln; ++i) e = fmaf(w_pow[st + i], w[i], e);
            orow[m] = logf(fmaxf(e, 1.1920928955078125e-07f));
        }
        __syncwarp();
    }
}

}  // namespace

int fbank_plan_create(FbankPlan** out, int sample_rate, int num_mel, int frame_len, int frame_shift,
                      float low_freq, float preemph, const float* window_host, const float* mel_dense_host) {
    (void)sample_rate;
    (void)low_freq;
    int nfft = 1;
    while (nfft < frame_len) nfft <<= 1;
    WB_REQUIRE(nfft <= MAX_NFFT && nfft >= 64, WB_ERR_UNSUPPORTED, "fbank: nfft %d unsupported", nfft);
    FbankPlan* p = new FbankPlan();
    memset(p, 0, sizeof(*p));
    p->num_mel = num_mel;
    p->frame_len = frame_len;
    p->frame_shift = frame_shift;
    p->nfft = nfft;
    p->preemph = preemph;
    const int half = nfft / 2;
    std::vector<float> tw(2 * half), twr(2 * (half + 1));   // radix-4 needs W^k up to k < 3 half / 4
    for (int k = 0; k < half; ++k) {
        const double a = -2.0 * M_PI * k / half;
        tw[2 * k] = (float)cos(a);
        tw[2 * k + 1] = (float)sin(a);
    }
    for (int k = 0; k <= half; ++k) {
        const double a = -2.0 * M_PI * k / nfft;
        twr[2 * k] = (float)cos(a);
        twr[2 * k + 1] = (float)sin(a);
    }
    std::vector<int> st(num_mel), ln(num_mel), off(num_mel);
    std::vector<float> w;
    const int nb = half + 1;
    for (int m = 0; m < num_mel; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < nb; ++k)
            if (mel_dense_host[(size_t)m * nb + k] != 0.f) {
                if (first < 0) first = k;
                last = k;
            }
        if (first < 0) {
            first = 0;
            last = -1;
        }
        st[m] = first;
        ln[m] = last - first + 1;
        off[m] = (int)w.size();
        for (int k = first; k <= last; ++k) w.push_back(mel_dense_host[(size_t)m * nb + k]);
    }
    if (w.empty()) w.push_back(0.f);
    p->mel_nnz = (int)w.size();
#define WB_UP(dst, vec, T)                                                                       \
    WB_CHECK_CUDA(cudaMalloc((void**)&dst, (vec).size() * sizeof(T)));                             \
    WB_CHECK_CUDA(cudaMemcpy(dst, (vec).data(), (vec).size() * sizeof(T), cudaMemcpyHostToDevice));
    std::vector<float> win(window_host, window_host + frame_len);
    WB_UP(p->window, win, float);
    WB_UP(p->twiddle, tw, float);
    WB_UP(p->twiddle_r, twr, float);
    WB_UP(p->mel_start, st, int);
    WB_UP(p->mel_len, ln, int);
    WB_UP(p->mel_off, off, int);
    WB_UP(p->mel_w, w, float);
#undef WB_UP
    *out = p;
    return WB_OK;
}

void fbank_plan_destroy(FbankPlan* p) {
    if (!p) return;
    cudaFree(p->window);
    cudaFree(p->twiddle);
    cudaFree(p->twiddle_r);
    cudaFree(p->mel_start);
    cudaFree(p->mel_len);
    cudaFree(p->mel_off);
    cudaFree(p->mel_w);
    delete p;
}

int fbank_forward(const FbankPlan* plan, const void* pcm, int is_int16, long long pcm_stride,
                  const int* num_samples_dev, int batch, float scale, float* out, long long out_frames_stride,
                  int max_frames, cudaStream_t stream) {
    WB_REQUIRE(plan != nullptr, WB_ERR_BAD_ARG, "fbank: null plan");
    if (batch <= 0 || max_frames <= 0) return WB_OK;
    FbankDev P;
    P.window = plan->window;
    P.twiddle = plan->twiddle;
    P.twiddle_r = plan->twiddle_r;
    P.mel_start = plan->mel_start;
    P.mel_len = plan->mel_len;
    P.mel_off = plan->mel_off;
    P.mel_w = plan->mel_w;
    P.num_mel = plan->num_mel;
    P.frame_len = plan->frame_len;
    P.frame_shift = plan->frame_shift;
    P.nfft = plan->nfft;
    P.preemph = plan->preemph;
    const int half = plan->nfft / 2;
    const int span = (FR - 1) * plan->frame_shift + plan->frame_len;
    const int esz = is_int16 ? 2 : 4;
    const int pcm_bytes = (span * esz + 15) & ~15;
    const size_t smem = 16 + pcm_bytes + (size_t)(FB_THREADS / 32) * (2 * half + half + 8) * sizeof(float);
    dim3 grid(ceil_div(max_frames, FR), batch);
    ProfScope _ps(PT_FBANK, stream,
                  (double)batch * ((double)max_frames * plan->frame_shift * esz + (double)max_frames * plan->num_mel * 4.0));
    if (is_int16) {
        if (smem > 48 * 1024)
            WB_CHECK_CUDA(cudaFuncSetAttribute(fbank_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem));
        fbank_kernel<int16_t><<<grid, FB_THREADS, smem, stream>>>(P, (const int16_t*)pcm, pcm_stride,
                                                                  num_samples_dev, scale, out,
                                                                  out_frames_stride, max_frames);
    } else {
        if (smem > 48 * 1024)
            WB_CHECK_CUDA(cudaFuncSetAttribute(fbank_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem));
        fbank_kernel<float><<<grid, FB_THREADS, smem, stream>>>(P, (const float*)pcm, pcm_stride,
                                                                num_samples_dev, scale, out, out_frames_stride,
                                                                max_frames);
    }
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
