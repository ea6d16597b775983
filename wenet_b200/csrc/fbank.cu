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
#include <stdlib.h>
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
    const int* tap_km;
    const float* tap_w;
    int taps_per_lane;
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


// ------------------------------------------------------------------------------------------------------------------
// Register-resident variant for the front-end every recipe uses (25 ms / 10 ms at 16 kHz: frame 400, hop 160, 512-point
// real FFT = 256-point complex FFT of z[n] = y[2n] + i y[2n+1]).  One warp per frame, 8 complex points per lane, the FFT
// as three in-register stages (radix 8, 8, 4) with two conflict-free transposes through shared memory:
//   n = 32 a + b, k = c + 8 d, d = g + 8 h, b = 4 e + f:
//   X[c + 8 (g + 8 h)] = sum_f W4^(f h) W32^(f g) sum_e W8^(e g) [ W256^(b c) sum_a W8^(a c) z[32 a + b] ]
//   stage A  lane b,         registers a -> c   (twiddle W256^(b c), per-lane constants)
//   stage B  lane (c, f),    registers e -> g   (twiddle W32^(f g))
//   stage C  lane (c, g>>1), registers (g&1, f) -> (g&1, h)
// Window, twiddles and the frame's samples live in registers; the round-1 kernel ran four radix-4 passes over a
// shared-memory buffer (2190 warp instructions and ~700 bank conflicts per frame, ncu profiles/r2_ncu_fbank.txt).
// Shared-memory index maps (float2 units) are chosen so that every 64-bit access of a warp touches each bank pair
// exactly twice (the minimum):  T1[36 c + b],  T2[c + 36 g + 8 f],  Z[24 (k >> 4) + (k & 15)].
constexpr int F5_FR = 16;          // frames per CTA (4 per warp)
constexpr int F5_THREADS = 128;
constexpr int F5_T1 = 384;         // float2 slots of the T1 / Z buffer (Z rows are padded to 24)
constexpr int F5_WARP_FLOATS = 2 * F5_T1 + 2 * 288;   // T1 / Z + T2 / power spectrum (288 float2)

__device__ __forceinline__ float2 cmul(const float2 a, const float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(const float2 a, const float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(const float2 a, const float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul_mi(const float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// in-place 4-point DFT, natural order in and out: y[m] = sum_j x[j] W4^(j m)
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = cmul_mi(csub(x1, x3));
    x0 = cadd(t0, t2);
    x1 = cadd(t1, t3);
    x2 = csub(t0, t2);
    x3 = csub(t1, t3);
}
// in-place 8-point DFT, natural order in and out: y[c] = sum_a x[a] W8^(a c)
__device__ __forceinline__ void dft8(float2 (&x)[8]) {
    constexpr float kR = 0.70710678118654752440f;
    float2 s0 = cadd(x[0], x[4]), s1 = cadd(x[1], x[5]), s2 = cadd(x[2], x[6]), s3 = cadd(x[3], x[7]);
    float2 d0 = csub(x[0], x[4]), d1 = csub(x[1], x[5]), d2 = csub(x[2], x[6]), d3 = csub(x[3], x[7]);
    // d_a * W8^a:  W8 = (1 - i) / sqrt 2,  W8^2 = -i,  W8^3 = (-1 - i) / sqrt 2
    d1 = make_float2(kR * (d1.x + d1.y), kR * (d1.y - d1.x));
    d2 = cmul_mi(d2);
    d3 = make_float2(kR * (d3.y - d3.x), -kR * (d3.x + d3.y));
    dft4(s0, s1, s2, s3);   // even outputs
    dft4(d0, d1, d2, d3);   // odd outputs
    x[0] = s0; x[2] = s1; x[4] = s2; x[6] = s3;
    x[1] = d0; x[3] = d1; x[5] = d2; x[7] = d3;
}

template <typename T>
struct PcmPair;   // the sample pair (2 n, 2 n + 1) of a frame as floats
template <>
struct PcmPair<int16_t> {
    static __device__ __forceinline__ float2 load(const int16_t* x, int n) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(x + 2 * n);
        return make_float2((float)(int16_t)(w & 0xffffu), (float)(int16_t)(w >> 16));
    }
};
template <>
struct PcmPair<float> {
    static __device__ __forceinline__ float2 load(const float* x, int n) {
        return *reinterpret_cast<const float2*>(x + 2 * n);
    }
};

template <typename T>
__global__ void __launch_bounds__(F5_THREADS)
fbank512_kernel(FbankDev P, const T* __restrict__ pcm, long long pcm_stride, const int* __restrict__ num_samples,
                float scale, float* __restrict__ out, long long out_frames_stride, int max_frames) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int b_utt = blockIdx.y;
    const int f0 = blockIdx.x * F5_FR;
    const int ns = num_samples[b_utt];
    const int flen = P.frame_len, shift = P.frame_shift;
    const int n_frames = (ns >= flen) ? 1 + (ns - flen) / shift : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* out_b = out + (long long)b_utt * out_frames_stride * P.num_mel;
    if (f0 >= n_frames) {   // frames past the utterance: zeros (processor.padding)
        for (int i = threadIdx.x; i < F5_FR * P.num_mel; i += F5_THREADS) {
            const int f = f0 + i / P.num_mel;
            if (f < max_frames) out_b[(long long)f * P.num_mel + (i % P.num_mel)] = 0.f;
        }
        return;
    }
    const int span = (F5_FR - 1) * shift + flen;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    T* s_pcm = reinterpret_cast<T*>(smem_raw + 16);
    const int pcm_bytes = (span * (int)sizeof(T) + 15) & ~15;
    float2* s_t1 = reinterpret_cast<float2*>(smem_raw + 16 + pcm_bytes) + warp * (F5_WARP_FLOATS / 2);
    float2* s_t2 = s_t1 + F5_T1;
    float* s_pow = reinterpret_cast<float*>(s_t2);   // the power spectrum reuses T2 once stage C has read it (264 floats)

    const int nf_here = min(F5_FR, n_frames - f0);
    const int need = (nf_here - 1) * shift + flen;
    const T* src = pcm + (long long)b_utt * pcm_stride + (long long)f0 * shift;
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
    }
    // per-lane constants, computed while the samples are in flight
    const int cB = lane >> 2, fB = lane & 3;      // stage B / C lane coordinates: c, and f (B) or g >> 1 (C)
    float2 win[8];      // window[64 a + 2 b], window[64 a + 2 b + 1]
    float2 twA[8];      // W256^(b c)
    float2 twB[8];      // W32^(f g)
    const float2* twh = reinterpret_cast<const float2*>(P.twiddle);   // W256^k, k < 256
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int j = 64 * a + 2 * lane;
        win[a] = make_float2(j < flen ? __ldg(P.window + j) : 0.f, j + 1 < flen ? __ldg(P.window + j + 1) : 0.f);
        twA[a] = __ldg(twh + ((lane * a) & 255));
        twB[a] = __ldg(twh + ((8 * fB * a) & 255));
    }
    if (bulk_ok) {
        mbar_wait(bar, 0);
    } else {
        for (int i = threadIdx.x; i < need; i += F5_THREADS) s_pcm[i] = src[i];
        __syncthreads();
    }

    const float inv_len = 1.0f / (float)flen;
    const int nw = (flen + 1) >> 1;     // sample pairs per frame (200)
    for (int fi = warp; fi < F5_FR; fi += F5_THREADS / 32) {
        const int f = f0 + fi;
        if (f >= max_frames) break;
        float* orow = out_b + (long long)f * P.num_mel;
        if (fi >= nf_here) {
            for (int m = lane; m < P.num_mel; m += 32) orow[m] = 0.f;
            continue;
        }
        const T* x = s_pcm + fi * shift;
        // 1. samples -> registers; mean (remove_dc_offset, kaldi.py:184-186)
        float2 v[8];
        float prev[8];   // sample 2 n - 1 (replicate pad at n = 0, kaldi.py:194-198)
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int n = 32 * a + lane;
            v[a] = make_float2(0.f, 0.f);
            prev[a] = 0.f;
            if (n < nw) {
                float2 cur = PcmPair<T>::load(x, n);
                if (2 * n + 1 >= flen) cur.y = 0.f;     // odd frame length: the last pair is half empty
                const float pv = (n > 0) ? PcmPair<T>::load(x, n - 1).y : cur.x;
                v[a] = make_float2(cur.x * scale, cur.y * scale);
                prev[a] = pv * scale;
                acc += v[a].x + ((2 * n + 1 < flen) ? v[a].y : 0.f);
            }
        }
        const float mean = warp_sum(acc) * inv_len;
        // 2. DC removal, pre-emphasis, window  ->  z[n] = y[2 n] + i y[2 n + 1]
        float2 z[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const float x0 = v[a].x - mean, x1 = v[a].y - mean, xm = prev[a] - mean;
            z[a] = make_float2((x0 - P.preemph * xm) * win[a].x, (x1 - P.preemph * x0) * win[a].y);   // win = 0 past the frame
        }
        // 3. stage A: 8-point DFT over a, twiddle W256^(b c), transpose
        dft8(z);
        s_t1[lane] = z[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) s_t1[36 * c + lane] = cmul(z[c], twA[c]);
        __syncwarp();
        // stage B: lane (c, f): 8-point DFT over e, twiddle W32^(f g), transpose
#pragma unroll
        for (int e = 0; e < 8; ++e) z[e] = s_t1[36 * cB + 4 * e + fB];
        dft8(z);
        s_t2[cB + 8 * fB] = z[0];
#pragma unroll
        for (int g = 1; g < 8; ++g) s_t2[cB + 36 * g + 8 * fB] = cmul(z[g], twB[g]);
        __syncwarp();
        // stage C: lane (c, gh): two 4-point DFTs over f;  k = c + 8 g + 64 h
#pragma unroll
        for (int gl = 0; gl < 2; ++gl) {
            const int g = 2 * fB + gl;
            float2 y0 = s_t2[cB + 36 * g], y1 = s_t2[cB + 36 * g + 8], y2 = s_t2[cB + 36 * g + 16], y3 = s_t2[cB + 36 * g + 24];
            dft4(y0, y1, y2, y3);
            z[4 * gl + 0] = y0;
            z[4 * gl + 1] = y1;
            z[4 * gl + 2] = y2;
            z[4 * gl + 3] = y3;
        }
        __syncwarp();   // all T1 / T2 reads done: T1 becomes Z, T2 becomes the power spectrum
#pragma unroll
        for (int gl = 0; gl < 2; ++gl)
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int k = cB + 8 * (2 * fB + gl) + 64 * h;
                s_t1[24 * (k >> 4) + (k & 15)] = z[4 * gl + h];
            }
        __syncwarp();
        // 4. real-FFT split + power:  X[k] = (Z[k] + conj(Z[256 - k])) / 2 - i W512^k (Z[k] - conj(Z[256 - k])) / 2
        const float2* twr = reinterpret_cast<const float2*>(P.twiddle_r);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int k = lane + 32 * j;
            if (k <= 256) {
                const int ka = k & 255, kb = (256 - k) & 255;
                const float2 za = s_t1[24 * (ka >> 4) + (ka & 15)];
                const float2 zb = s_t1[24 * (kb >> 4) + (kb & 15)];
                const float er = 0.5f * (za.x + zb.x), ei = 0.5f * (za.y - zb.y);
                const float orr = 0.5f * (za.x - zb.x), oi = 0.5f * (za.y + zb.y);
                const float2 w = __ldg(twr + k);
                // -i (or + i oi) = oi - i or
                const float xr = er + (oi * w.x + orr * w.y);
                const float xi = ei + (oi * w.y - orr * w.x);
                s_pow[k] = xr * xr + xi * xi;
            }
        }
        __syncwarp();
        // 5. mel + log (kaldi.py:620-633).  The triangles of the upper bins are 8 x wider than those of the lower ones, so
        //    the bins are dealt to the lanes by a longest-first greedy partition made at plan creation (<= 4 bins per
        //    lane, ~equal tap counts) instead of bin m -> lane m % 32.
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = __ldg(P.tap_km + lane * 4 + j);     // this lane's j-th mel bin or -1
            if (m >= 0) {
                const int st = __ldg(P.mel_start + m), ln = __ldg(P.mel_len + m);
                const float* w = P.mel_w + __ldg(P.mel_off + m);
                float e = 0.f;
#pragma unroll 4
                for (int i = 0; i < ln; ++i) e = fmaf(s_pow[st + i], __ldg(w + i), e);
                orow[m] = logf(fmaxf(e, 1.1920928955078125e-07f));
            }
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
    // fbank512_kernel: mel bins dealt to the 32 lanes, longest filter first, always to the least loaded lane that still
    // has a free slot (4 per lane): tap_km[lane * 4 + j] = j-th bin of the lane or -1
    std::vector<int> tkm(32 * 4, -1);
    std::vector<float> tw_(1, 0.f);
    if (num_mel <= 128) {   // (the fast kernel is only dispatched for <= 128 bins)
        std::vector<int> order(num_mel), load(32, 0), cnt(32, 0);
        for (int m = 0; m < num_mel; ++m) order[m] = m;
        for (int i = 0; i < num_mel; ++i)          // selection sort by length desc (num_mel is small)
            for (int j = i + 1; j < num_mel; ++j)
                if (ln[order[j]] > ln[order[i]]) std::swap(order[i], order[j]);
        for (int i = 0; i < num_mel; ++i) {
            int best = -1;
            for (int l = 0; l < 32; ++l)
                if (cnt[l] < 4 && (best < 0 || load[l] < load[best])) best = l;
            tkm[best * 4 + cnt[best]++] = order[i];
            load[best] += ln[order[i]] + 4;        // + fixed cost per bin (loads of start / len / off, log, store)
        }
    }
    p->taps_per_lane = 4;
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
    WB_UP(p->tap_km, tkm, int);
    WB_UP(p->tap_w, tw_, float);
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
    cudaFree(p->tap_km);
    cudaFree(p->tap_w);
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
    P.tap_km = plan->tap_km;
    P.tap_w = plan->tap_w;
    P.taps_per_lane = plan->taps_per_lane;
    P.num_mel = plan->num_mel;
    P.frame_len = plan->frame_len;
    P.frame_shift = plan->frame_shift;
    P.nfft = plan->nfft;
    P.preemph = plan->preemph;
    const int esz = is_int16 ? 2 : 4;
    ProfScope _ps(PT_FBANK, stream,
                  (double)batch * ((double)max_frames * plan->frame_shift * esz + (double)max_frames * plan->num_mel * 4.0));
    // the recipes' front-end (512-point FFT, even hop so that sample pairs stay aligned): register-resident kernel
    static const bool no_fast = getenv("WB_FBANK_GENERIC") != nullptr;
    if (plan->nfft == 512 && plan->frame_shift % 2 == 0 && plan->frame_len > 256 && plan->num_mel <= 128 && !no_fast) {
        const int span = (F5_FR - 1) * plan->frame_shift + plan->frame_len;
        const int pcm_bytes = (span * esz + 15) & ~15;
        const size_t smem = 16 + pcm_bytes + (size_t)(F5_THREADS / 32) * F5_WARP_FLOATS * sizeof(float);
        dim3 grid(ceil_div(max_frames, F5_FR), batch);
        if (is_int16) {
            if (smem > 48 * 1024)
                WB_CHECK_CUDA(cudaFuncSetAttribute(fbank512_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fbank512_kernel<int16_t><<<grid, F5_THREADS, smem, stream>>>(P, (const int16_t*)pcm, pcm_stride, num_samples_dev,
                                                                         scale, out, out_frames_stride, max_frames);
        } else {
            if (smem > 48 * 1024)
                WB_CHECK_CUDA(cudaFuncSetAttribute(fbank512_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fbank512_kernel<float><<<grid, F5_THREADS, smem, stream>>>(P, (const float*)pcm, pcm_stride, num_samples_dev, scale,
                                                                       out, out_frames_stride, max_frames);
        }
        count_launch();
        WB_CHECK_LAUNCH();
        return WB_OK;
    }
    const int half = plan->nfft / 2;
    const int span = (FR - 1) * plan->frame_shift + plan->frame_len;
    const int pcm_bytes = (span * esz + 15) & ~15;
    const size_t smem = 16 + pcm_bytes + (size_t)(FB_THREADS / 32) * (2 * half + half + 8) * sizeof(float);
    dim3 grid(ceil_div(max_frames, FR), batch);
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
