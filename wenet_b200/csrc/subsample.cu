// Conv2dSubsampling4 front: GlobalCMVN + Conv2d(1->d, 3x3, stride 2) + ReLU as a direct kernel
// (C_in = 1: 9 MACs per output, bandwidth-bound on its bf16 output), and the im2col gather that
// turns Conv2d(d->d, 3x3, stride 2) into a tcgen05 GEMM with K = 9*d.
// Replaces wenet/models/transformer/cmvn.py:36-47 and subsampling.py:203-228 (first half).
//
// Activations are channels-last: conv1 output row (b, t1, f1) holds d contiguous bf16 channels, so
// each (kh, kw) tap of conv2's im2col row is one contiguous d*2-byte vector.
#include "common.cuh"
#include "kernels.h"

namespace wb {

namespace {

// grid (max_t1, B); block d/2 threads; thread owns channels (2*tid, 2*tid+1)
__global__ void conv1_kernel(const float* __restrict__ feats, long long feat_stride_b, int idim,
                             const int* __restrict__ t1_len, const long long* __restrict__ off1,
                             const float* __restrict__ mean, const float* __restrict__ istd,
                             const float* __restrict__ w, const float* __restrict__ bias, int d,
                             __nv_bfloat16* __restrict__ out1, int split3) {
    extern __shared__ __align__(8) float s_in[];  // [3][idim_pad] (even pitch: the pair loads below are 8-byte aligned)
    const int idim_pad = (idim + 1) & ~1;
    const int b = blockIdx.y, t1 = blockIdx.x;
    if (t1 >= t1_len[b]) return;
    const int F1 = (idim - 3) / 2 + 1;
    const float* src = feats + (long long)b * feat_stride_b + (long long)(2 * t1) * idim;
    for (int i = threadIdx.x; i < 3 * idim; i += blockDim.x) {
        float v = src[i];
        const int r = i / idim, f = i - r * idim;
        if (mean != nullptr) v = (v - mean[f]) * istd[f];
        s_in[r * idim_pad + f] = v;
    }
    __syncthreads();
    const int c0 = 2 * threadIdx.x;
    if (c0 >= d) return;
    // the thread's two channels ride in one packed fp32x2 register pair: 9 FFMA2 per output instead of 18 FFMA (the
    // kernel is bound by instruction issue: 9 MACs per bf16 written)
    float2 wv[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wv[k] = make_float2(w[k * d + c0], w[k * d + c0 + 1]);
    const float2 bv = make_float2(bias[c0], bias[c0 + 1]);
    const int ldo = split3 ? 3 * d : d;
    __nv_bfloat16* orow = out1 + (off1[b] + (long long)t1 * F1) * ldo + c0;
    for (int f1 = 0; f1 < F1; ++f1) {
        float2 acc = bv;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            // inputs 2 f1 .. 2 f1 + 2 of row kh: an aligned pair (broadcast LDS.64) + one scalar
            const float2 x01 = *reinterpret_cast<const float2*>(s_in + kh * idim_pad + 2 * f1);
            const float x2 = s_in[kh * idim_pad + 2 * f1 + 2];
            acc = f2_fma(wv[kh * 3 + 0], make_float2(x01.x, x01.x), acc);
            acc = f2_fma(wv[kh * 3 + 1], make_float2(x01.y, x01.y), acc);
            acc = f2_fma(wv[kh * 3 + 2], make_float2(x2, x2), acc);
        }
        const float a0 = fmaxf(acc.x, 0.f), a1 = fmaxf(acc.y, 0.f);
        const uint32_t hi = pack_bf16x2(a0, a1);
        *reinterpret_cast<uint32_t*>(orow + (long long)f1 * ldo) = hi;
        if (split3) {
            *reinterpret_cast<uint32_t*>(orow + (long long)f1 * ldo + d) = pack_bf16x2(a0 - bf16_lo(hi), a1 - bf16_hi(hi));
            *reinterpret_cast<uint32_t*>(orow + (long long)f1 * ldo + 2 * d) = hi;
        }
    }
}

// grid (max_t2, B); 256 threads; copies 16-byte vectors
__global__ void im2col_kernel(const uint4* __restrict__ out1, const long long* __restrict__ off1,
                              const int* __restrict__ t2_len, const long long* __restrict__ off2, int F1,
                              int F2, int d, uint4* __restrict__ a2) {
    const int b = blockIdx.y, t2 = blockIdx.x;
    if (t2 >= t2_len[b]) return;
    const int dv = d / 8;  // uint4 per channel vector
    const int per_row = 9 * dv;
    const int total = F2 * per_row;
    const long long o1 = off1[b], o2 = off2[b];
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int f2 = i / per_row;
        const int r = i - f2 * per_row;
        const int tap = r / dv, cv = r - tap * dv;
        const int kh = tap / 3, kw = tap - kh * 3;
        const long long srow = o1 + (long long)(2 * t2 + kh) * F1 + (2 * f2 + kw);
        const long long drow = o2 + (long long)t2 * F2 + f2;
        a2[drow * per_row + tap * dv + cv] = out1[srow * dv + cv];
    }
}

}  // namespace

int subsample_conv1(const float* feats, long long feat_stride_b, int idim, const int* t1_len,
                    const long long* off1, int batch, int max_t1, const float* cmvn_mean,
                    const float* cmvn_istd, const float* w, const float* bias, int d, void* out1_bf16,
                    int split3, cudaStream_t stream) {
    if (batch <= 0 || max_t1 <= 0) return WB_OK;
    WB_REQUIRE(d % 64 == 0 && d <= 2048, WB_ERR_UNSUPPORTED, "conv1: d=%d unsupported", d);
    dim3 grid(max_t1, batch);
    ProfScope _ps(PT_CONV1, stream, (double)batch * max_t1 * (((idim - 3) / 2 + 1) * (double)d * 2.0 + 2.0 * idim * 4.0));
    conv1_kernel<<<grid, d / 2, 3 * ((idim + 1) & ~1) * sizeof(float), stream>>>(
        feats, feat_stride_b, idim, t1_len, off1, cmvn_mean, cmvn_istd, w, bias, d,
        reinterpret_cast<__nv_bfloat16*>(out1_bf16), split3);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int subsample_im2col(const void* out1_bf16, const long long* off1, const int* t2_len, const long long* off2,
                     int batch, int max_t2, int F1, int F2, int d, void* a2_bf16, int split3,
                     cudaStream_t stream) {
    if (batch <= 0 || max_t2 <= 0) return WB_OK;
    WB_REQUIRE(split3 == 0, WB_ERR_UNSUPPORTED, "im2col: split3 not supported");
    WB_REQUIRE(d % 8 == 0, WB_ERR_BAD_ARG, "im2col: d %% 8");
    dim3 grid(max_t2, batch);
    ProfScope _ps(PT_IM2COL, stream, (double)batch * max_t2 * F2 * 9.0 * d * 4.0);
    im2col_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(out1_bf16), off1, t2_len, off2, F1,
                                            F2, d, reinterpret_cast<uint4*>(a2_bf16));
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
