// ConvolutionModule tail: depthwise Conv1d (K taps, causal or symmetric) + LayerNorm-over-channels or
// folded BatchNorm + SiLU, fused in shared memory; output bf16 feeds pointwise_conv2 (a GEMM).
// Replaces wenet/models/transformer/convolution.py:119-147 (masked_fill / pad-or-cache /
// depthwise_conv / norm / activation).  pointwise_conv1 + GLU (:138-139) is the GEMM's GLU epilogue.
//
// HBM-bound: reads 2*d B and writes 2*d B per frame (+ (K-1)/TT halo re-read).
#include "common.cuh"
#include "kernels.h"

namespace wb {

namespace {

constexpr int TT = 32;        // output frames per CTA
constexpr int DW_THREADS = 256;
constexpr int MAX_K = 31;

struct DwDev {
    const __nv_bfloat16* g;
    long long ldg;
    const int* seq_start;
    const int* seq_len;
    const int* out_start;
    int lead, d, ksize, causal;
    const float* w;
    const float* bias;
    int norm_type;
    const float* gamma;
    const float* beta;
    float eps;
    const float* pad_vec;
    int pad_until;
    __nv_bfloat16* out;
    long long ldo;
    int split3;
};

// KT: compile-time kernel size (8 / 15 are the recipe values), 0 = run-time loop
template <int KT>
__global__ void __launch_bounds__(DW_THREADS)
dwconv_kernel(DwDev P) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int b = blockIdx.y;
    const int n_in = P.seq_len[b];
    const int n_out = n_in - P.lead;
    const int t0 = blockIdx.x * TT;
    if (t0 >= n_out) return;
    const int d = P.d, K = (KT > 0) ? KT : P.ksize;
    const int left = P.causal ? (K - 1) : (K - 1) / 2;
    const int rows_in = TT + K - 1;
    __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(smem_raw);           // [rows_in][d]
    float* s_out = reinterpret_cast<float*>(smem_raw + ((size_t)rows_in * d * 2 + 15) / 16 * 16);  // [TT][d]
    const long long base = P.seq_start[b];

    // stage input rows: input position p = lead + t0 - left + r
    const int dv = d / 8;
    for (int i = threadIdx.x; i < rows_in * dv; i += DW_THREADS) {
        const int r = i / dv, cv = i - r * dv;
        const int p = P.lead + t0 - left + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (p >= 0 && p < n_in) {
            v = *reinterpret_cast<const uint4*>(P.g + (base + p) * P.ldg + cv * 8);
        } else if (P.pad_vec != nullptr &&
                   ((p < 0 && P.causal) || (p >= n_in && !P.causal && p < P.pad_until + P.lead))) {
            // frames the reference zero-fills BEFORE pointwise_conv1 (causal left pad,
            // convolution.py:122-124; masked batch padding, :119-120) reach the depthwise conv as
            // GLU(pointwise_conv1(0)) = GLU(bias), not 0
            const float* pv = P.pad_vec + cv * 8;
            v = make_uint4(pack_bf16x2(pv[0], pv[1]), pack_bf16x2(pv[2], pv[3]), pack_bf16x2(pv[4], pv[5]),
                           pack_bf16x2(pv[6], pv[7]));
        }
        *reinterpret_cast<uint4*>(s_in + (size_t)r * d + cv * 8) = v;
    }
    __syncthreads();

    const int nt = min(TT, n_out - t0);
    // depthwise conv: thread = (channel pair, frame group); weights live in registers
    {
        const int pairs = d >> 1;
        const int ngroups = (DW_THREADS / pairs) > 0 ? (DW_THREADS / pairs) : 1;
        for (int idx = threadIdx.x; idx < pairs * ngroups; idx += DW_THREADS) {
            const int c = 2 * (idx % pairs);
            const int grp = idx / pairs;
            constexpr int KR = (KT > 0) ? KT : MAX_K;
            float w0[KR], w1[KR];
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                w0[k] = (k < K) ? P.w[(size_t)c * K + k] : 0.f;
                w1[k] = (k < K) ? P.w[(size_t)(c + 1) * K + k] : 0.f;
            }
            const float b0 = P.bias[c], b1 = P.bias[c + 1];
            float sc0 = 1.f, sc1 = 1.f, sh0 = 0.f, sh1 = 0.f;
            if (P.norm_type == 1) {
                sc0 = P.gamma[c]; sc1 = P.gamma[c + 1]; sh0 = P.beta[c]; sh1 = P.beta[c + 1];
            }
            for (int t = grp; t < nt; t += ngroups) {
                float a0 = b0, a1 = b1;
#pragma unroll
                for (int k = 0; k < KR; ++k) {
                    if (KT > 0 || k < K) {
                        const uint32_t xx = *reinterpret_cast<const uint32_t*>(s_in + (size_t)(t + k) * d + c);
                        a0 = fmaf(w0[k], bf16_lo(xx), a0);
                        a1 = fmaf(w1[k], bf16_hi(xx), a1);
                    }
                }
                if (P.norm_type == 1) {  // folded BatchNorm (eval): y = x*scale + shift, then SiLU
                    a0 = silu_f(fmaf(a0, sc0, sh0));
                    a1 = silu_f(fmaf(a1, sc1, sh1));
                }
                *reinterpret_cast<float2*>(s_out + (size_t)t * d + c) = make_float2(a0, a1);
            }
        }
    }
    __syncthreads();

    // norm (LayerNorm over channels) + SiLU + store: warp per frame
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = warp; t < nt; t += DW_THREADS / 32) {
        const float* row = s_out + (size_t)t * d;
        float mean = 0.f, rstd = 1.f;
        if (P.norm_type == 0) {
            float s = 0.f;
            for (int c = lane; c < d; c += 32) s += row[c];
            mean = warp_sum(s) / (float)d;
            float q = 0.f;
            for (int c = lane; c < d; c += 32) {
                const float z = row[c] - mean;
                q += z * z;
            }
            rstd = rsqrtf(warp_sum(q) / (float)d + P.eps);
        }
        __nv_bfloat16* o = P.out + ((long long)P.out_start[b] + t0 + t) * P.ldo;
        for (int c = 2 * lane; c < d; c += 64) {
            float y0 = row[c], y1 = row[c + 1];
            if (P.norm_type == 0) {
                y0 = silu_f((y0 - mean) * rstd * P.gamma[c] + P.beta[c]);
                y1 = silu_f((y1 - mean) * rstd * P.gamma[c + 1] + P.beta[c + 1]);
            }
            const uint32_t pk = pack_bf16x2(y0, y1);
            *reinterpret_cast<uint32_t*>(o + c) = pk;
            if (P.split3) {
                *reinterpret_cast<uint32_t*>(o + d + c) = pack_bf16x2(y0 - bf16_lo(pk), y1 - bf16_hi(pk));
                *reinterpret_cast<uint32_t*>(o + 2 * d + c) = pk;
            }
        }
    }
}

}  // namespace

int dwconv_norm_silu(const DwConvArgs& a, cudaStream_t stream) {
    if (a.batch <= 0 || a.max_len <= 0) return WB_OK;
    WB_REQUIRE(a.ksize >= 1 && a.ksize <= MAX_K, WB_ERR_UNSUPPORTED, "dwconv: kernel size %d unsupported", a.ksize);
    WB_REQUIRE(a.d % 8 == 0 && a.ldg % 8 == 0 && a.ldo % 2 == 0, WB_ERR_BAD_ARG, "dwconv: alignment");
    WB_REQUIRE(a.causal || (a.ksize % 2 == 1), WB_ERR_BAD_ARG, "dwconv: symmetric kernel must be odd");
    DwDev P;
    P.g = reinterpret_cast<const __nv_bfloat16*>(a.g);
    P.ldg = a.ldg;
    P.seq_start = a.seq_start;
    P.seq_len = a.seq_len;
    P.out_start = a.out_start;
    P.lead = a.lead;
    P.d = a.d;
    P.ksize = a.ksize;
    P.causal = a.causal;
    P.w = a.w;
    P.bias = a.bias;
    P.norm_type = a.norm_type;
    P.gamma = a.gamma;
    P.beta = a.beta;
    P.eps = a.eps;
    P.pad_vec = a.pad_vec;
    P.pad_until = a.pad_until;
    P.out = reinterpret_cast<__nv_bfloat16*>(a.out);
    P.ldo = a.ldo;
    P.split3 = a.split3;
    const int rows_in = TT + a.ksize - 1;
    const size_t smem = ((size_t)rows_in * a.d * 2 + 15) / 16 * 16 + (size_t)TT * a.d * sizeof(float);
    dim3 grid(ceil_div(a.max_len, TT), a.batch);
    ProfScope _ps(PT_DWCONV, stream, (double)a.batch * a.max_len * a.d * 4.0);
#define WB_DW(KT)                                                                                              \
    do {                                                                                                       \
        static size_t smem_set = 0;                                                                            \
        if (smem > 48 * 1024 && smem > smem_set) {                                                             \
            WB_CHECK_CUDA(cudaFuncSetAttribute(dwconv_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)smem));                                                    \
            smem_set = smem;                                                                                   \
        }                                                                                                      \
        dwconv_kernel<KT><<<grid, DW_THREADS, smem, stream>>>(P);                                              \
    } while (0)
    if (a.ksize == 8) WB_DW(8);
    else if (a.ksize == 15) WB_DW(15);
    else WB_DW(0);
#undef WB_DW
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
