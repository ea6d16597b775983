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

// KT: compile-time kernel size (8 / 15 are the recipe values), 0 = run-time loop.
//
// One warp owns FPW consecutive output frames for ALL channels, in passes of 128 channels (4 per lane): the
// (FPW + K - 1) input rows of a pass are read once from the staged tile and each value feeds up to FPW accumulators
// (sliding window in registers), the conv outputs stay in registers, LayerNorm statistics are two warp reductions per
// frame, and the 4-channel groups go through the quad sigmoid.  No intermediate fp32 tile, one __syncthreads.
constexpr int FPW = TT / (DW_THREADS / 32);   // 4 frames per warp
constexpr int MAX_PASSES = 4;                 // d <= 512

// NP: passes of 128 channels (d <= 128 * NP)
template <int KT, int NP>
__global__ void __launch_bounds__(DW_THREADS)
dwconv_kernel(DwDev P) {
    pdl_launch_dependents();
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
    float* s_w = reinterpret_cast<float*>(smem_raw + ((size_t)rows_in * d * 2 + 15) / 16 * 16);   // [K][d] (tap-major)
    const long long base = P.seq_start[b];

    // depthwise weights [d][K] -> shared memory, tap-major, so a lane fetches its 4 channels of tap k as one float4
    for (int i = threadIdx.x; i < d * K; i += DW_THREADS) {
        const int c = i / K, k = i - c * K;
        s_w[k * d + c] = P.w[i];
    }

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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw0 = warp * FPW;            // first frame (within the tile) of this warp
    if (tw0 >= nt) return;
    constexpr int KR = (KT > 0) ? KT : MAX_K;
    float acc[NP][FPW][4];
#pragma unroll
    for (int ps = 0; ps < NP; ++ps) {
        const int c0 = ps * 128 + 4 * lane;
        const bool act = c0 < d;   // d % 4 == 0: a lane's four channels are all inside or all outside
        {
            float w[4][KR];
#pragma unroll
            for (int k = 0; k < KR; ++k) {
                float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (act && (KT > 0 || k < K)) w4 = *reinterpret_cast<const float4*>(s_w + k * d + c0);
                w[0][k] = w4.x;
                w[1][k] = w4.y;
                w[2][k] = w4.z;
                w[3][k] = w4.w;
            }
            const float4 bb = act ? *reinterpret_cast<const float4*>(P.bias + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int f = 0; f < FPW; ++f) {
                acc[ps][f][0] = bb.x;
                acc[ps][f][1] = bb.y;
                acc[ps][f][2] = bb.z;
                acc[ps][f][3] = bb.w;
            }
#pragma unroll
            for (int r = 0; r < FPW + KR - 1; ++r) {
                if (KT > 0 || r < FPW + K - 1) {
                    const uint2 xx = act ? *reinterpret_cast<const uint2*>(s_in + (size_t)(tw0 + r) * d + c0) : make_uint2(0u, 0u);
                    const float x0 = bf16_lo(xx.x), x1 = bf16_hi(xx.x), x2 = bf16_lo(xx.y), x3 = bf16_hi(xx.y);
#pragma unroll
                    for (int f = 0; f < FPW; ++f) {
                        const int k = r - f;   // tap index of input row r for output frame tw0 + f
                        if (k >= 0 && k < KR && (KT > 0 || k < K)) {
                            acc[ps][f][0] = fmaf(w[0][k], x0, acc[ps][f][0]);
                            acc[ps][f][1] = fmaf(w[1][k], x1, acc[ps][f][1]);
                            acc[ps][f][2] = fmaf(w[2][k], x2, acc[ps][f][2]);
                            acc[ps][f][3] = fmaf(w[3][k], x3, acc[ps][f][3]);
                        }
                    }
                }
            }
        }
    }

    const float inv_d = 1.0f / (float)d;
#pragma unroll
    for (int f = 0; f < FPW; ++f) {
        const int t = tw0 + f;
        if (t >= nt) break;   // warp-uniform
        float mean = 0.f, rstd = 1.f;
        if (P.norm_type == 0) {
            float sm = 0.f;
#pragma unroll
            for (int ps = 0; ps < NP; ++ps) sm += (acc[ps][f][0] + acc[ps][f][1]) + (acc[ps][f][2] + acc[ps][f][3]);
            mean = warp_sum(sm) * inv_d;
            float q = 0.f;
#pragma unroll
            for (int ps = 0; ps < NP; ++ps) {
                if (ps * 128 + 4 * lane < d) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float z = acc[ps][f][j] - mean;
                        q = fmaf(z, z, q);
                    }
                }
            }
            rstd = rsqrtf(warp_sum(q) * inv_d + P.eps);
        }
        __nv_bfloat16* o = P.out + ((long long)P.out_start[b] + t0 + t) * P.ldo;
#pragma unroll
        for (int ps = 0; ps < NP; ++ps) {
            const int c0 = ps * 128 + 4 * lane;
            if (c0 < d) {
                const float4 g4 = *reinterpret_cast<const float4*>(P.gamma + c0);
                const float4 b4 = *reinterpret_cast<const float4*>(P.beta + c0);
                float y0, y1, y2, y3;
                if (P.norm_type == 0) {
                    y0 = (acc[ps][f][0] - mean) * rstd * g4.x + b4.x;
                    y1 = (acc[ps][f][1] - mean) * rstd * g4.y + b4.y;
                    y2 = (acc[ps][f][2] - mean) * rstd * g4.z + b4.z;
                    y3 = (acc[ps][f][3] - mean) * rstd * g4.w + b4.w;
                } else {  // folded BatchNorm (eval): y = x * scale + shift
                    y0 = fmaf(acc[ps][f][0], g4.x, b4.x);
                    y1 = fmaf(acc[ps][f][1], g4.y, b4.y);
                    y2 = fmaf(acc[ps][f][2], g4.z, b4.z);
                    y3 = fmaf(acc[ps][f][3], g4.w, b4.w);
                }
                float s0, s1, s2, s3;
                sigmoid4(y0, y1, y2, y3, s0, s1, s2, s3);
                y0 *= s0;
                y1 *= s1;
                y2 *= s2;
                y3 *= s3;
                const uint32_t p01 = pack_bf16x2(y0, y1), p23 = pack_bf16x2(y2, y3);
                *reinterpret_cast<uint2*>(o + c0) = make_uint2(p01, p23);
                if (P.split3) {
                    *reinterpret_cast<uint2*>(o + d + c0) =
                        make_uint2(pack_bf16x2(y0 - bf16_lo(p01), y1 - bf16_hi(p01)),
                                   pack_bf16x2(y2 - bf16_lo(p23), y3 - bf16_hi(p23)));
                    *reinterpret_cast<uint2*>(o + 2 * d + c0) = make_uint2(p01, p23);
                }
            }
        }
    }
}

}  // namespace

int dwconv_norm_silu(const DwConvArgs& a, cudaStream_t stream) {
    if (a.batch <= 0 || a.max_len <= 0) return WB_OK;
    WB_REQUIRE(a.ksize >= 1 && a.ksize <= MAX_K, WB_ERR_UNSUPPORTED, "dwconv: kernel size %d unsupported", a.ksize);
    WB_REQUIRE(a.d % 8 == 0 && a.d <= 128 * MAX_PASSES && a.ldg % 8 == 0 && a.ldo % 4 == 0, WB_ERR_UNSUPPORTED,
               "dwconv: d=%d must be a multiple of 8 (<= %d) with 8-byte aligned rows", a.d, 128 * MAX_PASSES);
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
    const size_t smem = ((size_t)rows_in * a.d * 2 + 15) / 16 * 16 + (size_t)a.ksize * a.d * sizeof(float);
    dim3 grid(ceil_div(a.max_len, TT), a.batch);
    ProfScope _ps(PT_DWCONV, stream, (double)a.batch * a.max_len * a.d * 4.0);
#define WB_DW(KT, NP)                                                                                              \
    do {                                                                                                           \
        /* function attributes are per device and the size depends on (K, d): set it whenever the opt-in is needed */ \
        if (smem > 48 * 1024)                                                                                      \
            WB_CHECK_CUDA(cudaFuncSetAttribute(dwconv_kernel<KT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)smem));                                                        \
        dwconv_kernel<KT, NP><<<grid, DW_THREADS, smem, stream>>>(P);                                              \
    } while (0)
#define WB_DW_NP(KT)                  \
    do {                              \
        if (a.d <= 128) WB_DW(KT, 1); \
        else if (a.d <= 256) WB_DW(KT, 2); \
        else WB_DW(KT, 4);            \
    } while (0)
    if (a.ksize == 8) WB_DW_NP(8);
    else if (a.ksize == 15) WB_DW_NP(15);
    else WB_DW(0, 4);
#undef WB_DW_NP
#undef WB_DW
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
