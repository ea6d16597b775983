// fp32 companions of the two non-GEMM stages that round to bf16 internally, for the PRECISE parity mode
// (wb_model_config.precise = 1): every GEMM of the encoder then runs as bf16x3 (A = [hi|lo|hi], B = [hi|hi|lo],
// fp32-grade products, fp32 accumulation in TMEM) and these two kernels keep scores / probabilities / convolution
// sums in fp32 on the CUDA cores.  The mode exists to show the path equals the reference's fp32 arithmetic to
// <= 1e-3 (BASELINE.json north_star); it is not the throughput path.
//
//   attention_f32:  RelPositionMultiHeadedAttention.forward (wenet/models/transformer/attention.py:364-438, rel_shift
//                   not applied :407-409) / MultiHeadedAttention.forward_attention (:133-178) with the masks of
//                   wenet/utils/mask.py:88-123,164-198 generated from (len, chunk, left).
//   dwconv_f32:     ConvolutionModule tail (convolution.py:119-147): pad / cache, depthwise conv, LayerNorm or folded
//                   BatchNorm, SiLU.
#include "common.cuh"
#include "kernels.h"
#include <math.h>

namespace wb {

namespace {

constexpr int AW = 8;     // query rows (warps) per CTA
constexpr int AKT = 32;   // keys per tile (one per lane)

__global__ void __launch_bounds__(AW * 32)
attention_f32_kernel(AttnF32Args P) {
    __shared__ float s_k[AKT][65];   // k' = k + p  (padded: lane j walks row j)
    __shared__ float s_v[AKT][64];
    __shared__ float s_c[AKT];       // u.k + v.p per key
    __shared__ float s_q[AW][64];
    const int b = blockIdx.z, h = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nq = P.q_len[b], nk = P.k_len[b];
    const int q0 = blockIdx.x * AW;
    if (q0 >= nq) return;
    const int i = q0 + warp;                       // query index inside the utterance
    const bool q_ok = i < nq;
    const long long qrow = (long long)P.q_start[b] + i;
    const long long krow0 = P.k_start[b];
    const int d = P.heads * 64;
    if (q_ok) {
        s_q[warp][lane] = P.q[qrow * P.ldq + h * 64 + lane];
        s_q[warp][lane + 32] = P.q[qrow * P.ldq + h * 64 + lane + 32];
    }
    // visible key range of this query (chunk mask, mask.py:88-123 with num_left_chunks :164-173)
    int j_lo = 0, j_hi = nk;
    if (P.chunk_size > 0) {
        const int ci = i / P.chunk_size;
        j_hi = min(nk, (ci + 1) * P.chunk_size);
        if (P.num_left_chunks >= 0) j_lo = max((ci - P.num_left_chunks) * P.chunk_size, 0);
    }
    // union over the CTA's queries, so that all warps walk the same tiles
    int t_lo = 0, t_hi = nk;
    if (P.chunk_size > 0) {
        const int c_first = q0 / P.chunk_size, c_last = min(q0 + AW - 1, nq - 1) / P.chunk_size;
        t_hi = min(nk, (c_last + 1) * P.chunk_size);
        if (P.num_left_chunks >= 0) t_lo = max((c_first - P.num_left_chunks) * P.chunk_size, 0);
    }
    float m = -INFINITY, l = 0.f, acc0 = 0.f, acc1 = 0.f;
    const int lk = threadIdx.x >> 3, le = (threadIdx.x & 7) * 8;   // loader: key lk, dims [le, le + 8)
    for (int kt = (t_lo / AKT) * AKT; kt < t_hi; kt += AKT) {
        __syncthreads();
        {
            const int j = kt + lk;
            float cpart = 0.f;
            if (j < nk) {
                const float* kr = P.k + (krow0 + j) * P.ldk + h * 64 + le;
                const float* vr = P.v + (krow0 + j) * P.ldv + h * 64 + le;
                const float* pr = P.pos_proj ? P.pos_proj + (long long)P.row_pos[krow0 + j] * d + h * 64 + le : nullptr;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float kv = kr[e];
                    const float pv = pr ? pr[e] : 0.f;
                    s_k[lk][le + e] = kv + pv;
                    s_v[lk][le + e] = vr[e];
                    if (pr) cpart += P.pos_u[h * 64 + le + e] * kv + P.pos_v[h * 64 + le + e] * pv;
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    s_k[lk][le + e] = 0.f;
                    s_v[lk][le + e] = 0.f;
                }
            }
            cpart += __shfl_xor_sync(0xffffffffu, cpart, 1);
            cpart += __shfl_xor_sync(0xffffffffu, cpart, 2);
            cpart += __shfl_xor_sync(0xffffffffu, cpart, 4);
            if ((threadIdx.x & 7) == 0) s_c[lk] = cpart;
        }
        __syncthreads();
        if (!q_ok) continue;
        const int j = kt + lane;
        float s = -INFINITY;
        if (j >= j_lo && j < j_hi) {
            float dot = 0.f;
#pragma unroll 16
            for (int e = 0; e < 64; ++e) dot = fmaf(s_q[warp][e], s_k[lane][e], dot);
            s = (dot + s_c[lane]) * P.scale;
        }
        const float tm = warp_max(s);
        if (tm == -INFINITY) continue;   // warp-uniform: no visible key in this tile
        const float mn = fmaxf(m, tm);
        const float corr = (m == -INFINITY) ? 0.f : expf(m - mn);
        const float p = (s == -INFINITY) ? 0.f : expf(s - mn);
        l = l * corr + warp_sum(p);
        acc0 *= corr;
        acc1 *= corr;
#pragma unroll 8
        for (int jj = 0; jj < AKT; ++jj) {
            const float pj = __shfl_sync(0xffffffffu, p, jj);
            acc0 = fmaf(pj, s_v[jj][lane], acc0);
            acc1 = fmaf(pj, s_v[jj][lane + 32], acc1);
        }
        m = mn;
    }
    if (!q_ok) return;
    // a fully masked query row gives zeros (attention.py:158-165: softmax of -inf rows is re-masked to 0)
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    const float o0 = acc0 * inv, o1 = acc1 * inv;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(P.out) + qrow * P.ldo + h * 64;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(o0), h1 = __float2bfloat16_rn(o1);
    o[lane] = h0;
    o[lane + 32] = h1;
    if (P.split3_out) {
        o[d + lane] = __float2bfloat16_rn(o0 - __bfloat162float(h0));
        o[d + lane + 32] = __float2bfloat16_rn(o1 - __bfloat162float(h1));
        o[2 * d + lane] = h0;
        o[2 * d + lane + 32] = h1;
    }
}

constexpr int DWP_THREADS = 256;
constexpr int DWP_MAXC = 4;   // channels per thread: d <= 1024

struct DwF32Dev {
    const __nv_bfloat16* g;
    long long ldg;
    const int* seq_start;
    const int* seq_len;
    const int* out_start;
    int lead, d, ksize, causal, in_split3;
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

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < DWP_THREADS / 32; ++w) t += red[w];
    return t;
}

// one CTA per output frame; thread owns channels tid, tid + 256, ...
__global__ void __launch_bounds__(DWP_THREADS)
dwconv_f32_kernel(DwF32Dev P) {
    __shared__ float red[DWP_THREADS / 32];
    const int b = blockIdx.y, t = blockIdx.x;
    const int n_in = P.seq_len[b];
    const int n_out = n_in - P.lead;
    if (t >= n_out) return;
    const int d = P.d, K = P.ksize;
    const int left = P.causal ? (K - 1) : (K - 1) / 2;
    const long long base = P.seq_start[b];
    float y[DWP_MAXC];
    float sm = 0.f;
#pragma unroll
    for (int ci = 0; ci < DWP_MAXC; ++ci) {
        const int c = threadIdx.x + ci * DWP_THREADS;
        y[ci] = 0.f;
        if (c >= d) continue;
        float acc = P.bias[c];
        for (int k = 0; k < K; ++k) {
            const int p = P.lead + t - left + k;
            float x = 0.f;
            if (p >= 0 && p < n_in) {
                const __nv_bfloat16* r = P.g + (base + p) * P.ldg;
                x = __bfloat162float(r[c]);
                if (P.in_split3) x += __bfloat162float(r[d + c]);
            } else if (P.pad_vec != nullptr &&
                       ((p < 0 && P.causal) || (p >= n_in && !P.causal && p < P.pad_until + P.lead))) {
                x = P.pad_vec[c];   // GLU(pointwise_conv1(0)) = GLU(bias): see convmod.cu
            }
            acc = fmaf(P.w[c * K + k], x, acc);
        }
        y[ci] = acc;
        sm += acc;
    }
    float mean = 0.f, rstd = 1.f;
    if (P.norm_type == 0) {
        mean = block_sum(sm, red) / (float)d;
        float q = 0.f;
#pragma unroll
        for (int ci = 0; ci < DWP_MAXC; ++ci)
            if (threadIdx.x + ci * DWP_THREADS < d) q += (y[ci] - mean) * (y[ci] - mean);
        rstd = 1.0f / sqrtf(block_sum(q, red) / (float)d + P.eps);
    }
    __nv_bfloat16* o = P.out + ((long long)P.out_start[b] + t) * P.ldo;
#pragma unroll
    for (int ci = 0; ci < DWP_MAXC; ++ci) {
        const int c = threadIdx.x + ci * DWP_THREADS;
        if (c >= d) continue;
        float z = (P.norm_type == 0) ? (y[ci] - mean) * rstd * P.gamma[c] + P.beta[c] : fmaf(y[ci], P.gamma[c], P.beta[c]);
        z = z / (1.0f + expf(-z));
        const __nv_bfloat16 hi = __float2bfloat16_rn(z);
        o[c] = hi;
        if (P.split3) {
            o[d + c] = __float2bfloat16_rn(z - __bfloat162float(hi));
            o[2 * d + c] = hi;
        }
    }
}

}  // namespace

int attention_f32(const AttnF32Args& a, cudaStream_t stream) {
    if (a.batch <= 0 || a.max_q_len <= 0) return WB_OK;
    WB_REQUIRE(a.q && a.k && a.v && a.out && a.q_start && a.q_len && a.k_start && a.k_len, WB_ERR_BAD_ARG,
               "attention_f32: null argument");
    WB_REQUIRE(a.pos_proj == nullptr || (a.row_pos && a.pos_u && a.pos_v), WB_ERR_BAD_ARG,
               "attention_f32: rel-pos tables incomplete");
    dim3 grid(ceil_div(a.max_q_len, AW), a.heads, a.batch);
    ProfScope _ps(PT_ATTENTION, stream, 0.0);
    attention_f32_kernel<<<grid, AW * 32, 0, stream>>>(a);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int dwconv_norm_silu_f32(const DwConvArgs& a, cudaStream_t stream) {
    if (a.batch <= 0 || a.max_len <= 0) return WB_OK;
    WB_REQUIRE(a.ksize >= 1 && a.d <= DWP_THREADS * DWP_MAXC, WB_ERR_UNSUPPORTED, "dwconv_f32: d=%d ksize=%d unsupported",
               a.d, a.ksize);
    WB_REQUIRE(a.causal || (a.ksize % 2 == 1), WB_ERR_BAD_ARG, "dwconv_f32: symmetric kernel must be odd");
    DwF32Dev P;
    P.g = reinterpret_cast<const __nv_bfloat16*>(a.g);
    P.ldg = a.ldg;
    P.seq_start = a.seq_start;
    P.seq_len = a.seq_len;
    P.out_start = a.out_start;
    P.lead = a.lead;
    P.d = a.d;
    P.ksize = a.ksize;
    P.causal = a.causal;
    P.in_split3 = a.in_split3;
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
    dim3 grid(a.max_len, a.batch);
    ProfScope _ps(PT_DWCONV, stream, 0.0);
    dwconv_f32_kernel<<<grid, DWP_THREADS, 0, stream>>>(P);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
