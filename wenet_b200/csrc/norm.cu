// Row-wise LayerNorm (fp32 statistics, warp per row) with fused bf16 down-cast for the next GEMM's
// A operand, plus small row utilities.
// Replaces torch.nn.LayerNorm on the path: wenet/models/transformer/encoder_layer.py:169-183
// (norm_ff_macaron / norm_mha / norm_conv / norm_ff / norm_final), encoder.py:111-112,176-177
// (after_norm), decoder_layer.py norm1-3, decoder.py after_norm.
// HBM-bound: reads 4*d B/row, writes 2*d (bf16) and/or 4*d (fp32) B/row.
#include "common.cuh"
#include "kernels.h"

namespace wb {

namespace {

constexpr int LN_WARPS = 8;

// VPL = float4 vectors per lane (d = 128 * VPL)
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_kernel(const float* __restrict__ x, long long ldx, int M, int d, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ out_bf16,
                 long long ld_bf16, int split3, float* out_f32, long long ld_f32) {
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * LN_WARPS + warp;
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        v[i] = xr[lane + 32 * i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        v[i].x -= mean;
        v[i].y -= mean;
        v[i].z -= mean;
        v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c4 = lane + 32 * i;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
        float4 y;
        y.x = v[i].x * rstd * g.x + b.x;
        y.y = v[i].y * rstd * g.y + b.y;
        y.z = v[i].z * rstd * g.z + b.z;
        y.w = v[i].w * rstd * g.w + b.w;
        if (out_f32) reinterpret_cast<float4*>(out_f32 + row * ld_f32)[c4] = y;
        if (out_bf16) {
            __nv_bfloat16* o = out_bf16 + row * ld_bf16 + 4 * c4;
            const uint32_t p0 = pack_bf16x2(y.x, y.y), p1 = pack_bf16x2(y.z, y.w);
            *reinterpret_cast<uint2*>(o) = make_uint2(p0, p1);
            if (split3) {
                const uint32_t l0 = pack_bf16x2(y.x - bf16_lo(p0), y.y - bf16_hi(p0));
                const uint32_t l1 = pack_bf16x2(y.z - bf16_lo(p1), y.w - bf16_hi(p1));
                *reinterpret_cast<uint2*>(o + d) = make_uint2(l0, l1);
                *reinterpret_cast<uint2*>(o + 2 * d) = make_uint2(p0, p1);
            }
        }
    }
}

// Two LayerNorms back to back on the same row, one read of x:  y = LN1(x) (fp32, optional write-back: the residual
// stream after norm_final, encoder_layer.py:262-263),  z = LN2(y) (bf16 and / or fp32): norm_final of layer l fused with
// norm_ff_macaron of layer l + 1, and norm_final of the last layer with after_norm (encoder.py:176-177).
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm2_kernel(const float* __restrict__ x, long long ldx, int M, int d, const float* __restrict__ g1,
                  const float* __restrict__ b1, const float* __restrict__ g2, const float* __restrict__ b2, float eps,
                  float* y_f32, long long ld_y, __nv_bfloat16* __restrict__ z_bf16, long long ld_zb, int split3,
                  float* __restrict__ z_f32, long long ld_zf) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * LN_WARPS + warp;
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    float4 v[VPL];
    const float inv_d = 1.0f / (float)d;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            if (pass == 0) v[i] = xr[lane + 32 * i];
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
        const float mean = warp_sum(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            v[i].x -= mean;
            v[i].y -= mean;
            v[i].z -= mean;
            v[i].w -= mean;
            q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
        const float* gamma = pass == 0 ? g1 : g2;
        const float* beta = pass == 0 ? b1 : b2;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c4 = lane + 32 * i;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
            const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
            v[i].x = v[i].x * rstd * g.x + b.x;
            v[i].y = v[i].y * rstd * g.y + b.y;
            v[i].z = v[i].z * rstd * g.z + b.z;
            v[i].w = v[i].w * rstd * g.w + b.w;
            if (pass == 0) {
                if (y_f32) reinterpret_cast<float4*>(y_f32 + row * ld_y)[c4] = v[i];
            } else {
                if (z_f32) reinterpret_cast<float4*>(z_f32 + row * ld_zf)[c4] = v[i];
                if (z_bf16) {
                    __nv_bfloat16* o = z_bf16 + row * ld_zb + 4 * c4;
                    const uint32_t p0 = pack_bf16x2(v[i].x, v[i].y), p1 = pack_bf16x2(v[i].z, v[i].w);
                    *reinterpret_cast<uint2*>(o) = make_uint2(p0, p1);
                    if (split3) {
                        const uint32_t l0 = pack_bf16x2(v[i].x - bf16_lo(p0), v[i].y - bf16_hi(p0));
                        const uint32_t l1 = pack_bf16x2(v[i].z - bf16_lo(p1), v[i].w - bf16_hi(p1));
                        *reinterpret_cast<uint2*>(o + d) = make_uint2(l0, l1);
                        *reinterpret_cast<uint2*>(o + 2 * d) = make_uint2(p0, p1);
                    }
                }
            }
        }
    }
}

__global__ void cast_rows_kernel(const float* __restrict__ x, long long ldx, int M, int d,
                                 __nv_bfloat16* __restrict__ out, long long ldo, int split3) {
    const long long n4 = (long long)M * (d / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / (d / 4);
        const int c4 = (int)(i - row * (d / 4));
        const float4 y = *reinterpret_cast<const float4*>(x + row * ldx + 4 * c4);
        __nv_bfloat16* o = out + row * ldo + 4 * c4;
        const uint32_t p0 = pack_bf16x2(y.x, y.y), p1 = pack_bf16x2(y.z, y.w);
        *reinterpret_cast<uint2*>(o) = make_uint2(p0, p1);
        if (split3) {
            const uint32_t l0 = pack_bf16x2(y.x - bf16_lo(p0), y.y - bf16_hi(p0));
            const uint32_t l1 = pack_bf16x2(y.z - bf16_lo(p1), y.w - bf16_hi(p1));
            *reinterpret_cast<uint2*>(o + d) = make_uint2(l0, l1);
            *reinterpret_cast<uint2*>(o + 2 * d) = make_uint2(p0, p1);
        }
    }
}

// pos_offset_dev (optional): the offset is read from device memory and added to pos_offset - lets a captured CUDA
// graph of the streaming chunk step advance through the utterance; positions are clamped to [0, max_pos)
__global__ void fill_row_pos_kernel(const int* __restrict__ seq_start, const int* __restrict__ seq_len,
                                    int pos_offset, const int* __restrict__ pos_offset_dev, int per_seq, int max_pos,
                                    int* __restrict__ row_pos) {
    const int b = blockIdx.y;
    const int s = seq_start[b], n = seq_len[b];
    const int off = pos_offset + (pos_offset_dev ? pos_offset_dev[per_seq ? b : 0] : 0);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        row_pos[s + t] = min(max(off + t, 0), max_pos - 1);
}

}  // namespace

int layernorm_rows(const float* x, long long ldx, int M, int d, const float* gamma, const float* beta,
                   float eps, void* out_bf16, long long ld_bf16, int split3, float* out_f32, long long ld_f32,
                   cudaStream_t stream) {
    if (M <= 0) return WB_OK;
    WB_REQUIRE(d % 128 == 0 && d <= 1280, WB_ERR_UNSUPPORTED, "layernorm: d=%d must be a multiple of 128, <= 1280", d);
    WB_REQUIRE(ldx % 4 == 0 && ld_bf16 % 4 == 0 && ld_f32 % 4 == 0, WB_ERR_BAD_ARG, "layernorm: pitches must be %%4");
    const int grid = ceil_div(M, LN_WARPS);
    ProfScope _ps(PT_LAYERNORM, stream, (double)M * d * (4.0 + (out_bf16 ? 2.0 : 0.0) + (out_f32 ? 4.0 : 0.0)));
    __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
#define WB_LN(V)                                                                                          \
    WB_CHECK_CUDA(launch_maybe_pdl(layernorm_kernel<V>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, x, ldx, M, d, gamma, beta, \
                                   eps, ob, ld_bf16, split3, out_f32, ld_f32))
    switch (d / 128) {
        case 1: WB_LN(1); break;
        case 2: WB_LN(2); break;
        case 3: WB_LN(3); break;
        case 4: WB_LN(4); break;
        case 5: WB_LN(5); break;
        case 6: WB_LN(6); break;
        case 8: WB_LN(8); break;
        case 10: WB_LN(10); break;   // Whisper-large (d = 1280)
        default:
            set_last_error("layernorm: d=%d unsupported", d);
            return WB_ERR_UNSUPPORTED;
    }
#undef WB_LN
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int layernorm2_rows(const float* x, long long ldx, int M, int d, const float* g1, const float* b1, const float* g2,
                    const float* b2, float eps, float* y_f32, long long ld_y, void* z_bf16, long long ld_zb, int split3,
                    float* z_f32, long long ld_zf, cudaStream_t stream) {
    if (M <= 0) return WB_OK;
    WB_REQUIRE(d % 128 == 0 && d <= 1024, WB_ERR_UNSUPPORTED, "layernorm2: d=%d must be a multiple of 128, <= 1024", d);
    WB_REQUIRE(ldx % 4 == 0 && ld_y % 4 == 0 && ld_zb % 4 == 0 && ld_zf % 4 == 0, WB_ERR_BAD_ARG, "layernorm2: pitches must be %%4");
    const int grid = ceil_div(M, LN_WARPS);
    ProfScope _ps(PT_LAYERNORM, stream,
                  (double)M * d * (4.0 + (y_f32 ? 4.0 : 0.0) + (z_bf16 ? 2.0 : 0.0) + (z_f32 ? 4.0 : 0.0)));
    __nv_bfloat16* zb = reinterpret_cast<__nv_bfloat16*>(z_bf16);
#define WB_LN2(V)                                                                                                   \
    layernorm2_kernel<V><<<grid, LN_WARPS * 32, 0, stream>>>(x, ldx, M, d, g1, b1, g2, b2, eps, y_f32, ld_y, zb, ld_zb, \
                                                             split3, z_f32, ld_zf)
    switch (d / 128) {
        case 1: WB_LN2(1); break;
        case 2: WB_LN2(2); break;
        case 3: WB_LN2(3); break;
        case 4: WB_LN2(4); break;
        case 5: WB_LN2(5); break;
        case 6: WB_LN2(6); break;
        case 8: WB_LN2(8); break;
        default:
            set_last_error("layernorm2: d=%d unsupported", d);
            return WB_ERR_UNSUPPORTED;
    }
#undef WB_LN2
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int cast_rows_bf16(const float* x, long long ldx, int M, int d, void* out_bf16, long long ld_bf16, int split3,
                   cudaStream_t stream) {
    if (M <= 0) return WB_OK;
    WB_REQUIRE(d % 4 == 0 && ldx % 4 == 0 && ld_bf16 % 4 == 0, WB_ERR_BAD_ARG, "cast_rows: d/pitches must be %%4");
    const long long n4 = (long long)M * (d / 4);
    const int grid = (int)((n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16);
    ProfScope _ps(PT_MISC, stream, (double)M * d * 6.0);
    cast_rows_kernel<<<grid, 256, 0, stream>>>(x, ldx, M, d, reinterpret_cast<__nv_bfloat16*>(out_bf16), ld_bf16,
                                               split3);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

int fill_row_pos(const int* seq_start, const int* seq_len, int batch, int pos_offset, int* row_pos, int max_len,
                 cudaStream_t stream, const int* pos_offset_dev, int max_pos, int per_seq_offset) {
    if (batch <= 0 || max_len <= 0) return WB_OK;
    dim3 grid(ceil_div(max_len, 256), batch);
    ProfScope _ps(PT_MISC, stream, 0.0);
    fill_row_pos_kernel<<<grid, 256, 0, stream>>>(seq_start, seq_len, pos_offset, pos_offset_dev, per_seq_offset, max_pos, row_pos);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
