// Fused position-wise feed-forward for d_model = 256:
//     x[M,256] += alpha * ( act(a[M,256] W1^T + b1) W2^T + b2 )   (W1 [ff,256], W2 [256,ff]; act = SiLU | ReLU)
// Replaces wenet/models/transformer/positionwise_feed_forward.py:50-58 (both Linear layers and the
// activation) + the 1/2-scaled residual add of encoder_layer.py:221-228, :254-259 in ONE kernel: the
// [M, ff] hidden activation never leaves the SM (the unfused pair wrote and re-read 2 x 196 MB per call
// at 64 x 30 s and spent its time in the SiLU epilogue instead of the tensor core).
//
// One persistent CTA per SM, 352 threads, a 128-row tile of `a` resident in shared memory:
//   warp 0      TMA producer : a-tile once per tile; per 64-wide ff chunk c the W1 rows [64c, 64c+64) (4 boxes
//                              64x64) and the W2 columns [64c, 64c+64) (one 256x64 box), 2-stage rings
//   warp 1      MMA issuer   : H_c = a W1_c^T  (tcgen05.mma 128x64x16, 16 per chunk) into a double-buffered TMEM
//                              slot; Y += hA_c W2_c^T (128x256x16, 4 per chunk) into TMEM columns [0,256).
//                              MMA1 of chunk c+1 is issued BEFORE MMA2 of chunk c, so the tensor core works while
//                              the SiLU warps transform chunk c.
//   warps 2..9  SiLU         : two groups of 4 warps on alternate chunks (de-phased: one group's MUFU burst overlaps the
//                              other's TMEM loads / shared-memory stores): tcgen05.ld H_c -> + b1 -> SiLU -> bf16 ->
//                              shared memory in the canonical K-major SWIZZLE_128B layout (A operand of MMA2).
//   end of tile : Y -> + b2, * alpha -> staged through (now idle) shared memory -> TMA reduce-add into x.
// Per chunk: tensor 2 x 512 cycles, MUFU (2 per element) 1024 cycles -> balanced; weights stream from L2
// (2 MB per tile).
#include "common.cuh"
#include "kernels.h"
#include <string.h>

namespace wb {

namespace {

constexpr int FF_D = 256;      // d_model handled by this kernel
constexpr int FF_C = 64;       // ff columns per chunk
constexpr int FF_M = 128;      // rows per tile

// shared memory map (bytes, all 1024-aligned)
constexpr int SM_A = 0;                          // a tile: 4 k-panels [128 x 64] bf16 = 64 KB
constexpr int SM_W1 = SM_A + 4 * 16384;          // 2 stages x (4 k-panels [64 x 64] = 32 KB)
constexpr int SM_W2 = SM_W1 + 2 * 32768;         // 2 stages x ([256 x 64] = 32 KB)
constexpr int SM_H = SM_W2 + 2 * 32768;          // 2 buffers x ([128 x 64] bf16 = 16 KB)
constexpr int SM_BAR = SM_H + 2 * 16384;         // barriers
constexpr int FF_SMEM = SM_BAR + 256 + 1024;     // = 230656 B
// final epilogue staging reuses the two hA buffers (8 warps x 4 KB = 32 KB)

struct FfnParams {
    int M, ff;
    int act;  // 0 = SiLU (conformer encoder), 1 = ReLU (transformer decoder)
    float alpha;
    const float* b1;
    const float* b2;
    int num_tiles;
};

__global__ void __launch_bounds__(352, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_x, FfnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    uint64_t* a_full = bars + 0;      // a tile landed
    uint64_t* a_empty = bars + 1;     // all MMA1 of the tile retired -> a may be overwritten
    uint64_t* w1_full = bars + 2;     // [2]
    uint64_t* w1_empty = bars + 4;    // [2]
    uint64_t* w2_full = bars + 6;     // [2]
    uint64_t* w2_empty = bars + 8;    // [2]
    uint64_t* h_full = bars + 10;     // [2] MMA1 result in TMEM slot
    uint64_t* h_empty = bars + 12;    // [2] SiLU warp group done reading the TMEM slot (4 arrivals)
    uint64_t* ha_full = bars + 14;    // [2] hA smem buffer written (4 arrivals)
    uint64_t* ha_empty = bars + 16;   // [2] MMA2 done reading hA buffer
    uint64_t* y_full = bars + 18;     // Y accumulator complete
    uint64_t* y_empty = bars + 19;    // Y drained by the epilogue (8 arrivals)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nchunks = p.ff / FF_C;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w1);
        tma_prefetch_desc(&tmap_w2);
        tma_prefetch_desc(&tmap_x);
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&w1_full[s], 1);
            mbar_init(&w1_empty[s], 1);
            mbar_init(&w2_full[s], 1);
            mbar_init(&w2_empty[s], 1);
            mbar_init(&h_full[s], 1);
            mbar_init(&h_empty[s], 4);
            mbar_init(&ha_full[s], 4);
            mbar_init(&ha_empty[s], 1);
        }
        mbar_init(y_full, 1);
        mbar_init(y_empty, 8);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tmem_y = tmem_base;            // columns [0, 256)
    const uint32_t tmem_h0 = tmem_base + 256;     // two 64-column slots at 256 and 320

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;  // running chunk counter (ring position of W1 / W2 stages)
            uint32_t tile_i = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tile_i) {
                if (tile_i > 0) mbar_wait(a_empty, (tile_i - 1) & 1);
                mbar_expect_tx(a_full, 4 * 16384);
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem + SM_A + kb * 16384, &tmap_a, a_full, kb * 64, t * FF_M);
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it & 1;
                    const uint32_t ph = (it >> 1) & 1;
                    mbar_wait(&w1_empty[s], ph ^ 1);
                    mbar_expect_tx(&w1_full[s], 32768);
                    for (int kb = 0; kb < 4; ++kb)
                        tma_load_2d(smem + SM_W1 + s * 32768 + kb * 8192, &tmap_w1, &w1_full[s], kb * 64, c * FF_C);
                }
            }
        }
    } else if (warp == 10) {
        // ===================== W2 producer =====================
        // A separate thread from the W1 producer: MMA1 runs two chunks ahead of MMA2, so a single in-order producer
        // would hold every W1 refill behind the (later) W2 stage release and expose the L2 latency once per chunk.
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
                for (int c = 0; c < nchunks; ++c, ++it) {
                    const int s = it & 1;
                    const uint32_t ph = (it >> 1) & 1;
                    mbar_wait(&w2_empty[s], ph ^ 1);
                    mbar_expect_tx(&w2_full[s], 32768);
                    tma_load_2d(smem + SM_W2 + s * 32768, &tmap_w2, &w2_full[s], c * FF_C, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc1 = make_idesc_bf16(FF_M, FF_C);    // H_c: 128 x 64
            constexpr uint32_t idesc2 = make_idesc_bf16(FF_M, FF_D);    // Y  : 128 x 256
            uint32_t it = 0;
            uint32_t tile_i = 0;
            const uint32_t a_addr = smem_u32(smem + SM_A);
            auto issue_mma1 = [&](uint32_t g) {   // g = global chunk index (ring position)
                const int s = g & 1;
                const uint32_t ph = (g >> 1) & 1;
                mbar_wait(&w1_full[s], ph);
                mbar_wait(&h_empty[s], ph ^ 1);
                tc_fence_after();
                const uint32_t w_addr = smem_u32(smem + SM_W1 + s * 32768);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(tmem_h0 + (uint32_t)(s * 64), make_smem_desc_sw128(a_addr + kb * 16384 + k * 32, 16, 1024),
                                 make_smem_desc_sw128(w_addr + kb * 8192 + k * 32, 16, 1024), idesc1, (kb | k) != 0 ? 1u : 0u);
                umma_commit(&w1_empty[s]);
                umma_commit(&h_full[s]);
            };
            for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tile_i) {
                mbar_wait(a_full, tile_i & 1);
                tc_fence_after();
                // MMA1 runs two chunks ahead of MMA2: the two SiLU warp groups work on alternate chunks, so while one
                // group is in its MUFU phase the other loads / stores, and the tensor core fills the gaps.
                issue_mma1(it);
                if (nchunks > 1) issue_mma1(it + 1);
                if (nchunks <= 2) umma_commit(a_empty);
                for (int c = 0; c < nchunks; ++c, ++it) {
                    if (c + 2 < nchunks) {
                        issue_mma1(it + 2);
                        if (c + 3 == nchunks) umma_commit(a_empty);   // the last MMA1 of this tile has been issued
                    }
                    const int s = it & 1;
                    const uint32_t ph = (it >> 1) & 1;
                    if (c == 0 && tile_i > 0) mbar_wait(y_empty, (tile_i - 1) & 1);   // previous tile's Y has been drained
                    mbar_wait(&w2_full[s], ph);
                    mbar_wait(&ha_full[s], ph);
                    tc_fence_after();
                    const uint32_t h_addr = smem_u32(smem + SM_H + s * 16384);
                    const uint32_t w2_addr = smem_u32(smem + SM_W2 + s * 32768);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(tmem_y, make_smem_desc_sw128(h_addr + k * 32, 16, 1024),
                                 make_smem_desc_sw128(w2_addr + k * 32, 16, 1024), idesc2, (c | k) != 0 ? 1u : 0u);
                    umma_commit(&w2_empty[s]);
                    umma_commit(&ha_empty[s]);
                }
                umma_commit(y_full);
            }
        }
    } else if (warp < 10) {
        // ===================== SiLU warps (2..9) + final epilogue =====================
        const int q = warp & 3;             // TMEM lane quarter
        const int half = (warp - 2) >> 2;   // warp group: chunks c = half, half + 2, ... ; column half in the final epilogue
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const int row_in_tile = q * 32 + lane;
        const int act = p.act;
        uint32_t it0 = 0;   // global chunk index of this tile's chunk 0
        uint32_t tile_i = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tile_i, it0 += (uint32_t)nchunks) {
            for (int c = half; c < nchunks; c += 2) {
                const uint32_t g = it0 + (uint32_t)c;
                const int s = g & 1;
                const uint32_t ph = (g >> 1) & 1;
                mbar_wait(&h_full[s], ph);
                tc_fence_after();
                uint32_t r0[32], r1[32];
                tmem_ld_32x32b_x32(tmem_h0 + lane_sel + (uint32_t)(s * 64), r0);
                tmem_ld_32x32b_x32(tmem_h0 + lane_sel + (uint32_t)(s * 64 + 32), r1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_empty[s]);   // TMEM slot may be overwritten by MMA1 of chunk g+2
                const float* bp = p.b1 + c * FF_C;
                uint32_t pk[32];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const uint32_t* rr = hh == 0 ? r0 : r1;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + hh * 32 + i));
                        float t0 = __uint_as_float(rr[i]) + b4.x, t1 = __uint_as_float(rr[i + 1]) + b4.y;
                        float t2 = __uint_as_float(rr[i + 2]) + b4.z, t3 = __uint_as_float(rr[i + 3]) + b4.w;
                        if (act == 0) {
                            float s0, s1, s2, s3;
                            sigmoid4(t0, t1, t2, t3, s0, s1, s2, s3);
                            t0 *= s0; t1 *= s1; t2 *= s2; t3 *= s3;
                        } else {
                            t0 = fmaxf(t0, 0.f); t1 = fmaxf(t1, 0.f); t2 = fmaxf(t2, 0.f); t3 = fmaxf(t3, 0.f);
                        }
                        pk[hh * 16 + (i >> 1)] = pack_bf16x2(t0, t1);
                        pk[hh * 16 + (i >> 1) + 1] = pack_bf16x2(t2, t3);
                    }
                }
                mbar_wait(&ha_empty[s], ph ^ 1);   // MMA2 of chunk g-2 has finished reading this buffer
                uint8_t* hrow = smem + SM_H + s * 16384 + row_in_tile * 128;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    *reinterpret_cast<uint4*>(hrow + ((u ^ (row_in_tile & 7)) << 4)) =
                        make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ha_full[s]);
            }
            // ---- final epilogue of the tile: Y + b2, * alpha -> TMA reduce-add into x ----
            mbar_wait(y_full, tile_i & 1);
            tc_fence_after();
            // staging = the two hA buffers (8 x 4 KB): y_full implies every MMA2 reading them has retired, and they are
            // rewritten only by these same warps in the next tile
            // (each warp stages in the 4 KB it also writes during the SiLU phase, so a warp that runs ahead into the next
            // tile can never overwrite a slower warp's staging while its TMA reduce is still reading it)
            uint8_t* stage = smem + SM_H + half * 16384 + q * 4096;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int col0 = (half * 4 + cc) * 32;
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_y + lane_sel + (uint32_t)col0, r);
                tmem_ld_wait();
                if (cc > 0) {
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                }
                uint8_t* srow = stage + lane * 128;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + 4 * u));
                    *reinterpret_cast<float4*>(srow + ((u ^ (lane & 7)) << 4)) =
                        make_float4(p.alpha * (__uint_as_float(r[4 * u]) + b4.x), p.alpha * (__uint_as_float(r[4 * u + 1]) + b4.y),
                                    p.alpha * (__uint_as_float(r[4 * u + 2]) + b4.z), p.alpha * (__uint_as_float(r[4 * u + 3]) + b4.w));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_reduce_add_2d(&tmap_x, stage, col0, t * FF_M + q * 32);
                    tma_store_commit();
                }
            }
            if (lane == 0) tma_store_wait_read<0>();   // staging (= hA buffers) is reused by the next tile's SiLU phase
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(y_empty);
        }
        if (lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

// ff must be a multiple of 128: an even chunk count keeps warp group g on hA buffer / TMEM slot g in every tile
bool ffn_fused_supported(int d, int ff) { return d == FF_D && ff % (2 * FF_C) == 0 && ff >= 2 * FF_C; }

int g_sm_reserve_ffn = 0;
void ffn_set_sm_reserve(int n) { g_sm_reserve_ffn = n < 0 ? 0 : n; }

int ffn_fused(const void* a_bf16, long long lda, const void* w1, const float* b1, const void* w2, const float* b2, int M,
              int d, int ff, float alpha, int act, float* x, long long ldx, cudaStream_t stream) {
    if (M <= 0) return WB_OK;
    WB_REQUIRE(ffn_fused_supported(d, ff), WB_ERR_UNSUPPORTED, "ffn_fused: d=%d ff=%d unsupported", d, ff);
    CUtensorMap ta, tw1, tw2, tx;
    int rc;
    if ((rc = make_tmap_2d_bf16(&ta, a_bf16, (uint64_t)M, (uint64_t)d, (uint64_t)lda, 128, 64)) != WB_OK) return rc;
    if ((rc = make_tmap_2d_bf16(&tw1, w1, (uint64_t)ff, (uint64_t)d, (uint64_t)d, 64, 64)) != WB_OK) return rc;
    if ((rc = make_tmap_2d_bf16(&tw2, w2, (uint64_t)d, (uint64_t)ff, (uint64_t)ff, 256, 64)) != WB_OK) return rc;
    if ((rc = make_tmap_2d(&tx, x, 4, (uint64_t)M, (uint64_t)d, (uint64_t)ldx, 32, 32)) != WB_OK) return rc;
    FfnParams p;
    p.M = M;
    p.ff = ff;
    p.alpha = alpha;
    p.act = act;
    p.b1 = b1;
    p.b2 = b2;
    p.num_tiles = ceil_div(M, FF_M);
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        WB_CHECK_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM));
        int dev = 0;
        WB_CHECK_CUDA(cudaGetDevice(&dev));
        WB_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_set = true;
    }
    const int usable = (num_sms - g_sm_reserve_ffn) > 1 ? (num_sms - g_sm_reserve_ffn) : 1;
    const int grid = p.num_tiles < usable ? p.num_tiles : usable;
    ProfScope _ps(PT_FFN_FUSED, stream, 4.0 * (double)M * (double)d * (double)ff);
    ffn_fused_kernel<<<grid, 352, FF_SMEM, stream>>>(ta, tw1, tw2, tx, p);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
