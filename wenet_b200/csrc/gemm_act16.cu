// Weight-stationary tcgen05 GEMM with SIXTEEN epilogue warps for the activation-heavy Linear of the feed-forward module
//   H[M, N] = act(A[M, K] W[N, K]^T + b),  K <= 256, bf16 in / bf16 out, act = SiLU (Conformer FFN w_1,
//   wenet/models/transformer/positionwise_feed_forward.py:50-58 with activation swish) or ReLU / GELU / identity.
//
// Why a second kernel.  In gemm_tcgen05_kernel<256, true, SiLU> the eight epilogue warps (two per SM sub-partition) spend
// ~5000 cycles on a 128 x 256 tile that the tensor core finishes in ~2830: ncu (profiles/r2_ncu_gemm_v29.txt, launch 2)
// shows the MUFU pipe 41 % and the issue slots 41 % busy - neither saturated, two warps per scheduler simply cannot hide
// their own tcgen05.ld -> ex2 -> rcp -> pack -> st.shared -> TMA-store chain.  Here FOUR warps per scheduler work on the
// epilogue: two groups of eight warps, group g draining accumulator stage g, i.e. the groups take alternate tiles and are
// naturally half a tile out of phase (one group is in its MUFU-heavy part while the other packs / stores).  Everything
// else is the weight-stationary schedule of gemm.cu: warp 0 TMA producer (resident [256 x K] weight panel + A ring),
// warp 1 tcgen05.mma issuer into a double-buffered TMEM accumulator.
// Budget: 576 threads -> 112 registers per thread, so the epilogue keeps ONE 32-column chunk in flight (no software
// prefetch: four warps per scheduler cover the tcgen05.ld latency), and 16 warps x 2 KB of output staging (32 rows x 64
// bytes, SWIZZLE_64B boxes) fit the 32 KB the eight-warp kernel uses for 4 KB buffers.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>
#include <string.h>

namespace wb {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kABytes = BM * BK * 2;        // 16 KB
constexpr int kBBytes = BN * BK * 2;        // 32 KB
constexpr int kStageArea = 4 * (kABytes + kBBytes);   // 192 KB: resident panel (num_kb x 32 KB) + A ring
constexpr int kOutBytes = 16 * 2048;
constexpr int kBiasBytes = 2 * 256 * 4;
constexpr int kSmemBytes = kStageArea + kOutBytes + kBiasBytes + 256;
constexpr int kThreads = 64 + 16 * 32;

struct Act16Params {
    int M, N, K;
    const float* bias;
    int num_m_tiles, num_n_tiles;
    int ring;   // A ring depth (slots of one k-block)
};

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_act16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_c, Act16Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();
    const int num_kb = (p.K + BK - 1) / BK;
    const int kRing = p.ring;
    uint8_t* smem_b = smem;
    uint8_t* smem_a = smem + num_kb * kBBytes;
    uint8_t* smem_out = smem + kStageArea;
    float* smem_bias = reinterpret_cast<float*>(smem + kStageArea + kOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStageArea + kOutBytes + kBiasBytes);
    uint64_t* full_bar = bars;          // [kRing <= 8]
    uint64_t* empty_bar = bars + 8;
    uint64_t* tmem_full = bars + 16;    // [2]
    uint64_t* tmem_empty = bars + 18;   // [2]
    uint64_t* b_full = bars + 20;
    uint64_t* b_empty = bars + 21;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_mt = p.num_m_tiles;
    const int num_tiles = num_mt * p.num_n_tiles;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        tma_prefetch_desc(&tmap_c);
        for (int s = 0; s < kRing; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(b_full, 1);
        mbar_init(b_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 8);   // the eight warps of the group that drains this stage
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_holder, 2 * BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    // each CTA owns a contiguous range of the n-major tile list: its weight panel changes at most a few times
    const int per_cta = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_begin = (int)blockIdx.x * per_cta;
    const int t_end = min(num_tiles, t_begin + per_cta);

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int cur_n = -1;
            uint32_t bemp_phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int n_tile = t / num_mt, m_tile = t % num_mt;
                if (n_tile != cur_n) {
                    if (cur_n >= 0) {
                        mbar_wait(b_empty, bemp_phase);
                        bemp_phase ^= 1;
                    }
                    mbar_expect_tx(b_full, (uint32_t)num_kb * kBBytes);
                    for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(smem_b + kb * kBBytes, &tmap_b, b_full, kb * BK, n_tile * BN);
                    cur_n = n_tile;
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], kABytes);
                    tma_load_2d(smem_a + stage * kABytes, &tmap_a, &full_bar[stage], kb * BK, m_tile * BM);
                    if (++stage == kRing) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int cur_n = -1;
            uint32_t bfull_phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int n_tile = t / num_mt;
                if (n_tile != cur_n) {
                    mbar_wait(b_full, bfull_phase);
                    bfull_phase ^= 1;
                    cur_n = n_tile;
                }
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
                    const uint32_t b_addr = smem_u32(smem_b + kb * kBBytes);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_f16(tmem_d, make_smem_desc_sw128(a_addr + k * 32, 16, 1024),
                                 make_smem_desc_sw128(b_addr + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (++stage == kRing) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full[acc]);
                if (t + 1 < t_end && (t + 1) / num_mt != n_tile) umma_commit(b_empty);   // panel may be replaced once these retire
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue: warps 2..17, two groups of eight =====================
        const int e = warp - 2;
        const int g = e >> 3;            // group = accumulator stage = tile parity within this CTA
        const int q = warp & 3;          // TMEM lane quarter this warp may access
        const int half = (e >> 2) & 1;   // column half of the tile (4 chunks of 32 columns)
        uint32_t acc_phase = 0;
        bool need_wait = false;
        uint8_t* sbuf_warp = smem_out + e * 2048;
        uint8_t* sbuf = sbuf_warp + lane * 64;
        const int sw = (lane >> 1) & 3;  // SWIZZLE_64B: 16-byte unit ^= bits 7..8 of the byte address (row pitch 64 B)
        float* sbias = smem_bias + g * 256 + half * 128;
        for (int t = t_begin + g; t < t_end; t += 2) {
            const int n_tile = t / num_mt, m_tile = t % num_mt;
            const int row0 = m_tile * BM + q * 32;
            float4 bpre = make_float4(0.f, 0.f, 0.f, 0.f);
            {
                const int bcol = n_tile * BN + half * 128 + 4 * lane;
                if (bcol + 0 < p.N) bpre.x = __ldg(p.bias + bcol + 0);
                if (bcol + 1 < p.N) bpre.y = __ldg(p.bias + bcol + 1);
                if (bcol + 2 < p.N) bpre.z = __ldg(p.bias + bcol + 2);
                if (bcol + 3 < p.N) bpre.w = __ldg(p.bias + bcol + 3);
            }
            mbar_wait(&tmem_full[g], acc_phase);
            tc_fence_after();
            // the four warps of a column half write identical values; nobody still reads the previous tile's slice: this
            // stage's accumulator only became full again after all eight warps had arrived on tmem_empty, past their reads
            *reinterpret_cast<float4*>(sbias + 4 * lane) = bpre;
            __syncwarp();
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BN + half * 128);
#pragma unroll 1
            for (int ci = 0; ci < 4; ++ci) {
                const int n0 = n_tile * BN + half * 128 + ci * 32;
                if (n0 >= p.N) break;   // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr0 + (uint32_t)(ci * 32), r);
                tmem_ld_wait_regs(r);
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sbias + ci * 32 + i);
                    const float2 lo = f2_add(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), make_float2(b4.x, b4.y));
                    const float2 hi = f2_add(make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), make_float2(b4.z, b4.w));
                    v[i] = lo.x;
                    v[i + 1] = lo.y;
                    v[i + 2] = hi.x;
                    v[i + 3] = hi.y;
                }
                if (EPI == EPI_BF16_SILU) {
                    silu_inplace(v);
                } else if (EPI == EPI_BF16_RELU) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                } else if (EPI == EPI_BF16_GELU) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
                }
                if (need_wait) {   // the previous TMA store of this warp must have finished reading the staging buffer
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                    need_wait = false;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<uint4*>(sbuf + ((u ^ sw) << 4)) =
                        make_uint4(pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                                   pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmap_c, sbuf_warp, n0, row0);   // rows >= M and columns >= N are clipped by the tensor map
                    tma_store_commit();
                }
                need_wait = true;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[g]);
            acc_phase ^= 1;
        }
        if (lane == 0) tma_store_wait<0>();   // shared memory must outlive the bulk stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

template <int EPI>
int launch_act16(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const Act16Params& p, int grid,
                 cudaStream_t stream) {
    WB_SET_MAX_DYN_SMEM((gemm_act16_kernel<EPI>), kSmemBytes);
    gemm_act16_kernel<EPI><<<grid, kThreads, kSmemBytes, stream>>>(ta, tb, tc, p);
    return WB_OK;
}

}  // namespace

int g_act16_flag = -1;

// Returns WB_OK after launching, or 1 when the shape / configuration is not handled here (the caller falls back to
// gemm_tcgen05_kernel).  WB_GEMM_ACT16=0 disables the kernel.
int gemm_act16_try(const void* A, long long lda, const CUtensorMap* tmap_b_one, int M, int N, int K, const float* bias, int epi,
                   void* out, long long ldc, int sm_reserve, cudaStream_t stream) {
    if (g_act16_flag < 0) {
        const char* e = getenv("WB_GEMM_ACT16");
        g_act16_flag = (e == nullptr) ? 1 : atoi(e);
    }
    if (!g_act16_flag) return 1;
    if (!(epi == EPI_BF16_SILU || epi == EPI_BF16_GELU) || bias == nullptr || tmap_b_one == nullptr) return 1;
    if (K > 256 || (K % BK) != 0 || (N % BN) != 0 || M < 16 * BM) return 1;
    if ((ldc * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0 || (lda % 8) != 0) return 1;
    CUtensorMap ta, tc;
    int rc = make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
    if (rc != WB_OK) return rc;
    rc = make_tmap_2d_bf16_sw64(&tc, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32);
    if (rc != WB_OK) return rc;
    Act16Params p;
    p.M = M;
    p.N = N;
    p.K = K;
    p.bias = bias;
    p.num_m_tiles = ceil_div(M, BM);
    p.num_n_tiles = N / BN;
    const int num_kb = K / BK;
    int ring = (kStageArea - num_kb * kBBytes) / kABytes;
    p.ring = ring > 8 ? 8 : ring;
    const int sms = current_device_sms();
    WB_REQUIRE(sms > 0, WB_ERR_CUDA, "gemm_act16: cannot query the SM count");
    const int usable = (sms - sm_reserve) > 1 ? (sms - sm_reserve) : 1;
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int grid = tiles < usable ? tiles : usable;
    ProfScope _ps(PT_GEMM, stream, 2.0 * (double)M * (double)N * (double)K);
    if (epi == EPI_BF16_SILU) rc = launch_act16<EPI_BF16_SILU>(ta, *tmap_b_one, tc, p, grid, stream);
    else rc = launch_act16<EPI_BF16_GELU>(ta, *tmap_b_one, tc, p, grid, stream);
    if (rc != WB_OK) return rc;
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
