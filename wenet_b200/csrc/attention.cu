// Variable-length multi-head attention on tcgen05 (d_k = 64), used for
//   * RelPositionMultiHeadedAttention  (wenet/models/transformer/attention.py:364-438, :133-178)
//       scores[i,j] = ((q_i+u).k_j + (q_i+v).p_j)/sqrt(dk)      (rel_shift is NOT applied, :407-409)
//                   = (q_i.(k_j+p_j) + (u.k_j + v.p_j))/sqrt(dk) = (q_i.k'_j + c_j)/sqrt(dk)
//     so ONE score GEMM against K' = K + P plus a per-key bias c_j (both produced by relpos_kprep)
//     replaces the reference's two (matrix_ac, matrix_bd).
//   * MultiHeadedAttention self-attn (causal) and MultiHeadedCrossAttention of the rescoring decoder
//     (attention.py:247-304, :441-520; masks decoder.py:179-185, mask.py:88-123).
// Masks are generated in-kernel from (k_len, chunk_size, num_left_chunks): key-padding mask, the
// streaming chunk mask (mask.py:subsequent_chunk_mask) and the causal mask (chunk_size = 1).
//
// One CTA = 128 queries of one (sequence, head); 128 threads, thread r owns query row r (TMEM lane r), so the softmax
// needs no cross-thread reduction.  Single pass over 64-key tiles with an online softmax (see the kernel comment):
// S = Q K'^T (tcgen05.mma 128x64x64 into TMEM) -> p = exp2(.) -> P (bf16) into shared memory in the canonical K-major
// SWIZZLE_128B layout -> O += P V (tcgen05.mma 128x64x64).  Q / K' / V tiles arrive by TMA (SWIZZLE_128B); V is consumed
// directly as an MN-major B operand.  Mask comparisons are executed only on tiles that cross a mask boundary.
// (Round 1 carried three earlier schedules - two-pass serial, two-pass pipelined, two threads per row - behind env
// switches; they measured 142 / 177 / 167 us per layer against 108 us for this one and were removed.)
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

namespace wb {

namespace {

constexpr int AT_M = 128;
constexpr int DK = 64;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KB

struct AttnDev {
    const float* kbias;
    int ld_kbias;
    const int* q_start;
    const int* q_len;
    const int* k_start;
    const int* k_len;
    int q_col0, k_col0, v_col0;
    int chunk_size, num_left_chunks;
    float scale_log2e;
    __nv_bfloat16* out;
    long long ldo;
    int out_col0;
    int split3_out;
    int split_width;
    int v_mode;
    int kbias_scaled;   // kbias already holds c * scale * log2(e): fetched by cp.async straight into shared memory
    // split-key mode (part_o != null): batch item z covers ONE piece of the keys of a query block; the kernel leaves the
    // unnormalised fp32 output and (reference point, sum) of its piece, attention_merge_parts combines the pieces
    float* part_o;      // [items][max_q][64]
    float2* part_ml;    // [items][max_q]  (m in the log2 domain, l)
    int part_max_q;
};

// KN = keys per tile.  128: 100 KB smem, 256 TMEM columns (S 128 | O 64 | L 16) -> 2 CTAs/SM.
//                      64:  51 KB smem, 128 TMEM columns (S 64 | O 64)         -> 4 CTAs/SM (row sums by FADD).
template <int KN>
struct AttnCfg {
    static constexpr int kKVBytes = KN * 128;
    static constexpr int kPBytes = KN * 256;
    static constexpr int kTmemCols = (KN == 128) ? 256 : 128;
    static constexpr int kSmem = 1024 /*align*/ + TILE_BYTES /*Q*/ + 2 * kKVBytes /*K,V*/ + kPBytes + 2048 /*ones*/ +
                                 2 * KN * 4 /*c*/ + 128 /*barriers*/;
};

// ----------------------------------------------------------------------------------------------
// Single-pass variant (default): online softmax, so Q K'^T is computed ONCE per key tile (the two-pass kernels spend a
// third of their instructions and half of their score MMAs on the row-max pass).  The running maximum is only
// raised when a tile exceeds it by more than 2^8 ("lazy rescaling"): probabilities are then at most 256 (exact in
// the fp32 row sum, same relative rounding in bf16) and the O accumulator in TMEM has to be rescaled
// (tcgen05.ld -> multiply -> tcgen05.st by the thread that owns the row) only on those rare tiles.  The result is
// sum_j 2^(s_j - m) v_j / sum_j 2^(s_j - m) for whatever reference point m was in use, i.e. the same softmax.
// Key padding is masked through the per-key bias (-1e30 for keys >= k_len), so only chunk / causal boundary tiles
// take the per-element mask path.  Layout as attention_kernel<64>: 128 threads, 51 KB smem, 128 TMEM columns, 4 CTAs/SM.
__global__ void __launch_bounds__(128, 4)
attention_online_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, AttnDev P) {
    constexpr int KN = 64;
    pdl_launch_dependents();
    pdl_wait();   // (ahead of everything: q_len / k_len tables may come from an earlier kernel of the chain)
    using Cfg = AttnCfg<KN>;
    const int b = blockIdx.z, h = blockIdx.y, qt = blockIdx.x;
    const int q_len = P.q_len[b];
    if (qt * AT_M >= q_len) return;
    const int q_start = P.q_start[b];
    const int k_start = P.k_start[b];
    const int k_len = P.k_len[b];

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;
    uint8_t* sV = sK + Cfg::kKVBytes;
    uint8_t* sP = sV + Cfg::kKVBytes;
    float* sC = reinterpret_cast<float*>(sP + Cfg::kPBytes + 2048);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::kPBytes + 2048 + 2 * KN * 4);
    uint64_t* bar_q = bars + 0;
    uint64_t* bar_k = bars + 1;
    uint64_t* bar_v = bars + 2;
    uint64_t* bar_s = bars + 3;
    uint64_t* bar_pv = bars + 4;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 6);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int qi = qt * AT_M + tid;
    // A warp whose 32 query rows all lie past the sequence end only keeps the CTA barriers company: no tcgen05.ld, no
    // exponentials, no P store (its P rows hold stale shared memory, which only reaches its own, never stored, O rows).
    // Autoregressive decoding runs this kernel with `beam` (10) queries per (utterance, head): three of the four warps
    // are idle there, and with 4 CTAs per SM the one live warp per CTA gets the MUFU pipe and the issue slots to itself.
#ifdef WB_ATT_NO_WARP_SKIP
    const bool warp_live = true;
#else
    const bool warp_live = (qt * AT_M + warp * 32) < q_len;
#endif

    if (tid == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        mbar_init(bar_q, 1);
        mbar_init(bar_k, 1);
        mbar_init(bar_v, 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_pv, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_holder, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + KN;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

    int row_lo = 0, row_hi = k_len;
    int cta_lo = 0, cta_hi = k_len;
    int full_lo = 0, full_hi = (k_len + KN - 1) / KN * KN;   // key padding is handled by the bias, not by compares
    if (P.chunk_size > 0) {
        const int c = P.chunk_size;
        row_hi = min((qi / c + 1) * c, k_len);
        row_lo = (P.num_left_chunks < 0) ? 0 : max((qi / c - P.num_left_chunks) * c, 0);
        const int q_first = qt * AT_M, q_last = min(qt * AT_M + AT_M - 1, q_len - 1), q_last_all = qt * AT_M + AT_M - 1;
        cta_hi = min((q_last / c + 1) * c, k_len);
        cta_lo = (P.num_left_chunks < 0) ? 0 : max((q_first / c - P.num_left_chunks) * c, 0);
        full_hi = min((q_first / c + 1) * c, k_len);
        if (full_hi == k_len) full_hi = (k_len + KN - 1) / KN * KN;
        full_lo = (P.num_left_chunks < 0) ? 0 : max((q_last_all / c - P.num_left_chunks) * c, 0);
    }
    const int kt0 = cta_lo / KN;
    const int kt1 = (cta_hi + KN - 1) / KN;

    const int kcol = P.k_col0 + h * DK, vcol = P.v_col0 + h * DK;
    const int kvrow0 = k_start;
    uint32_t ph = 0;   // all per-tile barriers flip once per tile
    constexpr uint32_t idesc_s = make_idesc_bf16(AT_M, KN, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(AT_M, DK, 1);

    if (tid == 0 && kt0 < kt1) {
        mbar_expect_tx(bar_q, TILE_BYTES);
        tma_load_2d(sQ, &tmap_q, bar_q, P.q_col0 + h * DK, q_start + qt * AT_M);
        mbar_expect_tx(bar_k, Cfg::kKVBytes);
        tma_load_2d(sK, &tmap_k, bar_k, kcol, kvrow0 + kt0 * KN);
        mbar_expect_tx(bar_v, Cfg::kKVBytes);
        tma_load_2d(sV, &tmap_v, bar_v, vcol, kvrow0 + kt0 * KN);
    }

    float m_used = -INFINITY;   // reference point of the exponentials accumulated so far
    float l_acc = 0.f;
    // per-key bias of a tile (log2 domain); keys past the sequence end (rows of the next utterance / TMA zero fill)
    // get -1e30 and never count.  The values of tile kt + 1 are fetched at the top of tile kt and parked in the other
    // half of sC before tile kt's only CTA barrier, so neither the load latency nor a second barrier is on the path.
    // The load is issued here and its value is first USED at the store into sC at the end of the tile (the scaling happens
    // there): with the multiply next to the load the in-order warp sat on the scoreboard at the top of every tile
    // (profiles/r2_ncu_attn_v29.txt: 6.6 % of the samples on this line, and the other warps behind it at the CTA barrier).
    auto fetch_c_raw = [&](int kt, bool& valid) -> float {
        const int j = kt * KN + tid;
        valid = (tid < KN) && (kt < kt1) && (j < k_len);
        return (valid && P.kbias != nullptr) ? __ldg(P.kbias + (long long)(k_start + j) * P.ld_kbias + h) : 0.f;
    };
    // Pre-scaled bias (the encoder path: relpos_kprep writes c * scale * log2 e): the next tile's 64 values go global ->
    // shared by cp.async (LDGSTS), no register and no scoreboard wait in the in-order instruction stream at all; keys past
    // the sequence end get -1e30 by a plain store.  cp.async.wait_all sits right before the tile's CTA barrier.
    auto fetch_c_async = [&](int kt) {
        if (tid < KN) {
            const int j = kt * KN + tid;
            float* dst = sC + (kt & 1) * KN + tid;
            if (kt < kt1 && j < k_len) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)),
                             "l"(P.kbias + (long long)(k_start + j) * P.ld_kbias + h)
                             : "memory");
            } else {
                *dst = -1.0e30f;
            }
        }
    };
    const bool c_async = P.kbias_scaled != 0;
    if (kt0 < kt1) {
        if (c_async) {
            fetch_c_async(kt0);
            asm volatile("cp.async.wait_all;" ::: "memory");
        } else if (tid < KN) {
            bool v0;
            const float c0 = fetch_c_raw(kt0, v0);
            sC[(kt0 & 1) * KN + tid] = v0 ? c0 * P.scale_log2e : -1.0e30f;
        }
    }
    __syncthreads();
    for (int kt = kt0; kt < kt1; ++kt) {
        const int j0 = kt * KN;
        bool c_next_valid = false;
        float c_next_raw = 0.f;
        if (c_async) fetch_c_async(kt + 1);
        else c_next_raw = fetch_c_raw(kt + 1, c_next_valid);
        if (tid == 0 && kt == kt0) {   // later score MMAs are issued one tile ahead, together with the P V MMA (below)
            mbar_wait(bar_q, 0);
            mbar_wait(bar_k, ph);
            tc_fence_after();
            const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
            for (int k = 0; k < DK / 16; ++k)
                umma_f16(tmem_s, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024),
                         idesc_s, k != 0);
            umma_commit(bar_s);
        }
        mbar_wait(bar_s, ph);
        tc_fence_after();
        if (tid == 0 && kt + 1 < kt1) {   // K buffer is free again
            mbar_expect_tx(bar_k, Cfg::kKVBytes);
            tma_load_2d(sK, &tmap_k, bar_k, kcol, kvrow0 + (kt + 1) * KN);
        }
        if (warp_live) {
        const float* cc = sC + (kt & 1) * KN;
        const bool tile_full = (j0 >= full_lo) && (j0 + KN <= full_hi);
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(tmem_s + lane_sel, r0);
        tmem_ld_32x32b_x32(tmem_s + lane_sel + 32u, r1);
        tmem_ld_wait();
        // scores (log2 domain) in place, tile maximum.  Two explicit loops: with the mask test inside one loop the
        // compiler predicates the compares instead of branching around them, and they cost issue slots on every tile.
        float mt0 = -INFINITY, mt1 = -INFINITY, mt2 = -INFINITY, mt3 = -INFINITY;
        if (tile_full) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t* r = hh ? r1 : r0;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 c4 = *reinterpret_cast<const float4*>(cc + hh * 32 + i);
                    const float s0 = fmaf(__uint_as_float(r[i]), P.scale_log2e, c4.x);
                    const float s1 = fmaf(__uint_as_float(r[i + 1]), P.scale_log2e, c4.y);
                    const float s2 = fmaf(__uint_as_float(r[i + 2]), P.scale_log2e, c4.z);
                    const float s3 = fmaf(__uint_as_float(r[i + 3]), P.scale_log2e, c4.w);
                    mt0 = fmaxf(mt0, s0);
                    mt1 = fmaxf(mt1, s1);
                    mt2 = fmaxf(mt2, s2);
                    mt3 = fmaxf(mt3, s3);
                    r[i] = __float_as_uint(s0);
                    r[i + 1] = __float_as_uint(s1);
                    r[i + 2] = __float_as_uint(s2);
                    r[i + 3] = __float_as_uint(s3);
                }
            }
        } else {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t* r = hh ? r1 : r0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int key = j0 + hh * 32 + i;
                    float sv = fmaf(__uint_as_float(r[i]), P.scale_log2e, cc[hh * 32 + i]);
                    if (key < row_lo || key >= row_hi) sv = -INFINITY;
                    mt0 = fmaxf(mt0, sv);
                    r[i] = __float_as_uint(sv);
                }
            }
        }
        const float mt = fmaxf(fmaxf(mt0, mt1), fmaxf(mt2, mt3));
        // lazy rescale: raise the reference point only if this tile exceeds it by more than 2^8 (padding-only tiles,
        // mt <= -1e30, never do)
        const bool raise = (mt > -1.0e29f) && (m_used == -INFINITY || mt > m_used + 8.0f);
        float alpha = 1.0f;
        if (raise) {
            alpha = (m_used == -INFINITY) ? 1.0f : fast_exp2(m_used - mt);
            l_acc *= alpha;
            m_used = mt;
        }
        if (kt != kt0 && __any_sync(0xffffffffu, raise && alpha != 1.0f)) {
            // O (TMEM, all P V MMAs so far have retired: bar_pv was awaited at the end of the previous tile) *= alpha
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t o[32];
                tmem_ld_32x32b_x32(tmem_o + lane_sel + (uint32_t)(c * 32), o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32b_x32(tmem_o + lane_sel + (uint32_t)(c * 32), o);
            }
            tmem_st_wait();
        }
        const float m_eff = (m_used == -INFINITY) ? 0.f : m_used;
        // probabilities of one 32-key half at a time, packed and stored straight away: 16 packed words live instead of 32
        // (the kernel runs at the 128-register cap of 4 CTAs / SM)
        uint8_t* prow = sP + tid * 128;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const uint32_t* r = hh ? r1 : r0;
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float p0 = fast_exp2(__uint_as_float(r[i]) - m_eff);
                const float p1 = fast_exp2(__uint_as_float(r[i + 1]) - m_eff);
                const float p2 = fast_exp2(__uint_as_float(r[i + 2]) - m_eff);
                const float p3 = fast_exp2(__uint_as_float(r[i + 3]) - m_eff);
                pk[i >> 1] = pack_bf16x2(p0, p1);
                pk[(i >> 1) + 1] = pack_bf16x2(p2, p3);
                l_acc += (p0 + p1) + (p2 + p3);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                *reinterpret_cast<uint4*>(prow + (((hh * 4 + u) ^ (tid & 7)) << 4)) =
                    make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
        }   // warp_live
        if (c_async) asm volatile("cp.async.wait_all;" ::: "memory");
        else if (tid < KN) sC[((kt + 1) & 1) * KN + tid] = c_next_valid ? c_next_raw * P.scale_log2e : -1.0e30f;
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            mbar_wait(bar_v, ph);
            tc_fence_after();
            const uint32_t pa = smem_u32(sP), va = smem_u32(sV);
#pragma unroll
            for (int ks = 0; ks < KN / 16; ++ks)
                umma_f16(tmem_o, make_smem_desc_sw128(pa + ks * 32, 16, 1024),
                         make_smem_desc_sw128(va + ks * 2048, 1024, 1024), idesc_o, (kt != kt0 || ks != 0) ? 1u : 0u);
            umma_commit(bar_pv);
            if (kt + 1 < kt1) {
                // S(kt+1) = Q K'(kt+1)^T right behind the P V MMAs: every thread pulled S(kt) out of TMEM before the
                // barrier above and K'(kt+1) was requested when S(kt) completed, so the next tile's scores are ready
                // (or nearly) when the threads come back from waiting for P V - one MMA round trip per tile less.
                mbar_wait(bar_k, ph ^ 1);
                tc_fence_after();
                const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
#pragma unroll
                for (int k2 = 0; k2 < DK / 16; ++k2)
                    umma_f16(tmem_s, make_smem_desc_sw128(qa + k2 * 32, 16, 1024),
                             make_smem_desc_sw128(ka + k2 * 32, 16, 1024), idesc_s, k2 != 0);
                umma_commit(bar_s);
            }
        }
        // P / V buffers, the S accumulator and O (possible rescale) are touched again next tile: wait for the P V MMAs
        mbar_wait(bar_pv, ph);
        tc_fence_after();
        if (tid == 0 && kt + 1 < kt1) {
            mbar_expect_tx(bar_v, Cfg::kKVBytes);
            tma_load_2d(sV, &tmap_v, bar_v, vcol, kvrow0 + (kt + 1) * KN);
        }
        ph ^= 1;
    }

    // ------------------------------- epilogue -------------------------------
    if (P.part_o != nullptr) {
        // split-key mode: piece (blockIdx.z, head h) of query row qi
        if (qi < q_len) {
            const long long pr = ((long long)blockIdx.z * gridDim.y + h) * P.part_max_q + qi;
            const bool any = (kt0 < kt1) && (m_used != -INFINITY);
            P.part_ml[pr] = make_float2(any ? m_used : -INFINITY, any ? l_acc : 0.f);
        }
        if (kt0 < kt1) {
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_o + lane_sel + (uint32_t)(c * 32), r);
                tmem_ld_wait();
                if (qi < q_len) {
                    float4* o = reinterpret_cast<float4*>(P.part_o + (((long long)blockIdx.z * gridDim.y + h) * P.part_max_q + qi) * 64 + c * 32);
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        o[u] = make_float4(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 1]), __uint_as_float(r[4 * u + 2]),
                                           __uint_as_float(r[4 * u + 3]));
                }
            }
        }
    } else if (kt0 < kt1) {
        const float inv = (l_acc > 0.f) ? 1.0f / l_acc : 0.f;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_o + lane_sel + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            if (qi < q_len) {
                __nv_bfloat16* o = P.out + (long long)(q_start + qi) * P.ldo + P.out_col0 + h * DK + c * 32;
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    pk[i] = pack_bf16x2(__uint_as_float(r[2 * i]) * inv, __uint_as_float(r[2 * i + 1]) * inv);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    reinterpret_cast<uint4*>(o)[u] = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                if (P.split3_out) {
                    uint32_t lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        lo[i] = pack_bf16x2(__uint_as_float(r[2 * i]) * inv - bf16_lo(pk[i]),
                                            __uint_as_float(r[2 * i + 1]) * inv - bf16_hi(pk[i]));
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        reinterpret_cast<uint4*>(o + P.split_width)[u] =
                            make_uint4(lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
                        reinterpret_cast<uint4*>(o + 2 * P.split_width)[u] =
                            make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                    }
                }
            }
        }
    } else if (qi < q_len) {
        __nv_bfloat16* o = P.out + (long long)(q_start + qi) * P.ldo + P.out_col0 + h * DK;
        for (int i = 0; i < DK; ++i) o[i] = __float2bfloat16_rn(0.f);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}


// ----------------------------------------------------------------------------------------------
// merge of the split-key pieces: one warp per (query block b, head, row), lane = channel pair.  Piece s of block b is
// batch item b * S + s of the attention launch.
__global__ void attention_merge_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml,
                                       const int* __restrict__ q_start, const int* __restrict__ q_len, int S, int H, int max_q,
                                       int total, __nv_bfloat16* __restrict__ out, long long ldo, int out_col0) {
    pdl_launch_dependents();
    pdl_wait();
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= total) return;   // total = B * H * max_q, ordered (b, h, row)
    const int row = w % max_q, bh = w / max_q, h = bh % H, b = bh / H;
    if (row >= q_len[b * S]) return;
    float M = -INFINITY;
    for (int s2 = 0; s2 < S; ++s2) M = fmaxf(M, part_ml[((long long)(b * S + s2) * H + h) * max_q + row].x);
    float L = 0.f;
    float2 acc = make_float2(0.f, 0.f);
    for (int s2 = 0; s2 < S; ++s2) {
        const long long pi = ((long long)(b * S + s2) * H + h) * max_q + row;
        const float2 ml = part_ml[pi];
        if (!(ml.y > 0.f)) continue;
        const float wgt = fast_exp2(ml.x - M);
        L += ml.y * wgt;
        const float2 ov = reinterpret_cast<const float2*>(part_o + pi * 64)[lane];
        acc.x = fmaf(ov.x, wgt, acc.x);
        acc.y = fmaf(ov.y, wgt, acc.y);
    }
    const float inv = (L > 0.f) ? 1.0f / L : 0.f;
    reinterpret_cast<uint32_t*>(out + (long long)(q_start[b * S] + row) * ldo + out_col0 + h * DK)[lane] =
        pack_bf16x2(acc.x * inv, acc.y * inv);
}

// ----------------------------------------------------------------------------------------------
__global__ void relpos_kprep_kernel(const __nv_bfloat16* __restrict__ k, long long ldk,
                                    const float* __restrict__ P, const int* __restrict__ row_pos,
                                    const float* __restrict__ bias_u, const float* __restrict__ bias_v, int M,
                                    int heads, __nv_bfloat16* __restrict__ kp, long long ldkp,
                                    float* __restrict__ kbias, float out_scale) {
    // one warp per (row, head): 64 channels = 2 per lane
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= (long long)M * heads) return;
    const long long row = gw / heads;
    const int h = (int)(gw - row * heads);
    const int d = heads * DK;
    const int col = h * DK + 2 * lane;
    const uint32_t kk = *reinterpret_cast<const uint32_t*>(k + row * ldk + col);
    const float k0 = bf16_lo(kk), k1 = bf16_hi(kk);
    const float2 p = *reinterpret_cast<const float2*>(P + (long long)row_pos[row] * d + col);
    const float2 u = *reinterpret_cast<const float2*>(bias_u + col);
    const float2 v = *reinterpret_cast<const float2*>(bias_v + col);
    *reinterpret_cast<uint32_t*>(kp + row * ldkp + col) = pack_bf16x2(k0 + p.x, k1 + p.y);
    float c = u.x * k0 + u.y * k1 + v.x * p.x + v.y * p.y;
    c = warp_sum(c);
    if (lane == 0) kbias[row * heads + h] = c * out_scale;
}

}  // namespace

int attention_forward(const AttnArgs& a, cudaStream_t stream) {
    if (a.batch <= 0 || a.max_q_len <= 0) return WB_OK;
    constexpr int KN = 64;
    CUtensorMap tq, tk, tv;
    int rc;
    // the maps cover the whole row width so that column offsets select the head
    if ((rc = make_tmap_2d_bf16(&tq, a.q, (uint64_t)a.q_rows, (uint64_t)a.ldq, (uint64_t)a.ldq, 128, 64)) != WB_OK) return rc;
    if ((rc = make_tmap_2d_bf16(&tk, a.k, (uint64_t)a.k_rows, (uint64_t)a.ldk, (uint64_t)a.ldk, KN, 64)) != WB_OK) return rc;
    if ((rc = make_tmap_2d_bf16(&tv, a.v, (uint64_t)a.v_rows, (uint64_t)a.ldv, (uint64_t)a.ldv, KN, 64)) != WB_OK) return rc;
    AttnDev P;
    P.kbias = a.kbias;
    P.ld_kbias = a.ld_kbias;
    P.q_start = a.q_start;
    P.q_len = a.q_len;
    P.k_start = a.k_start;
    P.k_len = a.k_len;
    P.q_col0 = a.q_col0;
    P.k_col0 = a.k_col0;
    P.v_col0 = a.v_col0;
    P.chunk_size = a.chunk_size;
    P.num_left_chunks = a.num_left_chunks;
    P.scale_log2e = a.scale * 1.4426950408889634f;
    P.out = reinterpret_cast<__nv_bfloat16*>(a.out);
    P.ldo = a.ldo;
    P.out_col0 = a.out_col0;
    P.split3_out = a.split3_out;
    P.split_width = a.heads * DK;
    P.v_mode = a.v_mode;
    P.kbias_scaled = (a.kbias != nullptr && a.kbias_scaled) ? 1 : 0;
    P.part_o = a.part_o;
    P.part_ml = reinterpret_cast<float2*>(a.part_ml);
    P.part_max_q = a.max_q_len;
    WB_REQUIRE(a.part_o == nullptr || (a.part_ml != nullptr && a.splits >= 1 && a.batch % a.splits == 0 && !a.split3_out),
               WB_ERR_BAD_ARG, "attention: split-key mode needs part_ml, batch = blocks x splits, no split3 output");
    WB_REQUIRE((a.ldo % 8) == 0 && (a.out_col0 % 8) == 0, WB_ERR_BAD_ARG, "attention: output pitch/offset must be %%8");
    WB_SET_MAX_DYN_SMEM(attention_online_kernel, AttnCfg<KN>::kSmem);
    dim3 grid(ceil_div(a.max_q_len, AT_M), a.heads, a.batch);
    ProfScope _ps(PT_ATTENTION, stream, 0.0);
    WB_CHECK_CUDA(launch_maybe_pdl(attention_online_kernel, grid, dim3(128), AttnCfg<KN>::kSmem, stream, tq, tk, tv, P));
    count_launch();
    WB_CHECK_LAUNCH();
    if (a.part_o != nullptr) {
        const int blocks = a.batch / a.splits;
        const int total = blocks * a.heads * a.max_q_len;
        WB_CHECK_CUDA(launch_maybe_pdl(attention_merge_kernel, dim3(ceil_div(total * 32, 256)), dim3(256), 0, stream, a.part_o,
                                       P.part_ml, a.q_start, a.q_len, a.splits, a.heads, a.max_q_len, total, P.out, a.ldo,
                                       a.out_col0));
        count_launch();
        WB_CHECK_LAUNCH();
    }
    return WB_OK;
}

int relpos_kprep(const void* k_bf16, long long ldk, const float* P, const int* row_pos, const float* bias_u,
                 const float* bias_v, int M, int heads, void* kprime_bf16, long long ldkp, float* kbias,
                 cudaStream_t stream, float kbias_scale) {
    if (M <= 0) return WB_OK;
    const long long warps = (long long)M * heads;
    const int block = 256;
    const long long grid = (warps * 32 + block - 1) / block;
    ProfScope _ps(PT_KPREP, stream, (double)M * heads * 64 * 8.0);
    relpos_kprep_kernel<<<(unsigned)grid, block, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(k_bf16), ldk, P, row_pos, bias_u, bias_v, M, heads,
        reinterpret_cast<__nv_bfloat16*>(kprime_bf16), ldkp, kbias, kbias_scale);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

}  // namespace wb
