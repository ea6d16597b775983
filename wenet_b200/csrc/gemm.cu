// tcgen05 GEMM for every Linear / pointwise-conv / im2col-conv on the path.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T + bias[N] )        A, B bf16 (K-major), fp32 accumulate
//
// Replaces: nn.Linear / Conv1d(k=1) / Conv2d-as-im2col in the reference
//   (wenet/models/transformer/positionwise_feed_forward.py:50-58, attention.py:74-77,109-131,
//    convolution.py:46-53,88-95, subsampling.py:194-195,203-228, ctc.py:44, decoder.py:96-103).
//
// Structure (one persistent CTA per SM, 320 threads):
//   warp 0      : TMA producer   — cp.async.bulk.tensor loads of A (128x64) and B (BNx64) tiles, SWIZZLE_128B, ring of
//                                  kStages smem slots (mbarrier full/empty).  K <= 256: the [BN x K] weight panel of
//                                  the current n-tile stays resident and only A streams.  conv mode: the A tile is
//                                  one strided 3-D TMA box of the channels-last conv1 output (implicit im2col).
//   warp 1      : MMA issuer     — one lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per
//                                  k-block into a TMEM accumulator; tcgen05.commit frees smem slots / publishes the
//                                  accumulator
//   warps 2..9  : epilogue       — two warps per TMEM lane quarter; tcgen05.ld 32x32b.x32 (thread == accumulator
//                                  row) software-pipelined one 32-column chunk ahead; bias slice prefetched before the
//                                  accumulator wait and read back from shared memory; activation / GLU / alpha;
//                                  staged through SWIZZLE_128B shared memory and written by TMA store (bf16, fp32) or
//                                  TMA reduce-add (fp32 residual stream: no read-modify-write in the SM); EPI_LSE keeps
//                                  only per-row log-sum-exp partials.  TMEM holds two accumulator stages so the
//                                  epilogue of tile i overlaps the main loop of tile i+1.
// The epilogue variant is a template parameter for BN = 256 (one specialised kernel per variant; the generic kernel
// with a run-time switch is ~150 KB of SASS and stalls on instruction fetch) and a run-time switch for BN = 128.
// The in-kernel stall accounting (clock64 around every mbarrier wait, read by wb_gemm_diag) that was used to tune this
// file is compiled in only with -DWB_GEMM_DIAG.
#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>
#include <string.h>

namespace wb {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row

// NC = CTAs per tile: 1, or 2 = a CTA pair (cluster of two) computing a 256 x BN tile with ONE tcgen05.mma.cta_group::2 per
// k-step: each CTA holds its own 128 A rows and HALF of the B tile (and half of a resident weight panel), so a stage is
// 32 KB instead of 48 KB (6 instead of 4 stages), every SM ingests half the B bytes per MMA, and the K <= 256
// weight-stationary mode has room for an 8-slot A ring (two A tiles in flight) next to its 64 KB half panel.
template <int BN, int NC = 1>
struct GemmCfg {
    static constexpr int kStages = (BN == 256) ? (NC == 2 ? 6 : 4) : 6;
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2 / NC;   // per CTA
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BN;  // double-buffered accumulator (power of two)
    // epilogue staging: 8 warps x (32 rows x 128 B), SWIZZLE_128B, read back by TMA stores
    static constexpr int kOutBytes = 8 * 4096;
    // bias staging: 2 tile parities x BN floats (<= 2 KB), then 256 B of barriers; the dynamic smem window is declared
    // 1024-aligned, so no alignment slack is needed
    static constexpr int kBiasBytes = 2 * 256 * 4;
    static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + kBiasBytes + 256 /*barriers*/;
    // weight-stationary mode (K <= kResMaxKB * 64): the whole [BN x K] weight panel of the current n-tile
    // stays in shared memory while the CTA streams A tiles past it -> per tile only the 128 x K A tile is
    // fetched from L2 (the per-SM L2 path, ~80 GB/s, is what bounds the K = 256 GEMMs otherwise)
    static constexpr int kResMaxKB = (BN == 256) ? (NC == 2 ? 8 : 4) : 8;
    // the A ring takes whatever the resident panel of num_kb k-blocks leaves of the stage area (at most 8 slots: K = 256
    // gives 4 slots = ONE A tile with BN = 256 but 8 slots = TWO A tiles with BN = 128 - the depth that hides the
    // L2 -> SM latency of the next tile's A loads behind the current tile's MMAs)
    static constexpr int kMaxRing = 8;
    static constexpr int res_ring(int num_kb) {
        const int r = (kStages * kStageBytes - num_kb * kBBytes) / kABytes;
        return r > kMaxRing ? kMaxRing : r;
    }
};

struct GemmParams {
    int M, N, K;
    int epi;
    float alpha;
    const float* bias;
    void* out;
    long long ldc;
    int split3;
    int use_tma_out;  // epilogue through shared memory + TMA store / reduce-add (all non-split3 cases)
    int res_ring;     // weight-stationary mode: A ring depth (GemmCfg::res_ring(num_kb))
    int ksplit;       // streaming mode, EPI_RESID_F32 only: the K range is cut into ksplit pieces, each its own tile of the
                      // schedule; the pieces meet in the TMA reduce-add of the fp32 output (bias from piece 0 only)
    int num_m_tiles, num_n_tiles;
    // conv mode (Conv2d 3x3 stride 2 as an implicit GEMM): the A tile of k-block (kh, kw, c-block) is one 3-D TMA box
    // {64 channels, 19 frequency taps (element stride 2), 6 time taps (element stride 2)} of the channels-last conv1
    // output = 114 rows (t2, f2) of the im2col matrix, which is never materialised.
    int conv;
    int cblocks;            // d / 64
    const int4* tile_tab;   // per m-tile: .x t coordinate of tap kh = 0, .y first output row, .z valid rows (<= 114)
    float2* lse_part;       // EPI_LSE: [M][2 * num_n_tiles]
    // EPI_RESID_LN: LayerNorm applied to the updated residual rows (bf16 result through tmap_c2)
    const float* ln_g;
    const float* ln_b;
    float ln_eps;
    // EPI_RESID_LN2: x = LayerNorm(ln1_g, ln1_b)(updated rows) (fp32), bf16 result = LayerNorm(ln_g, ln_b)(x)
    const float* ln1_g;
    const float* ln1_b;
};
constexpr int kConvRows = 114;   // 6 x 19

// Stall accounting (wb_gemm_diag): cycles one lane of each role spent waiting, summed over CTAs and launches.
//   0 producer: ring slot not free      1 MMA: operands not landed     2 MMA: accumulator stage not drained
//   3 epilogue: accumulator not ready   4 epilogue: staging buffer still being read by a TMA store
//   5 epilogue: total time in the tile loop (warp 2)   6 CTA lifetime   7 tiles
// Compiled in only with -DWB_GEMM_DIAG (tools/bench_ops.py builds that way); the production kernel carries no clock reads
// and no atomics.
#ifdef WB_GEMM_DIAG
__device__ unsigned long long g_gemm_diag[12];   // 8: epilogue tcgen05.ld wait, 9: bias + activation, 10: staging stores + TMA issue
#define WB_TIMED_WAIT(slot, call)            \
    do {                                     \
        const long long _t0 = clock64();     \
        call;                                \
        diag[slot] += clock64() - _t0;       \
    } while (0)
#define WB_DIAG(...) __VA_ARGS__
#else
#define WB_TIMED_WAIT(slot, call) \
    do {                          \
        call;                     \
    } while (0)
#define WB_DIAG(...)
#endif

// EPI >= 0 fixes the epilogue at compile time (the kernel is ~150 KB of SASS with all six variants behind a runtime
// switch and then stalls on instruction fetch); EPI = -1 keeps the runtime switch (BN = 128: small / test models only).
template <int BN, bool BRES, int EPI, int NC>
__global__ void __launch_bounds__(320, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_c2,
                    GemmParams p) {
    using Cfg = GemmCfg<BN, NC>;
    pdl_launch_dependents();   // a PDL successor may set itself up (and fetch its weights) while this grid works
    const int rank = (NC == 2) ? (int)cluster_ctarank() : 0;          // 0 = leader (issues the MMAs)
    const int cta_id = (NC == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // tile-schedule id of this CTA (pair)
    const int num_ctas = (NC == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B tiles need 1024-B alignment
    const int num_kb = (p.K + BK - 1) / BK;
    const int kRing = BRES ? p.res_ring : Cfg::kStages;   // A (or A+B) ring depth
    uint8_t* smem_a = BRES ? smem + num_kb * Cfg::kBBytes : smem;
    uint8_t* smem_b = BRES ? smem : smem + Cfg::kStages * Cfg::kABytes;
    uint8_t* smem_out = smem + Cfg::kStages * Cfg::kStageBytes;   // 1024-aligned (stage sizes are multiples of 1024)
    float* smem_bias = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kOutBytes + Cfg::kBiasBytes);
    uint64_t* full_bar = bars;                // [kRing]
    uint64_t* empty_bar = bars + 8;           // [kRing]
    uint64_t* tmem_full = bars + 16;          // [2]
    uint64_t* tmem_empty = bars + 18;         // [2]
    uint64_t* b_full = bars + 20;             // resident weight panel landed
    uint64_t* b_empty = bars + 21;            // all MMAs reading the resident panel retired
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 22);
    static_assert(Cfg::kStages <= 8 && Cfg::kMaxRing <= 8, "barrier layout");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_mt = (NC == 2) ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles;   // m-tiles of the schedule (pairs of 128-row tiles)
    const int ksplit = BRES ? 1 : p.ksplit;
    const int kb_per = (num_kb + ksplit - 1) / ksplit;
    const int num_tiles = num_mt * p.num_n_tiles * ksplit;
    const int epi = (EPI >= 0) ? EPI : p.epi;
    WB_DIAG(long long diag[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long cta_t0 = clock64();)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        if (p.use_tma_out) tma_prefetch_desc(&tmap_c);
        if (p.conv || EPI == EPI_RESID_LN || EPI == EPI_RESID_LN2) tma_prefetch_desc(&tmap_c2);
        for (int s = 0; s < kRing; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(b_full, 1);
        mbar_init(b_empty, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 8 * NC);  // one arrive per epilogue warp (of both CTAs of a pair, on the leader's)
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        if (NC == 2) {
            tmem_alloc_pair(tmem_holder, Cfg::kTmemCols);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_holder, Cfg::kTmemCols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (NC == 2) cluster_sync_all();   // the peer's barriers must exist before anything is signalled across the pair
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    // tile schedule.  streaming mode: t = blockIdx.x + i * gridDim.x, n fastest.  weight-stationary mode: each
    // CTA owns a contiguous range of the n-major tile list, so its weight panel changes at most a few times.
    const int per_cta = (num_tiles + num_ctas - 1) / num_ctas;
    const int t_begin = BRES ? cta_id * per_cta : cta_id;
    const int t_end = BRES ? min(num_tiles, t_begin + per_cta) : num_tiles;
    const int t_step = BRES ? 1 : num_ctas;
    // (pair mode: this CTA's 128-row tile is 2 x (schedule m-tile) + rank; an odd last tile leaves the peer with rows past
    //  M: its TMA loads are zero-filled and its stores clipped)
#define WB_TILE_COORDS(t)                                                                                   \
    const int n_tile = BRES ? (t) / num_mt : (t) % p.num_n_tiles;                                          \
    const int mt_sp = BRES ? (t) % num_mt : (t) / p.num_n_tiles;   /* streaming: m-tile fastest, then K piece */ \
    const int k_piece = BRES ? 0 : mt_sp / num_mt;                                                         \
    const int m_tile = (BRES ? mt_sp : mt_sp - k_piece * num_mt) * NC + rank;                              \
    const int kb0 = k_piece * kb_per, kb1 = min(num_kb, kb0 + kb_per);

    if (warp == 0) {
        // ===================== TMA producer =====================
        // pair mode: both CTAs load (their A rows, their half of B); every load completes on the LEADER's barrier, on which
        // the leader alone posts the expected bytes of both; each CTA waits for its own (multicast-released) slots
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int cur_n = -1;
            uint32_t bemp_phase = 0;
            const int b_row0 = (NC == 2) ? rank * (BN / 2) : 0;
            // PDL: the weights are never written by a kernel, so the first tile's weight boxes are requested BEFORE waiting
            // for the predecessor grid (the activations, `tile_tab` and everything the epilogue touches come after it).
            // Streaming mode: the first ring pass' B boxes (with their barriers' byte counts); weight-stationary mode: the
            // panel is loaded ahead of the first A tile anyway, the wait sits behind it (`pdl_pending`).
            int pre_b = 0;
            bool pdl_pending = true;
            if (!BRES && t_begin < t_end && !p.conv) {
                WB_TILE_COORDS(t_begin)
                (void)m_tile;
                pre_b = min(kRing, kb1 - kb0);
                for (int s = 0; s < pre_b; ++s) {
                    if (rank == 0) mbar_expect_tx(&full_bar[s], Cfg::kStageBytes * NC);
                    if (NC == 2)
                        tma_load_2d_pair(smem_b + s * Cfg::kBBytes, &tmap_b, mapa_u32(&full_bar[s], 0), (kb0 + s) * BK,
                                         n_tile * BN + b_row0);
                    else
                        tma_load_2d(smem_b + s * Cfg::kBBytes, &tmap_b, &full_bar[s], (kb0 + s) * BK, n_tile * BN);
                }
            }
            if (!BRES) {
                pdl_wait();
                pdl_pending = false;
            }
            for (int t = t_begin; t < t_end; t += t_step) {
                WB_TILE_COORDS(t)
                if (BRES && n_tile != cur_n) {
                    if (cur_n >= 0) {  // the previous panel must not be overwritten while MMAs still read it
                        mbar_wait(b_empty, bemp_phase);
                        bemp_phase ^= 1;
                    }
                    if (rank == 0) mbar_expect_tx(b_full, (uint32_t)num_kb * Cfg::kBBytes * NC);
                    for (int kb = 0; kb < num_kb; ++kb) {
                        if (NC == 2)
                            tma_load_2d_pair(smem_b + kb * Cfg::kBBytes, &tmap_b, mapa_u32(b_full, 0), kb * BK,
                                             n_tile * BN + b_row0);
                        else
                            tma_load_2d(smem_b + kb * Cfg::kBBytes, &tmap_b, b_full, kb * BK, n_tile * BN);
                    }
                    cur_n = n_tile;
                }
                if (pdl_pending) {
                    pdl_wait();
                    pdl_pending = false;
                }
                int conv_t = 0;
                if (!BRES && p.conv) conv_t = (m_tile < p.num_m_tiles) ? __ldg(&p.tile_tab[m_tile]).x : 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    const bool b_done = pre_b > 0;   // this slot's byte count and B box were issued ahead of the PDL wait
                    if (b_done) --pre_b;
                    WB_TIMED_WAIT(0, mbar_wait(&empty_bar[stage], phase ^ 1));
                    const uint32_t full_addr = (NC == 2) ? mapa_u32(&full_bar[stage], 0) : 0u;
                    if (!BRES && p.conv) {
                        const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
                        const int kh = tap / 3, kw = tap - 3 * kh;
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], (kConvRows * 128 + Cfg::kBBytes) * NC);
                        if (NC == 2)
                            tma_load_3d_pair(smem_a + stage * Cfg::kABytes, &tmap_a, full_addr, cb * BK, kw, conv_t + kh);
                        else
                            tma_load_3d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], cb * BK, kw, conv_t + kh);
                    } else {
                        if (rank == 0 && !b_done) mbar_expect_tx(&full_bar[stage], (BRES ? Cfg::kABytes : Cfg::kStageBytes) * NC);
                        if (NC == 2)
                            tma_load_2d_pair(smem_a + stage * Cfg::kABytes, &tmap_a, full_addr, kb * BK, m_tile * BM);
                        else
                            tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * BK, m_tile * BM);
                    }
                    if (!BRES && !b_done) {
                        if (NC == 2)
                            tma_load_2d_pair(smem_b + stage * Cfg::kBBytes, &tmap_b, full_addr, kb * BK, n_tile * BN + b_row0);
                        else
                            tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * BK, n_tile * BN);
                    }
                    if (++stage == kRing) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0 && rank == 0) {   // (pair mode: the leader issues the M = 256 MMAs for both CTAs)
            constexpr uint32_t idesc = make_idesc_bf16(BM * NC, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int cur_n = -1;
            uint32_t bfull_phase = 0;
            for (int t = t_begin; t < t_end; t += t_step) {
                WB_TILE_COORDS(t)
                (void)m_tile;
                if (BRES && n_tile != cur_n) {
                    mbar_wait(b_full, bfull_phase);
                    bfull_phase ^= 1;
                    cur_n = n_tile;
                }
                WB_TIMED_WAIT(2, mbar_wait(&tmem_empty[acc], acc_phase ^ 1));
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    WB_TIMED_WAIT(1, mbar_wait(&full_bar[stage], phase));
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::kABytes);
                    const uint32_t b_addr = smem_u32(smem_b + (BRES ? kb : stage) * Cfg::kBBytes);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
                        if (NC == 2) umma_f16_pair(tmem_d, adesc, bdesc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
                        else umma_f16(tmem_d, adesc, bdesc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
                    }
                    // smem slot free (in both CTAs of a pair) once these MMAs retire
                    if (NC == 2) umma_commit_pair(&empty_bar[stage]);
                    else umma_commit(&empty_bar[stage]);
                    if (++stage == kRing) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (NC == 2) umma_commit_pair(&tmem_full[acc]);   // accumulator complete
                else umma_commit(&tmem_full[acc]);
                if (BRES && t + t_step < t_end) {
                    const int next_n = (t + t_step) / num_mt;
                    if (next_n != n_tile) {   // panel may be replaced once these MMAs retire
                        if (NC == 2) umma_commit_pair(b_empty);
                        else umma_commit(b_empty);
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        // two warps per TMEM lane quarter; each takes half of the tile's 32-column chunks
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        constexpr int kChunksPerWarp = BN / 64;
        int acc = 0;
        uint32_t acc_phase = 0;
        bool need_wait = false;
        int tile_par = 0;
        pdl_wait();   // residual stream / output buffers belong to the predecessor grid until it has completed
        WB_DIAG(const long long epi_t0 = clock64();)
        for (int t = t_begin; t < t_end; t += t_step) {
            WB_TILE_COORDS(t)
            long long row_base = (long long)m_tile * BM;
            int n_in = 32;   // valid rows in this warp's 32-row slice
            if (!BRES && p.conv) {
                const int4 tt = (m_tile < p.num_m_tiles) ? __ldg(&p.tile_tab[m_tile]) : make_int4(0, 0, 0, 0);
                row_base = tt.y;
                n_in = min(32, max(0, tt.z - q * 32));
            }
            const long long row = row_base + q * 32 + lane;
            const bool row_ok = (!BRES && p.conv) ? (lane < n_in) : (row < p.M);
            // conv tiles hold 114 rows: slices of 32 and 18 rows leave through TMA boxes of that height, the ragged
            // last tile of an utterance falls back to guarded register stores
            const bool warp_tma = p.use_tma_out && (n_in == 32 || n_in == 18);
            const CUtensorMap* cmap = (n_in == 18) ? &tmap_c2 : &tmap_c;
            if (EPI == EPI_RESID_LN || EPI == EPI_RESID_LN2) {
                // ---- residual update + LayerNorm(s) of the updated row in one epilogue (N == BN == 256: the tile holds whole
                // rows).  Thread = one row x 128 columns (its TMEM lane, this warp's column half).
                //   pass 1: y = x_old + alpha * (acc + bias), parked back in TMEM in place of the accumulator; running
                //           (mean, M2) merged chunk by chunk (Chan's update).  EPI_RESID_LN: x = y (fp32, staged, TMA store)
                //   exchange (mean, M2) with the warp that owns the other column half (named barrier of the two warps)
                //   EPI_RESID_LN2 only - pass 2: x = LayerNorm_1(y) (fp32 store), again parked in TMEM, statistics of x
                //   last pass: tcgen05.ld -> (v - mean) * rstd * gamma + beta -> bf16 -> staged -> TMA store
                constexpr bool kTwo = (EPI == EPI_RESID_LN2);
                const uint32_t taddr0 =
                    tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN) + (uint32_t)(half * 128);
                // x_old is fetched coalesced (instruction j: lane l reads 16 bytes of row 4 j + l / 8, a warp covers 4 rows x
                // 128 B), one chunk ahead, and transposed to "thread = row" through this warp's staging buffer
                const int sub_r = lane >> 3, sub_c = lane & 7;
                const long long wrow0 = row_base + q * 32;
                const float* xwarp = reinterpret_cast<const float*>(p.out) + wrow0 * p.ldc + half * 128;
                float4 xo[8];
                auto load_x = [&](int ci) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int rr = 4 * j + sub_r;
                        xo[j] = (wrow0 + rr < p.M)
                                    ? __ldcg(reinterpret_cast<const float4*>(xwarp + (long long)rr * p.ldc + ci * 32) + sub_c)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                };
                load_x(0);
                WB_TIMED_WAIT(3, mbar_wait(&tmem_full[acc], acc_phase));
                tc_fence_after();
                uint32_t r[32];
                tmem_ld_32x32b_x32(taddr0, r);
                const int sw = lane & 7;
                uint8_t* sbuf_warp = smem_out + (warp - 2) * 4096;
                uint8_t* sbuf = sbuf_warp + lane * 128;
                const int row0 = (int)wrow0;
                auto staging_ready = [&]() {
                    if (need_wait) {
                        WB_TIMED_WAIT(4, if (lane == 0) tma_store_wait_read<0>(); __syncwarp());
                        need_wait = false;
                    }
                };
                // running statistics over the chunks seen so far (ci is a compile-time constant after unrolling)
                auto stats_merge = [&](const uint32_t (&v)[32], int ci, float& mean, float& m2) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        s0 += __uint_as_float(v[i]);
                        s1 += __uint_as_float(v[i + 1]);
                        s2 += __uint_as_float(v[i + 2]);
                        s3 += __uint_as_float(v[i + 3]);
                    }
                    const float cm = ((s0 + s1) + (s2 + s3)) * (1.0f / 32.0f);
                    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float d0 = __uint_as_float(v[i]) - cm, d1 = __uint_as_float(v[i + 1]) - cm;
                        const float d2 = __uint_as_float(v[i + 2]) - cm, d3 = __uint_as_float(v[i + 3]) - cm;
                        q0 = fmaf(d0, d0, q0);
                        q1 = fmaf(d1, d1, q1);
                        q2 = fmaf(d2, d2, q2);
                        q3 = fmaf(d3, d3, q3);
                    }
                    const float na = 32.0f * ci, nt = na + 32.0f;
                    const float delta = cm - mean;
                    mean += delta * (32.0f / nt);
                    m2 += ((q0 + q1) + (q2 + q3)) + delta * delta * (na * 32.0f / nt);
                };
                // fp32 chunk (thread = row) -> staging -> TMA store into x
                auto store_x = [&](const uint32_t (&v)[32], int n0) {
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        *reinterpret_cast<uint4*>(sbuf + ((u ^ sw) << 4)) =
                            make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmap_c, sbuf_warp, n0, row0);
                        tma_store_commit();
                    }
                    need_wait = true;
                };
                float2* sx = reinterpret_cast<float2*>(smem_bias);   // [column half][tile row] (no bias staging in this variant)
                auto row_stats = [&](float mean, float m2, float& mu, float& rstd) {
                    sx[half * 128 + q * 32 + lane] = make_float2(mean, m2);
                    named_bar_sync(1 + q, 64);
                    const float2 oth = sx[(half ^ 1) * 128 + q * 32 + lane];
                    named_bar_sync(1 + q, 64);   // the slot is rewritten by the next exchange
                    const float dm = oth.x - mean;
                    mu = 0.5f * (mean + oth.x);
                    rstd = rsqrtf((m2 + oth.y + dm * dm * 64.0f) * (1.0f / 256.0f) + p.ln_eps);
                };
                float mean = 0.f, m2 = 0.f;
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const int n0 = half * 128 + ci * 32;
                    staging_ready();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int rr = 4 * j + sub_r;
                        *reinterpret_cast<float4*>(sbuf_warp + rr * 128 + ((sub_c ^ (rr & 7)) << 4)) = xo[j];
                    }
                    __syncwarp();
                    if (ci + 1 < 4) load_x(ci + 1);
                    tmem_ld_wait_regs(r);
                    uint32_t y[32];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + u);
                        const float4 x4 = *reinterpret_cast<const float4*>(sbuf + ((u ^ sw) << 4));
                        y[4 * u] = __float_as_uint(fmaf(p.alpha, __uint_as_float(r[4 * u]) + b4.x, x4.x));
                        y[4 * u + 1] = __float_as_uint(fmaf(p.alpha, __uint_as_float(r[4 * u + 1]) + b4.y, x4.y));
                        y[4 * u + 2] = __float_as_uint(fmaf(p.alpha, __uint_as_float(r[4 * u + 2]) + b4.z, x4.z));
                        y[4 * u + 3] = __float_as_uint(fmaf(p.alpha, __uint_as_float(r[4 * u + 3]) + b4.w, x4.w));
                    }
                    tmem_st_32x32b_x32(taddr0 + (uint32_t)(ci * 32), y);
                    if (ci + 1 < 4) tmem_ld_32x32b_x32(taddr0 + (uint32_t)((ci + 1) * 32), r);
                    stats_merge(y, ci, mean, m2);
                    // (a thread reads and rewrites only its own row of the buffer: no barrier between the two)
                    if (!kTwo) store_x(y, n0);
                    else __syncwarp();   // every lane has read its row before the next chunk's x_old overwrites the buffer
                }
                float mu, rstd;
                row_stats(mean, m2, mu, rstd);
                tmem_st_wait();
                tmem_ld_32x32b_x32(taddr0, r);
                if (kTwo) {
                    mean = 0.f;
                    m2 = 0.f;
#pragma unroll
                    for (int ci = 0; ci < 4; ++ci) {
                        const int n0 = half * 128 + ci * 32;
                        tmem_ld_wait_regs(r);
                        uint32_t y[32];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln1_g + n0) + u);
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln1_b + n0) + u);
                            const float a0 = rstd * g4.x, a1 = rstd * g4.y, a2 = rstd * g4.z, a3 = rstd * g4.w;
                            y[4 * u] = __float_as_uint(fmaf(__uint_as_float(r[4 * u]), a0, fmaf(-mu, a0, b4.x)));
                            y[4 * u + 1] = __float_as_uint(fmaf(__uint_as_float(r[4 * u + 1]), a1, fmaf(-mu, a1, b4.y)));
                            y[4 * u + 2] = __float_as_uint(fmaf(__uint_as_float(r[4 * u + 2]), a2, fmaf(-mu, a2, b4.z)));
                            y[4 * u + 3] = __float_as_uint(fmaf(__uint_as_float(r[4 * u + 3]), a3, fmaf(-mu, a3, b4.w)));
                        }
                        tmem_st_32x32b_x32(taddr0 + (uint32_t)(ci * 32), y);
                        if (ci + 1 < 4) tmem_ld_32x32b_x32(taddr0 + (uint32_t)((ci + 1) * 32), r);
                        stats_merge(y, ci, mean, m2);
                        staging_ready();
                        store_x(y, n0);
                    }
                    row_stats(mean, m2, mu, rstd);
                    tmem_st_wait();
                    tmem_ld_32x32b_x32(taddr0, r);
                }
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    const int n0 = half * 128 + ci * 32;
                    tmem_ld_wait_regs(r);
                    uint32_t pk[16];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln_g + n0) + u);
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln_b + n0) + u);
                        const float a0 = rstd * g4.x, a1 = rstd * g4.y, a2 = rstd * g4.z, a3 = rstd * g4.w;
                        const float o0 = fmaf(__uint_as_float(r[4 * u]), a0, fmaf(-mu, a0, b4.x));
                        const float o1 = fmaf(__uint_as_float(r[4 * u + 1]), a1, fmaf(-mu, a1, b4.y));
                        const float o2 = fmaf(__uint_as_float(r[4 * u + 2]), a2, fmaf(-mu, a2, b4.z));
                        const float o3 = fmaf(__uint_as_float(r[4 * u + 3]), a3, fmaf(-mu, a3, b4.w));
                        pk[2 * u] = pack_bf16x2(o0, o1);
                        pk[2 * u + 1] = pack_bf16x2(o2, o3);
                    }
                    if (ci + 1 < 4) tmem_ld_32x32b_x32(taddr0 + (uint32_t)((ci + 1) * 32), r);
                    if ((ci & 1) == 0) staging_ready();
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        *reinterpret_cast<uint4*>(sbuf + ((((ci & 1) * 4 + u) ^ sw) << 4)) =
                            make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
                    if (ci & 1) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmap_c2, sbuf_warp, half * 128 + (ci & ~1) * 32, row0);
                            tma_store_commit();
                        }
                        need_wait = true;
                    }
                }
            } else {
            // this warp's slice of the bias (kChunksPerWarp x 32 columns) is fetched before the accumulator wait and
            // parked in shared memory (one copy per tile parity; the four warps of a column half write identical
            // values), so the chunk loop reads it with broadcast LDS instead of an L2 round trip per chunk
            float4 bpre = make_float4(0.f, 0.f, 0.f, 0.f);
            const int bcol = n_tile * BN + half * (kChunksPerWarp * 32) + 4 * lane;
            const bool bias_on = (p.bias != nullptr) && (k_piece == 0);   // K pieces after the first add no bias
            if (bias_on && 4 * lane < kChunksPerWarp * 32) {
                if (bcol + 0 < p.N) bpre.x = __ldg(p.bias + bcol + 0);
                if (bcol + 1 < p.N) bpre.y = __ldg(p.bias + bcol + 1);
                if (bcol + 2 < p.N) bpre.z = __ldg(p.bias + bcol + 2);
                if (bcol + 3 < p.N) bpre.w = __ldg(p.bias + bcol + 3);
            }
            float* sbias = smem_bias + (tile_par * 256 + half * (kChunksPerWarp * 32));
            WB_TIMED_WAIT(3, mbar_wait(&tmem_full[acc], acc_phase));
            tc_fence_after();
            if (4 * lane < kChunksPerWarp * 32) *reinterpret_cast<float4*>(sbias + 4 * lane) = bpre;
            __syncwarp();
            tile_par ^= 1;
            const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            // software pipeline over the warp's 32-column chunks: the tcgen05.ld of chunk c+1 is in flight while chunk c
            // goes through bias / activation / staging (two register sets, loop fully unrolled)
            uint32_t rbuf[2][32];
            float lse_m = -INFINITY, lse_s = 0.f;   // EPI_LSE: running (max, sum) over this warp's columns, log2 domain
            constexpr int c_begin_rel = 0;
            const int c0 = half * kChunksPerWarp;
            if (n_tile * BN + c0 * 32 < p.N) tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c0 * 32), rbuf[0]);
#pragma unroll
            for (int ci = c_begin_rel; ci < kChunksPerWarp; ++ci) {
                const int c = c0 + ci;
                const int n0 = n_tile * BN + c * 32;
                if (n0 >= p.N) break;  // warp-uniform
                uint32_t (&r)[32] = rbuf[ci & 1];
                WB_TIMED_WAIT(8, tmem_ld_wait_regs(r));
                WB_DIAG(const long long t_math0 = clock64();)
                if (ci + 1 < kChunksPerWarp && n0 + 32 < p.N)
                    tmem_ld_32x32b_x32(taddr0 + (uint32_t)((c + 1) * 32), rbuf[(ci + 1) & 1]);
                float v[32];
                const bool full = (n0 + 32 <= p.N);
                if (bias_on) {
                    const float* sb = sbias + ci * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(sb + i);
                        const float2 lo = f2_add(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])),
                                                 make_float2(b4.x, b4.y));
                        const float2 hi = f2_add(make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])),
                                                 make_float2(b4.z, b4.w));
                        v[i] = lo.x;
                        v[i + 1] = lo.y;
                        v[i + 2] = hi.x;
                        v[i + 3] = hi.y;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                }
                if (epi == EPI_LSE) {
                    constexpr float kLog2e = 1.4426950408889634f;
                    float cm = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] = (full || n0 + i < p.N) ? v[i] * kLog2e : -INFINITY;
                        cm = fmaxf(cm, v[i]);
                    }
                    const float mn = fmaxf(lse_m, cm);
                    float acc_s = (lse_m == -INFINITY) ? 0.f : lse_s * fast_exp2(lse_m - mn);
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc_s += fast_exp2(v[i] - mn);
                    lse_s = acc_s;
                    lse_m = mn;
                    continue;
                }
                if (warp_tma) {
                    // ---- staged epilogue: registers -> swizzled smem (row = lane, 128 B) -> TMA ----
                    // rows >= M and columns >= N are clipped by the tensor map, so no guards are needed.
                    uint8_t* sbuf = smem_out + (warp - 2) * 4096 + lane * 128;
                    // the previous TMA store of this warp must have finished reading the staging buffer before it is
                    // overwritten; the check sits after the activation math so that latency is hidden behind it
                    auto staging_ready = [&]() {
                        if (need_wait) {
                            WB_TIMED_WAIT(4, if (lane == 0) tma_store_wait_read<0>(); __syncwarp());
                            need_wait = false;
                        }
                    };
                    const int sw = lane & 7;
                    WB_DIAG(long long t_st0 = 0;)
                    bool flush = false;
                    int out_col = 0;
                    if (epi == EPI_F32 || epi == EPI_RESID_F32) {
                        WB_DIAG(diag[9] += clock64() - t_math0;)
                        staging_ready();
                        WB_DIAG(t_st0 = clock64();)
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            *reinterpret_cast<float4*>(sbuf + ((u ^ sw) << 4)) =
                                make_float4(p.alpha * v[4 * u], p.alpha * v[4 * u + 1], p.alpha * v[4 * u + 2],
                                            p.alpha * v[4 * u + 3]);
                        flush = true;
                        out_col = n0;
                    } else if (epi == EPI_GLU_BF16) {
                        float g[16];
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            float2 sa, sb2;
                            sigmoid4_f2(make_float2(v[16 + i], v[17 + i]), make_float2(v[18 + i], v[19 + i]), sa, sb2);
                            const float2 ga = f2_mul(sa, make_float2(v[i], v[i + 1]));
                            const float2 gb = f2_mul(sb2, make_float2(v[i + 2], v[i + 3]));
                            g[i] = ga.x;
                            g[i + 1] = ga.y;
                            g[i + 2] = gb.x;
                            g[i + 3] = gb.y;
                        }
                        WB_DIAG(diag[9] += clock64() - t_math0;)
                        staging_ready();
                        WB_DIAG(t_st0 = clock64();)
#pragma unroll
                        for (int u = 0; u < 2; ++u)
                            *reinterpret_cast<uint4*>(sbuf + ((((c & 3) * 2 + u) ^ sw) << 4)) =
                                make_uint4(pack_bf16x2(g[8 * u], g[8 * u + 1]), pack_bf16x2(g[8 * u + 2], g[8 * u + 3]),
                                           pack_bf16x2(g[8 * u + 4], g[8 * u + 5]), pack_bf16x2(g[8 * u + 6], g[8 * u + 7]));
                        flush = ((c & 3) == 3) || (n0 + 32 >= p.N);
                        out_col = (n_tile * BN + (c & ~3) * 32) >> 1;
                    } else {
                        if (epi == EPI_BF16_SILU) {
                            silu_inplace(v);
                        } else if (epi == EPI_BF16_RELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        } else if (epi == EPI_BF16_GELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
                        }
                        if (p.alpha != 1.0f) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                        }
                        WB_DIAG(diag[9] += clock64() - t_math0;)
                        staging_ready();
                        WB_DIAG(t_st0 = clock64();)
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            *reinterpret_cast<uint4*>(sbuf + ((((c & 1) * 4 + u) ^ sw) << 4)) = make_uint4(
                                pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                                pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
                        flush = ((c & 1) == 1) || (n0 + 32 >= p.N);
                        out_col = n_tile * BN + (c & ~1) * 32;
                    }
                    if (flush) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            const void* src = smem_out + (warp - 2) * 4096;
                            const int row0 = (int)row_base + q * 32;
                            if (epi == EPI_RESID_F32)
                                tma_reduce_add_2d(cmap, src, out_col, row0);
                            else
                                tma_store_2d(cmap, src, out_col, row0);
                            tma_store_commit();
                        }
                        need_wait = true;
                    }
                    WB_DIAG(diag[10] += clock64() - t_st0;)
                    continue;
                }
                if (!row_ok) continue;

                switch (epi) {
                    case EPI_BF16:
                    case EPI_BF16_SILU:
                    case EPI_BF16_RELU:
                    case EPI_BF16_GELU: {
                        if (epi == EPI_BF16_SILU) {
                            silu_inplace(v);
                        } else if (epi == EPI_BF16_RELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        } else if (epi == EPI_BF16_GELU) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
                        }
                        if (p.alpha != 1.0f) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
                        }
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldc + n0;
                        const bool vec = full && ((p.ldc & 7) == 0) && ((p.N & 7) == 0);
                        if (vec) {
                            uint32_t pk[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                reinterpret_cast<uint4*>(o)[i] =
                                    make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                            if (p.split3) {
                                uint32_t lo[16];
#pragma unroll
                                for (int i = 0; i < 16; ++i)
                                    lo[i] = pack_bf16x2(v[2 * i] - bf16_lo(pk[i]),
                                                        v[2 * i + 1] - bf16_hi(pk[i]));
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    reinterpret_cast<uint4*>(o + p.N)[i] =
                                        make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                                    reinterpret_cast<uint4*>(o + 2 * (long long)p.N)[i] =
                                        make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
                                }
                            }
                        } else {
                            for (int i = 0; i < 32; ++i) {
                                if (n0 + i < p.N) {
                                    const __nv_bfloat16 h = __float2bfloat16_rn(v[i]);
                                    o[i] = h;
                                    if (p.split3) {
                                        o[p.N + i] = __float2bfloat16_rn(v[i] - __bfloat162float(h));
                                        o[2 * (long long)p.N + i] = h;
                                    }
                                }
                            }
                        }
                    } break;
                    case EPI_GLU_BF16: {
                        // weight rows are packed in groups of 32: [16 value rows | 16 gate rows]
                        const int on = p.N >> 1;
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldc + (n0 >> 1);
                        float g[16];
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            sigmoid4(v[16 + i], v[17 + i], v[18 + i], v[19 + i], g[i], g[i + 1], g[i + 2], g[i + 3]);
                            g[i] *= v[i];
                            g[i + 1] *= v[i + 1];
                            g[i + 2] *= v[i + 2];
                            g[i + 3] *= v[i + 3];
                        }
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(g[2 * i], g[2 * i + 1]);
                        reinterpret_cast<uint4*>(o)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        reinterpret_cast<uint4*>(o)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        if (p.split3) {
                            uint32_t lo[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                lo[i] = pack_bf16x2(g[2 * i] - bf16_lo(pk[i]), g[2 * i + 1] - bf16_hi(pk[i]));
                            reinterpret_cast<uint4*>(o + on)[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                            reinterpret_cast<uint4*>(o + on)[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                            reinterpret_cast<uint4*>(o + 2 * on)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            reinterpret_cast<uint4*>(o + 2 * on)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        }
                    } break;
                    case EPI_RESID_F32: {
                        float* o = reinterpret_cast<float*>(p.out) + row * p.ldc + n0;
                        if (full && ((p.ldc & 3) == 0)) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float4 x = reinterpret_cast<float4*>(o)[i];
                                x.x += p.alpha * v[4 * i];
                                x.y += p.alpha * v[4 * i + 1];
                                x.z += p.alpha * v[4 * i + 2];
                                x.w += p.alpha * v[4 * i + 3];
                                reinterpret_cast<float4*>(o)[i] = x;
                            }
                        } else {
                            for (int i = 0; i < 32; ++i)
                                if (n0 + i < p.N) o[i] += p.alpha * v[i];
                        }
                    } break;
                    case EPI_F32:
                    default: {
                        float* o = reinterpret_cast<float*>(p.out) + row * p.ldc + n0;
                        if (full && ((p.ldc & 3) == 0)) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                reinterpret_cast<float4*>(o)[i] =
                                    make_float4(p.alpha * v[4 * i], p.alpha * v[4 * i + 1],
                                                p.alpha * v[4 * i + 2], p.alpha * v[4 * i + 3]);
                        } else {
                            for (int i = 0; i < 32; ++i)
                                if (n0 + i < p.N) o[i] = p.alpha * v[i];
                        }
                    } break;
                }
            }
            if (epi == EPI_LSE && row_ok)
                p.lse_part[row * (2 * p.num_n_tiles) + n_tile * 2 + half] = make_float2(lse_m, lse_s);
            }   // EPI != EPI_RESID_LN / EPI_RESID_LN2
            // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (NC == 2 && rank != 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));   // the leader's barrier
                else mbar_arrive(&tmem_empty[acc]);
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (p.use_tma_out && lane == 0) tma_store_wait<0>();  // smem must outlive the bulk stores
        WB_DIAG(diag[5] = clock64() - epi_t0;)
    }

#undef WB_TILE_COORDS
#ifdef WB_GEMM_DIAG
    if (lane == 0 && warp <= 2) {   // one lane per role: producer (0), MMA issuer (1), first epilogue warp (2)
        for (int i = 0; i < 12; ++i)
            if (i != 6 && i != 7 && diag[i] != 0) atomicAdd(&g_gemm_diag[i], (unsigned long long)diag[i]);
        if (warp == 0) {
            atomicAdd(&g_gemm_diag[6], (unsigned long long)(clock64() - cta_t0));
            int nt = 0;
            for (int t = t_begin; t < t_end; t += t_step) ++nt;
            atomicAdd(&g_gemm_diag[7], (unsigned long long)nt);
        }
    }
#endif
    tc_fence_before();
    if (NC == 2) cluster_sync_all();   // neither CTA may leave while the pair's MMAs still read its shared memory
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (NC == 2) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
        else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

int g_sm_reserve = 0;  // SMs left free for concurrently running latency-bound kernels on other streams

template <int BN, bool BRES, int EPI, int NC>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tc2,
                const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, NC>;
    WB_SET_MAX_DYN_SMEM((gemm_tcgen05_kernel<BN, BRES, EPI, NC>), Cfg::kSmemBytes);
    const int num_sms = current_device_sms();
    WB_REQUIRE(num_sms > 0, WB_ERR_CUDA, "gemm: cannot query the SM count of the current device");
    const int usable = (num_sms - g_sm_reserve) > 1 ? (num_sms - g_sm_reserve) : 1;
    ProfScope _ps((EPI == EPI_RESID_LN || EPI == EPI_RESID_LN2) ? PT_GEMM_LN : PT_GEMM, stream,
                  2.0 * (double)p.M * (double)p.N * (double)p.K);
    if (NC == 2) {
        // one CTA pair (cluster of two, same TPC) per two SMs; the schedule walks pairs of 128-row tiles
        const int tiles = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
        const int pairs = tiles < usable / 2 ? tiles : (usable / 2 > 0 ? usable / 2 : 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(320);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_active() ? 2 : 1;
        WB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, BRES, EPI, NC>, ta, tb, tc, tc2, p));
    } else {
        const int tiles = p.num_m_tiles * p.num_n_tiles * (BRES ? 1 : p.ksplit);
        const int grid = tiles < usable ? tiles : usable;
        if (pdl_active()) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(320);
            cfg.dynamicSmemBytes = Cfg::kSmemBytes;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            WB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, BRES, EPI, NC>, ta, tb, tc, tc2, p));
        } else {
            gemm_tcgen05_kernel<BN, BRES, EPI, NC><<<grid, 320, Cfg::kSmemBytes, stream>>>(ta, tb, tc, tc2, p);
        }
    }
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}

// CTA-pair (cta_group::2) tiles for the 256-column GEMMs with more than one 128-row tile.  Measured
// (profiles/r2_ops_gemm_2cta.txt): the streaming K > 256 shapes gain (FFN2 67.6 -> 63.5 us), the weight-stationary
// K <= 256 shapes lose (FFN1 82 -> 94 us, QKV 33 -> 35 us: the leader's MMAs wait for the slower of two epilogues /
// producers every tile), so the default is pair tiles for K > 256 only.  WB_GEMM_2CTA=0: never, =2: always.
int g_gemm_2cta = -1;
bool gemm_use_pair(int M, int bn, int K) {
    if (g_gemm_2cta < 0) {
        const char* e = getenv("WB_GEMM_2CTA");
        g_gemm_2cta = (e == nullptr) ? 1 : atoi(e);
    }
    if (g_gemm_2cta == 0 || bn != 256 || M <= BM) return false;
    return g_gemm_2cta >= 2 || K > 4 * BK;
}

}  // namespace

void gemm_set_sm_reserve(int n) { g_sm_reserve = n < 0 ? 0 : n; }

// Tile width: 256 columns wherever N allows.  (A 128-column tile for K <= 256 would leave room for an 8-slot A ring -
// two A tiles in flight - but tcgen05.mma 128x128x16 takes as long as 128x256x16 with cta_group::1 (~170 SM cycles under
// the power cap, profiles/r2_ops_v18_diag.txt): FFN1 92 vs 86 us, QKV 39 vs 31 us.  WB_GEMM_BN128=1 re-enables the
// experiment.)
static int g_bn128_small_k = -1;
int gemm_bn_for(int N, int K, int epi) {
    // the GLU epilogue packs four 32-column chunks (= 64 output columns, one 128-byte TMA box row) per warp: 256-wide tiles
    if (epi == EPI_GLU_BF16) return 256;
    if (g_bn128_small_k < 0) {
        const char* e = getenv("WB_GEMM_BN128");
        g_bn128_small_k = (e != nullptr && atoi(e) != 0) ? 1 : 0;   // measured slower (r2_ops_v18_diag.txt): off by default
    }
    if (g_bn128_small_k && K <= 256) return 128;
    return (N % 256 == 0 || N >= 1024) ? 256 : 128;
}

int make_weight_tmap(WeightMaps* out, const void* w, int N, int K, int epi) {
    const int bn = gemm_bn_for(N, K, epi);
    int rc = make_tmap_2d_bf16(&out->one, w, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)bn, BK);
    if (rc != WB_OK) return rc;
    // CTA-pair mode: each CTA of the pair fetches half of the tile's weight rows
    return make_tmap_2d_bf16(&out->pair, w, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)(bn / 2), BK);
}

struct LnFuse {
    const float* gamma1;   // EPI_RESID_LN2: the LayerNorm whose fp32 result replaces x
    const float* beta1;
    const float* gamma;
    const float* beta;
    float eps;
    void* out_bf16;
    long long ld;
};

static int gemm_impl(const void* A, long long lda, const WeightMaps* tmap_b_opt, const void* B, int M, int N,
                     int K, const float* bias, int epi, float alpha, void* out, long long ldc, int split3,
                     float2* lse_part, cudaStream_t stream, const LnFuse* ln = nullptr, bool allow_ksplit = false) {
    if (M <= 0) return WB_OK;
    WB_REQUIRE(N > 0 && K > 0, WB_ERR_BAD_ARG, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    WB_REQUIRE((K % 8) == 0 && (lda % 8) == 0, WB_ERR_BAD_ARG,
               "gemm: K (%d) and lda (%lld) must be multiples of 8 (TMA 16-byte row pitch)", K, lda);
    WB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0, WB_ERR_BAD_ARG, "gemm: A not 16B aligned");
    if (epi == EPI_GLU_BF16) {
        WB_REQUIRE((N % 256) == 0 && (ldc % 8) == 0, WB_ERR_BAD_ARG, "gemm GLU: N %% 256 and ldc %% 8 required");
    }
    const bool ln_epi = (epi == EPI_RESID_LN || epi == EPI_RESID_LN2);
    if (epi == EPI_RESID_F32 || epi == EPI_F32 || ln_epi) {
        WB_REQUIRE(!split3, WB_ERR_BAD_ARG, "gemm: split3 only for bf16 outputs");
    }
    int bn = gemm_bn_for(N, K, epi);
    // Few-row GEMMs (autoregressive decoding: M = batch x beam rows): with 256-column tiles only ceil(M/128) * N/256 CTAs
    // stream the whole weight matrix (15 CTAs for M = 320, N = 1280); 128-column tiles double the CTAs that share it.
    // The prebuilt `pair` map of a weight (128-row boxes) is exactly the B map such a tile needs.
    bool narrow = false;
    if (bn == 256 && epi != EPI_GLU_BF16 && epi != EPI_LSE && !ln_epi && !split3) {
        const int tiles256 = ceil_div(M, BM) * ceil_div(N, 256);
        const int sms = current_device_sms();
        if (M <= 4 * BM && 2 * tiles256 <= sms) {
            bn = 128;
            narrow = true;
        }
    }
    if (ln_epi) {
        WB_REQUIRE(ln != nullptr && N == 256 && bn == 256 && bias != nullptr, WB_ERR_BAD_ARG,
                   "gemm: the fused residual + LayerNorm epilogue needs N == 256 and a bias");
        WB_REQUIRE(((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(ln->gamma) |
                     reinterpret_cast<uintptr_t>(ln->beta) | reinterpret_cast<uintptr_t>(out) |
                     reinterpret_cast<uintptr_t>(ln->gamma1) | reinterpret_cast<uintptr_t>(ln->beta1)) & 15) == 0 &&
                       (epi == EPI_RESID_LN || (ln->gamma1 != nullptr && ln->beta1 != nullptr)) &&
                       (ldc % 4) == 0 && (ln->ld % 8) == 0 && (reinterpret_cast<uintptr_t>(ln->out_bf16) & 15) == 0,
                   WB_ERR_BAD_ARG, "gemm: fused residual + LayerNorm operands must be 16-byte aligned");
    }
    const bool pair = gemm_use_pair(M, bn, K);
    CUtensorMap ta, tb_local;
    int rc = make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK);
    if (rc != WB_OK) return rc;
    const CUtensorMap* tb = tmap_b_opt ? ((pair || narrow) ? &tmap_b_opt->pair : &tmap_b_opt->one) : nullptr;
    if (tb == nullptr) {
        rc = make_tmap_2d_bf16(&tb_local, B, (uint64_t)N, (uint64_t)K, (uint64_t)K, (uint32_t)(pair ? bn / 2 : bn), BK);
        if (rc != WB_OK) return rc;
        tb = &tb_local;
    }
    GemmParams p;
    p.M = M;
    p.N = N;
    p.K = K;
    p.epi = epi;
    p.alpha = alpha;
    p.bias = bias;
    p.out = out;
    p.ldc = ldc;
    p.split3 = split3;
    p.num_m_tiles = ceil_div(M, BM);
    p.num_n_tiles = ceil_div(N, bn);
    // output tensor map: 32-row x 128-byte boxes (one per epilogue warp and column group)
    CUtensorMap tc;
    memset(&tc, 0, sizeof(tc));
    const bool f32_out = (epi == EPI_RESID_F32 || epi == EPI_F32 || ln_epi);
    const int out_cols = (epi == EPI_GLU_BF16) ? N / 2 : N;
    const int eb = f32_out ? 4 : 2;
    p.use_tma_out = (epi != EPI_LSE && !split3 && (ldc * eb) % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    if (p.use_tma_out) {
        rc = make_tmap_2d(&tc, out, eb, (uint64_t)M, (uint64_t)out_cols, (uint64_t)ldc, 32, f32_out ? 32 : 64);
        if (rc != WB_OK) return rc;
    }
    p.conv = 0;
    p.cblocks = 0;
    p.tile_tab = nullptr;
    p.lse_part = lse_part;
    p.ln_g = ln ? ln->gamma : nullptr;
    p.ln_b = ln ? ln->beta : nullptr;
    p.ln_eps = ln ? ln->eps : 0.f;
    p.ln1_g = ln ? ln->gamma1 : nullptr;
    p.ln1_b = ln ? ln->beta1 : nullptr;
    CUtensorMap tc2 = tc;
    if (ln_epi) {
        WB_REQUIRE(p.use_tma_out, WB_ERR_BAD_ARG, "gemm: fused residual + LayerNorm needs a TMA-compatible x");
        rc = make_tmap_2d(&tc2, ln->out_bf16, 2, (uint64_t)M, (uint64_t)N, (uint64_t)ln->ld, 32, 64);
        if (rc != WB_OK) return rc;
    }
    const int num_kb = ceil_div(K, BK);
    // split-K for few-row residual GEMMs with a long K (decoding: FFN w_2 with K = 5120 is 30 tiles of 80 k-blocks each -
    // 30 CTAs streaming 2.6 MB apiece): the pieces accumulate in the TMA reduce-add of the fp32 residual stream
    p.ksplit = 1;
    // (opt-in, gemm_resid_splitk: the order of the fp32 reduce-adds is not fixed, and the streaming encoder promises
    //  bit-reproducible chunks)
    if (allow_ksplit && epi == EPI_RESID_F32 && p.use_tma_out && bn == 128 && num_kb > GemmCfg<128>::kResMaxKB && M <= 4 * BM) {
        const int tiles = p.num_m_tiles * p.num_n_tiles;
        int ks = current_device_sms() / (tiles > 0 ? tiles : 1);
        ks = ks > 8 ? 8 : ks;
        ks = ks > num_kb / 4 ? num_kb / 4 : ks;
        if (ks >= 2) p.ksplit = ks;
    }
    p.res_ring = (bn == 256) ? (pair ? GemmCfg<256, 2>::res_ring(num_kb) : GemmCfg<256>::res_ring(num_kb))
                             : GemmCfg<128>::res_ring(num_kb);
    // activation-heavy weight-stationary shapes (FFN w_1 + SiLU): the sixteen-epilogue-warp kernel of gemm_act16.cu
    if (bn == 256 && !pair && num_kb <= 4 && !split3 && alpha == 1.0f && (epi == EPI_BF16_SILU || epi == EPI_BF16_GELU)) {
        const int r16 = gemm_act16_try(A, lda, tb, M, N, K, bias, epi, out, ldc, g_sm_reserve, stream);
        if (r16 != 1) return r16;
    }
    if (bn == 256) {
        const bool res = num_kb <= 4;   // K <= 256: weight-stationary
        switch (epi) {
#define WB_GEMM_CASE(E)                                                                              \
    case E:                                                                                          \
        if (pair)                                                                                    \
            return res ? launch_gemm<256, true, E, 2>(ta, *tb, tc, tc2, p, stream)                   \
                       : launch_gemm<256, false, E, 2>(ta, *tb, tc, tc2, p, stream);                 \
        return res ? launch_gemm<256, true, E, 1>(ta, *tb, tc, tc2, p, stream)                       \
                   : launch_gemm<256, false, E, 1>(ta, *tb, tc, tc2, p, stream);
            WB_GEMM_CASE(EPI_BF16)
            WB_GEMM_CASE(EPI_BF16_SILU)
            WB_GEMM_CASE(EPI_BF16_RELU)
            WB_GEMM_CASE(EPI_BF16_GELU)
            WB_GEMM_CASE(EPI_RESID_F32)
            WB_GEMM_CASE(EPI_GLU_BF16)
            WB_GEMM_CASE(EPI_F32)
            WB_GEMM_CASE(EPI_LSE)
            WB_GEMM_CASE(EPI_RESID_LN)
            WB_GEMM_CASE(EPI_RESID_LN2)
#undef WB_GEMM_CASE
            default:
                WB_REQUIRE(false, WB_ERR_BAD_ARG, "gemm: unknown epilogue %d", epi);
        }
    }
    WB_REQUIRE(!ln_epi, WB_ERR_BAD_ARG, "gemm: fused residual + LayerNorm needs 256-column tiles");
    if (num_kb <= GemmCfg<128>::kResMaxKB) {
        switch (epi) {
#define WB_GEMM_CASE(E) \
    case E:             \
        return launch_gemm<128, true, E, 1>(ta, *tb, tc, tc, p, stream);
            WB_GEMM_CASE(EPI_BF16)
            WB_GEMM_CASE(EPI_BF16_SILU)
            WB_GEMM_CASE(EPI_BF16_RELU)
            WB_GEMM_CASE(EPI_BF16_GELU)
            WB_GEMM_CASE(EPI_RESID_F32)
            WB_GEMM_CASE(EPI_GLU_BF16)
            WB_GEMM_CASE(EPI_F32)
            WB_GEMM_CASE(EPI_LSE)
#undef WB_GEMM_CASE
            default:
                WB_REQUIRE(false, WB_ERR_BAD_ARG, "gemm: unknown epilogue %d", epi);
        }
    }
    return launch_gemm<128, false, -1, 1>(ta, *tb, tc, tc, p, stream);
}

int gemm_bf16(const void* A, long long lda, const WeightMaps* tmap_b_opt, const void* B, int M, int N,
              int K, const float* bias, int epi, float alpha, void* out, long long ldc, int split3,
              cudaStream_t stream) {
    WB_REQUIRE(epi != EPI_LSE && epi != EPI_RESID_LN && epi != EPI_RESID_LN2, WB_ERR_BAD_ARG, "gemm: this epilogue has its own entry point");
    return gemm_impl(A, lda, tmap_b_opt, B, M, N, K, bias, epi, alpha, out, ldc, split3, nullptr, stream);
}

// out_f32 += alpha * (A B^T + bias) with the K range cut into pieces when only a few CTAs would otherwise stream a long K
// (few-row decoding GEMMs); the pieces meet in the TMA reduce-add, so the fp32 summation order is not reproducible
int gemm_resid_splitk(const void* A, long long lda, const WeightMaps* tmap_b_opt, const void* B, int M, int N, int K,
                      const float* bias, float alpha, float* out, long long ldc, cudaStream_t stream) {
    return gemm_impl(A, lda, tmap_b_opt, B, M, N, K, bias, EPI_RESID_F32, alpha, out, ldc, 0, nullptr, stream, nullptr, true);
}

static int g_fuse_ln = -1;
bool gemm_resid_ln_supported(int N) {
    if (g_fuse_ln < 0) {
        const char* e = getenv("WB_FUSE_LN");
        g_fuse_ln = (e == nullptr) ? 1 : atoi(e);
    }
    return g_fuse_ln != 0 && N == 256;
}

int gemm_resid_ln(const void* A, long long lda, const WeightMaps* tmap_b_opt, const void* B, int M, int N, int K,
                  const float* bias, float alpha, float* x, long long ldx, const float* gamma1, const float* beta1,
                  const float* gamma, const float* beta, float eps, void* ln_out_bf16, long long ld_ln,
                  cudaStream_t stream) {
    WB_REQUIRE(N == 256, WB_ERR_UNSUPPORTED, "gemm_resid_ln: N = %d (only 256)", N);
    LnFuse ln;
    ln.gamma1 = gamma1;
    ln.beta1 = beta1;
    ln.gamma = gamma;
    ln.beta = beta;
    ln.eps = eps;
    ln.out_bf16 = ln_out_bf16;
    ln.ld = ld_ln;
    return gemm_impl(A, lda, tmap_b_opt, B, M, N, K, bias, gamma1 ? EPI_RESID_LN2 : EPI_RESID_LN, alpha, x, ldx, 0, nullptr,
                     stream, &ln);
}

int lse_parts(int N, int K) { return 2 * ceil_div(N, gemm_bn_for(N, K, EPI_LSE)); }

int gemm_lse_partials(const void* A, long long lda, const WeightMaps* tmap_b_opt, const void* B, int M, int N, int K,
                      const float* bias, float2* part, cudaStream_t stream) {
    WB_REQUIRE(part != nullptr, WB_ERR_BAD_ARG, "gemm_lse_partials: null output");
    return gemm_impl(A, lda, tmap_b_opt, B, M, N, K, bias, EPI_LSE, 1.0f, part /*unused as matrix*/, 0, 0, part, stream);
}

// Conv2d(d -> d, 3x3, stride 2) + bias + ReLU over the channels-last conv1 output, as an implicit GEMM:
//   out2[(tile row), n] = relu(bias[n] + sum_{kh,kw,c} out1[t1 = 2 t2 + kh][f1 = 2 f2 + kw][c] * W[n][(kh,kw,c)])
// out1: [T1_total][F1][d] bf16 (utterances stacked along t), out2: [rows_out][d] bf16, tile_tab_dev: one int4 per
// 114-row tile (see GemmParams).
int gemm_conv2_implicit(const void* out1, long long t1_total, int F1, int d, const WeightMaps* tmap_w, const float* bias,
                        const void* tile_tab_dev, int num_tiles, long long rows_out, void* out2, cudaStream_t stream) {
    if (num_tiles <= 0) return WB_OK;
    WB_REQUIRE(d % 256 == 0 && F1 == 39, WB_ERR_UNSUPPORTED, "conv2 implicit GEMM: d=%d F1=%d unsupported", d, F1);
    CUtensorMap ta, tc, tc18;
    const uint64_t dims[3] = {(uint64_t)d, (uint64_t)F1, (uint64_t)t1_total};
    const uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)F1 * d * 2};
    const uint32_t box[3] = {64, 37, 11};     // traversal extents: ceil(37/2) = 19 taps in f, ceil(11/2) = 6 in t
    const uint32_t estr[3] = {1, 2, 2};
    int rc;
    if ((rc = make_tmap_3d_bf16(&ta, out1, dims, strides, box, estr)) != WB_OK) return rc;
    if ((rc = make_tmap_2d(&tc, out2, 2, (uint64_t)rows_out, (uint64_t)d, (uint64_t)d, 32, 64)) != WB_OK) return rc;
    if ((rc = make_tmap_2d(&tc18, out2, 2, (uint64_t)rows_out, (uint64_t)d, (uint64_t)d, 18, 64)) != WB_OK) return rc;
    GemmParams p;
    p.M = (int)rows_out;
    p.N = d;
    p.K = 9 * d;
    p.epi = EPI_BF16_RELU;
    p.alpha = 1.0f;
    p.bias = bias;
    p.out = out2;
    p.ldc = d;
    p.split3 = 0;
    p.use_tma_out = 1;
    p.res_ring = 0;
    p.ksplit = 1;
    p.num_m_tiles = num_tiles;
    p.num_n_tiles = d / 256;
    p.lse_part = nullptr;
    p.conv = 1;
    p.cblocks = d / 64;
    p.tile_tab = reinterpret_cast<const int4*>(tile_tab_dev);
    if (gemm_use_pair(2 * BM, 256, 9 * d) && num_tiles > 1)
        return launch_gemm<256, false, EPI_BF16_RELU, 2>(ta, tmap_w->pair, tc, tc18, p, stream);
    return launch_gemm<256, false, EPI_BF16_RELU, 1>(ta, tmap_w->one, tc, tc18, p, stream);
}

int gemm_diag(unsigned long long* out8, int reset) {
#ifdef WB_GEMM_DIAG
    if (out8) WB_CHECK_CUDA(cudaMemcpyFromSymbol(out8, g_gemm_diag, sizeof(unsigned long long) * 12));
    if (reset) {
        unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        WB_CHECK_CUDA(cudaMemcpyToSymbol(g_gemm_diag, z, sizeof(z)));
    }
    return WB_OK;
#else
    (void)out8;
    (void)reset;
    set_last_error("gemm_diag: the stall accounting is compiled out (build with NVCC_EXTRA=-DWB_GEMM_DIAG)");
    return WB_ERR_UNSUPPORTED;
#endif
}

}  // namespace wb
