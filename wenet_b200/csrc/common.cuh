// wenet_b200 — shared device/host helpers for the sm_100a kernels.
// Inline-PTX wrappers for mbarrier / TMA (cp.async.bulk.tensor) / tcgen05 (UMMA + TMEM).
// No CUTLASS/CuTe dependency: descriptors are built by hand (bit layouts follow the PTX ISA
// "tcgen05 matrix descriptor" / "instruction descriptor" tables).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string.h>
#include "../../include/wenet_b200.h"  // wb_status codes

namespace wb {

// ----------------------------------------------------------------------------------------------
// error plumbing (C-ABI never throws; see include/wenet_b200.h)
// ----------------------------------------------------------------------------------------------

void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define WB_CHECK_CUDA(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            wb::set_last_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__,     \
                               (int)_e, cudaGetErrorString(_e), #expr);                  \
            return WB_ERR_CUDA;                                                      \
        }                                                                                \
    } while (0)

#define WB_CHECK_LAUNCH() WB_CHECK_CUDA(cudaGetLastError())

#define WB_REQUIRE(cond, code, ...)                                                      \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            wb::set_last_error(__VA_ARGS__);                                             \
            return (code);                                                               \
        }                                                                                \
    } while (0)

// launch counter: every kernel launch made by this library bumps it (bench.py reports it)
extern std::atomic<unsigned long long> g_launch_count;
// ---- programmatic dependent launch (PDL) ----
// The latency-bound chains (autoregressive decoding: ~450 dependent launches per step; streaming chunks) spend a good part of
// every GEMM on its launch latency, prologue (barriers, TMEM allocation, descriptor prefetch) and the first weight fetch.
// Inside a PdlScope the GEMM launches carry cudaLaunchAttributeProgrammaticStreamSerialization: the kernel may start while
// its predecessor is still running, does everything that does not depend on it (prologue + the first weight tiles, which
// nothing ever writes), and only then executes griddepcontrol.wait (= predecessor complete and visible).  Predecessors
// call griddepcontrol.launch_dependents at their top.  Kernels launched without the attribute are ordinary stream
// successors, whatever their predecessor did.  WB_PDL=0 turns the attribute off.
extern thread_local int g_pdl_depth;
bool pdl_allowed();
struct PdlScope {
    bool on;
    explicit PdlScope(bool enable = true) : on(enable) { if (on) ++g_pdl_depth; }
    ~PdlScope() { if (on) --g_pdl_depth; }
};
bool pdl_stream_allowed();   // WB_PDL_STREAM=0: the streaming chunk paths launch the ordinary way
static inline bool pdl_active() { return g_pdl_depth > 0 && pdl_allowed(); }
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Launch of a kernel that executes pdl_wait() before it touches anything an earlier kernel of the stream wrote (or still
// reads): inside a PdlScope with the programmatic-serialization attribute, otherwise an ordinary launch.  NOTE for such
// kernels: ahead of pdl_wait() not even the output of the last-but-one kernel is safe (the direct predecessor may itself
// still be waiting for it) - only data that was complete before the last ordinary launch, i.e. weights and tables.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_maybe_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                           Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_active() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

static inline void count_launch(int n = 1) { g_launch_count.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// Optional per-kernel-family profiler (CUDA events on the launching stream around every launch).
// Off by default; bench.py switches it on to obtain live per-kernel durations and roofline numbers.
enum ProfTag : int {
    PT_GEMM = 0, PT_ATTENTION, PT_LAYERNORM, PT_DWCONV, PT_CONV1, PT_IM2COL, PT_KPREP, PT_FBANK, PT_LOGSOFTMAX_TOPK,
    PT_GREEDY, PT_PREFIX_BEAM, PT_EMBED, PT_GATHER_LOGPROB, PT_RESCORE, PT_MISC, PT_GEMM_LN, PT_COUNT
};
extern int g_prof_on;
void prof_begin(int tag, cudaStream_t st, double work);
void prof_end(cudaStream_t st);
void prof_reset();
int prof_collect(double* ms, double* work, long long* launches);
struct ProfScope {
    cudaStream_t st;
    bool on;
    ProfScope(int tag, cudaStream_t s, double work) : st(s), on(g_prof_on != 0) {
        if (on) prof_begin(tag, st, work);
    }
    ~ProfScope() {
        if (on) prof_end(st);
    }
};

// Opt a kernel into > 48 KB of dynamic shared memory.  Function attributes are PER DEVICE: the "done" bits are kept
// per device ordinal (a process may drive several GPUs, e.g. one B200ASRModel per device) and updated atomically
// (entry points may be called from several host threads).
#define WB_SET_MAX_DYN_SMEM(kernel, bytes)                                                                        \
    do {                                                                                                          \
        static std::atomic<unsigned long long> _done[2] = {};                                                     \
        int _dev = 0;                                                                                             \
        WB_CHECK_CUDA(cudaGetDevice(&_dev));                                                                      \
        const unsigned long long _bit = 1ull << (_dev & 63);                                                      \
        std::atomic<unsigned long long>& _w = _done[(_dev >> 6) & 1];                                             \
        if (!(_w.load(std::memory_order_acquire) & _bit)) {                                                       \
            WB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            _w.fetch_or(_bit, std::memory_order_release);                                                         \
        }                                                                                                         \
    } while (0)
// number of SMs of the CURRENT device (cached per device ordinal)
int current_device_sms();

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// single-instruction MUFU approximations (ex2.approx.ftz / rcp.approx.ftz, ~2 ulp — far below the bf16 rounding
// of everything they feed).  NB: __expf / exp2f / __frcp_rn expand to 5-70 SASS instructions each (denormal
// range fix-ups, IEEE-rounded reciprocal), which made the SiLU / GLU GEMM epilogues ALU-bound.
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_f(float x) { return fast_rcp(1.0f + fast_exp2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
// Four sigmoids for five MUFU ops instead of eight: with y_i = 1 + 2^(-x_i log2 e), one reciprocal of the product
// y0 y1 y2 y3 gives every 1/y_i by multiplications (1/y0 = r y2 y3 y1 ...).  y_i is clamped to 2^30 + 1 so the
// product stays below 2^127; the clamp changes sigmoid(x) only for x < -20.8, by less than 1e-9 absolute.  The
// SiLU / GLU GEMM epilogues are bound by the 16-lane MUFU pipe (ncu stall_mio), not by the FMA pipe that takes
// the extra multiplications.  Relative error ~3 ulp (one rcp.approx + three roundings).
__device__ __forceinline__ void sigmoid4(const float x0, const float x1, const float x2, const float x3, float& s0,
                                         float& s1, float& s2, float& s3) {
    constexpr float kNegLog2e = -1.4426950408889634f;
    constexpr float kCap = 1073741824.0f;   // 2^30
    const float y0 = 1.0f + fminf(fast_exp2(kNegLog2e * x0), kCap);
    const float y1 = 1.0f + fminf(fast_exp2(kNegLog2e * x1), kCap);
    const float y2 = 1.0f + fminf(fast_exp2(kNegLog2e * x2), kCap);
    const float y3 = 1.0f + fminf(fast_exp2(kNegLog2e * x3), kCap);
    const float p01 = y0 * y1, p23 = y2 * y3;
    const float r = fast_rcp(p01 * p23);
    const float r01 = r * p23, r23 = r * p01;   // 1 / (y0 y1), 1 / (y2 y3)
    s0 = r01 * y1;
    s1 = r01 * y0;
    s2 = r23 * y3;
    s3 = r23 * y2;
}
// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 retire two fp32 operations per issue slot) ----------------
// The activation epilogues are bound by instruction issue and by the 16-lane MUFU pipe, not by the FMA pipe: every pair
// of multiplies / adds that shares one issue slot shortens them.
__device__ __forceinline__ float2 f2_mul(const float2 a, const float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2_add(const float2 a, const float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2_fma(const float2 a, const float2 b, const float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
// Four sigmoids, packed: same algebra as sigmoid4 (one reciprocal of y0 y1 y2 y3, y_i = 1 + min(2^(-x_i log2 e), 2^30)),
// arranged so that every product except two is a two-wide instruction:
//   (pa, pb) = (y0 y2, y1 y3);  r = 1 / (pa pb);  (qa, qb) = (r pb, r pa);
//   (1/y0, 1/y1) = (qa y2, qb y3);  (1/y2, 1/y3) = (qa y0, qb y1).
// 5 MUFU + 4 FMNMX + 9 FMA-pipe issue slots per four values (scalar form: 5 + 4 + 17).
__device__ __forceinline__ void sigmoid4_f2(const float2 xa, const float2 xb, float2& sa, float2& sb) {
    constexpr float kNegLog2e = -1.4426950408889634f;
    constexpr float kCap = 1073741824.0f;   // 2^30
    const float2 ta = f2_mul(xa, make_float2(kNegLog2e, kNegLog2e));
    const float2 tb = f2_mul(xb, make_float2(kNegLog2e, kNegLog2e));
    const float2 ea = make_float2(fminf(fast_exp2(ta.x), kCap), fminf(fast_exp2(ta.y), kCap));
    const float2 eb = make_float2(fminf(fast_exp2(tb.x), kCap), fminf(fast_exp2(tb.y), kCap));
    const float2 ya = f2_add(ea, make_float2(1.0f, 1.0f));
    const float2 yb = f2_add(eb, make_float2(1.0f, 1.0f));
    const float2 p = f2_mul(ya, yb);
    const float r = fast_rcp(p.x * p.y);
    const float2 q = make_float2(r * p.y, r * p.x);
    sa = f2_mul(q, yb);
    sb = f2_mul(q, ya);
}
// in-place SiLU over a register array whose length is a multiple of 4
template <int N>
__device__ __forceinline__ void silu_inplace(float (&v)[N]) {
    static_assert(N % 4 == 0, "silu_inplace: N % 4");
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        const float2 xa = make_float2(v[i], v[i + 1]), xb = make_float2(v[i + 2], v[i + 3]);
        float2 sa, sb;
        sigmoid4_f2(xa, xb, sa, sb);
        const float2 oa = f2_mul(xa, sa), ob = f2_mul(xb, sb);
        v[i] = oa.x;
        v[i + 1] = oa.y;
        v[i + 2] = ob.x;
        v[i + 3] = ob.y;
    }
}
// One-MUFU variants: sigmoid(x) = 0.5 + 0.5 tanh(x/2) with tanh.approx.f32 (max rel. error 2^-11, i.e. <= 2.5e-4
// absolute on the sigmoid).  The SiLU / GLU GEMM epilogues are MUFU-throughput bound (16 ops/clk/SM, ncu
// stall_mio) and this form makes FFN1 21 % faster (110 -> 87 us), but the systematic 2.5e-4 error flipped a
// near-tie CTC frame against the reference goldens, so the model path keeps the accurate two-MUFU forms above;
// these are kept for experiments only.
__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_t(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ float silu_t(float x) {
    const float h = 0.5f * x;
    return fmaf(h, fast_tanh(h), h);
}

// exact (erf) GELU of torch.nn.GELU(): 0.5 x (1 + erf(x / sqrt 2)); erff is libdevice's polynomial form (FMA pipe)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- TMA -------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// ---- CTA pair (cta_group::2) variants -----------------------------------------------------------
// Two CTAs of a cluster (ranks 0 / 1, same TPC) execute ONE tcgen05.mma of M = 256: each contributes its own 128 rows
// of A and HALF of the B tile from its own shared memory and receives its 128 accumulator lanes in its own TMEM.  The
// leader (rank 0) issues the MMAs; TMA loads of both CTAs complete on the LEADER's mbarrier; tcgen05.commit multicasts
// its arrival to the same barrier offset in both CTAs.
// named barrier over `nthreads` threads of the CTA (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared-memory object of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion is signalled on an mbarrier given by its shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                 int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// TMA stores (shared -> global), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
// element-wise  global += shared  performed by the TMA unit / L2 (no read-modify-write in the SM)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// 1-D bulk copy global -> shared (no tensor map), completes on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- tcgen05 / TMEM ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_holder)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers fp16/bf16 operands, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_holder, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA of a CTA pair (issued by one thread of the leader CTA)
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
            smem_u32(bar))
        : "memory");
}
// TMEM -> registers: each thread of the warp reads 32 consecutive fp32 columns of "its" lane
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// wait for outstanding tcgen05.ld and tie the destination registers to the wait, so uses of `r` cannot be scheduled
// above it when the load was issued earlier (software-pipelined epilogues)
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// registers -> TMEM (32 lanes x 32 columns, thread t writes lane t), e.g. rescaling an accumulator in place
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :
        : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, SWIZZLE_128B canonical layouts (PTX ISA "matrix descriptor"):
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset>>4 bits [46,48) version = 1 (Blackwell)
//   bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major operand tile [rows][64 bf16] (128 B per row, TMA SWIZZLE_128B): 8-row groups are
// 1024 B apart (SBO = 1024), LBO is unused for swizzled K-major layouts (set to 1).
// MN-major operand tile [k rows][64 bf16 along MN]: one 64-wide MN atom, 8-k groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
//   [4,6) D format (1 = F32)  [7,10) A format (1 = BF16)  [10,13) B format (1 = BF16)
//   [15] A major (0 = K)      [16] B major (0 = K, 1 = MN)
//   [17,23) N >> 3            [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int b_mn_major = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#endif  // __CUDACC__

// ----------------------------------------------------------------------------------------------
// host: TMA descriptor creation (driver entry point fetched at run time; no libcuda link)
// ----------------------------------------------------------------------------------------------
// 2-D row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box = box_rows x
// box_cols (box_cols*2 bytes must be 128 for SWIZZLE_128B).
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                      uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols);
int make_tmap_2d_bf16_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                           uint32_t box_rows);
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                      const uint32_t box[3], const uint32_t estr[3]);
// same for fp32 (elem_bytes = 4, box_cols = 32) or bf16 (elem_bytes = 2, box_cols = 64): 128-byte inner box
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t ld_elems, uint32_t box_rows, uint32_t box_cols);

}  // namespace wb
