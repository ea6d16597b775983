// Host-side runtime glue: last-error string, launch counter, TMA descriptor encoding.
#include <stdlib.h>
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>
#include <utility>

namespace wb {

std::atomic<unsigned long long> g_launch_count{0};
thread_local int g_pdl_depth = 0;
bool pdl_stream_allowed() {
    static const bool on = [] {
        const char* e = getenv("WB_PDL_STREAM");
        return e == nullptr || atoi(e) != 0;
    }();
    return on;
}
bool pdl_allowed() {
    static const bool on = [] {
        const char* e = getenv("WB_PDL");
        return e == nullptr || atoi(e) != 0;
    }();
    return on;
}

static thread_local char g_err[1024] = {0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_last_error() { return g_err; }

// ---------------------------------------------------------------- profiler
int current_device_sms() {
    static std::atomic<int> cache[128] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    std::atomic<int>& slot = cache[dev & 127];
    int n = slot.load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
        slot.store(n, std::memory_order_relaxed);
    }
    return n;
}

int g_prof_on = 0;
namespace {
struct ProfRec {
    int tag;
    double work;
    cudaEvent_t e0, e1;
};
std::vector<ProfRec> g_recs;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_free_events;
std::mutex g_prof_mu;
// prof_begin/prof_end pairs are issued by one host thread per stream; the record index travels in TLS
thread_local long g_cur_rec = -1;
}  // namespace

void prof_begin(int tag, cudaStream_t st, double work) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    r.tag = tag;
    r.work = work;
    if (!g_free_events.empty()) {
        r.e0 = g_free_events.back().first;
        r.e1 = g_free_events.back().second;
        g_free_events.pop_back();
    } else {
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
    }
    cudaEventRecord(r.e0, st);
    g_recs.push_back(r);
    g_cur_rec = (long)g_recs.size() - 1;
}
void prof_end(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_cur_rec >= 0 && g_cur_rec < (long)g_recs.size()) cudaEventRecord(g_recs[g_cur_rec].e1, st);
    g_cur_rec = -1;
}
void prof_reset() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_recs) g_free_events.push_back({r.e0, r.e1});
    g_recs.clear();
}
int prof_collect(double* ms, double* work, long long* launches) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int t = 0; t < PT_COUNT; ++t) {
        ms[t] = 0;
        work[t] = 0;
        launches[t] = 0;
    }
    for (auto& r : g_recs) {
        float e = 0.f;
        if (cudaEventElapsedTime(&e, r.e0, r.e1) == cudaSuccess) {
            ms[r.tag] += e;
            work[r.tag] += r.work;
            launches[r.tag] += 1;
        }
    }
    return 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (PFN_encodeTiled)p;
    });
    return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                 uint32_t box_rows, uint32_t box_cols) {
    PFN_encodeTiled fn = get_encode_fn();
    WB_REQUIRE(fn != nullptr, WB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no driver?)");
    WB_REQUIRE(elem_bytes == 2 || elem_bytes == 4, WB_ERR_BAD_ARG, "tmap: element size %d", elem_bytes);
    WB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, WB_ERR_BAD_ARG, "tmap: base not 16B aligned");
    WB_REQUIRE((ld_elems * elem_bytes) % 16 == 0, WB_ERR_BAD_ARG, "tmap: row pitch %llu B not a multiple of 16",
               (unsigned long long)(ld_elems * elem_bytes));
    WB_REQUIRE(box_cols * elem_bytes == 128 && box_rows <= 256, WB_ERR_BAD_ARG, "tmap: bad box %u x %u", box_rows,
               box_cols);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * (uint64_t)elem_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WB_REQUIRE(r == CUDA_SUCCESS, WB_ERR_CUDA,
               "cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r, base,
               (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols);
    return WB_OK;
}

// 3-D bf16 tensor [d2][d1][d0] (d0 innermost, contiguous) with element (traversal) strides: used for the
// TMA-side im2col of Conv2d(3x3, stride 2): box {64 channels, 19 frequency taps (stride 2), 6 time taps (stride 2)}
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                      const uint32_t box[3], const uint32_t estr[3]) {
    PFN_encodeTiled fn = get_encode_fn();
    WB_REQUIRE(fn != nullptr, WB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no driver?)");
    WB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && strides_bytes[0] % 16 == 0 && strides_bytes[1] % 16 == 0,
               WB_ERR_BAD_ARG, "tmap3d: alignment");
    cuuint64_t gdim[3] = {dims[0], dims[1], dims[2]};
    cuuint64_t gstride[2] = {strides_bytes[0], strides_bytes[1]};
    cuuint32_t b[3] = {box[0], box[1], box[2]};
    cuuint32_t e[3] = {estr[0], estr[1], estr[2]};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, b, e,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WB_REQUIRE(r == CUDA_SUCCESS, WB_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d): dims %llu %llu %llu box %u %u %u",
               (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0], box[1],
               box[2]);
    return WB_OK;
}

// bf16 [rows][cols] with 64-byte boxes (32 columns) and SWIZZLE_64B: the 2 KB per-warp output staging of gemm_act16.cu
int make_tmap_2d_bf16_sw64(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                           uint32_t box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    WB_REQUIRE(fn != nullptr, WB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no driver?)");
    WB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld_elems * 2) % 16 == 0 && box_rows <= 256, WB_ERR_BAD_ARG,
               "tmap sw64: alignment");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * 2};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WB_REQUIRE(r == CUDA_SUCCESS, WB_ERR_CUDA, "cuTensorMapEncodeTiled(sw64) failed (%d)", (int)r);
    return WB_OK;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
    return make_tmap_2d(out, base, 2, rows, cols, ld_elems, box_rows, box_cols);
}

}  // namespace wb
