// Host-side runtime glue: last-error string, launch counter, TMA descriptor encoding.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <mutex>

namespace wb {

unsigned long long g_launch_count = 0;

static thread_local char g_err[1024] = {0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_last_error() { return g_err; }

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

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
    PFN_encodeTiled fn = get_encode_fn();
    WB_REQUIRE(fn != nullptr, WB_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no driver?)");
    WB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, WB_ERR_BAD_ARG, "tmap: base not 16B aligned");
    WB_REQUIRE((ld_elems * 2) % 16 == 0, WB_ERR_BAD_ARG, "tmap: row pitch %llu B not a multiple of 16",
               (unsigned long long)(ld_elems * 2));
    WB_REQUIRE(box_cols * 2 == 128 && box_rows <= 256, WB_ERR_BAD_ARG, "tmap: bad box %u x %u", box_rows,
               box_cols);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    WB_REQUIRE(r == CUDA_SUCCESS, WB_ERR_CUDA,
               "cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r, base,
               (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows, box_cols);
    return WB_OK;
}

}  // namespace wb
