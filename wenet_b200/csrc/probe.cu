// Hardware probes used by tests/tools (not on the product path): behaviour of TMA tiled loads with
// element (traversal) strides, needed by the TMA-side im2col of Conv2dSubsampling4's second conv.
#include "common.cuh"
#include "kernels.h"

namespace wb {
namespace {
__global__ void probe_tma3d_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, uint32_t expect_bytes,
                                   uint32_t copy_bytes, uint8_t* out, int* status) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xDEADBEEFu;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, expect_bytes);
        tma_load_3d(smem, &tm, bar, c0, c1, c2);
        int ok = 0;
        for (int it = 0; it < 2000000 && !ok; ++it) ok = mbar_try_wait(bar, 0) ? 1 : 0;   // bounded: never hangs
        status[0] = ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (int)copy_bytes; i += blockDim.x) out[i] = smem[i];
}
}  // namespace

// loads one box at (c0, c1, c2) and copies the first copy_bytes of shared memory back; status = 1 if exactly
// expect_bytes arrived (the mbarrier completed), 0 if it timed out
int probe_tma3d(const void* base, const uint64_t dims[3], const uint64_t strides_bytes[2], const uint32_t box[3],
                const uint32_t estr[3], int c0, int c1, int c2, uint32_t expect_bytes, uint32_t copy_bytes, uint8_t* out_dev,
                int* status_dev, cudaStream_t stream) {
    CUtensorMap tm;
    int rc = make_tmap_3d_bf16(&tm, base, dims, strides_bytes, box, estr);
    if (rc != WB_OK) return rc;
    WB_CHECK_CUDA(cudaFuncSetAttribute(probe_tma3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024 + 1024));
    probe_tma3d_kernel<<<1, 128, 34 * 1024 + 1024, stream>>>(tm, c0, c1, c2, expect_bytes, copy_bytes, out_dev, status_dev);
    count_launch();
    WB_CHECK_LAUNCH();
    return WB_OK;
}
}  // namespace wb

extern "C" int wb_probe_tma3d(const void* base, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                              const uint32_t* estr, int c0, int c1, int c2, uint32_t expect_bytes, uint32_t copy_bytes,
                              uint8_t* out_dev, int* status_dev, void* stream) {
    return wb::probe_tma3d(base, dims, strides_bytes, box, estr, c0, c1, c2, expect_bytes, copy_bytes, out_dev, status_dev,
                           (cudaStream_t)stream);
}
