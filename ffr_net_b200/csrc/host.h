// Host-side helpers shared by the C-ABI entry points: error reporting and TMA tensor-map encoding.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

namespace ffr {

// Thread-local message behind ffr_last_error().
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define FFR_CHECK_ARG(cond, ...)                            \
    do {                                                    \
        if (!(cond)) return ::ffr::set_error(-1, __VA_ARGS__); \
    } while (0)

#define FFR_CUDA(expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return ::ffr::set_error(static_cast<int>(_e), "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

// Called after every kernel launch: checks the launch and counts it (ffr_launch_count()).
int launch_status(const char* what);
long long launch_count();

// 2-D bf16 row-major matrix [rows, cols] with row pitch `ld` elements; box = [box_rows, 64 cols]
// (64 bf16 = 128 B = one SWIZZLE_128B row). Out-of-bounds elements read as zero.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols = 64);
int make_tmap_h9_pixel_bf16(CUtensorMap* out, const void* base, uint64_t n_img, uint64_t cols, uint64_t ld_elems,
                            uint32_t box_imgs);
int make_tmap_pixel_bf16(CUtensorMap* out, const void* base, uint64_t n_img, uint32_t G, uint64_t cols,
                         uint64_t ld_elems, uint32_t box_imgs);

int num_sms();

// Programmatic dependent launch (ptx.cuh: pdl_wait / pdl_launch_dependents) for kernels that call pdl_sync(): on by
// default, ffr_debug_set_pdl(0) turns the launch attribute off (plain stream order; A/B runs and tests).
enum PdlKind { PDL_GEMM = 1, PDL_SIMT = 2 };     // bit mask: tcgen05 GEMM kernels / the memory-bound SIMT kernels
bool pdl_enabled(int kind);
void set_pdl_mask(int mask);

// Kernel launch with optional cluster size and the programmatic-stream-serialization attribute. ONLY for kernels that
// execute pdl_wait() before their first access to global memory another kernel may write (or still read).
template <typename... P, typename... A>
inline cudaError_t launch_ex(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                             int pdl_kind, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster_x;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled(pdl_kind)) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

}  // namespace ffr
