#include "host.h"

#include <atomic>
#include <cstring>
#include <mutex>

namespace ffr {

char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

// cuTensorMapEncodeTiled is a driver API; fetch it through the runtime so the library has no
// hard link against libcuda (it must load, and export its symbols, on a box without a driver).
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
    });
    return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
    encode_tiled_fn enc = get_encode();
    if (!enc) return set_error(-2, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(-1, "TMA base pointer not 16-B aligned");
    if ((ld_elems * 2) % 16 != 0) return set_error(-1, "TMA row pitch %llu B not a multiple of 16",
                                                   (unsigned long long)(ld_elems * 2));
    if (box_rows > 256 || box_cols * 2 > 128) return set_error(-1, "TMA box %ux%u too large", box_rows, box_cols);
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(-3, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
                         (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_rows,
                         box_cols);
    return 0;
}

// H9 activation matrix [n_img * 81 rows][ld] viewed as (channel, w, h, image): a box of `box_imgs` images at ONE pixel
// lands in shared memory as box_imgs rows of 64 channels, i.e. exactly like a 2-D box of consecutive rows.
int make_tmap_h9_pixel_bf16(CUtensorMap* out, const void* base, uint64_t n_img, uint64_t cols, uint64_t ld_elems,
                            uint32_t box_imgs) {
    return make_tmap_pixel_bf16(out, base, n_img, 9, cols, ld_elems, box_imgs);
}

// General form: a flat map with G x G rows per image (H9: G = 9; halo-shared backbone maps: G = S + 1). Coordinates
// outside [0, G) are filled with zeros by TMA, which is exactly the zero padding of a tap that leaves the image.
int make_tmap_pixel_bf16(CUtensorMap* out, const void* base, uint64_t n_img, uint32_t G, uint64_t cols,
                         uint64_t ld_elems, uint32_t box_imgs) {
    encode_tiled_fn enc = get_encode();
    if (!enc) return set_error(-2, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(-1, "TMA base pointer not 16-B aligned");
    if ((ld_elems * 2) % 16 != 0) return set_error(-1, "TMA row pitch %llu B not a multiple of 16",
                                                   (unsigned long long)(ld_elems * 2));
    if (box_imgs > 256) return set_error(-1, "TMA box of %u images too large", box_imgs);
    cuuint64_t gdim[4] = {cols, G, G, n_img};
    cuuint64_t gstr[3] = {ld_elems * 2, (uint64_t)G * ld_elems * 2, (uint64_t)G * G * ld_elems * 2};
    cuuint32_t box[4] = {64, 1, 1, box_imgs};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(-3, "cuTensorMapEncodeTiled (H9 pixel view) failed (%d) n=%llu cols=%llu ld=%llu box=%u", (int)r,
                         (unsigned long long)n_img, (unsigned long long)cols, (unsigned long long)ld_elems, box_imgs);
    return 0;
}

static std::atomic<long long> g_launches{0};

int launch_status(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(static_cast<int>(e), "%s launch failed: %s", what, cudaGetErrorString(e));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// Measured on the bs512 eval step (tools/ab_bench.py, profiles/r02_ab_pdl_pair_512.json; medians of 10 interleaved runs):
// no PDL 11.09 ms, GEMM kernels only 10.95 ms, SIMT kernels only 11.37 ms, both 11.30 ms. Early-scheduled CTAs of the
// memory-bound kernels sit next to the running GEMM CTAs and cost more than their launch latency saves.
constexpr int kPdlDefault = PDL_GEMM;
static std::atomic<int> g_pdl{kPdlDefault};
bool pdl_enabled(int kind) { return (g_pdl.load(std::memory_order_relaxed) & kind) != 0; }
void set_pdl_mask(int mask) { g_pdl.store(mask < 0 ? kPdlDefault : mask, std::memory_order_relaxed); }

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace ffr
