// Hardware-semantics probe (debug entry point, not on the product path):
// does a SWIZZLE_128B K-major UMMA descriptor whose start address is offset by a whole number of 128-byte rows
// (not a multiple of the 1024-byte swizzle atom) read rows [r0, r0+128) of a TMA-written tile?
// variant 0: base_offset field = 0; variant 1: base_offset = (start_address >> 7) & 7.
// The answer decides how the sliding-window convolution kernel addresses its taps.
#include "../../include/ffr_sm100_probe.h"
#include "host.h"
#include "ptx.cuh"

namespace ffr {

__global__ void __launch_bounds__(128, 1)
rowshift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out,
                      int row_off, int variant) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 256 rows x 128 B
    uint8_t* sB = smem + 256 * 128;     // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 256 * 128 + 64 * 128);
    uint64_t* mma_bar = bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(mma_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<64>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 256 * 128 + 64 * 128);
        tma_load_2d(sA, &tmA, bar, 0, 0);
        tma_load_2d(sB, &tmB, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA) + row_off * 128;
        const uint32_t b_addr = smem_u32(sB);
        const uint32_t bo = variant ? ((a_addr >> 7) & 7u) : 0u;
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, umma_smem_desc_sw128(a_addr + k * 32, bo), umma_smem_desc_sw128(b_addr + k * 32),
                      idesc, k > 0);
        umma_commit(mma_bar);
    }
    __syncwarp();
    mbar_wait(mma_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + c0 + (static_cast<uint32_t>(warp * 32) << 16), v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<64>(tmem_base);
    }
}

}  // namespace ffr

extern "C" FFR_API int ffr_debug_rowshift_probe(const void* a, const void* w, float* out, int row_off, int variant,
                                                ffr_stream_t stream) {
    using namespace ffr;
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d_bf16(&tmA, a, 256, 64, 64, 256);
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tmB, w, 64, 64, 64, 64);
    if (rc) return rc;
    const int smem = 1024 + 256 * 128 + 64 * 128 + 64;
    FFR_CUDA(cudaFuncSetAttribute(rowshift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    rowshift_probe_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tmA, tmB, out, row_off, variant);
    return launch_status("rowshift_probe_kernel");
}

// ----------------------------------------------------------------------------------------------------------
// Micro-benchmark (debug): issue rate of tcgen05.mma.kind::f16 (SS mode, K = 16 per instruction) as a function of
// the instruction shape and of how many independent accumulators the instruction stream alternates between.
// Operands are whatever is in shared memory (timing only). out[0] = cycles for `iters` MMAs on one CTA.
// ----------------------------------------------------------------------------------------------------------
namespace ffr {
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(long long* out, int M, int N, int n_acc, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1) {   // warp-uniform issue loop, one elected lane issues (see ptx.cuh: elect_one_sync)
        const uint32_t idesc = umma_idesc_bf16(M, N);
        const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(smem));
        const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem + 128 * 128));
        if (elect_one_sync()) {
            for (int i = 0; i < 8; ++i) umma_bf16(tmem_base, a_desc, b_desc, idesc, 1);
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        const long long t0 = clock64();
        const uint32_t mask = n_acc - 1;
        for (int i = 0; i < iters; i += 4) {
            const uint32_t d0 = tmem_base + ((i >> 2) & mask) * N;
            if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(d0, a_desc + 2 * k, b_desc + 2 * k, idesc, 1);
            }
            __syncwarp();
        }
        if (elect_one_sync()) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 1);
        const long long t1 = clock64();
        if (threadIdx.x == 32) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base); }
}
}  // namespace ffr

extern "C" FFR_API int ffr_debug_mma_bench(long long* out_cycles, int M, int N, int n_acc, int iters, int grid,
                                           ffr_stream_t stream) {
    using namespace ffr;
    FFR_CHECK_ARG((M == 64 || M == 128) && N >= 16 && N <= 256 && N % 16 == 0 && n_acc >= 1 && n_acc * N <= 512,
                  "mma_bench: bad shape");
    const int smem = 1024 + 128 * 128 + 256 * 128;
    FFR_CUDA(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_bench_kernel<<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(out_cycles, M, N, n_acc, iters);
    return launch_status("mma_bench_kernel");
}

// ----------------------------------------------------------------------------------------------------------
// Probe (debug): MN-major operands (the weight-gradient GEMM contracts over pixels, i.e. over the ROWS of the
// row-major activation matrices). a: [96][128] bf16 (k rows, m contiguous), b: [96][64] bf16 (k rows, n contiguous);
// out[128][64] = sum_{k<64} a[k][m] * b[r0 + k][n]. variant 0: LBO = MN-block stride, SBO = 8-row K-group stride;
// variant 1: the two swapped.
// ----------------------------------------------------------------------------------------------------------
namespace ffr {
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1)
mn_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int r0,
                int variant) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                   // 2 blocks x [64 k rows][128 B]
    uint8_t* sB = smem + 2 * 8192;        // [96 k rows][128 B]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * 8192 + 96 * 128);
    uint64_t* mma_bar = bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mma_bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<64>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 1) {
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(bar, 2 * 8192 + 96 * 128);
            tma_load_2d(sA, &tmA, bar, 0, 0);
            tma_load_2d(sA + 8192, &tmA, bar, 64, 0);
            tma_load_2d(sB, &tmB, bar, 0, 0);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        tc_fence_after();
        // instruction descriptor with both operands MN-major
        // variant bit 1: A holds fp16 (a_format = F16), bit 2: B holds fp16 — mixed fp16 x bf16 operands in one MMA
        uint32_t idesc = umma_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
        if (variant & 2) idesc &= ~(7u << 7);
        if (variant & 4) idesc &= ~(7u << 10);
        variant &= 1;
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB) + r0 * 128;
        if (elect_one_sync()) {
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = variant ? umma_desc_mn_sw128(a0 + ks * 2048, 1024, 8192)
                                            : umma_desc_mn_sw128(a0 + ks * 2048, 8192, 1024);
                const uint64_t bd = variant ? umma_desc_mn_sw128(b0 + ks * 2048, 1024, 8192)
                                            : umma_desc_mn_sw128(b0 + ks * 2048, 8192, 1024);
                umma_bf16(tmem_base, ad, bd, idesc, ks > 0);
            }
            umma_commit(mma_bar);
        }
        __syncwarp();
    }
    mbar_wait(mma_bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + c0 + (static_cast<uint32_t>(warp * 32) << 16), v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tmem_base); }
}
}  // namespace ffr

extern "C" FFR_API int ffr_debug_mn_probe(const void* a, const void* b, float* out, int r0, int variant,
                                          ffr_stream_t stream) {
    using namespace ffr;
    CUtensorMap tmA, tmB;
    int rc = make_tmap_2d_bf16(&tmA, a, 96, 128, 128, 64);     // box: 64 k-rows x 64 m
    if (rc) return rc;
    rc = make_tmap_2d_bf16(&tmB, b, 96, 64, 64, 96);           // box: 96 k-rows x 64 n
    if (rc) return rc;
    const int smem = 1024 + 2 * 8192 + 96 * 128 + 64;
    FFR_CUDA(cudaFuncSetAttribute(mn_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mn_probe_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(tmA, tmB, out, r0, variant);
    return launch_status("mn_probe_kernel");
}
