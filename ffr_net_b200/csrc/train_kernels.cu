// RecNet training kernels (models/recnet.py ConvLayer in train mode + its backward, models/trainer.py:154-187):
//   - wgrad_kernel: weight gradient of a reflection-padded 3x3 conv as a tcgen05 GEMM that contracts over PIXELS.
//     Both operands are read straight from their row-major (pixel-major) H9 matrices as MN-major UMMA operands
//     (profiles/r01_probe_mn_major.json); a tap is a TMA row-coordinate shift of the x operand.
//   - pack_conv3x3_kernel: fp16 (forward) / bf16 (dgrad) K-major packings of a conv weight; clip_adam_kernel.
//   (BatchNorm / PReLU forward and backward live in bn_train_kernels.cu.)
#include "host.h"
#include "kernels.h"
#include "ptx.cuh"

#include <cmath>
#include <cuda_fp16.h>

namespace ffr {

// ------------------------------------------------------------------------------------------------------------
// wgrad: dW[co][ci][t] += sum_p dz[p][co] * x[p + shift_t][ci]
// ------------------------------------------------------------------------------------------------------------
static int g_wgrad_splits = 0;      // > 0: override of the split heuristic (ffr_debug_set_wgrad_splits, tuning only)
constexpr int WG_THREADS = 352;     // warps 0-7 epilogue, 8 TMEM alloc, 9 TMA, 10 MMA (same roles as conv_gemm.cu)
constexpr int WG_STAGES = 4;
constexpr int WG_BN = 256;          // ci tile
constexpr int WG_STAGE_BYTES = 2 * 8192 + (WG_BN / 64) * 8192;   // dz: 2 x [64 p][64 co], x: 4 x [64 p][64 ci]

struct WgradParams {
    int P;                 // pixel rows (n * 81)
    int Cout, Cin;         // real channel counts (gradient tensor is [Cout][Cin][3][3] fp32)
    int m_tiles, n_tiles;  // over padded Cout (128) and padded Cin (256)
    int splits, kb_per_split, kb_total;
    int tap_shift[9];
    int pix_iblocks;       // > 0: pixel-major contraction (k-block = 64 images at one interior pixel, 4-D TMA boxes)
    float* ws;             // staging, fp32 [slabs][ntaps][cout_p][cin_p] (tap-major so that a thread's 32 columns are contiguous)
    int cout_p, cin_p;     // m_tiles * 128, n_tiles * WG_BN
    int ntaps;             // 9 (3x3 on the H9 grid) or 1 (plain X^T Y contraction over rows)
    int slabs;             // 1: splits of the pixel axis meet in one staging slab (vector reductions, atomics);
                           // == splits: every split owns a slab (plain stores) and the finish kernel adds the slabs
                           // in a fixed order — deterministic
    uint32_t idesc_xor;    // fp16 operands instead of bf16
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr) {   // LBO = 8192 B between 64-wide MN blocks, SBO = 1024 B
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(8192 >> 4) << 16) |
           (static_cast<uint64_t>(1024 >> 4) << 32) | (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDZ, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + WG_STAGES;
    uint64_t* tfull_bar = bars + 2 * WG_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 9 && lane == 0) { tma_prefetch_desc(&tmDZ); tma_prefetch_desc(&tmX); }
    if (warp == 10 && lane == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 256); }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc<2 * WG_BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int num_work = p.ntaps * p.m_tiles * p.n_tiles * p.splits;

    // work -> (split, n_tile, m_tile, tap); tap slowest so concurrently running CTAs share the dz / x tiles in L2
    if (warp == 9) {
        int stage = 0; uint32_t phase = 0;
        for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
            int w = work;
            const int split = w % p.splits; w /= p.splits;
            const int n_tile = w % p.n_tiles; w /= p.n_tiles;
            const int m_tile = w % p.m_tiles; w /= p.m_tiles;
            const int shift = p.tap_shift[w];
            const int kb0 = split * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            const int tap_r = w / 3, tap_s = w - tap_r * 3;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* s = smem + stage * WG_STAGE_BYTES;
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
                    if (p.pix_iblocks > 0) {       // 64 images at interior pixel q: dz at (h, w), x at the tap's source
                        const int q = kb / p.pix_iblocks, img0 = (kb - q * p.pix_iblocks) * 64;
                        const int qh = q / 7, hh = qh + 1, ww = q - qh * 7 + 1;
                        tma_load_4d(s, &tmDZ, &full_bar[stage], m_tile * 128, ww, hh, img0);
                        tma_load_4d(s + 8192, &tmDZ, &full_bar[stage], m_tile * 128 + 64, ww, hh, img0);
#pragma unroll
                        for (int j = 0; j < WG_BN / 64; ++j)
                            tma_load_4d(s + 16384 + j * 8192, &tmX, &full_bar[stage], n_tile * WG_BN + j * 64,
                                        ww + tap_s - 1, hh + tap_r - 1, img0);
                    } else {
                        tma_load_2d(s, &tmDZ, &full_bar[stage], m_tile * 128, kb * 64);
                        tma_load_2d(s + 8192, &tmDZ, &full_bar[stage], m_tile * 128 + 64, kb * 64);
#pragma unroll
                        for (int j = 0; j < WG_BN / 64; ++j)
                            tma_load_2d(s + 16384 + j * 8192, &tmX, &full_bar[stage], n_tile * WG_BN + j * 64,
                                        kb * 64 + shift);
                    }
                }
                __syncwarp();
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 10) {
        const uint32_t idesc = (umma_idesc_bf16(128, WG_BN) | (1u << 15) | (1u << 16)) ^ p.idesc_xor;   // both operands MN-major
        int stage = 0; uint32_t phase = 0; int it = 0;
        for (int work = blockIdx.x; work < num_work; work += gridDim.x, ++it) {
            const int split = work % p.splits;
            const int kb0 = split * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            const int acc = it & 1;
            mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * WG_BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t a_desc = desc_mn_sw128(smem_u32(smem + stage * WG_STAGE_BYTES));
                const uint64_t b_desc = desc_mn_sw128(smem_u32(smem + stage * WG_STAGE_BYTES + 16384));
                const uint32_t first = (kb > kb0) ? 1u : 0u;
                if (elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)      // 16 pixel rows = 2048 B = +128 in the (addr >> 4) field
                        umma_bf16(d_tmem, a_desc + 128 * ks, b_desc + 128 * ks, idesc, ks > 0 ? 1u : first);
                    umma_commit(&empty_bar[stage]);
                    if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < 8) {
        const int quad = warp & 3, chalf = warp >> 2;
        int it = 0;
        for (int work = blockIdx.x; work < num_work; work += gridDim.x, ++it) {
            int w = work / p.splits;
            const int split = work - w * p.splits;
            const int n_tile = w % p.n_tiles; w /= p.n_tiles;
            const int m_tile = w % p.m_tiles; w /= p.m_tiles;
            const int tap = w;
            const int acc = it & 1;
            const int co = m_tile * 128 + quad * 32 + lane;
            mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * WG_BN + chalf * (WG_BN / 2) + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < WG_BN / 2; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + c0, v);
                tmem_ld_wait();
                const int ci0 = n_tile * WG_BN + chalf * (WG_BN / 2) + c0;
                // one 128-byte run per thread: plain vector stores when the tile is complete, vector reductions
                // (REDG.ADD.F32x4) when the pixel axis is split over several CTAs
                const int slab = (p.slabs > 1) ? split : 0;
                float* o = p.ws + (((long long)slab * p.ntaps + tap) * p.cout_p + co) * p.cin_p + ci0;
                if (p.splits == 1 || p.slabs > 1) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        reinterpret_cast<float4*>(o)[q] =
                            make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                        __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        red_add_v4(o + q * 4, __uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                   __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc<2 * WG_BN>(tmem_base); }
}

// staging [slabs][ntaps][cout_p][cin_p] -> dw [Cout][Cin][ntaps] (OIHW for 3x3); one CTA per output channel. The slabs are
// added in a fixed order; accumulate != 0 adds to dw instead of overwriting it. bias_col >= 0: that input column is
// not part of dw but goes to db[co] (a ones-column appended to the activation operand yields the bias gradient).
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ ws, int slabs, int ntaps, int cout_p,
                                                           int cin_p, int Cout, int Cin, int ld_w, int bias_col,
                                                           float* __restrict__ dw, float* __restrict__ db, int accumulate) {
    // The staging buffer is tap-major ([slab][tap][co][ci], ci contiguous: coalesced reads); dw is OIHW, i.e. the ntaps
    // values of one (co, ci) are contiguous. A chunk of 256 input channels is transposed through shared memory so that
    // the writes are contiguous runs of 256 * ntaps floats instead of ntaps stride-36-byte scatters per thread.
    __shared__ float tile[256 * 9];
    const int co = blockIdx.x;
    const long long slab_stride = (long long)ntaps * cout_p * cin_p;
    for (int ci0 = 0; ci0 < Cin; ci0 += 256) {
        const int ci = ci0 + threadIdx.x;
        if (ci < Cin) {
            // slabs outermost with the taps unrolled: nine independent loads in flight per thread (the tap-outer order had
            // one); every tap still adds its slabs in ascending order: same bits
            float a[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const float* q = ws + (long long)co * cin_p + ci;
            const long long tap_stride = (long long)cout_p * cin_p;
            if (ntaps == 9) {
                for (int sl = 0; sl < slabs; ++sl) {
                    float v[9];
#pragma unroll
                    for (int t = 0; t < 9; ++t) v[t] = __ldg(q + sl * slab_stride + t * tap_stride);
#pragma unroll
                    for (int t = 0; t < 9; ++t) a[t] += v[t];
                }
            } else {
                for (int t = 0; t < ntaps; ++t)
                    for (int sl = 0; sl < slabs; ++sl) a[t] += __ldg(q + sl * slab_stride + t * tap_stride);
            }
#pragma unroll
            for (int t = 0; t < 9; ++t)
                if (t < ntaps) tile[threadIdx.x * ntaps + t] = a[t];
            if (ci == bias_col) {
                if (accumulate) db[co] += tile[threadIdx.x * ntaps]; else db[co] = tile[threadIdx.x * ntaps];
            }
        }
        __syncthreads();
        // output columns [ci0, min(ci0 + 256, Cin)) except the bias column
        const int n_ci = min(256, Cin - ci0);
        const int skip = (bias_col >= ci0 && bias_col < ci0 + n_ci) ? bias_col - ci0 : -1;
        float* o = dw + ((long long)co * ld_w + ci0) * ntaps;
        for (int i = threadIdx.x; i < n_ci * ntaps; i += 256) {
            if (skip >= 0 && i / ntaps == skip) continue;
            if (accumulate) o[i] += tile[i]; else o[i] = tile[i];
        }
        __syncthreads();
    }
}

// Work items = 9 taps x m_tiles x n_tiles x splits on a persistent grid: pick the split count that minimises
// (waves over the SMs) x (k-blocks per item + a fixed per-item cost), so the big layers run in one or two full waves.
static int wgrad_pick_splits(int base_work, int kb_total, int sms) {
    int best = 1;
    long long best_cost = -1;
    const int max_splits = kb_total / 8 > 1 ? kb_total / 8 : 1;
    for (int s = 1; s <= max_splits; ++s) {
        const int kps = (kb_total + s - 1) / s;
        const int real = (kb_total + kps - 1) / kps;
        const long long waves = ((long long)base_work * real + sms - 1) / sms;
        const long long cost = waves * (kps + 4) + (real > 1 ? 1 : 0);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = real; }
    }
    return best;
}

// dz: [P][ld_dz] (zeros on halo rows), x: [P][ld_x] H9 (with halo) — both bf16, or both fp16 (f16 != 0);
// dw: fp32 [Cout][Cin][ntaps]; ws: fp32 staging of wgrad_workspace_floats() elements.
// deterministic != 0: one staging slab per split of the pixel axis, added in a fixed order by the finish kernel.
long long wgrad_workspace_floats(int P, int Cout, int Cin, int ntaps, int G, int deterministic, int* splits_out) {
    const int m_tiles = (Cout + 127) / 128, n_tiles = (Cin + WG_BN - 1) / WG_BN;
    const int n_img = P / 81;
    const bool pix = (ntaps == 9) && (G == 9) && (P % 81 == 0) && pixmajor_profitable_k64(n_img);
    const int kb_total = pix ? 49 * ((n_img + 63) / 64) : (P + 63) / 64;
    int splits = wgrad_pick_splits(ntaps * m_tiles * n_tiles, kb_total, num_sms());
    if (g_wgrad_splits > 0) splits = g_wgrad_splits;
    const int kps = (kb_total + splits - 1) / splits;
    splits = (kb_total + kps - 1) / kps;
    if (splits_out) *splits_out = splits;
    return (long long)(deterministic ? splits : 1) * ntaps * (m_tiles * 128) * (n_tiles * WG_BN);
}

int wgrad_launch_ex(const void* dz, int ld_dz, const void* x, int ld_x, int x_ch0, int P, int Cout, int Cin, int G,
                    int ntaps, int f16, int deterministic, int accumulate, int ld_w, int bias_col, float* dw, float* db,
                    float* ws, cudaStream_t stream) {
    WgradParams p;
    p.P = P; p.Cout = Cout; p.Cin = Cin; p.ws = ws;
    p.ntaps = ntaps;
    p.idesc_xor = f16 ? ((1u << 7) | (1u << 10)) : 0u;
    p.m_tiles = (Cout + 127) / 128;
    p.n_tiles = (Cin + WG_BN - 1) / WG_BN;
    p.cout_p = p.m_tiles * 128;
    p.cin_p = p.n_tiles * WG_BN;
    const int n_img = P / 81;
    const bool pix = (ntaps == 9) && (G == 9) && (P % 81 == 0) && pixmajor_profitable_k64(n_img);
    p.pix_iblocks = pix ? (n_img + 63) / 64 : 0;
    p.kb_total = pix ? 49 * p.pix_iblocks : (P + 63) / 64;
    const int base_work = ntaps * p.m_tiles * p.n_tiles;
    int splits = 1;
    wgrad_workspace_floats(P, Cout, Cin, ntaps, G, deterministic, &splits);
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.slabs = deterministic ? p.splits : 1;
    for (int t = 0; t < 9; ++t) p.tap_shift[t] = 0;
    if (ntaps == 9)
        for (int r = 0; r < 3; ++r)
            for (int s = 0; s < 3; ++s) p.tap_shift[r * 3 + s] = (r - 1) * G + (s - 1);
    CUtensorMap tmDZ, tmX;
    // the x map starts at channel x_ch0 of a possibly wider (concatenated) matrix and exposes Cin padded to 64 columns
    const int cin_cols = (Cin + 63) / 64 * 64;
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x) + x_ch0;
    int rc;
    if (pix) {
        rc = make_tmap_h9_pixel_bf16(&tmDZ, dz, (uint64_t)n_img, (uint64_t)ld_dz, (uint64_t)ld_dz, 64);
        if (rc) return rc;
        rc = make_tmap_h9_pixel_bf16(&tmX, xb, (uint64_t)n_img, (uint64_t)cin_cols, (uint64_t)ld_x, 64);
    } else {
        rc = make_tmap_2d_bf16(&tmDZ, dz, (uint64_t)P, (uint64_t)ld_dz, (uint64_t)ld_dz, 64);
        if (rc) return rc;
        rc = make_tmap_2d_bf16(&tmX, xb, (uint64_t)P, (uint64_t)cin_cols, (uint64_t)ld_x, 64);
    }
    if (rc) return rc;
    const int smem = 1024 + WG_STAGES * WG_STAGE_BYTES + 256;
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    if (p.splits > 1 && p.slabs == 1)
        FFR_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)ntaps * p.cout_p * p.cin_p, stream));
    const int num_work = base_work * p.splits;
    const int grid = num_work < num_sms() ? num_work : num_sms();
    wgrad_kernel<<<grid, WG_THREADS, smem, stream>>>(tmDZ, tmX, p);
    rc = launch_status("wgrad_kernel");
    if (rc) return rc;
    wgrad_finish_kernel<<<Cout, 256, 0, stream>>>(ws, p.slabs, ntaps, p.cout_p, p.cin_p, Cout, Cin, ld_w, bias_col, dw, db,
                                                  accumulate);
    return launch_status("wgrad_finish_kernel");
}

void set_wgrad_splits(int s) { g_wgrad_splits = s; }

// ------------------------------------------------------------------------------------------------------------
// Weight packing for one ConvLayer in a single launch (done once per optimizer step, cached by the host):
//   fwd  [cout_p][9*cin_p]  : fwd[co][t*cin_p + ci]          = w[co][ci][t]
//   dgrad[cin_p][9*cout_p]  : dgrad[ci][(8-t)*cout_p + co]   = w[co][ci][t]   (spatially flipped, transposed)
// Padded rows/columns are written as zeros.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_conv3x3_kernel(const float* __restrict__ w, int cout, int cin, int cout_p, int cin_p,
                    __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dgrad, int fwd_f16) {
    const long long total = (long long)cout_p * cin_p * 9;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % cin_p);
        const int t = (int)((i / cin_p) % 9);
        const int co = (int)(i / ((long long)cin_p * 9));
        const float v = (co < cout && ci < cin) ? w[((long long)co * cin + ci) * 9 + t] : 0.f;
        const __nv_bfloat16 b = __float2bfloat16(v);
        if (fwd_f16) reinterpret_cast<__half*>(fwd)[i] = __float2half_rn(v);
        else fwd[i] = b;                                                 // i == (co*9 + t)*cin_p + ci
        if (dgrad) dgrad[((long long)ci * 9 + (8 - t)) * cout_p + co] = b;
    }
}

// Tiled version: a 32 (co) x 32 (ci) x 9 block of the OIHW tensor goes through shared memory, so that the fp32 reads
// are contiguous runs of 288 floats and both packed outputs are written as 64-byte rows of 16-bit pairs (the direct
// kernel above reads with a 36-byte stride and scatters single 2-byte elements of the transposed copy).
constexpr int PK_T = 32;
constexpr int PK_PITCH = PK_T * 9 + 1;
__global__ void __launch_bounds__(256)
pack_conv3x3_tiled_kernel(const float* __restrict__ w, int cout, int cin, int cout_p, int cin_p,
                          __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dgrad, int fwd_f16) {
    __shared__ float tile[PK_T * PK_PITCH];
    const int ci0 = blockIdx.x * PK_T, co0 = blockIdx.y * PK_T;
    for (int e = threadIdx.x; e < PK_T * PK_T * 9; e += 256) {
        const int r = e / (PK_T * 9), col = e - r * (PK_T * 9);
        const int co = co0 + r, ci = ci0 + col / 9;
        tile[r * PK_PITCH + col] = (co < cout && ci < cin) ? __ldg(w + ((long long)co * cin + ci0) * 9 + col) : 0.f;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < PK_T * 9 * (PK_T / 2); e += 256) {
        const int pr = e % (PK_T / 2), t = (e / (PK_T / 2)) % 9, r = e / ((PK_T / 2) * 9);
        {   // forward packing: fwd[co][t][ci], pair of input channels (2 pr, 2 pr + 1) of output channel r
            const int co = co0 + r, ci = ci0 + 2 * pr;
            if (co < cout_p && ci < cin_p) {
                const float v0 = tile[r * PK_PITCH + (2 * pr) * 9 + t], v1 = tile[r * PK_PITCH + (2 * pr + 1) * 9 + t];
                uint32_t u;
                if (fwd_f16) { const __half2 h = __floats2half2_rn(v0, v1); u = *reinterpret_cast<const uint32_t*>(&h); }
                else u = pack_bf16x2(v0, v1);
                *reinterpret_cast<uint32_t*>(fwd + ((long long)co * 9 + t) * cin_p + ci) = u;
            }
        }
        if (dgrad) {   // flipped / transposed packing: dgrad[ci][8 - t][co], pair of output channels of input channel r
            const int ci = ci0 + r, co = co0 + 2 * pr;
            if (ci < cin_p && co < cout_p) {
                const float v0 = tile[(2 * pr) * PK_PITCH + r * 9 + t], v1 = tile[(2 * pr + 1) * PK_PITCH + r * 9 + t];
                *reinterpret_cast<uint32_t*>(dgrad + ((long long)ci * 9 + (8 - t)) * cout_p + co) = pack_bf16x2(v0, v1);
            }
        }
    }
}

int pack_conv3x3_launch_ex(const float* w, int cout, int cin, int cout_p, int cin_p, void* fwd, void* dgrad, int fwd_f16,
                           cudaStream_t stream) {
    if (cout_p % 2 == 0 && cin_p % 2 == 0 && cout_p > 0 && cin_p > 0) {
        const dim3 tgrid((cin_p + PK_T - 1) / PK_T, (cout_p + PK_T - 1) / PK_T);
        pack_conv3x3_tiled_kernel<<<tgrid, 256, 0, stream>>>(w, cout, cin, cout_p, cin_p, reinterpret_cast<__nv_bfloat16*>(fwd),
                                                            reinterpret_cast<__nv_bfloat16*>(dgrad), fwd_f16);
        return launch_status("pack_conv3x3_tiled_kernel");
    }
    const long long total = (long long)cout_p * cin_p * 9;
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    pack_conv3x3_kernel<<<grid, 256, 0, stream>>>(w, cout, cin, cout_p, cin_p, reinterpret_cast<__nv_bfloat16*>(fwd),
                                                  reinterpret_cast<__nv_bfloat16*>(dgrad), fwd_f16);
    return launch_status("pack_conv3x3_kernel");
}

int pack_conv3x3_launch(const float* w, int cout, int cin, int cout_p, int cin_p, void* fwd, void* dgrad,
                        cudaStream_t stream) {
    return pack_conv3x3_launch_ex(w, cout, cin, cout_p, cin_p, fwd, dgrad, 0, stream);
}

}  // namespace ffr

// ------------------------------------------------------------------------------------------------------------
// Fused clip_grad_value_ + Adam over a table of tensors in ONE launch (models/trainer.py:182-187 with
// torch.optim.Adam semantics, no amsgrad): g = clamp(g, +-clip) (written back), optional L2 weight decay,
// m = b1 m + (1-b1) g, v = b2 v + (1-b2) g^2, p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).
// ------------------------------------------------------------------------------------------------------------
namespace ffr {
struct AdamTensor { float* p; float* g; float* m; float* v; long long n; };
constexpr int ADAM_CHUNK = 4096;   // elements per CTA (256 threads x 4 float4)

// hyper[0] = learning rate, hyper[1] = step count (1-based, as float): device-resident so that a captured CUDA graph
// of the training step stays valid when the host scheduler changes the rate / the step advances.
__global__ void __launch_bounds__(256)
clip_adam_kernel(const AdamTensor* __restrict__ tab, const int2* __restrict__ chunks, const float* __restrict__ hyper,
                 float b1, float b2, float eps, float wd, float clip) {
    const int2 ch = chunks[blockIdx.x];
    const AdamTensor t = tab[ch.x];
    const float lr = hyper[0], step = hyper[1];
    const float lr_over_bc1 = lr / (1.f - powf(b1, step));
    const float inv_sqrt_bc2 = rsqrtf(1.f - powf(b2, step));
    const long long base = (long long)ch.y * ADAM_CHUNK;
    // full chunks of 16-byte aligned tensors: 128-bit accesses, the four loads of a group issued together (the tensors
    // are views of flat buffers at arbitrary element offsets, hence the run-time alignment test); same arithmetic
    if (base + ADAM_CHUNK <= t.n &&
        ((reinterpret_cast<uintptr_t>(t.p + base) | reinterpret_cast<uintptr_t>(t.g + base) |
          reinterpret_cast<uintptr_t>(t.m + base) | reinterpret_cast<uintptr_t>(t.v + base)) & 15) == 0) {
        float4* p4 = reinterpret_cast<float4*>(t.p + base);
        float4* g4 = reinterpret_cast<float4*>(t.g + base);
        float4* m4 = reinterpret_cast<float4*>(t.m + base);
        float4* v4 = reinterpret_cast<float4*>(t.v + base);
#pragma unroll 2
        for (int k = 0; k < ADAM_CHUNK / 1024; ++k) {
            const int i = k * 256 + threadIdx.x;
            const float4 gi = g4[i], pi = p4[i], mi = m4[i], vi = v4[i];
            float gg[4] = {gi.x, gi.y, gi.z, gi.w}, pp[4] = {pi.x, pi.y, pi.z, pi.w};
            float mm[4] = {mi.x, mi.y, mi.z, mi.w}, vv[4] = {vi.x, vi.y, vi.z, vi.w};
            float gc[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float g = fminf(fmaxf(gg[e], -clip), clip);
                gc[e] = g;
                const float p = pp[e];
                g = fmaf(wd, p, g);
                const float m = fmaf(b1, mm[e], (1.f - b1) * g);
                const float v = fmaf(b2, vv[e], (1.f - b2) * g * g);
                mm[e] = m;
                vv[e] = v;
                pp[e] = p - lr_over_bc1 * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
            }
            g4[i] = make_float4(gc[0], gc[1], gc[2], gc[3]);
            m4[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
            v4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
            p4[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        }
        return;
    }
    for (int k = 0; k < ADAM_CHUNK / 256; ++k) {
        const long long i = base + k * 256 + threadIdx.x;
        if (i >= t.n) break;
        float g = fminf(fmaxf(t.g[i], -clip), clip);
        t.g[i] = g;
        const float p = t.p[i];
        g = fmaf(wd, p, g);
        const float m = fmaf(b1, t.m[i], (1.f - b1) * g);
        const float v = fmaf(b2, t.v[i], (1.f - b2) * g * g);
        t.m[i] = m;
        t.v[i] = v;
        t.p[i] = p - lr_over_bc1 * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
    }
}

int clip_adam_launch(const void* table, const int* chunks, int n_chunks, const float* hyper, float b1, float b2,
                     float eps, float wd, float clip, cudaStream_t stream) {
    if (n_chunks == 0) return 0;
    clip_adam_kernel<<<n_chunks, 256, 0, stream>>>(reinterpret_cast<const AdamTensor*>(table),
                                                  reinterpret_cast<const int2*>(chunks), hyper, b1, b2, eps, wd, clip);
    return launch_status("clip_adam_kernel");
}
}  // namespace ffr
