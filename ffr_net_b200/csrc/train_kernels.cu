// RecNet training kernels (models/recnet.py ConvLayer in train mode + its backward, models/trainer.py:154-187):
//   - wgrad_kernel: weight gradient of a reflection-padded 3x3 conv as a tcgen05 GEMM that contracts over PIXELS.
//     Both operands are read straight from their row-major (pixel-major) H9 matrices as MN-major UMMA operands
//     (profiles/r01_probe_mn_major.json); a tap is a TMA row-coordinate shift of the x operand.
//   - bn_prelu_fwd_kernel: batch-statistics BatchNorm + PReLU (+ residual) on the raw conv output, scattered into H9.
//   - bn_prelu_bwd_reduce_kernel / bn_prelu_bwd_dz_kernel: gradient fold of the reflection mirrors, PReLU and
//     BatchNorm backward (per-channel reductions, then dz).
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
    const int co = blockIdx.x;
    const long long slab_stride = (long long)ntaps * cout_p * cin_p;
    for (int ci = threadIdx.x; ci < Cin; ci += blockDim.x) {
        float v[9];
        for (int t = 0; t < ntaps; ++t) {
            const float* q = ws + ((long long)t * cout_p + co) * cin_p + ci;
            float a = 0.f;
            for (int sl = 0; sl < slabs; ++sl) a += __ldg(q + sl * slab_stride);
            v[t] = a;
        }
        if (ci == bias_col) {
            if (accumulate) db[co] += v[0]; else db[co] = v[0];
            continue;
        }
        float* o = dw + ((long long)co * ld_w + ci) * ntaps;
        for (int t = 0; t < ntaps; ++t) { if (accumulate) o[t] += v[t]; else o[t] = v[t]; }
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

int wgrad_launch(const void* dz, int ld_dz, const void* x, int ld_x, int x_ch0, int P, int Cout, int Cin, int G,
                 float* dw, float* ws, cudaStream_t stream) {
    return wgrad_launch_ex(dz, ld_dz, x, ld_x, x_ch0, P, Cout, Cin, G, 9, 0, 0, 0, Cin, -1, dw, nullptr, ws, stream);
}

void set_wgrad_splits(int s) { g_wgrad_splits = s; }

// ------------------------------------------------------------------------------------------------------------
// BatchNorm(batch statistics) + PReLU (+ residual) forward on the raw conv output z (rows of the H9 grid).
// stats: [2][C] = per-channel sum and sum of squares over the n*49 valid rows (accumulated by the conv epilogue).
// Writes a = prelu(gamma*(z-mean)*rstd + beta) (+ res) to every destination of the scatter table (self + mirrors).
// grid.x covers rows*C/8 work items of 8 channels.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_prelu_fwd_kernel(const __nv_bfloat16* __restrict__ z, int ldz, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ slope, const __nv_bfloat16* __restrict__ res, int ldres,
                    __nv_bfloat16* __restrict__ out, int ldo, const int2* __restrict__ scatter, int scatter_n,
                    int n_img, int C) {
    const int c8n = C / 8;
    const long long total = (long long)n_img * 49 * c8n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % c8n);
        const long long pr = i / c8n;
        const int n = (int)(pr / 49), pix = (int)(pr - (long long)n * 49);
        const int r_local = (pix / 7 + 1) * 9 + (pix % 7 + 1);
        const long long row = (long long)n * 81 + r_local;
        const uint4 zv = __ldg(reinterpret_cast<const uint4*>(z + row * ldz) + c8);
        float v[8] = {bf16lo(zv.x), bf16hi(zv.x), bf16lo(zv.y), bf16hi(zv.y), bf16lo(zv.z), bf16hi(zv.z), bf16lo(zv.w), bf16hi(zv.w)};
        float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (res) {
            const uint4 rv = __ldg(reinterpret_cast<const uint4*>(res + row * ldres) + c8);
            r[0] = bf16lo(rv.x); r[1] = bf16hi(rv.x); r[2] = bf16lo(rv.y); r[3] = bf16hi(rv.y);
            r[4] = bf16lo(rv.z); r[5] = bf16hi(rv.z); r[6] = bf16lo(rv.w); r[7] = bf16hi(rv.w);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c8 * 8 + j;
            float y = (v[j] - mean[c]) * rstd[c] * gamma[c] + beta[c];
            y = fmaxf(y, 0.f) + slope[c] * fminf(y, 0.f);
            v[j] = y + r[j];
        }
        const uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        for (int k = 0; k < scatter_n; ++k) {
            const int2 e = __ldg(scatter + r_local * scatter_n + k);
            if (e.x >= 0) *(reinterpret_cast<uint4*>(out + ((long long)n * 81 + e.x) * ldo + e.y) + c8) = o;
        }
    }
}

int bn_prelu_fwd_launch(const void* z, int ldz, const float* mean, const float* rstd, const float* gamma,
                        const float* beta, const float* slope, const void* res, int ldres, void* out, int ldo,
                        const int* scatter, int scatter_n, int n_img, int C, cudaStream_t stream) {
    FFR_CHECK_ARG(C % 8 == 0, "bn_prelu_fwd: C=%d", C);
    const long long total = (long long)n_img * 49 * (C / 8);
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    bn_prelu_fwd_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(z), ldz, mean, rstd, gamma, beta,
                                                  slope, reinterpret_cast<const __nv_bfloat16*>(res), ldres,
                                                  reinterpret_cast<__nv_bfloat16*>(out), ldo,
                                                  reinterpret_cast<const int2*>(scatter), scatter_n, n_img, C);
    return launch_status("bn_prelu_fwd_kernel");
}

// ------------------------------------------------------------------------------------------------------------
// Backward, pass 1. da_h9 is the gradient w.r.t. the H9 OUTPUT tensor (own row + mirror rows, possibly a channel
// slot of a wider matrix): fold it (sum over the scatter destinations), then
//   dy = da * prelu'(y);   sums[0][c] += dy;  sums[1][c] += dy * zhat;  sums[2][c] += da * min(y, 0)
// and store dy (bf16, plain rows) for pass 2; optionally store the folded da as the residual-branch gradient.
// One CTA handles a slab of valid rows for 64 channels; per-channel partial sums are reduced in shared memory.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_prelu_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ da, int ldda, const int2* __restrict__ scatter, int scatter_n,
                           const __nv_bfloat16* __restrict__ z, int ldz, const float* __restrict__ mean,
                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                           const float* __restrict__ beta, const float* __restrict__ slope,
                           __nv_bfloat16* __restrict__ dy, int lddy, __nv_bfloat16* __restrict__ dres, int lddres,
                           float* __restrict__ sums, int n_img, int C) {
    // 256 threads = 32 row lanes x 8 channel groups of 8 (16-byte accesses; a row's 64-channel slab is one 128-byte line)
    __shared__ float red[3][32][65];
    const int c0 = blockIdx.y * 64;
    const int cg = (threadIdx.x & 7) * 8;            // first channel of this thread's group within the slab
    const int rlane = threadIdx.x >> 3;              // 32 row lanes
    const long long rows = (long long)n_img * 49;
    float s0[8], s1[8], s2[8], m[8], rs[8], g[8], b[8], sl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + cg + j;
        s0[j] = s1[j] = s2[j] = 0.f;
        m[j] = mean[c]; rs[j] = rstd[c]; g[j] = gamma[c]; b[j] = beta[c]; sl[j] = slope[c];
    }
    for (long long pr = (long long)blockIdx.x * 32 + rlane; pr < rows; pr += (long long)gridDim.x * 32) {
        const int n = (int)(pr / 49), pix = (int)(pr - (long long)n * 49);
        const int r_local = (pix / 7 + 1) * 9 + (pix % 7 + 1);
        const long long row = (long long)n * 81 + r_local;
        const uint4 zv = __ldg(reinterpret_cast<const uint4*>(z + row * ldz + c0 + cg));
        float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < scatter_n; ++k) {
            const int2 e = __ldg(scatter + r_local * scatter_n + k);
            if (e.x >= 0) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(da + ((long long)n * 81 + e.x) * ldda + e.y + c0 + cg));
                a[0] += bf16lo(u.x); a[1] += bf16hi(u.x); a[2] += bf16lo(u.y); a[3] += bf16hi(u.y);
                a[4] += bf16lo(u.z); a[5] += bf16hi(u.z); a[6] += bf16lo(u.w); a[7] += bf16hi(u.w);
            }
        }
        const float zz[8] = {bf16lo(zv.x), bf16hi(zv.x), bf16lo(zv.y), bf16hi(zv.y),
                             bf16lo(zv.z), bf16hi(zv.z), bf16lo(zv.w), bf16hi(zv.w)};
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float zh = (zz[j] - m[j]) * rs[j];
            const float y = zh * g[j] + b[j];
            d[j] = a[j] * (y > 0.f ? 1.f : sl[j]);
            s0[j] += d[j];
            s1[j] += d[j] * zh;
            s2[j] += a[j] * fminf(y, 0.f);
        }
        *reinterpret_cast<uint4*>(dy + row * lddy + c0 + cg) =
            make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
        if (dres)
            *reinterpret_cast<uint4*>(dres + row * lddres + c0 + cg) =
                make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[0][rlane][cg + j] = s0[j]; red[1][rlane][cg + j] = s1[j]; red[2][rlane][cg + j] = s2[j];
    }
    __syncthreads();
    if (threadIdx.x < 192) {
        const int q = threadIdx.x / 64, c = threadIdx.x % 64;
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) t += red[q][r][c];
        atomicAdd(sums + q * C + c0 + c, t);
    }
}

// Pass 2: dz = gamma * rstd * (dy - sum(dy)/cnt - zhat * sum(dy*zhat)/cnt) on the valid rows of dz_h9 and ZERO on its
// halo rows (the dgrad / wgrad GEMMs read dz as a zero-padded map); the residual-branch gradient buffer gets its
// halo rows zeroed here as well (its valid rows were written by pass 1).
__global__ void __launch_bounds__(256)
bn_prelu_bwd_dz_kernel(const __nv_bfloat16* __restrict__ dy, int lddy, const __nv_bfloat16* __restrict__ z, int ldz,
                       const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                       const float* __restrict__ sums, __nv_bfloat16* __restrict__ dz, int lddz,
                       __nv_bfloat16* __restrict__ dres, int lddres, int n_img, int C) {
    const int c8n = C / 8;
    const float inv_cnt = 1.0f / (float)(n_img * 49);
    const long long total = (long long)n_img * 81 * c8n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % c8n);
        const long long row = i / c8n;
        const int pos = (int)(row % 81);
        const int hp = pos / 9, wp = pos - hp * 9;
        if (hp == 0 || hp == 8 || wp == 0 || wp == 8) {
            *(reinterpret_cast<uint4*>(dz + row * lddz) + c8) = make_uint4(0, 0, 0, 0);
            if (dres) *(reinterpret_cast<uint4*>(dres + row * lddres) + c8) = make_uint4(0, 0, 0, 0);
            continue;
        }
        const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dy + row * lddy) + c8);
        const uint4 zv = __ldg(reinterpret_cast<const uint4*>(z + row * ldz) + c8);
        const float d[8] = {bf16lo(dv.x), bf16hi(dv.x), bf16lo(dv.y), bf16hi(dv.y), bf16lo(dv.z), bf16hi(dv.z), bf16lo(dv.w), bf16hi(dv.w)};
        const float zz[8] = {bf16lo(zv.x), bf16hi(zv.x), bf16lo(zv.y), bf16hi(zv.y), bf16lo(zv.z), bf16hi(zv.z), bf16lo(zv.w), bf16hi(zv.w)};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c8 * 8 + j;
            const float zh = (zz[j] - mean[c]) * rstd[c];
            o[j] = gamma[c] * rstd[c] * (d[j] - sums[c] * inv_cnt - zh * sums[C + c] * inv_cnt);
        }
        *(reinterpret_cast<uint4*>(dz + row * lddz) + c8) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
}

int bn_prelu_bwd_launch(const void* da, int ldda, const int* scatter, int scatter_n, const void* z, int ldz,
                        const float* mean, const float* rstd, const float* gamma, const float* beta, const float* slope,
                        void* dy, int lddy, void* dres, int lddres, float* sums, void* dz, int lddz, int n_img, int C,
                        cudaStream_t stream) {
    FFR_CHECK_ARG(C % 64 == 0, "bn_prelu_bwd: C=%d", C);
    FFR_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 3 * C, stream));
    const long long rows = (long long)n_img * 49;
    // ~8 CTAs per SM in total: enough rows in flight to cover the HBM latency, few enough atomics at the end
    int gx = (int)((rows + 63) / 64);
    const int cap = (num_sms() * 8 + C / 64 - 1) / (C / 64);
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(gx, C / 64);
    bn_prelu_bwd_reduce_kernel<<<grid, 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(da), ldda, reinterpret_cast<const int2*>(scatter), scatter_n,
        reinterpret_cast<const __nv_bfloat16*>(z), ldz, mean, rstd, gamma, beta, slope,
        reinterpret_cast<__nv_bfloat16*>(dy), lddy, reinterpret_cast<__nv_bfloat16*>(dres), lddres, sums, n_img, C);
    int rc = launch_status("bn_prelu_bwd_reduce_kernel");
    if (rc) return rc;
    const long long total = (long long)n_img * 81 * (C / 8);
    int g2 = (int)((total + 255) / 256);
    if (g2 > num_sms() * 8) g2 = num_sms() * 8;
    bn_prelu_bwd_dz_kernel<<<g2, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), lddy,
                                                   reinterpret_cast<const __nv_bfloat16*>(z), ldz, mean, rstd, gamma, sums,
                                                   reinterpret_cast<__nv_bfloat16*>(dz), lddz,
                                                   reinterpret_cast<__nv_bfloat16*>(dres), lddres, n_img, C);
    return launch_status("bn_prelu_bwd_dz_kernel");
}

// ------------------------------------------------------------------------------------------------------------
// Layout converters with gradients: fp32 NCHW (n,C,7,7) <-> bf16 H9.
//   nchw_to_h9: scatter each pixel to its own row and mirrors, channel slot ch0 of a matrix with row pitch ld.
//   h9_to_nchw_fold: out[n][c][pix] = sum over the scatter destinations of pix (gradient fold), or just the own row.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nchw_to_h9_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int ld, int ch0, int n_img, int C,
                  int mirror) {
    __shared__ float tile[49][65];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 64 * 49; i += 256) {
        const int c = i / 49, pix = i - c * 49;
        tile[pix][c] = (c0 + c < C) ? x[((long long)n * C + c0 + c) * 49 + pix] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 81 * 32; i += 256) {
        const int pos = i >> 5, cp = (i & 31) * 2;
        const int hp = pos / 9, wp = pos - hp * 9;
        const int sh = hp == 0 ? 1 : (hp == 8 ? 5 : hp - 1), sw = wp == 0 ? 1 : (wp == 8 ? 5 : wp - 1);
        const int pix = sh * 7 + sw;
        const bool halo = (hp == 0 || hp == 8 || wp == 0 || wp == 8);
        *reinterpret_cast<uint32_t*>(out + ((long long)n * 81 + pos) * ld + ch0 + c0 + cp) =
            (halo && !mirror) ? 0u : pack_bf16x2(tile[pix][cp], tile[pix][cp + 1]);
    }
}

__global__ void __launch_bounds__(256)
h9_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, int ld, int ch0, float* __restrict__ y, int n_img, int C,
                  int fold) {
    __shared__ float tile[49][65];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 49 * 32; i += 256) {
        const int pix = i >> 5, cp = (i & 31) * 2;
        const int h = pix / 7, w = pix - h * 7;
        const int mh = fold ? ((h == 1) ? -2 : ((h == 5) ? 2 : 0)) : 0;
        const int mw = fold ? ((w == 1) ? -2 : ((w == 5) ? 2 : 0)) : 0;
        const long long base = (long long)n * 81 + (h + 1) * 9 + (w + 1);
        float a = 0.f, b = 0.f;
        auto add = [&](long long row) {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(in + row * ld + ch0 + c0 + cp);
            a += bf16lo(u); b += bf16hi(u);
        };
        add(base);
        if (mh) add(base + mh * 9);
        if (mw) add(base + mw);
        if (mh && mw) add(base + mh * 9 + mw);
        tile[pix][cp] = a; tile[pix][cp + 1] = b;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 49; i += 256) {
        const int c = i / 49, pix = i - c * 49;
        if (c0 + c < C) y[((long long)n * C + c0 + c) * 49 + pix] = tile[pix][c];
    }
}

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

int pack_conv3x3_launch_ex(const float* w, int cout, int cin, int cout_p, int cin_p, void* fwd, void* dgrad, int fwd_f16,
                           cudaStream_t stream) {
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

int nchw_to_h9_launch(const float* x, void* out, int ld, int ch0, int n_img, int C, int mirror, cudaStream_t stream) {
    dim3 grid((C + 63) / 64, n_img);
    nchw_to_h9_kernel<<<grid, 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), ld, ch0, n_img, C, mirror);
    return launch_status("nchw_to_h9_kernel");
}

int h9_to_nchw_launch(const void* in, int ld, int ch0, float* y, int n_img, int C, int fold, cudaStream_t stream) {
    dim3 grid((C + 63) / 64, n_img);
    h9_to_nchw_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), ld, ch0, y, n_img, C, fold);
    return launch_status("h9_to_nchw_kernel");
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
