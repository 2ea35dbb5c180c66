// RecNet (feature rectification, models/recnet.py) — the kernels that are not implicit GEMMs.
//
// Layouts (DESIGN.md "RecNet"):
//   H9  : a 7x7 map with its 1-pixel REFLECTION halo materialised, bf16 [n*81][C]; pixel (h,w) at row (h+1)*9+(w+1).
//         A 3x3 "ReflectionPad2d(1) + Conv2d(pad 0)" (recnet.py:64-65,78-82) is then a shifted-row GEMM with G = 9;
//         producers write every valid pixel to its own row and to the (up to 3) halo rows that mirror it.
//   XT  : X^T per sample, bf16 [n*128][512] (rows hw < 49 valid) — the A operand of feat_channel = M_channel @ X.
//   H5  : input of the last Conv4Channel linear, bf16 [n*512][64] (32 valid columns).
#include "host.h"
#include "kernels.h"
#include "ptx.cuh"

namespace ffr {

__device__ __forceinline__ float warp_sum_r(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int reflect_src(int p) {  // padded index 0..8 -> source index 0..6 (ReflectionPad2d(1))
    return p == 0 ? 1 : (p == 8 ? 5 : p - 1);
}

// PrepParams (kernels.h): x fp32 [n][512][49]; w0aT [49][32], w0bT [512][32] = Conv4Channel.0.weight split and
// transposed; b0; slope1/4/7 [512] (PReLU over the 512 rows, recnet.py:374); A1 = W3 W2, c1 = W3 b2 + b3, A2 = W6 W5,
// c2 = W6 b5 + b6 (folded 32x32 maps); outputs s0 / cm / xt / h5 (layouts above), optional fp32 ss_space [n][49][49].

// One CTA per sample: self-similarity (recnet.py:220-236), layout fan-out of X, and the thin part of the channel
// rectifier Conv4Channel (recnet.py:372-385) up to the input of its last Linear.
//   ss_space[i][j]  = <x[:,i], x[:,j]> / (max(|x[:,i]|,eps) max(|x[:,j]|,eps))
//   ss_channel      = Xh Xh^T with Xh = row-normalised X is never materialised: Linear(561->32) applied to
//                     cat(X, ss_channel) equals X W0a^T + Xh (Xh^T W0b^T) + b0 (associativity), a 49x32 intermediate.
//   Linear(32->512) followed by Linear(512->32) has no non-linearity in between and is applied as the folded 32x32 map.
__global__ void __launch_bounds__(512, 1) recnet_prep_kernel(const PrepParams p) {
    extern __shared__ float sm[];
    float* xs = sm;                      // [512][49]
    float* inv_c = xs + 512 * 49;        // [512]
    float* inv_s = inv_c + 512;          // [64]
    float* Gs = inv_s + 64;              // [49][49]
    float* T = Gs + 49 * 49 + 3;         // [49][32]   (49*49 + 3 = 2404: keeps T, W0a, A1s, A2s 16-byte aligned)
    float* W0a = T + 49 * 32;            // [49][32]
    float* A1s = W0a + 49 * 32;          // [32][32]
    float* A2s = A1s + 1024;             // [32][32]
    float* misc = A2s + 1024;            // b0, c1, c2 [3][32]
    float* Gpart = misc + 96;            // [8][49*49] per-channel-slice partial Grams
    const int n = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* x = p.x + (long long)n * 512 * 49;

    for (int i = tid; i < 512 * 49; i += 512) xs[i] = x[i];
    for (int i = tid; i < 49 * 32; i += 512) W0a[i] = p.w0aT[i];
    for (int i = tid; i < 1024; i += 512) { A1s[i] = p.A1[i]; A2s[i] = p.A2[i]; }
    if (tid < 32) { misc[tid] = p.b0[tid]; misc[32 + tid] = p.c1[tid]; misc[64 + tid] = p.c2[tid]; }
    __syncthreads();

    {   // row norms over HW (F.normalize(dim=2) of (N,C,HW), eps 1e-12)
        float ss = 0.f;
        for (int hw = 0; hw < 49; ++hw) { const float v = xs[tid * 49 + hw]; ss = fmaf(v, v, ss); }
        inv_c[tid] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }
    for (int hw = warp; hw < 49; hw += 16) {   // column norms over C
        float ss = 0.f;
        for (int c = lane; c < 512; c += 32) { const float v = xs[c * 49 + hw]; ss = fmaf(v, v, ss); }
        ss = warp_sum_r(ss);
        if (lane == 0) inv_s[hw] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }
    __syncthreads();

    // The three phases below are bound by shared-memory instruction throughput (every FMA wants an operand from
    // smem), so each thread register-tiles its outputs: 4 Gram columns per row value, 7 pixels per weight value.

    // spatial self-similarity Gram (49x49 over C): work item = (7x7 output block, one eighth of the channels) on
    // 392 threads — 49 FMAs per 14 shared-memory loads (the 4-column version spent half of the kernel here on LSU
    // wavefronts); the eight channel slices go to their own shared-memory planes and are summed in a fixed order
    // (deterministic: the embedding of an image must not depend on the run or on its batch)
    if (tid < 392) {
        const int item = tid % 49, split = tid / 49;
        const int bi = (item / 7) * 7, bj = (item % 7) * 7;
        float acc[7][7];
#pragma unroll
        for (int a = 0; a < 7; ++a)
#pragma unroll
            for (int q = 0; q < 7; ++q) acc[a][q] = 0.f;
#pragma unroll 2
        for (int c = split * 64; c < split * 64 + 64; ++c) {
            const float* xr = xs + c * 49;
            float va[7], vb[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) { va[q] = xr[bi + q]; vb[q] = xr[bj + q]; }
#pragma unroll
            for (int a = 0; a < 7; ++a)
#pragma unroll
                for (int q = 0; q < 7; ++q) acc[a][q] = fmaf(va[a], vb[q], acc[a][q]);
        }
        float* gp = Gpart + split * 2401;
#pragma unroll
        for (int a = 0; a < 7; ++a)
#pragma unroll
            for (int q = 0; q < 7; ++q) gp[(bi + a) * 49 + bj + q] = acc[a][q];
    }
    __syncthreads();
    for (int o = tid; o < 49 * 49; o += 512) {
        const int i = o / 49, j = o - i * 49;
        float g = 0.f;
#pragma unroll
        for (int sp = 0; sp < 8; ++sp) g += Gpart[sp * 2401 + o];
        Gs[o] = g * inv_s[i] * inv_s[j];
    }
    // T[hw][j] = sum_c Xh[c][hw] * W0b[j][c]: work item = (7 consecutive pixels, j, half of the channel range);
    // the two halves are combined with shared-memory atomics (T zero-initialised first)
    for (int o = tid; o < 49 * 32; o += 512) T[o] = 0.f;
    __syncthreads();
    if (tid < 448) {
        const int j = tid & 31, hg = (tid >> 5) % 7, half = tid / 224;
        float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int c_lo = half * 256;
#pragma unroll 4
        for (int c = c_lo; c < c_lo + 256; ++c) {
            const float wj = __ldg(p.w0bT + c * 32 + j) * inv_c[c];
            const float* xr = xs + c * 49 + hg * 7;
#pragma unroll
            for (int q = 0; q < 7; ++q) acc[q] = fmaf(xr[q], wj, acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 7; ++q) atomicAdd(&T[(hg * 7 + q) * 32 + j], acc[q]);
    }
    __syncthreads();

    // ---- layout fan-out of X (bf16) ----
    {
        const int c2 = (tid & 255) * 2;            // channel pair
        const int half = tid >> 8;                 // two row streams
        for (int hw = half; hw < 49; hw += 2) {
            const uint32_t v = pack_bf16x2(xs[c2 * 49 + hw], xs[(c2 + 1) * 49 + hw]);
            *reinterpret_cast<uint32_t*>(p.xt + ((long long)n * 128 + hw) * 512 + c2) = v;
        }
        for (int pos = half; pos < 81; pos += 2) {
            const int hp = pos / 9, wp = pos - hp * 9;
            const int hw = reflect_src(hp) * 7 + reflect_src(wp);
            const uint32_t v = pack_bf16x2(xs[c2 * 49 + hw], xs[(c2 + 1) * 49 + hw]);
            *reinterpret_cast<uint32_t*>(p.s0 + ((long long)n * 81 + pos) * 576 + c2) = v;
            *reinterpret_cast<uint32_t*>(p.cm + ((long long)n * 81 + pos) * 1536 + 1024 + c2) = v;
        }
    }
    // ss_space as 49 extra channels of the Conv4Space input: channel i at pixel j holds Gram[i][j]
    for (int o = tid; o < 81 * 49; o += 512) {
        const int pos = o / 49, i = o - pos * 49;
        const int hp = pos / 9, wp = pos - hp * 9;
        const int j = reflect_src(hp) * 7 + reflect_src(wp);
        p.s0[((long long)n * 81 + pos) * 576 + 512 + i] = __float2bfloat16(Gs[i * 49 + j]);
    }
    if (p.ss_space)
        for (int o = tid; o < 49 * 49; o += 512) p.ss_space[(long long)n * 2401 + o] = Gs[o];

    // ---- thin channel-rectifier chain, one thread per channel row ----
    {
        const int c = tid;
        float h[32], g[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { h[j] = misc[j]; g[j] = 0.f; }
        for (int hw = 0; hw < 49; ++hw) {
            const float xv = xs[c * 49 + hw];
            const float4* wa = reinterpret_cast<const float4*>(W0a + hw * 32);
            const float4* tt = reinterpret_cast<const float4*>(T + hw * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a4 = wa[q], t4 = tt[q];
                h[q * 4 + 0] = fmaf(xv, a4.x, h[q * 4 + 0]); h[q * 4 + 1] = fmaf(xv, a4.y, h[q * 4 + 1]);
                h[q * 4 + 2] = fmaf(xv, a4.z, h[q * 4 + 2]); h[q * 4 + 3] = fmaf(xv, a4.w, h[q * 4 + 3]);
                g[q * 4 + 0] = fmaf(xv, t4.x, g[q * 4 + 0]); g[q * 4 + 1] = fmaf(xv, t4.y, g[q * 4 + 1]);
                g[q * 4 + 2] = fmaf(xv, t4.z, g[q * 4 + 2]); g[q * 4 + 3] = fmaf(xv, t4.w, g[q * 4 + 3]);
            }
        }
        const float ic = inv_c[c];
        const float s1 = p.slope1[c], s4 = p.slope4[c], s7 = p.slope7[c];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float v = fmaf(ic, g[j], h[j]);
            h[j] = v > 0.f ? v : v * s1;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float a = misc[32 + j];
            const float4* ar = reinterpret_cast<const float4*>(A1s + j * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a4 = ar[q];
                a = fmaf(a4.x, h[q * 4 + 0], a); a = fmaf(a4.y, h[q * 4 + 1], a);
                a = fmaf(a4.z, h[q * 4 + 2], a); a = fmaf(a4.w, h[q * 4 + 3], a);
            }
            g[j] = a > 0.f ? a : a * s4;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float a = misc[64 + j];
            const float4* ar = reinterpret_cast<const float4*>(A2s + j * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a4 = ar[q];
                a = fmaf(a4.x, g[q * 4 + 0], a); a = fmaf(a4.y, g[q * 4 + 1], a);
                a = fmaf(a4.z, g[q * 4 + 2], a); a = fmaf(a4.w, g[q * 4 + 3], a);
            }
            h[j] = a > 0.f ? a : a * s7;
        }
        uint4* o = reinterpret_cast<uint4*>(p.h5 + ((long long)n * 512 + c) * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            o[q] = make_uint4(pack_bf16x2(h[q * 8 + 0], h[q * 8 + 1]), pack_bf16x2(h[q * 8 + 2], h[q * 8 + 3]),
                              pack_bf16x2(h[q * 8 + 4], h[q * 8 + 5]), pack_bf16x2(h[q * 8 + 6], h[q * 8 + 7]));
#pragma unroll
        for (int q = 4; q < 8; ++q) o[q] = make_uint4(0, 0, 0, 0);
    }
}

// ----------------------------------------------------------------------------------------------
// Warp-MMA version of recnet_prep_kernel (same inputs, outputs and algebra). The SIMT kernel above executes ~300 k warp
// instructions per sample (49x49x512 Gram, 49x32x512 T, 512x49x64 first chain layer, two 32x32 maps per row: every FMA
// with a shared-memory operand) at 40 % issue efficiency: 320-360 us for 512 samples, the largest non-GEMM kernel of the
// eval step. Here X is staged ONCE as bf16 in the pixel-major layout XT [64 pixels][512 channels] — the precision every
// other consumer of X in the eval path has (s0 / cm / xt ARE this matrix) — and
//   Gram  G = XT XT^T          (M 64 x N 56 x K 512)   m16n8k16 bf16, fp32 accumulate        warps 0-7
//   T     = XT (W0b^T / |x_c|) (M 64 x N 32 x K 512)   m16n8k16 bf16                          warps 8-15
//   H     = X [W0a | T]        (M 512 x N 64 x K 64)   m16n8k16 bf16, 32 rows per warp
//   the two composed 32x32 maps                        m16n8k8 tf32, A straight from the previous accumulators
// run on the tensor cores through ldmatrix fragments; row / column norms come from the staged values (so the Gram of
// the staged matrix has a unit diagonal), the fan-out copies 4-byte channel pairs from XT without conversions.
// The pad rows 49..63 of XT are zero: they are the zero K / M / N padding of all four contractions.
// ----------------------------------------------------------------------------------------------
constexpr int PM_XP = 520;     // XT pitch (bf16): 1040 B = 260 words = 4 mod 32 -> ldmatrix rows hit disjoint bank groups
constexpr int PM_BP = 40;      // W0b^T / |x_c| pitch (bf16), [512][32]
constexpr int PM_WP = 72;      // [W0a | T] pitch (bf16), [64][64]
constexpr int PM_AP = 36;      // composed 32x32 maps pitch (fp32)

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16_k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_tf32_k8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// One composed 32x32 map + PReLU on the accumulator fragments of a warp's two 16-row tiles: out = prelu(in A^T + cvec).
// The tf32 A fragment of k-step s is the C fragment of n-tile s with the K slots permuted (slot t <-> column 2t,
// slot t + 4 <-> column 2t + 1); the B fragment reads the map with the same permutation, so no data moves between lanes.
__device__ __forceinline__ void chain_map_tf32(float (&h)[2][4][4], const float* As, const float* cvec, const float (&slope)[2][2],
                                               int g, int t) {
    float o[2][4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float c0 = cvec[8 * q + 2 * t], c1 = cvec[8 * q + 2 * t + 1];
#pragma unroll
        for (int m = 0; m < 2; ++m) { o[m][q][0] = c0; o[m][q][1] = c1; o[m][q][2] = c0; o[m][q][3] = c1; }
    }
#pragma unroll
    for (int s2 = 0; s2 < 4; ++s2) {
        uint32_t a[2][4];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            a[m][0] = to_tf32(h[m][s2][0]); a[m][1] = to_tf32(h[m][s2][2]);
            a[m][2] = to_tf32(h[m][s2][1]); a[m][3] = to_tf32(h[m][s2][3]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 bv = *reinterpret_cast<const float2*>(As + (8 * q + g) * PM_AP + 8 * s2 + 2 * t);
            const uint32_t b0 = to_tf32(bv.x), b1 = to_tf32(bv.y);
            mma_tf32_k8(o[0][q], a[0], b0, b1);
            mma_tf32_k8(o[1][q], a[1], b0, b1);
        }
    }
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float v = o[m][q][e];
                h[m][q][e] = v > 0.f ? v : v * slope[m][e >> 1];
            }
}

__global__ void __launch_bounds__(512, 1) recnet_prep_mma_kernel(const PrepParams p) {
    extern __shared__ __align__(16) uint8_t pm_smem[];
    __nv_bfloat16* xt = reinterpret_cast<__nv_bfloat16*>(pm_smem);                  // [64][PM_XP]
    __nv_bfloat16* Bs = xt + 64 * PM_XP;                                            // [512][PM_BP]
    __nv_bfloat16* Wc = Bs + 512 * PM_BP;                                           // [64][PM_WP]
    float* Gs = reinterpret_cast<float*>(Wc + 64 * PM_WP);                          // [49][49] (+3)
    float* A1s = Gs + 2404;                                                         // [32][PM_AP]
    float* A2s = A1s + 32 * PM_AP;
    float* inv_c = A2s + 32 * PM_AP;                                                // [512]
    float* inv_s = inv_c + 512;                                                     // [64]
    float* misc = inv_s + 64;                                                       // b0, c1, c2 [3][32]
    const int n = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const float* x = p.x + (long long)n * 512 * 49;

    // ---- stage X as bf16 XT[hw][c] (coalesced reads), zero pad rows, small operands ----
    for (int i = tid; i < 512 * 49; i += 512) {
        const int c = i / 49, hw = i - c * 49;
        xt[hw * PM_XP + c] = __float2bfloat16_rn(__ldg(x + i));
    }
    for (int i = tid; i < 15 * (PM_XP / 2); i += 512) reinterpret_cast<uint32_t*>(xt + 49 * PM_XP)[i] = 0u;   // rows 49..63
    for (int i = tid; i < 64 * 32; i += 512) {           // W0a half of [W0a | T]; rows >= 49 zero
        const int hw = i >> 5, j = i & 31;
        Wc[hw * PM_WP + j] = __float2bfloat16_rn(hw < 49 ? p.w0aT[hw * 32 + j] : 0.f);
    }
    for (int i = tid; i < 1024; i += 512) {
        A1s[(i >> 5) * PM_AP + (i & 31)] = p.A1[i];
        A2s[(i >> 5) * PM_AP + (i & 31)] = p.A2[i];
    }
    if (tid < 32) { misc[tid] = p.b0[tid]; misc[32 + tid] = p.c1[tid]; misc[64 + tid] = p.c2[tid]; }
    __syncthreads();

    // ---- norms of the staged matrix: rows of X (over hw, F.normalize(dim=2) of (N,C,HW)) and pixel columns (over C) ----
    {
        float ss = 0.f;
        for (int hw = 0; hw < 49; ++hw) { const float v = __bfloat162float(xt[hw * PM_XP + tid]); ss = fmaf(v, v, ss); }
        inv_c[tid] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }
    for (int hw = warp; hw < 64; hw += 16) {
        float ss = 0.f;
        const uint32_t* row = reinterpret_cast<const uint32_t*>(xt + hw * PM_XP);
        for (int k = lane; k < 256; k += 32) { const uint32_t u = row[k]; const float a = bf16lo(u), b = bf16hi(u); ss = fmaf(a, a, fmaf(b, b, ss)); }
        ss = warp_sum_r(ss);
        if (lane == 0) inv_s[hw] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }
    __syncthreads();
    for (int i = tid; i < 512 * 32; i += 512) {           // B operand of T: W0b^T[c][j] / |x_c|
        const int c = i >> 5, j = i & 31;
        Bs[c * PM_BP + j] = __float2bfloat16_rn(__ldg(p.w0bT + i) * inv_c[c]);
    }
    __syncthreads();

    const uint32_t xt_s = smem_u32(xt), bs_s = smem_u32(Bs), wc_s = smem_u32(Wc);
    if (warp < 8) {
        // ---- spatial Gram: rows i = 16 mt .. +15, columns j = 32 ng .. +31 (n-tiles 4 ng .. 4 ng + 3) ----
        const int mt = warp >> 1, ng = warp & 1;
        float acc[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[q][e] = 0.f;
        // A[m = hw][k = c] = XT[hw][c]: non-transposed; B[k = c][n = hw'] = XT[hw'][c]: memory [n][k], non-transposed
        const uint32_t a_addr = xt_s + (uint32_t)(((16 * mt + (lane & 7) + ((lane >> 3) & 1) * 8) * PM_XP + (lane >> 4) * 8) * 2);
        const uint32_t b_addr = xt_s + (uint32_t)(((32 * ng + (lane & 7) + (lane >> 4) * 8) * PM_XP + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll 4
        for (int ks = 0; ks < 32; ++ks) {
            uint32_t a[4], b01[4], b23[4];
            ldsm_x4(a, a_addr + ks * 32);
            ldsm_x4(b01, b_addr + ks * 32);
            ldsm_x4(b23, b_addr + 16 * PM_XP * 2 + ks * 32);
            mma_bf16_k16(acc[0], a, b01[0], b01[1]);
            mma_bf16_k16(acc[1], a, b01[2], b01[3]);
            mma_bf16_k16(acc[2], a, b23[0], b23[1]);
            mma_bf16_k16(acc[3], a, b23[2], b23[3]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 16 * mt + g + (e >> 1) * 8, j = 32 * ng + 8 * q + 2 * t + (e & 1);
                if (i < 49 && j < 49) Gs[i * 49 + j] = acc[q][e] * inv_s[i] * inv_s[j];
            }
    } else {
        // ---- T[hw][j] = sum_c XT[hw][c] * Bs[c][j]: rows 16 mt .. +15, columns 16 nh .. +15 ----
        const int w8 = warp - 8, mt = w8 >> 1, nh = w8 & 1;
        float acc[2][4];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[q][e] = 0.f;
        const uint32_t a_addr = xt_s + (uint32_t)(((16 * mt + (lane & 7) + ((lane >> 3) & 1) * 8) * PM_XP + (lane >> 4) * 8) * 2);
        // B[k = c][n = j] = Bs[c][j]: memory [k][n] -> transposed load; matrices (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15)
        const uint32_t b_addr = bs_s + (uint32_t)((((lane & 7) + ((lane >> 3) & 1) * 8) * PM_BP + 16 * nh + (lane >> 4) * 8) * 2);
#pragma unroll 4
        for (int ks = 0; ks < 32; ++ks) {
            uint32_t a[4], b[4];
            ldsm_x4(a, a_addr + ks * 32);
            ldsm_x4_t(b, b_addr + ks * 16 * PM_BP * 2);
            mma_bf16_k16(acc[0], a, b[0], b[1]);
            mma_bf16_k16(acc[1], a, b[2], b[3]);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
                const int hw = 16 * mt + g + (e >> 1) * 8, j = 16 * nh + 8 * q + 2 * t;
                // rows >= 49 are exact zeros (zero rows of XT)
                *reinterpret_cast<uint32_t*>(Wc + hw * PM_WP + 32 + j) = pack_bf16x2(acc[q][e], acc[q][e + 1]);
            }
    }
    __syncthreads();

    // ---- layout fan-out of X: 4-byte channel pairs straight from XT ----
    {
        const int cp = tid & 255;                  // channel pair (2 cp, 2 cp + 1)
        const int half = tid >> 8;                 // two row streams
        for (int hw = half; hw < 49; hw += 2) {
            const uint32_t v = reinterpret_cast<const uint32_t*>(xt + hw * PM_XP)[cp];
            *reinterpret_cast<uint32_t*>(p.xt + ((long long)n * 128 + hw) * 512 + 2 * cp) = v;
        }
        for (int pos = half; pos < 81; pos += 2) {
            const int hp = pos / 9, wp = pos - hp * 9;
            const int hw = reflect_src(hp) * 7 + reflect_src(wp);
            const uint32_t v = reinterpret_cast<const uint32_t*>(xt + hw * PM_XP)[cp];
            *reinterpret_cast<uint32_t*>(p.s0 + ((long long)n * 81 + pos) * 576 + 2 * cp) = v;
            *reinterpret_cast<uint32_t*>(p.cm + ((long long)n * 81 + pos) * 1536 + 1024 + 2 * cp) = v;
        }
    }
    // ss_space as 49 extra channels of the Conv4Space input: channel i at pixel j holds Gram[i][j]
    for (int o = tid; o < 81 * 49; o += 512) {
        const int pos = o / 49, i = o - pos * 49;
        const int hp = pos / 9, wp = pos - hp * 9;
        const int j = reflect_src(hp) * 7 + reflect_src(wp);
        p.s0[((long long)n * 81 + pos) * 576 + 512 + i] = __float2bfloat16(Gs[i * 49 + j]);
    }
    if (p.ss_space)
        for (int o = tid; o < 49 * 49; o += 512) p.ss_space[(long long)n * 2401 + o] = Gs[o];

    // ---- channel-rectifier chain on the warp's 32 channel rows: H = X [W0a | T] (K = 64 pixels, 49 real) ----
    {
        const int m_base = 32 * warp;
        float acc[2][8][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][q][e] = 0.f;
        // A[m = c][k = hw] = XT[hw][c]: memory [k][m] -> transposed load; matrices (m 0-7, k 0-7), (m 8-15, k 0-7), (m 0-7, k 8-15), (m 8-15, k 8-15)
        const uint32_t a_addr = xt_s + (uint32_t)((((lane & 7) + (lane >> 4) * 8) * PM_XP + m_base + ((lane >> 3) & 1) * 8) * 2);
        // B[k = hw][n = j] = Wc[hw][j]: memory [k][n] -> transposed load, two n-tiles per instruction
        const uint32_t b_addr = wc_s + (uint32_t)((((lane & 7) + ((lane >> 3) & 1) * 8) * PM_WP + (lane >> 4) * 8) * 2);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a0[4], a1[4];
            ldsm_x4_t(a0, a_addr + ks * 16 * PM_XP * 2);
            ldsm_x4_t(a1, a_addr + ks * 16 * PM_XP * 2 + 32);
#pragma unroll
            for (int pq = 0; pq < 4; ++pq) {
                uint32_t b[4];
                ldsm_x4_t(b, b_addr + ks * 16 * PM_WP * 2 + pq * 32);
                mma_bf16_k16(acc[0][2 * pq], a0, b[0], b[1]);
                mma_bf16_k16(acc[0][2 * pq + 1], a0, b[2], b[3]);
                mma_bf16_k16(acc[1][2 * pq], a1, b[0], b[1]);
                mma_bf16_k16(acc[1][2 * pq + 1], a1, b[2], b[3]);
            }
        }
        // h = X W0a^T + b0 + (X T) / |x_c|, PReLU over the channel rows (recnet.py:373-374)
        float h[2][4][4];
        float sl1[2][2], sl4[2][2], sl7[2][2];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int c = m_base + 16 * m + g + 8 * r;
                sl1[m][r] = p.slope1[c]; sl4[m][r] = p.slope4[c]; sl7[m][r] = p.slope7[c];
                const float ic = inv_c[c];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int e2 = 0; e2 < 2; ++e2) {
                        const int e = 2 * r + e2;
                        const float v = fmaf(ic, acc[m][q + 4][e], acc[m][q][e] + misc[8 * q + 2 * t + e2]);
                        h[m][q][e] = v > 0.f ? v : v * sl1[m][r];
                    }
            }
        chain_map_tf32(h, A1s, misc + 32, sl4, g, t);
        chain_map_tf32(h, A2s, misc + 64, sl7, g, t);
        // H5 rows: 32 real columns as bf16 pairs, columns 32..63 zero
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int c = m_base + 16 * m + g + 8 * r;
                __nv_bfloat16* o = p.h5 + ((long long)n * 512 + c) * 64;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint32_t*>(o + 8 * q + 2 * t) = pack_bf16x2(h[m][q][2 * r], h[m][q][2 * r + 1]);
                *reinterpret_cast<uint4*>(o + 32 + 8 * t) = make_uint4(0, 0, 0, 0);
            }
    }
}

static int g_prep_mma = 1;      // ffr_debug_set_prep_mma(0): the SIMT kernel (A/B runs, tests)
void set_prep_mma(int on) { g_prep_mma = on; }

int recnet_prep_launch(const PrepParams& p, int n, cudaStream_t stream) {
    const int smem = (512 * 49 + 512 + 64 + 49 * 49 + 3 + 49 * 32 * 2 + 2048 + 96 + 8 * 2401) * (int)sizeof(float);
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(recnet_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    if (g_prep_mma) {
        const int smem2 = (64 * PM_XP + 512 * PM_BP + 64 * PM_WP) * 2 + (2404 + 2 * 32 * PM_AP + 512 + 64 + 96) * (int)sizeof(float);
        static bool attr2 = false;
        if (!attr2) {
            FFR_CUDA(cudaFuncSetAttribute(recnet_prep_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
            attr2 = true;
        }
        recnet_prep_mma_kernel<<<n, 512, smem2, stream>>>(p);
        return launch_status("recnet_prep_mma_kernel");
    }
    recnet_prep_kernel<<<n, 512, smem, stream>>>(p);
    return launch_status("recnet_prep_kernel");
}

// ----------------------------------------------------------------------------------------------
// feat_space = X (512x49) @ M_space (49x49)  (recnet.py:409), written as bf16 into slot [0,512) of the Conv4Merge
// input (H9 with mirrors) and optionally as fp32 NCHW. mspace: fp32 rows of the H9 grid, [n*81][64]; row = pixel j,
// column = channel i holds M_space[n, i, j] (the conv output is NHWC; recnet.py:405 views it as (N, HW, HW)).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) feat_space_kernel(const float* __restrict__ x, const float* __restrict__ mspace,
                                                         __nv_bfloat16* __restrict__ cm, float* __restrict__ out_nchw) {
    __shared__ float Ms[49 * 49];     // Ms[i*49 + j] = M_space[n,i,j]
    extern __shared__ __align__(16) float fs_xs[];   // [512][49]: the sample's X, staged with coalesced 16-byte loads
    const int n = blockIdx.x, tid = threadIdx.x;
    for (int o = tid; o < 49 * 49; o += 256) {
        const int j = o / 49, i = o - j * 49;
        const int pos = (j / 7 + 1) * 9 + (j % 7 + 1);
        Ms[i * 49 + j] = mspace[((long long)n * 81 + pos) * 64 + i];
    }
    {   // (a thread's two rows are 196 bytes apart from its neighbour's: read straight from global memory every one of
        // the 98 loads of a warp touched 32 lines)
        const float4* src = reinterpret_cast<const float4*>(x + (long long)n * 512 * 49);
        float4* dst = reinterpret_cast<float4*>(fs_xs);
        for (int o = tid; o < 512 * 49 / 4; o += 256) dst[o] = __ldg(src + o);
    }
    __syncthreads();
    const int c2 = tid * 2;
    float xa[49], xb[49];
    const float* xr = fs_xs + c2 * 49;
#pragma unroll
    for (int i = 0; i < 49; ++i) { xa[i] = xr[i]; xb[i] = xr[49 + i]; }
    for (int j = 0; j < 49; ++j) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < 49; ++i) {
            const float m = Ms[i * 49 + j];
            a = fmaf(xa[i], m, a);
            b = fmaf(xb[i], m, b);
        }
        if (out_nchw) {
            out_nchw[((long long)n * 512 + c2) * 49 + j] = a;
            out_nchw[((long long)n * 512 + c2 + 1) * 49 + j] = b;
        }
        const uint32_t v = pack_bf16x2(a, b);
        const int h = j / 7, w = j - h * 7;
        const int mh = (h == 1) ? -2 : ((h == 5) ? 2 : 0);
        const int mw = (w == 1) ? -2 : ((w == 5) ? 2 : 0);
        const long long base = (long long)n * 81 + (h + 1) * 9 + (w + 1);
        *reinterpret_cast<uint32_t*>(cm + base * 1536 + c2) = v;
        if (mh) *reinterpret_cast<uint32_t*>(cm + (base + mh * 9) * 1536 + c2) = v;
        if (mw) *reinterpret_cast<uint32_t*>(cm + (base + mw) * 1536 + c2) = v;
        if (mh && mw) *reinterpret_cast<uint32_t*>(cm + (base + mh * 9 + mw) * 1536 + c2) = v;
    }
}

int feat_space_launch(const float* x, const float* mspace, void* cm, float* out_nchw, int n, cudaStream_t stream) {
    const int smem = 512 * 49 * (int)sizeof(float);
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(feat_space_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    feat_space_kernel<<<n, 256, smem, stream>>>(x, mspace, reinterpret_cast<__nv_bfloat16*>(cm), out_nchw);
    return launch_status("feat_space_kernel");
}

// Warp-MMA version of feat_space_kernel, fed by the XT matrix recnet_prep already wrote (bf16 [n*128][512], rows = pixels):
//   FS^T[j][c] = sum_i M_space[i][j] * XT[i][c]      (M 64 x N 512 x K 64; 49 real pixels, pad rows / columns zero)
// m16n8k16 bf16 MMAs through ldmatrix (both operands are stored k-major: transposed loads); a thread's accumulators are
// channel PAIRS of one pixel, exactly the 4-byte stores of the H9 fan-out. 8 warps x 64 channels, two passes of two
// 16-pixel tiles. M_space is rounded to bf16 for the contraction (its output is stored as bf16).
constexpr int FS_MP = 72;      // M_space pitch (bf16), [64][64]
__global__ void __launch_bounds__(256) feat_space_mma_kernel(const __nv_bfloat16* __restrict__ xt_g, const float* __restrict__ mspace,
                                                             __nv_bfloat16* __restrict__ cm, float* __restrict__ out_nchw) {
    extern __shared__ __align__(16) uint8_t fsm_smem[];
    __nv_bfloat16* xts = reinterpret_cast<__nv_bfloat16*>(fsm_smem);      // [64][PM_XP]
    __nv_bfloat16* Ms = xts + 64 * PM_XP;                                  // [64][FS_MP]: Ms[i][j] = M_space[n, i, j]
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    {
        const uint4* src = reinterpret_cast<const uint4*>(xt_g + (long long)n * 128 * 512);
        for (int o = tid; o < 49 * 64; o += 256) {
            const int hw = o >> 6, q = o & 63;
            reinterpret_cast<uint4*>(xts + hw * PM_XP)[q] = __ldg(src + hw * 64 + q);
        }
        for (int o = tid; o < 15 * (PM_XP / 2); o += 256) reinterpret_cast<uint32_t*>(xts + 49 * PM_XP)[o] = 0u;
        for (int o = tid; o < 64 * 64; o += 256) {
            const int j = o >> 6, i = o & 63;
            float v = 0.f;
            if (i < 49 && j < 49) v = mspace[((long long)n * 81 + (j / 7 + 1) * 9 + (j % 7 + 1)) * 64 + i];
            Ms[i * FS_MP + j] = __float2bfloat16_rn(v);
        }
    }
    __syncthreads();
    const uint32_t xt_s = smem_u32(xts), ms_s = smem_u32(Ms);
    const int n_base = 64 * warp;
    // A[m = j][k = i] = Ms[i][j]: memory [k][m] -> transposed; B[k = i][n = c] = XT[i][c]: memory [k][n] -> transposed
    const uint32_t a_addr = ms_s + (uint32_t)((((lane & 7) + (lane >> 4) * 8) * FS_MP + ((lane >> 3) & 1) * 8) * 2);
    const uint32_t b_addr = xt_s + (uint32_t)((((lane & 7) + ((lane >> 3) & 1) * 8) * PM_XP + n_base + (lane >> 4) * 8) * 2);
#pragma unroll 1
    for (int mp = 0; mp < 2; ++mp) {
        float acc[2][8][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[m][q][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t a0[4], a1[4];
            ldsm_x4_t(a0, a_addr + (ks * 16 * FS_MP + 32 * mp) * 2);
            ldsm_x4_t(a1, a_addr + (ks * 16 * FS_MP + 32 * mp + 16) * 2);
#pragma unroll
            for (int pq = 0; pq < 4; ++pq) {
                uint32_t b[4];
                ldsm_x4_t(b, b_addr + (ks * 16 * PM_XP + 16 * pq) * 2);
                mma_bf16_k16(acc[0][2 * pq], a0, b[0], b[1]);
                mma_bf16_k16(acc[0][2 * pq + 1], a0, b[2], b[3]);
                mma_bf16_k16(acc[1][2 * pq], a1, b[0], b[1]);
                mma_bf16_k16(acc[1][2 * pq + 1], a1, b[2], b[3]);
            }
        }
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int j = 16 * (2 * mp + m) + g + 8 * r;
                if (j >= 49) continue;
                const int h = j / 7, w = j - h * 7;
                const int mh = (h == 1) ? -2 : ((h == 5) ? 2 : 0);
                const int mw = (w == 1) ? -2 : ((w == 5) ? 2 : 0);
                const long long base = (long long)n * 81 + (h + 1) * 9 + (w + 1);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int c = n_base + 8 * q + 2 * t;
                    const float a = acc[m][q][2 * r], b = acc[m][q][2 * r + 1];
                    if (out_nchw) {
                        out_nchw[((long long)n * 512 + c) * 49 + j] = a;
                        out_nchw[((long long)n * 512 + c + 1) * 49 + j] = b;
                    }
                    const uint32_t v = pack_bf16x2(a, b);
                    *reinterpret_cast<uint32_t*>(cm + base * 1536 + c) = v;
                    if (mh) *reinterpret_cast<uint32_t*>(cm + (base + mh * 9) * 1536 + c) = v;
                    if (mw) *reinterpret_cast<uint32_t*>(cm + (base + mw) * 1536 + c) = v;
                    if (mh && mw) *reinterpret_cast<uint32_t*>(cm + (base + mh * 9 + mw) * 1536 + c) = v;
                }
            }
    }
}

int feat_space_xt_launch(const void* xt, const float* mspace, void* cm, float* out_nchw, int n, cudaStream_t stream) {
    const int smem = (64 * PM_XP + 64 * FS_MP) * 2;
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(feat_space_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    feat_space_mma_kernel<<<n, 256, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(xt), mspace,
                                                    reinterpret_cast<__nv_bfloat16*>(cm), out_nchw);
    return launch_status("feat_space_mma_kernel");
}

// ----------------------------------------------------------------------------------------------
// Rows of a haloed/flat grid -> fp32 NCHW (S x S valid pixels at row (h+off)*G + (w+off)), from bf16 or fp32 rows,
// optional per-channel affine. grid = (C/64, n).
// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) rows_to_nchw_kernel(const T* __restrict__ rows, int ld, int ch0,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift, float* __restrict__ y,
                                                           int S, int G, int off, int rows_per_img, int C) {
    extern __shared__ float tile[];  // [S*S][65]
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const int P = S * S;
    for (int i = threadIdx.x; i < P * 64; i += blockDim.x) {
        const int c = i & 63, pix = i >> 6;
        const int ph = pix / S, pw = pix - ph * S;
        const long long row = (long long)n * rows_per_img + (ph + off) * G + (pw + off);
        float v = static_cast<float>(rows[row * ld + ch0 + c0 + c]);
        if (scale) v = v * scale[c0 + c] + shift[c0 + c];
        tile[pix * 65 + c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * 64; i += blockDim.x) {
        const int pix = i % P, c = i / P;
        y[((long long)n * C + c0 + c) * P + pix] = tile[pix * 65 + c];
    }
}

int rows_to_nchw_launch(const void* rows, int is_f32, int ld, int ch0, const float* scale, const float* shift, float* y,
                        int n, int S, int G, int off, int rows_per_img, int C, cudaStream_t stream) {
    FFR_CHECK_ARG(C % 64 == 0, "rows_to_nchw: C=%d", C);
    dim3 grid(C / 64, n);
    const size_t smem = (size_t)S * S * 65 * sizeof(float);
    FFR_CHECK_ARG(smem <= 48 * 1024, "rows_to_nchw: map %dx%d too large", S, S);
    if (is_f32)
        rows_to_nchw_kernel<float><<<grid, 256, smem, stream>>>(reinterpret_cast<const float*>(rows), ld, ch0, scale,
                                                                shift, y, S, G, off, rows_per_img, C);
    else
        rows_to_nchw_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(rows),
                                                                        ld, ch0, scale, shift, y, S, G, off,
                                                                        rows_per_img, C);
    return launch_status("rows_to_nchw_kernel");
}

// out[i] = in[i] * scale   (AvgPool2d(7) finish: pooled sums -> means, recnet.py:423)
__global__ void scale_f32_kernel(const float* __restrict__ in, float* __restrict__ out, long long count, float scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = in[i] * scale;
}

int scale_f32_launch(const float* in, float* out, long long count, float scale, cudaStream_t stream) {
    scale_f32_kernel<<<(int)((count + 255) / 256), 256, 0, stream>>>(in, out, count, scale);
    return launch_status("scale_f32_kernel");
}

}  // namespace ffr

// ------------------------------------------------------------------------------------------------------------
// selfSimilarity (recnet.py:226-236) as a standalone forward op in fp32 (used for the no-grad targets of the
// training loss, models/trainer.py:157): x (n,512,7,7) ->
//   ss_space  (n,49,49)  : cosine between pixel columns  (normalised over C,  eps 1e-12)
//   ss_channel(n,512,512): cosine between channel rows   (normalised over HW, eps 1e-12)
// ------------------------------------------------------------------------------------------------------------
namespace ffr {

// one CTA per sample; x in shared memory; 4 Gram columns per thread (as in recnet_prep_kernel)
__global__ void __launch_bounds__(512, 1) selfsim_space_kernel(const float* __restrict__ x, float* __restrict__ ss) {
    extern __shared__ float sm[];
    float* xs = sm;                 // [512][49]
    float* inv_s = xs + 512 * 49;   // [64]
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* xp = x + (long long)n * 512 * 49;
    for (int i = tid; i < 512 * 49; i += 512) xs[i] = xp[i];
    __syncthreads();
    for (int hw = warp; hw < 49; hw += 16) {
        float s = 0.f;
        for (int c = lane; c < 512; c += 32) { const float v = xs[c * 49 + hw]; s = fmaf(v, v, s); }
        s = warp_sum_r(s);
        if (lane == 0) inv_s[hw] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    }
    __syncthreads();
    for (int o = tid; o < 49 * 13; o += 512) {
        const int i = o / 13, j0 = (o - i * 13) * 4;
        const int j1 = min(j0 + 1, 48), j2 = min(j0 + 2, 48), j3 = min(j0 + 3, 48);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
        for (int c = 0; c < 512; ++c) {
            const float* xr = xs + c * 49;
            const float xi = xr[i];
            a0 = fmaf(xi, xr[j0], a0); a1 = fmaf(xi, xr[j1], a1);
            a2 = fmaf(xi, xr[j2], a2); a3 = fmaf(xi, xr[j3], a3);
        }
        float* o_ = ss + (long long)n * 2401 + i * 49;
        const float si = inv_s[i];
        o_[j0] = a0 * si * inv_s[j0];
        if (j0 + 1 < 49) o_[j0 + 1] = a1 * si * inv_s[j1];
        if (j0 + 2 < 49) o_[j0 + 2] = a2 * si * inv_s[j2];
        if (j0 + 3 < 49) o_[j0 + 3] = a3 * si * inv_s[j3];
    }
}

// grid (16, n): CTA (tile, sample) computes a 128x128 tile of the 512x512 channel Gram; 8x8 outputs per thread
__global__ void __launch_bounds__(256) selfsim_channel_kernel(const float* __restrict__ x, float* __restrict__ ss) {
    extern __shared__ __align__(16) float ssm[];
    float (*a_s)[132] = reinterpret_cast<float (*)[132]>(ssm);              // [k][row], rows of the tile, normalised
    float (*b_s)[132] = reinterpret_cast<float (*)[132]>(ssm + 49 * 132);   // pitch 132 keeps float4 alignment
    const int n = blockIdx.y, tr = blockIdx.x >> 2, tc = blockIdx.x & 3;
    const int tid = threadIdx.x;
    const float* xp = x + (long long)n * 512 * 49;
    if (tid < 256) {
        const int which = tid >> 7, r = tid & 127;               // 128 threads load+normalise A rows, 128 B rows
        const float* row = xp + (long long)((which ? tc : tr) * 128 + r) * 49;
        float v[49], s = 0.f;
#pragma unroll
        for (int k = 0; k < 49; ++k) { v[k] = __ldg(row + k); s = fmaf(v[k], v[k], s); }
        const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
#pragma unroll
        for (int k = 0; k < 49; ++k) (which ? b_s : a_s)[k][r] = v[k] * inv;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;                       // 16 x 16 threads, 8 x 8 outputs each
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k = 0; k < 49; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&a_s[k][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&a_s[k][ty * 8 + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&b_s[k][tx * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&b_s[k][tx * 8 + 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    float* out = ss + ((long long)n * 512 + tr * 128 + ty * 8) * 512 + tc * 128 + tx * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        *reinterpret_cast<float4*>(out + (long long)i * 512) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(out + (long long)i * 512 + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
}

int self_similarity_launch(const float* x, int n, float* ss_space, float* ss_channel, cudaStream_t stream) {
    if (n == 0) return 0;
    if (ss_space) {
        const int smem = (512 * 49 + 64) * (int)sizeof(float);
        static bool attr = false;
        if (!attr) {
            FFR_CUDA(cudaFuncSetAttribute(selfsim_space_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr = true;
        }
        selfsim_space_kernel<<<n, 512, smem, stream>>>(x, ss_space);
        int rc = launch_status("selfsim_space_kernel");
        if (rc) return rc;
    }
    if (ss_channel) {
        const int smem = 2 * 49 * 132 * (int)sizeof(float);
        static bool attr = false;
        if (!attr) {
            FFR_CUDA(cudaFuncSetAttribute(selfsim_channel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr = true;
        }
        selfsim_channel_kernel<<<dim3(16, n), 256, smem, stream>>>(x, ss_channel);
        return launch_status("selfsim_channel_kernel");
    }
    return 0;
}

}  // namespace ffr
