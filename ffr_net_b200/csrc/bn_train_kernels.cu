// Train-mode BatchNorm2d + PReLU (+ residual) around the reflection-padded 3x3 convolutions of RecNet
// (models/recnet.py:78-85 ConvLayer.forward with nn.BatchNorm2d in training mode, :213-218 ResidualBlock) and their
// backward (autograd of the same lines under models/trainer.py:180), for G "groups" of n images each in ONE launch:
// the two RecNet calls of a training iteration (trainer.py:144-145) are batched, BatchNorm statistics stay per call.
//
// Precision (DESIGN.md "Training numerics"): the raw conv output z is fp32; activations are written as fp16 hi + lo
// (x = hi + lo, ~21 mantissa bits) for the next convolution's forward GEMM plus a bf16 copy that the weight-gradient
// GEMM contracts with the bf16 dz (tcgen05 kind::f16 needs equal operand formats; profiles/r02_probe_mixed_formats.json);
// activation gradients are fp32 everywhere except dz, which is rounded to bf16 once, after the BatchNorm backward has
// removed its common mode. Every reduction is a fixed-order two-stage sum: no floating-point atomics.
#include "../../include/ffr_sm100.h"
#include "host.h"
#include "ptx.cuh"

#include <cuda_fp16.h>

namespace ffr {

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ float h_lo(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xFFFFu))); }
__device__ __forceinline__ float h_hi(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16))); }

// hi = fp16(x), lo = fp16(x - hi) for 8 values -> two 16-byte vectors
__device__ __forceinline__ void split_hilo8(const float (&v)[8], uint4& hi, uint4& lo) {
    float r[8];
    uint32_t h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half a = __float2half_rn(v[2 * j]), b = __float2half_rn(v[2 * j + 1]);
        r[2 * j] = v[2 * j] - __half2float(a);
        r[2 * j + 1] = v[2 * j + 1] - __half2float(b);
        h[j] = (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(pack_h2(r[0], r[1]), pack_h2(r[2], r[3]), pack_h2(r[4], r[5]), pack_h2(r[6], r[7]));
}

__device__ __forceinline__ int h9_row_of_pixel(int pix) { return (pix / 7 + 1) * 9 + (pix % 7 + 1); }

// ------------------------------------------------------------------------------------------------------------
// (A) statistics: per-(M tile, quadrant) partial sums of the conv epilogue -> per-group mean / rstd, running stats.
//   part [R][2][C]; mode 0: row-major tiles, partial row r covers matrix rows [32 r, 32 r + 32) -> image 32 r / 81;
//   mode 1: pixel-major tiles, r = ((pixel * iblocks + ib) * 4 + quad) -> images [ib*128 + quad*32, +32).
//   Partial rows beyond the data hold zeros (invalid rows are zeroed by the epilogue), so any group they map to is fine.
// ------------------------------------------------------------------------------------------------------------
constexpr int BN_MAX_GROUPS = 2;

// block = 8 channels x 128 row slices (grid C / 8): every slice adds its rows (r = slice, slice + 128, ...) per group in
// double, the slices are then added in a fixed order. The kernel is pure latency (a few hundred partial rows per
// channel): with 32 channels x 32 slices on C / 32 CTAs it took 21 us per layer — 15 launches per training step.
constexpr int BNF_SLICES = 128;
constexpr int BNF_CH = 8;
__global__ void __launch_bounds__(BNF_SLICES * BNF_CH)
bn_finalize_kernel(const float* __restrict__ part, int R, int mode, int iblocks, int n_per_group, int G, int C,
                   int C_real, float momentum, float eps, float* __restrict__ running_mean,
                   float* __restrict__ running_var, long long* __restrict__ nbt, float* __restrict__ mr) {
    __shared__ double red[BNF_SLICES][BN_MAX_GROUPS][2][BNF_CH];
    const int cl = threadIdx.x % BNF_CH, slice = threadIdx.x / BNF_CH;
    const int c = blockIdx.x * BNF_CH + cl;
    if (blockIdx.x == 0 && threadIdx.x == 0 && nbt != nullptr) *nbt += G;
    double s[BN_MAX_GROUPS], ss[BN_MAX_GROUPS];
#pragma unroll
    for (int g = 0; g < BN_MAX_GROUPS; ++g) { s[g] = 0.0; ss[g] = 0.0; }
    if (c < C) {
#pragma unroll 4
        for (int r = slice; r < R; r += BNF_SLICES) {
            const int img = (mode == 0) ? (r * 32) / 81 : ((r >> 2) % iblocks) * 128 + (r & 3) * 32;
            int rg = img / n_per_group;
            if (rg >= G) rg = G - 1;
            const double a = (double)__ldg(part + ((long long)r * 2) * C + c);
            const double b = (double)__ldg(part + ((long long)r * 2 + 1) * C + c);
#pragma unroll
            for (int g = 0; g < BN_MAX_GROUPS; ++g)
                if (g == rg) { s[g] += a; ss[g] += b; }
        }
    }
#pragma unroll
    for (int g = 0; g < BN_MAX_GROUPS; ++g) { red[slice][g][0][cl] = s[g]; red[slice][g][1][cl] = ss[g]; }
    __syncthreads();
    // stage 2: thread (g, stat, channel) adds the 128 slices in order; stage 3: one thread per channel finishes
    __shared__ double tot[BN_MAX_GROUPS][2][BNF_CH];
    if (threadIdx.x < BN_MAX_GROUPS * 2 * BNF_CH) {
        const int ch = threadIdx.x % BNF_CH, st = (threadIdx.x / BNF_CH) % 2, g = threadIdx.x / (2 * BNF_CH);
        double t = 0.0;
#pragma unroll 8
        for (int k = 0; k < BNF_SLICES; ++k) t += red[k][g][st][ch];
        tot[g][st][ch] = t;
    }
    __syncthreads();
    if (slice != 0 || c >= C) return;
    const double cnt = (double)n_per_group * 49.0;
    for (int g = 0; g < G; ++g) {
        const double ts = tot[g][0][cl], tss = tot[g][1][cl];
        const double mean = ts / cnt;
        double var = tss / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        mr[((long long)g * 2) * C + c] = (float)mean;
        mr[((long long)g * 2 + 1) * C + c] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean != nullptr && c < C_real) {      // nn.BatchNorm2d: momentum update, unbiased variance
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * (cnt / fmax(cnt - 1.0, 1.0)));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// (B) forward apply: a = prelu(gamma * (z - mean_g) * rstd_g + beta) (+ res), written
//   - as fp16 hi (+ lo at column lo_off) to every destination of the H9 scatter table (own row + reflection mirrors),
//   - optionally as a bf16 copy to the same destinations (operand of the weight-gradient GEMM),
//   - optionally as fp32 (own row only; sigmoid applied when `sigmoid` is set: M_space, recnet.py:370).
// One thread = one valid pixel x 8 channels.
// ------------------------------------------------------------------------------------------------------------
struct BnActFwd {
    const float* z; int ldz;
    const float* mr;                  // [G][2][C]
    const float* gamma; const float* beta; const float* slope;   // [C] (padded with zeros)
    const __half* res; int ldres; int res_lo_off;                // residual input (hi at res, lo at +res_lo_off; 0 = no lo)
    __half* out_h; int ldo; int lo_off;                          // lo_off == 0: hi only
    __nv_bfloat16* out_b; int ldb;
    float* out_f; int ldf; int sigmoid;
    const int2* scatter; int scatter_n;                          // [81][scatter_n] (row within the image, channel offset)
    int n_img, n_per_group, C;
    int C_real;                                                  // gamma / beta / slope hold C_real entries (0 beyond)
};

__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const BnActFwd p) {
    const int c8n = p.C / 8;
    const long long total = (long long)p.n_img * 49 * c8n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % c8n);
        const long long pr = i / c8n;
        const int n = (int)(pr / 49), pix = (int)(pr - (long long)n * 49);
        const int r_local = h9_row_of_pixel(pix);
        const long long row = (long long)n * 81 + r_local;
        const int g = n / p.n_per_group;
        const float* mean = p.mr + ((long long)g * 2) * p.C;
        const float* rstd = mean + p.C;
        const int c0 = c8 * 8;
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(p.z + row * p.ldz + c0));
        const float4 z1 = __ldg(reinterpret_cast<const float4*>(p.z + row * p.ldz + c0 + 4));
        float v[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (p.res != nullptr) {
            const uint4 rh = __ldg(reinterpret_cast<const uint4*>(p.res + row * p.ldres + c0));
            r[0] = h_lo(rh.x); r[1] = h_hi(rh.x); r[2] = h_lo(rh.y); r[3] = h_hi(rh.y);
            r[4] = h_lo(rh.z); r[5] = h_hi(rh.z); r[6] = h_lo(rh.w); r[7] = h_hi(rh.w);
            if (p.res_lo_off) {
                const uint4 rl = __ldg(reinterpret_cast<const uint4*>(p.res + row * p.ldres + p.res_lo_off + c0));
                r[0] += h_lo(rl.x); r[1] += h_hi(rl.x); r[2] += h_lo(rl.y); r[3] += h_hi(rl.y);
                r[4] += h_lo(rl.z); r[5] += h_hi(rl.z); r[6] += h_lo(rl.w); r[7] += h_hi(rl.w);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            const bool real = c < p.C_real;
            float y = (v[j] - mean[c]) * rstd[c] * (real ? p.gamma[c] : 0.f) + (real ? p.beta[c] : 0.f);
            y = fmaxf(y, 0.f) + (real ? p.slope[c] : 0.f) * fminf(y, 0.f);
            v[j] = y + r[j];
        }
        if (p.out_f != nullptr) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = p.sigmoid ? 1.0f / (1.0f + expf(-v[j])) : v[j];
            *reinterpret_cast<float4*>(p.out_f + row * p.ldf + c0) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(p.out_f + row * p.ldf + c0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (p.out_h == nullptr && p.out_b == nullptr) continue;
        uint4 hi, lo;
        split_hilo8(v, hi, lo);
        const uint4 bf = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        for (int k = 0; k < p.scatter_n; ++k) {
            const int2 e = __ldg(p.scatter + r_local * p.scatter_n + k);
            if (e.x < 0) continue;
            const long long drow = (long long)n * 81 + e.x;
            if (p.out_h != nullptr) {
                *reinterpret_cast<uint4*>(p.out_h + drow * p.ldo + e.y + c0) = hi;
                if (p.lo_off) *reinterpret_cast<uint4*>(p.out_h + drow * p.ldo + p.lo_off + e.y + c0) = lo;
            }
            if (p.out_b != nullptr) *reinterpret_cast<uint4*>(p.out_b + drow * p.ldb + e.y + c0) = bf;
        }
    }
}

// Fast path of (B): C / 8 is a power of two <= 256, so a 256-thread block covers whole pixel rows, a thread keeps its 8
// channels — and their affine / PReLU parameters in registers — for the whole grid-stride loop, the statistics are
// reloaded only when the group changes, and all index arithmetic is 32-bit (the generic kernel spends most of its
// issue slots on 64-bit divisions and 40 scalar parameter loads per 8 outputs). Same arithmetic per element.
__global__ void __launch_bounds__(256) bn_act_fwd_pow2_kernel(const BnActFwd p, const int c8_shift) {
    const unsigned c0 = (threadIdx.x & ((1u << c8_shift) - 1u)) * 8u;
    const unsigned rpb = 256u >> c8_shift;
    const unsigned rows = (unsigned)p.n_img * 49u;
    float gm[8], bt[8], sl[8], mean[8], rstd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool real = (int)(c0 + j) < p.C_real;
        gm[j] = real ? p.gamma[c0 + j] : 0.f;
        bt[j] = real ? p.beta[c0 + j] : 0.f;
        sl[j] = real ? p.slope[c0 + j] : 0.f;
        mean[j] = 0.f; rstd[j] = 0.f;
    }
    int g_cur = -1;
    for (unsigned pr = blockIdx.x * rpb + (threadIdx.x >> c8_shift); pr < rows; pr += gridDim.x * rpb) {
        const unsigned n = pr / 49u, pix = pr - n * 49u;
        const unsigned r_local = (pix / 7u + 1u) * 9u + (pix % 7u + 1u);
        const unsigned row = n * 81u + r_local;
        const int g = (int)(n / (unsigned)p.n_per_group);
        if (g != g_cur) {
            const float4* mp = reinterpret_cast<const float4*>(p.mr + ((size_t)g * 2) * p.C + c0);
            const float4* rp = reinterpret_cast<const float4*>(p.mr + ((size_t)g * 2 + 1) * p.C + c0);
            const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), r0 = __ldg(rp), r1 = __ldg(rp + 1);
            mean[0] = m0.x; mean[1] = m0.y; mean[2] = m0.z; mean[3] = m0.w;
            mean[4] = m1.x; mean[5] = m1.y; mean[6] = m1.z; mean[7] = m1.w;
            rstd[0] = r0.x; rstd[1] = r0.y; rstd[2] = r0.z; rstd[3] = r0.w;
            rstd[4] = r1.x; rstd[5] = r1.y; rstd[6] = r1.z; rstd[7] = r1.w;
            g_cur = g;
        }
        const float* zp = p.z + (size_t)row * p.ldz + c0;
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(zp));
        const float4 z1 = __ldg(reinterpret_cast<const float4*>(zp + 4));
        float v[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (p.res != nullptr) {
            const __half* rsp = p.res + (size_t)row * p.ldres + c0;
            const uint4 rh = __ldg(reinterpret_cast<const uint4*>(rsp));
            r[0] = h_lo(rh.x); r[1] = h_hi(rh.x); r[2] = h_lo(rh.y); r[3] = h_hi(rh.y);
            r[4] = h_lo(rh.z); r[5] = h_hi(rh.z); r[6] = h_lo(rh.w); r[7] = h_hi(rh.w);
            if (p.res_lo_off) {
                const uint4 rl = __ldg(reinterpret_cast<const uint4*>(rsp + p.res_lo_off));
                r[0] += h_lo(rl.x); r[1] += h_hi(rl.x); r[2] += h_lo(rl.y); r[3] += h_hi(rl.y);
                r[4] += h_lo(rl.z); r[5] += h_hi(rl.z); r[6] += h_lo(rl.w); r[7] += h_hi(rl.w);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float y = (v[j] - mean[j]) * rstd[j] * gm[j] + bt[j];
            y = fmaxf(y, 0.f) + sl[j] * fminf(y, 0.f);
            v[j] = y + r[j];
        }
        if (p.out_f != nullptr) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = p.sigmoid ? 1.0f / (1.0f + expf(-v[j])) : v[j];
            float* op = p.out_f + (size_t)row * p.ldf + c0;
            *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (p.out_h == nullptr && p.out_b == nullptr) continue;
        uint4 hi, lo;
        split_hilo8(v, hi, lo);
        const uint4 bf = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        for (int k = 0; k < p.scatter_n; ++k) {
            const int2 e = __ldg(p.scatter + r_local * p.scatter_n + k);
            if (e.x < 0) continue;
            const size_t drow = (size_t)(n * 81u + (unsigned)e.x);
            if (p.out_h != nullptr) {
                __half* dp = p.out_h + drow * p.ldo + e.y + c0;
                *reinterpret_cast<uint4*>(dp) = hi;
                if (p.lo_off) *reinterpret_cast<uint4*>(dp + p.lo_off) = lo;
            }
            if (p.out_b != nullptr) *reinterpret_cast<uint4*>(p.out_b + drow * p.ldb + e.y + c0) = bf;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// (C) backward. The gradient w.r.t. the layer's OUTPUT a arrives from up to three fp32 sources:
//   da   : on the H9 grid (own row + mirror rows) — the dgrad GEMM of the consuming convolution; folded through the
//          scatter table; channel slot da_ch0 of a matrix with pitch ldda;
//   dadd : own rows only (already folded): the residual branch of a ResidualBlock, or a loss gradient;
//   dv   : [n_img][lddv] * dv_scale broadcast over the 49 pixels (AvgPool2d(7), recnet.py:423).
// pass 1 folds them into afold (fp32, own rows), and reduces per (group, channel)
//   s0 = sum dy, s1 = sum dy * zhat, s2 = sum a * min(y, 0)   with dy = a * prelu'(y)
// into per-CTA partials (plain stores); the finalize kernel sums the partials in a fixed order, writes the BatchNorm /
// PReLU parameter gradients and the per-group sums; pass 2 recomputes dy in fp32 and writes
//   dz = gamma * rstd * (dy - s0/cnt - zhat * s1/cnt)  as bf16 (zeros on the halo rows).
// ------------------------------------------------------------------------------------------------------------
struct BnActBwd {
    const float* da; int ldda; int da_ch0;
    const int2* scatter; int scatter_n;
    const float* dadd; int ldadd; int dadd_ch0;
    const float* dv; int lddv; float dv_scale;
    const float* z; int ldz;
    const float* mr; const float* gamma; const float* beta; const float* slope;
    float* afold; int ldaf;              // fp32 [n_img*81][ldaf], own rows written by pass 1
    float* partial;                      // [gridDim.x of pass 1][3][C]
    float* gsum;                         // [G][2][C]
    float* dgamma; float* dbeta; float* dslope; int accumulate; int C_real;
    __nv_bfloat16* dz; int lddz;
    int n_img, n_per_group, G, C, ctas_per_group;
};

__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(const BnActBwd p) {
    __shared__ float red[3][32][65];
    const int c0 = blockIdx.y * 64;
    const int cg = (threadIdx.x & 7) * 8;
    const int rlane = threadIdx.x >> 3;
    const int g = blockIdx.x / p.ctas_per_group, cta_in_g = blockIdx.x - g * p.ctas_per_group;
    const unsigned rows_g = (unsigned)p.n_per_group * 49u;
    const float* mean = p.mr + ((long long)g * 2) * p.C;
    const float* rstd = mean + p.C;
    float s0[8], s1[8], s2[8], m[8], rs[8], gm[8], b[8], sl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + cg + j;
        const bool real = c < p.C_real;
        s0[j] = s1[j] = s2[j] = 0.f;
        m[j] = mean[c]; rs[j] = rstd[c];
        gm[j] = real ? p.gamma[c] : 0.f; b[j] = real ? p.beta[c] : 0.f; sl[j] = real ? p.slope[c] : 0.f;
    }
    for (unsigned pr = (unsigned)cta_in_g * 32u + rlane; pr < rows_g; pr += (unsigned)p.ctas_per_group * 32u) {
        const unsigned nl = pr / 49u, pix = pr - nl * 49u;
        const int n = g * p.n_per_group + (int)nl;
        if (n >= p.n_img) break;
        const int r_local = (int)((pix / 7u + 1u) * 9u + (pix % 7u + 1u));
        const size_t row = (size_t)((unsigned)n * 81u + (unsigned)r_local);
        float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (p.da != nullptr) {
            for (int k = 0; k < p.scatter_n; ++k) {
                const int2 e = __ldg(p.scatter + r_local * p.scatter_n + k);
                if (e.x < 0) continue;
                const float* src = p.da + (size_t)((unsigned)n * 81u + (unsigned)e.x) * p.ldda + p.da_ch0 + c0 + cg;
                const float4 u0 = __ldg(reinterpret_cast<const float4*>(src));
                const float4 u1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                a[0] += u0.x; a[1] += u0.y; a[2] += u0.z; a[3] += u0.w;
                a[4] += u1.x; a[5] += u1.y; a[6] += u1.z; a[7] += u1.w;
            }
        }
        if (p.dadd != nullptr) {
            const float* src = p.dadd + row * p.ldadd + p.dadd_ch0 + c0 + cg;
            const float4 u0 = __ldg(reinterpret_cast<const float4*>(src));
            const float4 u1 = __ldg(reinterpret_cast<const float4*>(src + 4));
            a[0] += u0.x; a[1] += u0.y; a[2] += u0.z; a[3] += u0.w;
            a[4] += u1.x; a[5] += u1.y; a[6] += u1.z; a[7] += u1.w;
        }
        if (p.dv != nullptr) {
            const float* src = p.dv + (long long)n * p.lddv + c0 + cg;
            const float4 u0 = __ldg(reinterpret_cast<const float4*>(src));
            const float4 u1 = __ldg(reinterpret_cast<const float4*>(src + 4));
            a[0] += u0.x * p.dv_scale; a[1] += u0.y * p.dv_scale; a[2] += u0.z * p.dv_scale; a[3] += u0.w * p.dv_scale;
            a[4] += u1.x * p.dv_scale; a[5] += u1.y * p.dv_scale; a[6] += u1.z * p.dv_scale; a[7] += u1.w * p.dv_scale;
        }
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(p.z + row * p.ldz + c0 + cg));
        const float4 z1 = __ldg(reinterpret_cast<const float4*>(p.z + row * p.ldz + c0 + cg + 4));
        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float zh = (zz[j] - m[j]) * rs[j];
            const float y = zh * gm[j] + b[j];
            const float d = a[j] * (y > 0.f ? 1.f : sl[j]);
            s0[j] += d;
            s1[j] += d * zh;
            s2[j] += a[j] * fminf(y, 0.f);
        }
        float* af = p.afold + row * p.ldaf + c0 + cg;
        *reinterpret_cast<float4*>(af) = make_float4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<float4*>(af + 4) = make_float4(a[4], a[5], a[6], a[7]);
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
        p.partial[((long long)blockIdx.x * 3 + q) * p.C + c0 + c] = t;
    }
}

// block = 8 channels x 32 slices of the per-CTA partial rows (grid C / 8), slices added in a fixed order
__global__ void __launch_bounds__(256) bn_act_bwd_finalize_kernel(const BnActBwd p) {
    __shared__ float red[32][3][8];
    const int cl = threadIdx.x & 7, slice = threadIdx.x >> 3;
    const int c = blockIdx.x * 8 + cl;
    float tg = 0.f, tb = 0.f, ts = 0.f;
    for (int g = 0; g < p.G; ++g) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        if (c < p.C) {
            for (int k = slice; k < p.ctas_per_group; k += 32) {
                const float* q = p.partial + ((long long)(g * p.ctas_per_group + k) * 3) * p.C + c;
                s0 += q[0]; s1 += q[p.C]; s2 += q[2 * p.C];
            }
        }
        __syncthreads();
        red[slice][0][cl] = s0; red[slice][1][cl] = s1; red[slice][2][cl] = s2;
        __syncthreads();
        if (slice == 0 && c < p.C) {
            s0 = s1 = s2 = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) { s0 += red[k][0][cl]; s1 += red[k][1][cl]; s2 += red[k][2][cl]; }
            p.gsum[((long long)g * 2) * p.C + c] = s0;
            p.gsum[((long long)g * 2 + 1) * p.C + c] = s1;
            tb += s0; tg += s1; ts += s2;
        }
    }
    if (slice == 0 && c < p.C_real) {
        if (p.accumulate) { p.dgamma[c] += tg; p.dbeta[c] += tb; p.dslope[c] += ts; }
        else              { p.dgamma[c] = tg;  p.dbeta[c] = tb;  p.dslope[c] = ts; }
    }
}

__global__ void __launch_bounds__(256) bn_act_bwd_dz_kernel(const BnActBwd p) {
    const int c8n = p.C / 8;
    const float inv_cnt = 1.0f / (float)(p.n_per_group * 49);
    const long long total = (long long)p.n_img * 81 * c8n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % c8n);
        const long long row = i / c8n;
        const int pos = (int)(row % 81);
        const int hp = pos / 9, wp = pos - hp * 9;
        uint4* dst = reinterpret_cast<uint4*>(p.dz + row * p.lddz) + c8;
        if (hp == 0 || hp == 8 || wp == 0 || wp == 8) { *dst = make_uint4(0, 0, 0, 0); continue; }
        const int n = (int)(row / 81);
        int g = n / p.n_per_group;
        if (g >= p.G) g = p.G - 1;
        const float* mean = p.mr + ((long long)g * 2) * p.C;
        const float* rstd = mean + p.C;
        const float* gs0 = p.gsum + ((long long)g * 2) * p.C;
        const float* gs1 = gs0 + p.C;
        const int c0 = c8 * 8;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.afold + row * p.ldaf + c0));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(p.afold + row * p.ldaf + c0 + 4));
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(p.z + row * p.ldz + c0));
        const float4 z1 = __ldg(reinterpret_cast<const float4*>(p.z + row * p.ldz + c0 + 4));
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            const bool real = c < p.C_real;
            const float gmm = real ? p.gamma[c] : 0.f;
            const float rs = rstd[c];
            const float zh = (zz[j] - mean[c]) * rs;
            const float y = zh * gmm + (real ? p.beta[c] : 0.f);
            const float d = a[j] * (y > 0.f ? 1.f : (real ? p.slope[c] : 0.f));
            o[j] = gmm * rs * (d - gs0[c] * inv_cnt - zh * gs1[c] * inv_cnt);
        }
        *dst = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
}

// Fast path of pass 2 (same conditions and layout of the work as bn_act_fwd_pow2_kernel, over all 81 grid rows)
__global__ void __launch_bounds__(256) bn_act_bwd_dz_pow2_kernel(const BnActBwd p, const int c8_shift) {
    const unsigned c0 = (threadIdx.x & ((1u << c8_shift) - 1u)) * 8u;
    const unsigned rpb = 256u >> c8_shift;
    const unsigned rows = (unsigned)p.n_img * 81u;
    const float inv_cnt = 1.0f / (float)(p.n_per_group * 49);
    float gm[8], bt[8], sl[8], mean[8], rstd[8], gs0[8], gs1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool real = (int)(c0 + j) < p.C_real;
        gm[j] = real ? p.gamma[c0 + j] : 0.f;
        bt[j] = real ? p.beta[c0 + j] : 0.f;
        sl[j] = real ? p.slope[c0 + j] : 0.f;
        mean[j] = rstd[j] = gs0[j] = gs1[j] = 0.f;
    }
    int g_cur = -1;
    for (unsigned row = blockIdx.x * rpb + (threadIdx.x >> c8_shift); row < rows; row += gridDim.x * rpb) {
        const unsigned n = row / 81u, pos = row - n * 81u;
        const unsigned hp = pos / 9u, wp = pos - hp * 9u;
        uint4* dst = reinterpret_cast<uint4*>(p.dz + (size_t)row * p.lddz + c0);
        if (hp == 0 || hp == 8 || wp == 0 || wp == 8) { *dst = make_uint4(0, 0, 0, 0); continue; }
        int g = (int)(n / (unsigned)p.n_per_group);
        if (g >= p.G) g = p.G - 1;
        if (g != g_cur) {
            const float* base = p.mr + ((size_t)g * 2) * p.C + c0;
            const float* gbase = p.gsum + ((size_t)g * 2) * p.C + c0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(base) + q);
                const float4 r = __ldg(reinterpret_cast<const float4*>(base + p.C) + q);
                const float4 a = __ldg(reinterpret_cast<const float4*>(gbase) + q);
                const float4 b = __ldg(reinterpret_cast<const float4*>(gbase + p.C) + q);
                mean[q * 4] = m.x; mean[q * 4 + 1] = m.y; mean[q * 4 + 2] = m.z; mean[q * 4 + 3] = m.w;
                rstd[q * 4] = r.x; rstd[q * 4 + 1] = r.y; rstd[q * 4 + 2] = r.z; rstd[q * 4 + 3] = r.w;
                gs0[q * 4] = a.x; gs0[q * 4 + 1] = a.y; gs0[q * 4 + 2] = a.z; gs0[q * 4 + 3] = a.w;
                gs1[q * 4] = b.x; gs1[q * 4 + 1] = b.y; gs1[q * 4 + 2] = b.z; gs1[q * 4 + 3] = b.w;
            }
            g_cur = g;
        }
        const float* ap = p.afold + (size_t)row * p.ldaf + c0;
        const float* zp = p.z + (size_t)row * p.ldz + c0;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(ap));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(ap + 4));
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(zp));
        const float4 z1 = __ldg(reinterpret_cast<const float4*>(zp + 4));
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float rs = rstd[j];
            const float zh = (zz[j] - mean[j]) * rs;
            const float y = zh * gm[j] + bt[j];
            const float d = a[j] * (y > 0.f ? 1.f : sl[j]);
            o[j] = gm[j] * rs * (d - gs0[j] * inv_cnt - zh * gs1[j] * inv_cnt);
        }
        *dst = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
}

// ------------------------------------------------------------------------------------------------------------
// AvgPool2d(7) over the valid rows of an fp32 H9 matrix: v[n][c] = mean over the 49 pixels (fixed summation order).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) h9_avgpool_kernel(const float* __restrict__ a, int lda, float* __restrict__ v,
                                                         int ldv, int n_img, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4n = C / 4;
    if (i >= (long long)n_img * c4n) return;
    const int n = (int)(i / c4n), c0 = (int)(i % c4n) * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pix = 0; pix < 49; ++pix) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(a + ((long long)n * 81 + h9_row_of_pixel(pix)) * lda + c0));
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    const float k = 1.0f / 49.0f;
    *reinterpret_cast<float4*>(v + (long long)n * ldv + c0) = make_float4(s.x * k, s.y * k, s.z * k, s.w * k);
}

// the same over a bf16 H9 matrix (the eval path's stored feature map), 8 channels per thread
__global__ void __launch_bounds__(256) h9_avgpool_bf16_kernel(const __nv_bfloat16* __restrict__ a, int lda,
                                                              float* __restrict__ v, int ldv, int n_img, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c8n = C / 8;
    if (i >= (long long)n_img * c8n) return;
    const int n = (int)(i / c8n), c0 = (int)(i % c8n) * 8;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int pix = 0; pix < 49; ++pix) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(a + ((long long)n * 81 + h9_row_of_pixel(pix)) * lda + c0));
        s[0] += bf16lo(t.x); s[1] += bf16hi(t.x); s[2] += bf16lo(t.y); s[3] += bf16hi(t.y);
        s[4] += bf16lo(t.z); s[5] += bf16hi(t.z); s[6] += bf16lo(t.w); s[7] += bf16hi(t.w);
    }
    const float k = 1.0f / 49.0f;
    float4* o = reinterpret_cast<float4*>(v + (long long)n * ldv + c0);
    o[0] = make_float4(s[0] * k, s[1] * k, s[2] * k, s[3] * k);
    o[1] = make_float4(s[4] * k, s[5] * k, s[6] * k, s[7] * k);
}

// fp32 NCHW (n,C,7,7) -> own rows of an fp32 H9 matrix (channel slot ch0, pitch ld): layout of the `dadd` gradient source
__global__ void __launch_bounds__(256) nchw_to_h9_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int ld,
                                                             int ch0, int C) {
    __shared__ float tile[49][65];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 64 * 49; i += 256) {
        const int c = i / 49, pix = i - c * 49;
        tile[pix][c] = (c0 + c < C) ? x[((long long)n * C + c0 + c) * 49 + pix] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 49 * 64; i += 256) {
        const int pix = i >> 6, c = i & 63;
        if (c0 + c < C) out[((long long)n * 81 + h9_row_of_pixel(pix)) * ld + ch0 + c0 + c] = tile[pix][c];
    }
}

}  // namespace ffr

using namespace ffr;
static inline cudaStream_t S_(ffr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

FFR_API int ffr_bn_finalize(const float* part, int part_rows, int pixmajor, int n_img, int n_per_group, int C,
                            int C_real, float momentum, float eps, float* running_mean, float* running_var,
                            long long* num_batches_tracked, float* mean_rstd, ffr_stream_t stream) {
    FFR_CHECK_ARG(part && mean_rstd && n_per_group > 0 && n_img % n_per_group == 0 && C % 8 == 0,
                  "ffr_bn_finalize: bad arguments");
    const int G = n_img / n_per_group;
    FFR_CHECK_ARG(G == 1 || n_per_group % 32 == 0, "ffr_bn_finalize: batched groups need n_per_group %% 32 == 0");
    FFR_CHECK_ARG(G <= BN_MAX_GROUPS, "ffr_bn_finalize: at most %d groups", BN_MAX_GROUPS);
    FFR_CHECK_ARG(!running_mean == !running_var, "ffr_bn_finalize: running_mean / running_var go together");
    const int iblocks = (n_img + 127) / 128;
    bn_finalize_kernel<<<(C + BNF_CH - 1) / BNF_CH, BNF_SLICES * BNF_CH, 0, S_(stream)>>>(part, part_rows, pixmajor, iblocks, n_per_group, G, C,
                                                               C_real, momentum, eps, running_mean, running_var,
                                                               num_batches_tracked, mean_rstd);
    return launch_status("bn_finalize_kernel");
}

FFR_API int ffr_bn_act_fwd(const float* z, int ldz, const float* mean_rstd, const float* gamma, const float* beta,
                           const float* slope, const void* res, int ldres, int res_lo_off, void* out_h, int ldo,
                           int lo_off, void* out_b, int ldb, float* out_f, int ldf, int sigmoid, const int* scatter,
                           int scatter_n, int n_img, int n_per_group, int C, int C_real, ffr_stream_t stream) {
    FFR_CHECK_ARG(z && mean_rstd && gamma && beta && slope && scatter, "ffr_bn_act_fwd: null pointer");
    FFR_CHECK_ARG(C % 8 == 0 && ldz % 4 == 0 && ldo % 8 == 0 && lo_off % 8 == 0 && ldb % 8 == 0 && ldf % 4 == 0 &&
                  ldres % 8 == 0 && res_lo_off % 8 == 0 && n_per_group > 0, "ffr_bn_act_fwd: bad pitch");
    BnActFwd p;
    p.z = z; p.ldz = ldz; p.mr = mean_rstd; p.gamma = gamma; p.beta = beta; p.slope = slope;
    p.res = reinterpret_cast<const __half*>(res); p.ldres = ldres; p.res_lo_off = res_lo_off;
    p.out_h = reinterpret_cast<__half*>(out_h); p.ldo = ldo; p.lo_off = lo_off;
    p.out_b = reinterpret_cast<__nv_bfloat16*>(out_b); p.ldb = ldb;
    p.out_f = out_f; p.ldf = ldf; p.sigmoid = sigmoid;
    p.scatter = reinterpret_cast<const int2*>(scatter); p.scatter_n = scatter_n;
    p.n_img = n_img; p.n_per_group = n_per_group; p.C = C; p.C_real = C_real;
    const long long total = (long long)n_img * 49 * (C / 8);
    if (total == 0) return 0;
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * 8) grid = num_sms() * 8;
    const int c8n = C / 8;
    if ((c8n & (c8n - 1)) == 0 && c8n <= 256 && (long long)n_img * 81 < (1ll << 24)) {
        int sh = 0;
        while ((1 << sh) < c8n) ++sh;
        bn_act_fwd_pow2_kernel<<<grid, 256, 0, S_(stream)>>>(p, sh);
        return launch_status("bn_act_fwd_pow2_kernel");
    }
    bn_act_fwd_kernel<<<grid, 256, 0, S_(stream)>>>(p);
    return launch_status("bn_act_fwd_kernel");
}

/* workspace: partial sums, fp32 [ffr_bn_act_bwd_partial_rows(n_img / n_per_group)][3][C] */
FFR_API int ffr_bn_act_bwd_partial_rows(int n_groups, int C) {
    int per = (num_sms() * 8 + C / 64 - 1) / (C / 64) / (n_groups > 0 ? n_groups : 1);
    if (per < 1) per = 1;
    return per * (n_groups > 0 ? n_groups : 1);
}

FFR_API int ffr_bn_act_bwd(const float* da, int ldda, int da_ch0, const int* scatter, int scatter_n, const float* dadd,
                           int ldadd, int dadd_ch0, const float* dv, int lddv, float dv_scale, const float* z, int ldz,
                           const float* mean_rstd, const float* gamma, const float* beta, const float* slope,
                           float* afold, int ldaf, float* partial, float* gsum, float* dgamma, float* dbeta,
                           float* dslope, int accumulate, int C_real, void* dz, int lddz, int n_img, int n_per_group,
                           int C, ffr_stream_t stream) {
    FFR_CHECK_ARG((da || dadd || dv) && z && mean_rstd && gamma && beta && slope && afold && partial && gsum && dgamma &&
                  dbeta && dslope && dz, "ffr_bn_act_bwd: null pointer");
    FFR_CHECK_ARG(!da || scatter, "ffr_bn_act_bwd: da needs the scatter table");
    FFR_CHECK_ARG(C % 64 == 0 && n_per_group > 0 && n_img % n_per_group == 0 && ldda % 4 == 0 && da_ch0 % 4 == 0 &&
                  ldadd % 4 == 0 && dadd_ch0 % 4 == 0 && lddv % 4 == 0 && ldz % 4 == 0 && ldaf % 4 == 0 && lddz % 8 == 0,
                  "ffr_bn_act_bwd: bad shape / pitch");
    BnActBwd p;
    p.da = da; p.ldda = ldda; p.da_ch0 = da_ch0;
    p.scatter = reinterpret_cast<const int2*>(scatter); p.scatter_n = scatter_n;
    p.dadd = dadd; p.ldadd = ldadd; p.dadd_ch0 = dadd_ch0;
    p.dv = dv; p.lddv = lddv; p.dv_scale = dv_scale;
    p.z = z; p.ldz = ldz; p.mr = mean_rstd; p.gamma = gamma; p.beta = beta; p.slope = slope;
    p.afold = afold; p.ldaf = ldaf; p.partial = partial; p.gsum = gsum;
    p.dgamma = dgamma; p.dbeta = dbeta; p.dslope = dslope; p.accumulate = accumulate; p.C_real = C_real;
    p.dz = reinterpret_cast<__nv_bfloat16*>(dz); p.lddz = lddz;
    p.n_img = n_img; p.n_per_group = n_per_group; p.G = n_img / n_per_group; p.C = C;
    p.ctas_per_group = ffr_bn_act_bwd_partial_rows(p.G, C) / p.G;
    const long long rows_g = (long long)n_per_group * 49;
    const int need = (int)((rows_g + 31) / 32);
    if (p.ctas_per_group > need) p.ctas_per_group = need;
    dim3 grid(p.ctas_per_group * p.G, C / 64);
    bn_act_bwd_reduce_kernel<<<grid, 256, 0, S_(stream)>>>(p);
    int rc = launch_status("bn_act_bwd_reduce_kernel");
    if (rc) return rc;
    bn_act_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, S_(stream)>>>(p);
    rc = launch_status("bn_act_bwd_finalize_kernel");
    if (rc) return rc;
    const long long total = (long long)n_img * 81 * (C / 8);
    int g2 = (int)((total + 255) / 256);
    if (g2 > num_sms() * 8) g2 = num_sms() * 8;
    const int c8n = C / 8;
    if ((c8n & (c8n - 1)) == 0 && c8n <= 256 && (long long)n_img * 81 < (1ll << 24)) {
        int sh = 0;
        while ((1 << sh) < c8n) ++sh;
        bn_act_bwd_dz_pow2_kernel<<<g2, 256, 0, S_(stream)>>>(p, sh);
        return launch_status("bn_act_bwd_dz_pow2_kernel");
    }
    bn_act_bwd_dz_kernel<<<g2, 256, 0, S_(stream)>>>(p);
    return launch_status("bn_act_bwd_dz_kernel");
}

FFR_API int ffr_nchw_to_h9_f32(const float* x, float* out, int ld, int ch0, int n, int C, ffr_stream_t stream) {
    FFR_CHECK_ARG(x && out, "ffr_nchw_to_h9_f32: null pointer");
    if (n == 0) return 0;
    nchw_to_h9_f32_kernel<<<dim3((C + 63) / 64, n), 256, 0, S_(stream)>>>(x, out, ld, ch0, C);
    return launch_status("nchw_to_h9_f32_kernel");
}

FFR_API int ffr_h9_avgpool(const float* a, int lda, float* v, int ldv, int n_img, int C, ffr_stream_t stream) {
    FFR_CHECK_ARG(a && v && C % 4 == 0 && lda % 4 == 0 && ldv % 4 == 0, "ffr_h9_avgpool: bad arguments");
    const long long total = (long long)n_img * (C / 4);
    if (total == 0) return 0;
    h9_avgpool_kernel<<<(int)((total + 255) / 256), 256, 0, S_(stream)>>>(a, lda, v, ldv, n_img, C);
    return launch_status("h9_avgpool_kernel");
}

FFR_API int ffr_h9_avgpool_bf16(const void* a, int lda, float* v, int ldv, int n_img, int C, ffr_stream_t stream) {
    FFR_CHECK_ARG(a && v && C % 8 == 0 && lda % 8 == 0 && ldv % 4 == 0, "ffr_h9_avgpool_bf16: bad arguments");
    const long long total = (long long)n_img * (C / 8);
    if (total == 0) return 0;
    h9_avgpool_bf16_kernel<<<(int)((total + 255) / 256), 256, 0, S_(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(a), lda, v, ldv, n_img, C);
    return launch_status("h9_avgpool_bf16_kernel");
}

}  // extern "C"
