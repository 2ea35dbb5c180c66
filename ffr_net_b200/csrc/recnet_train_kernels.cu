// RecNet training path, everything that is not a ConvLayer (models/recnet.py:398-423 forward and its autograd):
//   - recnet_prep_train_kernel   selfSimilarity (:399), cat fan-outs (:401-402, :420), the whole channel rectifier
//                                Conv4Channel (:372-386, :406) and feat_channel = M_channel @ X with the flip / cat
//                                of :410-417, in fp32, one CTA per sample;
//   - feat_space_train_kernel    feat_space = X @ M_space (:409) and its backward w.r.t. M_space (through the sigmoid);
//   - fc_bwd_gather_kernel       gradient of the flip / cat / reflection fan-out of feat_channel -> GEMM operand;
//   - chan_bwd_kernel (+ reduce / compose kernels)   backward of the thin Conv4Channel chain;
// The dense backward contractions of the channel rectifier (dM = dFC X^T, dh7 = dM W8, dW8 = dM^T h7) run on the
// tcgen05 GEMMs of conv_gemm.cu / train_kernels.cu (ffr_net_b200/recnet_train.py drives the sequence).
// The input feature map X comes from the frozen backbone (models/trainer.py:62-63) and needs no gradient.
//
// Activation outputs are fp16 hi + lo (+ a bf16 copy for the weight-gradient GEMMs), see bn_train_kernels.cu.
#include "../../include/ffr_sm100.h"
#include "host.h"
#include "ptx.cuh"

#include <cuda_fp16.h>

namespace ffr {

constexpr int XS = 52;   // shared-memory row pitch of X (49 pixels + 3 zero pads): rows stay 16-byte aligned

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int reflect_src_t(int p) { return p == 0 ? 1 : (p == 8 ? 5 : p - 1); }
__device__ __forceinline__ int h9_row(int pix) { return (pix / 7 + 1) * 9 + (pix % 7 + 1); }

struct H16 { __half hi, lo; };
__device__ __forceinline__ H16 split_h(float v) {
    H16 r;
    r.hi = __float2half_rn(v);
    r.lo = __float2half_rn(v - __half2float(r.hi));
    return r;
}

// One activation destination: fp16 hi/lo matrix + bf16 copy, both in the H9 layout
struct ActDst {
    __half* h; int ldh; int lo_off;       // lo_off == 0: hi only
    __nv_bfloat16* b; int ldb;            // may be nullptr
};
__device__ __forceinline__ void act_store(const ActDst& d, long long row, int ch, float v) {
    const H16 s = split_h(v);
    d.h[row * d.ldh + ch] = s.hi;
    if (d.lo_off) d.h[row * d.ldh + d.lo_off + ch] = s.lo;
    if (d.b) d.b[row * d.ldb + ch] = __float2bfloat16_rn(v);
}
// own row + reflection mirrors of interior pixel (h, w) of image n
template <typename F>
__device__ __forceinline__ void for_each_mirror(int n, int h, int w, F&& f) {
    const int mh = (h == 1) ? -2 : ((h == 5) ? 2 : 0);
    const int mw = (w == 1) ? -2 : ((w == 5) ? 2 : 0);
    const long long base = (long long)n * 81 + (h + 1) * 9 + (w + 1);
    f(base);
    if (mh) f(base + mh * 9);
    if (mw) f(base + mw);
    if (mh && mw) f(base + mh * 9 + mw);
}

struct PrepTrain {
    const float* x;                       // [n][512][49] fp32 (NCHW)
    const float* w0;                      // Conv4Channel.0.weight [32][561]
    const float* b0;                      // [32]
    const float* slope1; const float* slope4; const float* slope7;   // [512] PReLU over the 512 rows (recnet.py:374)
    const float* A1; const float* c1; const float* A2; const float* c2;   // composed 32x32 maps (chan_compose_kernel)
    const float* w8; const float* b8;     // Conv4Channel.8 [512][32], [512]
    ActDst s0;                            // Conv4Space input: X | ss_space | 0, 576 channels
    ActDst cm;                            // Conv4Merge input, slot [1024,1536) <- X
    float* g0; float* g1; float* g2;      // pre-PReLU activations of the chain, [n*512][32] each (saved for backward)
    __nv_bfloat16* h7b;                   // [n*512][64]: h7 | 1 | 0...   (B operand of the dW8 / db8 contraction)
    __nv_bfloat16* xk;                    // [n*512][64]: X rows, 49 valid (B operand of dM = dFC X^T)
    __half* mch2;                         // [n*512][1024] fp16: M_channel = sigmoid(.) as [hi | lo] (A operand of
                                          // feat_channel = M_channel @ X; the hi part also serves the backward's m (1 - m))
    __half* x3;                           // [n*64][1536] fp16: X^T rows (pixel hw, 49 valid) as [hi | lo | hi] over the
                                          // 512 channels: per-sample B operand of the same GEMM, run as three "taps" with
                                          // A column offsets (0, 0, 512): hi.hi + hi.lo + lo.hi
    float* inv_c;                         // [n*512] 1 / max(|X_c|, eps)
    float* tmat;                          // [n][49][32] T = Xh^T W0b^T (saved for backward)
    float* ss_space;                      // optional [n][49][49]
};

// Conv4Channel on cat(X, ss_channel): ss_channel = Xh Xh^T (Xh = rows of X normalised over HW) is not materialised:
// Linear(561->32)(cat(X, Xh Xh^T)) = X W0a^T + Xh (Xh^T W0b^T) + b0 by associativity. Linear(32->512) directly followed
// by Linear(512->32) (no activation in between, recnet.py:375-380) is applied as the composed 32x32 map.
__global__ void __launch_bounds__(512, 1) recnet_prep_train_kernel(const PrepTrain p) {
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                      // [512][XS]
    float* inv_c = xs + 512 * XS;        // [512]
    float* inv_s = inv_c + 512;          // [64]
    float* Gs = inv_s + 64;              // [49][49] (+3)
    float* T = Gs + 49 * 49 + 3;         // [49][32]
    float* W0a = T + 49 * 32;            // [49][32]
    float* A1s = W0a + 49 * 32;          // [32][32]
    float* A2s = A1s + 1024;             // [32][32]
    float* misc = A2s + 1024;            // b0, c1, c2 [3][32]
    float* big = misc + 96;              // [8][2401] partial Grams, later W8 [512][32] + b8 [512]
    const int n = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* x = p.x + (long long)n * 512 * 49;

    for (int i = tid; i < 512 * XS; i += 512) {
        const int c = i / XS, hw = i - c * XS;
        xs[i] = (hw < 49) ? x[c * 49 + hw] : 0.f;
    }
    for (int i = tid; i < 49 * 32; i += 512) { const int hw = i >> 5, k = i & 31; W0a[i] = p.w0[k * 561 + hw]; }
    if (tid < 32) { misc[tid] = p.b0[tid]; misc[32 + tid] = p.c1[tid]; misc[64 + tid] = p.c2[tid]; }
    __syncthreads();

    {   // row norms over HW (F.normalize(dim=2) of (N,C,HW), eps 1e-12)
        float ss = 0.f;
        for (int hw = 0; hw < 49; ++hw) { const float v = xs[tid * XS + hw]; ss = fmaf(v, v, ss); }
        const float iv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        inv_c[tid] = iv;
        p.inv_c[(long long)n * 512 + tid] = iv;
    }
    for (int hw = warp; hw < 49; hw += 16) {   // column norms over C
        float ss = 0.f;
        for (int c = lane; c < 512; c += 32) { const float v = xs[c * XS + hw]; ss = fmaf(v, v, ss); }
        ss = warp_sum_t(ss);
        if (lane == 0) inv_s[hw] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    }
    __syncthreads();

    // spatial self-similarity Gram (49x49 over C): (7x7 output block, one eighth of the channels) per thread, the eight
    // channel slices are summed in a fixed order
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
            const float* xr = xs + c * XS;
            float va[7], vb[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) { va[q] = xr[bi + q]; vb[q] = xr[bj + q]; }
#pragma unroll
            for (int a = 0; a < 7; ++a)
#pragma unroll
                for (int q = 0; q < 7; ++q) acc[a][q] = fmaf(va[a], vb[q], acc[a][q]);
        }
        float* gp = big + split * 2401;
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
        for (int sp = 0; sp < 8; ++sp) g += big[sp * 2401 + o];
        Gs[o] = g * inv_s[i] * inv_s[j];
    }
    __syncthreads();                      // all reads of the partial Grams are done: `big` is free
    // T[hw][j] = sum_c Xh[c][hw] * W0b[j][c]. W0b (Conv4Channel.0.weight[:, 49:], row pitch 561) is staged in `big` as
    // [c][j] with coalesced global reads; work item = (7 consecutive pixels, j, half of the channels); the two halves land
    // in T and in the (not yet loaded) A1/A2 area and are added in a fixed order
    float* W0bs = big;                    // [512][33] (pitch 33: conflict-free for lanes over c (fill) and over j (use))
    for (int i = tid; i < 32 * 512; i += 512) {
        const int j = i >> 9, c = i & 511;
        W0bs[c * 33 + j] = __ldg(p.w0 + j * 561 + 49 + c) * inv_c[c];
    }
    __syncthreads();
    if (tid < 448) {
        const int j = tid & 31, hg = (tid >> 5) % 7, half = tid / 224;
        float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int c_lo = half * 256;
#pragma unroll 4
        for (int c = c_lo; c < c_lo + 256; ++c) {
            const float wj = W0bs[c * 33 + j];
            const float* xr = xs + c * XS + hg * 7;
#pragma unroll
            for (int q = 0; q < 7; ++q) acc[q] = fmaf(xr[q], wj, acc[q]);
        }
        float* dstp = half ? A1s : T;     // A1s..A2s: 2048 floats, loaded from global only after this phase
#pragma unroll
        for (int q = 0; q < 7; ++q) dstp[(hg * 7 + q) * 32 + j] = acc[q];
    }
    __syncthreads();
    for (int o = tid; o < 49 * 32; o += 512) {
        const float t = T[o] + A1s[o];
        T[o] = t;
        p.tmat[(long long)n * 1568 + o] = t;
    }
    __syncthreads();                      // `big` is free again: stage W8 / b8 for the last Linear; load the composed maps
    for (int i = tid; i < 1024; i += 512) { A1s[i] = p.A1[i]; A2s[i] = p.A2[i]; }
    float* W8s = big;                     // [512][32]
    float* b8s = big + 512 * 32;          // [512]
    for (int i = tid; i < 512 * 32; i += 512) W8s[i] = p.w8[i];
    b8s[tid] = p.b8[tid];

    // ---- layout fan-out of X ----
    {
        for (int pos = 0; pos < 81; ++pos) {       // thread = channel: 512 consecutive channels per row
            const int hp = pos / 9, wp = pos - hp * 9;
            const int hw = reflect_src_t(hp) * 7 + reflect_src_t(wp);
            const float v = xs[tid * XS + hw];
            const long long row = (long long)n * 81 + pos;
            act_store(p.s0, row, tid, v);
            act_store(p.cm, row, 1024 + tid, v);
        }
        for (int hw = 0; hw < 64; ++hw) {          // X^T rows for the feat_channel GEMM: [hi | lo | hi], rows 49..63 zero
            const H16 sx = split_h(hw < 49 ? xs[tid * XS + hw] : 0.f);
            __half* xr = p.x3 + ((long long)n * 64 + hw) * 1536 + tid;
            xr[0] = sx.hi; xr[512] = sx.lo; xr[1024] = sx.hi;
        }
        __nv_bfloat16* xkr = p.xk + ((long long)n * 512 + tid) * 64;
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int hw = q * 8 + e * 2;
                w[e] = pack_bf16x2(hw < 49 ? xs[tid * XS + hw] : 0.f, hw + 1 < 49 ? xs[tid * XS + hw + 1] : 0.f);
            }
            *reinterpret_cast<uint4*>(xkr + q * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    // ss_space as 49 extra channels of the Conv4Space input (channel 512 + i at pixel j holds Gram[i][j]) + zero pad
    for (int o = tid; o < 81 * 64; o += 512) {
        const int pos = o >> 6, i = o & 63;
        const int hp = pos / 9, wp = pos - hp * 9;
        const int j = reflect_src_t(hp) * 7 + reflect_src_t(wp);
        act_store(p.s0, (long long)n * 81 + pos, 512 + i, i < 49 ? Gs[i * 49 + j] : 0.f);
    }
    if (p.ss_space)
        for (int o = tid; o < 49 * 49; o += 512) p.ss_space[(long long)n * 2401 + o] = Gs[o];
    __syncthreads();                      // W8s / b8s visible

    // ---- channel-rectifier chain, one thread per channel row c ----
    const int c = tid;
    float h[32], g[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { h[j] = misc[j]; g[j] = 0.f; }
    for (int hw = 0; hw < 49; ++hw) {
        const float xv = xs[c * XS + hw];
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
    const long long crow = (long long)n * 512 + c;
    {
        const float ic = inv_c[c];
        const float s1 = p.slope1[c], s4 = p.slope4[c], s7 = p.slope7[c];
        float4* o0 = reinterpret_cast<float4*>(p.g0 + crow * 32);
        float4* o1 = reinterpret_cast<float4*>(p.g1 + crow * 32);
        float4* o2 = reinterpret_cast<float4*>(p.g2 + crow * 32);
#pragma unroll
        for (int j = 0; j < 32; ++j) h[j] = fmaf(ic, g[j], h[j]);                 // g0
#pragma unroll
        for (int q = 0; q < 8; ++q) o0[q] = make_float4(h[q * 4], h[q * 4 + 1], h[q * 4 + 2], h[q * 4 + 3]);
#pragma unroll
        for (int j = 0; j < 32; ++j) h[j] = h[j] > 0.f ? h[j] : h[j] * s1;        // h1
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
            g[j] = a;                                                              // g1
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) o1[q] = make_float4(g[q * 4], g[q * 4 + 1], g[q * 4 + 2], g[q * 4 + 3]);
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = g[j] > 0.f ? g[j] : g[j] * s4;        // h4
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
            h[j] = a;                                                              // g2
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) o2[q] = make_float4(h[q * 4], h[q * 4 + 1], h[q * 4 + 2], h[q * 4 + 3]);
#pragma unroll
        for (int j = 0; j < 32; ++j) h[j] = h[j] > 0.f ? h[j] : h[j] * s7;        // h7
        uint4* ob = reinterpret_cast<uint4*>(p.h7b + crow * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            ob[q] = make_uint4(pack_bf16x2(h[q * 8 + 0], h[q * 8 + 1]), pack_bf16x2(h[q * 8 + 2], h[q * 8 + 3]),
                               pack_bf16x2(h[q * 8 + 4], h[q * 8 + 5]), pack_bf16x2(h[q * 8 + 6], h[q * 8 + 7]));
        ob[4] = make_uint4(pack_bf16x2(1.0f, 0.f), 0u, 0u, 0u);                   // ones column -> bias gradient
#pragma unroll
        for (int q = 5; q < 8; ++q) ob[q] = make_uint4(0u, 0u, 0u, 0u);
    }

    // ---- M_channel[c][j] = sigmoid(h7[c] . W8[j] + b8[j]) (recnet.py:385-386, :406), written as fp16 [hi | lo]:
    //      feat_channel = M_channel @ X (:410) runs on the tcgen05 GEMM as three K = 512 taps (hi.hi + hi.lo + lo.hi) ----
    __half* mrow = p.mch2 + crow * 1024;
#pragma unroll 1
    for (int j0 = 0; j0 < 512; j0 += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            float m2[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int j = j0 + e + u;
                const float4* wr = reinterpret_cast<const float4*>(W8s + j * 32);
                float a0 = b8s[j], a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 w4 = wr[q];
                    a0 = fmaf(w4.x, h[q * 4 + 0], a0); a1 = fmaf(w4.y, h[q * 4 + 1], a1);
                    a2 = fmaf(w4.z, h[q * 4 + 2], a2); a3 = fmaf(w4.w, h[q * 4 + 3], a3);
                }
                m2[u] = 1.0f / (1.0f + __expf(-((a0 + a1) + (a2 + a3))));
            }
            const H16 s0 = split_h(m2[0]), s1 = split_h(m2[1]);
            hi[e >> 1] = (uint32_t)__half_as_ushort(s0.hi) | ((uint32_t)__half_as_ushort(s1.hi) << 16);
            lo[e >> 1] = (uint32_t)__half_as_ushort(s0.lo) | ((uint32_t)__half_as_ushort(s1.lo) << 16);
        }
        *reinterpret_cast<uint4*>(mrow + j0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(mrow + 512 + j0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// feat_channel rows [n*512][64] fp32 (row = channel c, column = pixel hw; the GEMM output) -> ChannelFlipMerge input
// with the flip / cat fan-out (recnet.py:416-417): slot [512,1024) <- feat_channel, slot [0,512) <- W-flipped, each with
// its reflection mirrors. grid (8 channel chunks of 64, n).
__global__ void __launch_bounds__(256) fc_scatter_kernel(const float* __restrict__ fcraw, const ActDst fm) {
    __shared__ float tile[49][65];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 64 * 64; i += 256) {
        const int c = i >> 6, hw = i & 63;
        if (hw < 49) tile[hw][c] = fcraw[((long long)n * 512 + c0 + c) * 64 + hw];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 49 * 64; i += 256) {
        const int hw = i >> 6, c = i & 63;
        const int hh = hw / 7, ww = hw - hh * 7;
        const float v = tile[hw][c];
        for_each_mirror(n, hh, ww, [&](long long row) { act_store(fm, row, 512 + c0 + c, v); });
        for_each_mirror(n, hh, 6 - ww, [&](long long row) { act_store(fm, row, c0 + c, v); });
    }
}

// ----------------------------------------------------------------------------------------------------------
// Composed maps of the channel rectifier: A = W_b W_a (32x32), cvec = W_b b_a + b_b for (a,b) = (2,3) and (5,6).
//   wa [512][32], ba [512], wb [32][512], bb [32]  ->  A [32][32] (A[i][k] = sum_j wb[i][j] wa[j][k]), cvec [32]
// ----------------------------------------------------------------------------------------------------------
struct ComposeArgs { const float* wa[2]; const float* ba[2]; const float* wb[2]; const float* bb[2]; float* A[2]; float* cv[2];
                     const float* w8; __nv_bfloat16* w8t; };   // optional: W8 [512][32] -> bf16 W8^T [64][512] (rows >= 32 zero)

// grid = 64 composition CTAs (composition q = block / 32, output row i = block % 32) + 16 CTAs for W8^T; a composition
// CTA has 33 columns (32 of A + cvec) x 8 slices of the 512-long contraction, added in a fixed order. (One CTA per
// composition with a serial 512-step dot product per thread took 62 us of pure load latency.)
__global__ void __launch_bounds__(264) chan_compose_kernel(const ComposeArgs p) {
    if (blockIdx.x >= 64) {          // K-major operand of dh7 = dM_pre @ W8 (backward)
        const int part = blockIdx.x - 64;                    // 16 parts of 2048 elements
        for (int o = part * 2048 + threadIdx.x; o < (part + 1) * 2048; o += 264) {
            const int k = o >> 9, j = o & 511;
            p.w8t[o] = __float2bfloat16_rn(k < 32 ? p.w8[j * 32 + k] : 0.f);
        }
        return;
    }
    __shared__ float red[8][33];
    const int q = blockIdx.x >> 5, i = blockIdx.x & 31;
    const int k = threadIdx.x % 33, slice = threadIdx.x / 33;        // k == 32: the cvec column
    const float* wa = p.wa[q]; const float* wb = p.wb[q] + i * 512; const float* ba = p.ba[q];
    float a = 0.f;
#pragma unroll 8
    for (int j = slice * 64; j < slice * 64 + 64; ++j) a = fmaf(wb[j], (k < 32) ? wa[j * 32 + k] : ba[j], a);
    red[slice][k] = a;
    __syncthreads();
    if (slice == 0) {
        float t = (k < 32) ? 0.f : p.bb[q][i];
#pragma unroll
        for (int s2 = 0; s2 < 8; ++s2) t += red[s2][k];
        if (k < 32) p.A[q][i * 32 + k] = t; else p.cv[q][i] = t;
    }
}

// backward of the composition: dA [32][32], dc [32] ->
//   dwb[i][j] = sum_k dA[i][k] wa[j][k] + dc[i] ba[j];  dwa[j][k] = sum_i wb[i][j] dA[i][k];  dba[j] = sum_i wb[i][j] dc[i];  dbb = dc
struct ComposeBwdArgs { const float* wa[2]; const float* ba[2]; const float* wb[2]; const float* dA[2]; const float* dc[2];
                        float* dwa[2]; float* dba[2]; float* dwb[2]; float* dbb[2]; int accumulate; };

__global__ void __launch_bounds__(512) chan_compose_bwd_kernel(const ComposeBwdArgs p) {
    __shared__ float dA[1024], dc[32];
    const int q = blockIdx.x, j = threadIdx.x;
    for (int i = threadIdx.x; i < 1024; i += 512) dA[i] = p.dA[q][i];
    if (threadIdx.x < 32) dc[threadIdx.x] = p.dc[q][threadIdx.x];
    __syncthreads();
    const float* wa = p.wa[q]; const float* wb = p.wb[q];
    float war[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) war[k] = wa[j * 32 + k];
    const float baj = p.ba[q][j];
    float dwa[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) dwa[k] = 0.f;
    float dba = 0.f;
    for (int i = 0; i < 32; ++i) {
        float s = dc[i] * baj;
#pragma unroll
        for (int k = 0; k < 32; ++k) s = fmaf(dA[i * 32 + k], war[k], s);
        float* o = p.dwb[q] + i * 512 + j;
        *o = p.accumulate ? *o + s : s;
        const float wbij = wb[i * 512 + j];
        dba = fmaf(wbij, dc[i], dba);
#pragma unroll
        for (int k = 0; k < 32; ++k) dwa[k] = fmaf(wbij, dA[i * 32 + k], dwa[k]);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        float* o = p.dwa[q] + j * 32 + k;
        *o = p.accumulate ? *o + dwa[k] : dwa[k];
    }
    { float* o = p.dba[q] + j; *o = p.accumulate ? *o + dba : dba; }
    if (j < 32) { float* o = p.dbb[q] + j; *o = p.accumulate ? *o + dc[j] : dc[j]; }
}

// ----------------------------------------------------------------------------------------------------------
// feat_space = X (512x49) @ M_space (49x49)  (recnet.py:409) -> slot [0,512) of the Conv4Merge input (+ mirrors) and an
// fp32 copy on the own rows of an H9 matrix (input of the spatial self-similarity loss).
// mspace: fp32 rows of the H9 grid, [n*81][64]; row = pixel j, column = channel i holds M_space[n, i, j].
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) feat_space_train_kernel(const float* __restrict__ x, const float* __restrict__ mspace,
                                                               const ActDst cm, float* __restrict__ fs_f32, int ldfs) {
    __shared__ float Ms[49 * 49];     // Ms[i*49 + j] = M_space[n,i,j]
    extern __shared__ __align__(16) float fst_xs[];  // [512][49]: the sample's X, staged with coalesced 16-byte loads
    const int n = blockIdx.x, tid = threadIdx.x;
    for (int o = tid; o < 49 * 49; o += 256) {
        const int j = o / 49, i = o - j * 49;
        Ms[i * 49 + j] = mspace[((long long)n * 81 + h9_row(j)) * 64 + i];
    }
    {
        const float4* src = reinterpret_cast<const float4*>(x + (long long)n * 512 * 49);
        float4* dst = reinterpret_cast<float4*>(fst_xs);
        for (int o = tid; o < 512 * 49 / 4; o += 256) dst[o] = __ldg(src + o);
    }
    __syncthreads();
    const int c2 = tid * 2;
    float xa[49], xb[49];
    const float* xr = fst_xs + c2 * 49;
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
        const int h = j / 7, w = j - h * 7;
        if (fs_f32) {
            float* o = fs_f32 + ((long long)n * 81 + (h + 1) * 9 + (w + 1)) * ldfs + c2;
            *reinterpret_cast<float2*>(o) = make_float2(a, b);
        }
        for_each_mirror(n, h, w, [&](long long row) { act_store(cm, row, c2, a); act_store(cm, row, c2 + 1, b); });
    }
}

// Backward w.r.t. M_space (X carries no gradient): dFS[c][j] = fold(dcm[.., c]) + dfs[.., c];
//   dM[i][j] = sum_c X[c][i] dFS[c][j];  dpre[i][j] = dM[i][j] * M[i][j] (1 - M[i][j])
// written to dmsp [n*81][64] (own row of pixel j, column i; columns 49..63 zero): the `dadd` source of the last
// Conv4Space layer's backward (the sigmoid follows the residual add, recnet.py:370).
// One CTA per sample, 392 threads = 49 (7x7 output blocks) x 8 channel slices; channels are staged 256 at a time.
__global__ void __launch_bounds__(448, 1)
feat_space_bwd_kernel(const float* __restrict__ x, const float* __restrict__ mspace, const float* __restrict__ dcm, int lddcm,
                      const float* __restrict__ dfs, int lddfs, float* __restrict__ dmsp) {
    extern __shared__ __align__(16) float fsm[];
    float* xs = fsm;                    // [256][49]
    float* ds = xs + 256 * 49;          // [49][257]  dFS[j][c] (pitch 257: conflict-free column walks)
    float* part = ds + 49 * 257;        // [8][2401]
    const int n = blockIdx.x, tid = threadIdx.x;
    float acc[7][7];
#pragma unroll
    for (int a = 0; a < 7; ++a)
#pragma unroll
        for (int q = 0; q < 7; ++q) acc[a][q] = 0.f;
    const int item = tid % 49, split = tid / 49;      // split 0..8 (only < 8 compute)
    const int bi = (item / 7) * 7, bj = (item % 7) * 7;
    for (int half = 0; half < 2; ++half) {
        const int cb = half * 256;
        __syncthreads();
        for (int i = tid; i < 256 * 49; i += 448) xs[i] = x[((long long)n * 512 + cb) * 49 + i];
        // four elements per step, all of their (up to five) loads issued before the first add: the fold is a chain of
        // dependent adds and with one element per step the kernel spent most of its time on one load latency after the
        // other. Same order of additions per element (own row, row mirror, column mirror, corner mirror).
        for (int i0 = tid; i0 < 49 * 256; i0 += 4 * 448) {
            float t[4][5];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 448;
#pragma unroll
                for (int e = 0; e < 5; ++e) t[u][e] = 0.f;
                if (i < 49 * 256) {
                    const int j = i >> 8, c = i & 255;
                    const int h = j / 7, w = j - h * 7;
                    const int mh = (h == 1) ? -2 : ((h == 5) ? 2 : 0);
                    const int mw = (w == 1) ? -2 : ((w == 5) ? 2 : 0);
                    const long long base = (long long)n * 81 + (h + 1) * 9 + (w + 1);
                    if (dfs) t[u][0] = __ldg(dfs + base * lddfs + cb + c);
                    if (dcm) {
                        const float* q = dcm + cb + c;
                        t[u][1] = __ldg(q + base * lddcm);
                        if (mh) t[u][2] = __ldg(q + (base + mh * 9) * lddcm);
                        if (mw) t[u][3] = __ldg(q + (base + mw) * lddcm);
                        if (mh && mw) t[u][4] = __ldg(q + (base + mh * 9 + mw) * lddcm);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 448;
                if (i < 49 * 256) {
                    const int j = i >> 8, c = i & 255;
                    const int h = j / 7, w = j - h * 7;
                    const bool mh = (h == 1) || (h == 5), mw = (w == 1) || (w == 5);
                    float v = t[u][0];
                    if (dcm) {
                        v += t[u][1];
                        if (mh) v += t[u][2];
                        if (mw) v += t[u][3];
                        if (mh && mw) v += t[u][4];
                    }
                    ds[j * 257 + c] = v;
                }
            }
        }
        __syncthreads();
        if (split < 8) {
#pragma unroll 2
            for (int c = split * 32; c < split * 32 + 32; ++c) {
                float va[7], vb[7];
#pragma unroll
                for (int q = 0; q < 7; ++q) { va[q] = xs[c * 49 + bi + q]; vb[q] = ds[(bj + q) * 257 + c]; }
#pragma unroll
                for (int a = 0; a < 7; ++a)
#pragma unroll
                    for (int q = 0; q < 7; ++q) acc[a][q] = fmaf(va[a], vb[q], acc[a][q]);
            }
        }
    }
    if (split < 8) {
        float* gp = part + split * 2401;
#pragma unroll
        for (int a = 0; a < 7; ++a)
#pragma unroll
            for (int q = 0; q < 7; ++q) gp[(bi + a) * 49 + bj + q] = acc[a][q];
    }
    __syncthreads();
    for (int o = tid; o < 49 * 64; o += 448) {
        const int j = o >> 6, i = o & 63;
        float v = 0.f;
        const long long row = (long long)n * 81 + h9_row(j);
        if (i < 49) {
            float g = 0.f;
#pragma unroll
            for (int sp = 0; sp < 8; ++sp) g += part[sp * 2401 + i * 49 + j];
            const float m = mspace[row * 64 + i];
            v = g * m * (1.0f - m);
        }
        dmsp[row * 64 + i] = v;
    }
}

// ----------------------------------------------------------------------------------------------------------
// Gradient of the flip / cat / reflection fan-out of feat_channel (recnet.py:416-417 + ReflectionPad2d):
//   dfc[c][(h,w)] = fold(dfm[(h,w)][512 + c]) + fold(dfm[(h,6-w)][c])     (fold = own row + mirror rows)
// dfm: fp32 [n*81][lddfm] on the H9 grid (dgrad of ChannelFlipMerge.0); dfc: bf16 [n*512][64] (49 valid columns), the
// A operand of dM = dFC X^T. grid (8 channel chunks of 64, n).
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fc_bwd_gather_kernel(const float* __restrict__ dfm, int lddfm,
                                                            __nv_bfloat16* __restrict__ dfc) {
    __shared__ float tile[49][65];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    for (int i = threadIdx.x; i < 49 * 64; i += 256) {
        const int pix = i >> 6, c = i & 63;
        const int h = pix / 7, w = pix - h * 7;
        // the (up to eight) loads of both folds first, then the adds in the order own row, row mirror, column mirror,
        // corner mirror (a chain of load-dependent adds costs one memory latency per term)
        float t[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        bool mhs[2], mws[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int we = e == 0 ? w : 6 - w;
            const int mh = (h == 1) ? -2 : ((h == 5) ? 2 : 0);
            const int mw = (we == 1) ? -2 : ((we == 5) ? 2 : 0);
            mhs[e] = mh != 0; mws[e] = mw != 0;
            const long long base = (long long)n * 81 + (h + 1) * 9 + (we + 1);
            const float* q = dfm + (e == 0 ? 512 : 0) + c0 + c;
            t[e][0] = __ldg(q + base * lddfm);
            if (mh) t[e][1] = __ldg(q + (base + mh * 9) * lddfm);
            if (mw) t[e][2] = __ldg(q + (base + mw) * lddfm);
            if (mh && mw) t[e][3] = __ldg(q + (base + mh * 9 + mw) * lddfm);
        }
        float v = 0.f;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            v += t[e][0];
            if (mhs[e]) v += t[e][1];
            if (mws[e]) v += t[e][2];
            if (mhs[e] && mws[e]) v += t[e][3];
        }
        tile[pix][c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 32; i += 256) {
        const int c = i >> 5, hp = (i & 31) * 2;
        const float a = hp < 49 ? tile[hp][c] : 0.f, b = hp + 1 < 49 ? tile[hp + 1][c] : 0.f;
        *reinterpret_cast<uint32_t*>(dfc + ((long long)n * 512 + c0 + c) * 64 + hp) = pack_bf16x2(a, b);
    }
}

// ----------------------------------------------------------------------------------------------------------
// Backward of the thin Conv4Channel chain for one sample per CTA (512 threads = channel rows):
//   dh7 [n*512][64] fp32 (columns 0..31) from the tcgen05 GEMM dM_pre @ W8; g0/g1/g2 saved by the forward kernel.
// Per-sample partial parameter gradients go to part[n][CH_PART] (fixed-order reduction by chan_reduce_kernel):
//   [0,1024) dA2, [1024,1056) dc2, [1056,2080) dA1, [2080,2112) dc1, [2112,2144) db0, [2144, 2144+32*561) dW0
// and the per-row PReLU slope gradients to dslope_part[n][3][512] (slope1, slope4, slope7).
// ----------------------------------------------------------------------------------------------------------
constexpr int CH_PART = 1024 + 32 + 1024 + 32 + 32 + 32 * 561;   // 20096

struct ChanBwd {
    const float* x; const float* dh7; const float* g0; const float* g1; const float* g2;
    const float* inv_c; const float* tmat;
    const float* A1; const float* A2;
    const float* slope1; const float* slope4; const float* slope7;
    float* part; float* dslope_part;
};

__global__ void __launch_bounds__(512, 1) chan_bwd_kernel(const ChanBwd p) {
    extern __shared__ __align__(16) float csm[];
    float* xs = csm;                    // [512][49]
    float* bufB = xs + 512 * 49;        // [512][32] right operand (h4, h1) / dg0
    float* bufA = bufB + 512 * 32;      // [512][8]  left operand chunk
    float* A1s = bufA + 512 * 8;        // [32][32]
    float* A2s = A1s + 1024;
    float* dT = A2s + 1024;             // [49][32]
    float* invc = dT + 1568;            // [512]
    const int n = blockIdx.x, tid = threadIdx.x, c = tid;
    const long long crow = (long long)n * 512 + c;
    float* part = p.part + (long long)n * CH_PART;
    for (int i = tid; i < 512 * 49; i += 512) xs[i] = p.x[(long long)n * 512 * 49 + i];
    for (int i = tid; i < 1024; i += 512) { A1s[i] = p.A1[i]; A2s[i] = p.A2[i]; }
    invc[tid] = p.inv_c[crow];

    float d[32], a[32];                 // running gradient / saved pre-activation
    {
        const float4* s = reinterpret_cast<const float4*>(p.dh7 + crow * 64);
        const float4* gq = reinterpret_cast<const float4*>(p.g2 + crow * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = s[q], w = gq[q];
            d[q * 4] = v.x; d[q * 4 + 1] = v.y; d[q * 4 + 2] = v.z; d[q * 4 + 3] = v.w;
            a[q * 4] = w.x; a[q * 4 + 1] = w.y; a[q * 4 + 2] = w.z; a[q * 4 + 3] = w.w;
        }
    }
    __syncthreads();

    // generic step: d = gradient w.r.t. PReLU output h = prelu(a; slope); returns dg (in d) and the slope gradient
    auto prelu_bwd = [&](float slope) {
        float ds = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            ds = fmaf(d[k], fminf(a[k], 0.f), ds);
            d[k] = a[k] > 0.f ? d[k] : d[k] * slope;
        }
        return ds;
    };
    // outer-product reduction over the 512 rows: out[i][k] = sum_c dg[c][i] * hprev[c][k]; dsum[i] = sum_c dg[c][i].
    // hprev is in bufB (all 32 columns), dg is staged 8 columns at a time in bufA.
    auto outer_reduce = [&](float* out_mat, float* out_vec) {
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += 8) {
            __syncthreads();
#pragma unroll
            for (int e = 0; e < 8; ++e) bufA[c * 8 + e] = d[i0 + e];
            __syncthreads();
            if (tid < 256) {
                const int i = tid >> 5, k = tid & 31;
                float s = 0.f;
                for (int r = 0; r < 512; ++r) s = fmaf(bufA[r * 8 + i], bufB[r * 32 + k], s);
                out_mat[(i0 + i) * 32 + k] = s;
            } else if (tid < 264) {
                const int i = tid - 256;
                float s = 0.f;
                for (int r = 0; r < 512; ++r) s += bufA[r * 8 + i];
                out_vec[i0 + i] = s;
            }
        }
        __syncthreads();
    };

    // ---- layer 7: h7 = prelu(g2; slope7), g2 = A2 h4 + c2, h4 = prelu(g1; slope4) ----
    const float ds7 = prelu_bwd(p.slope7[c]);                    // d = dg2
    {
        const float4* gq = reinterpret_cast<const float4*>(p.g1 + crow * 32);
        const float s4 = p.slope4[c];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 w = gq[q];
            a[q * 4] = w.x; a[q * 4 + 1] = w.y; a[q * 4 + 2] = w.z; a[q * 4 + 3] = w.w;
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) bufB[c * 32 + k] = a[k] > 0.f ? a[k] : a[k] * s4;     // h4
    }
    outer_reduce(part, part + 1024);                             // dA2, dc2
    {
        float t[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) t[k] = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float di = d[i];
            const float4* ar = reinterpret_cast<const float4*>(A2s + i * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a4 = ar[q];
                t[q * 4] = fmaf(a4.x, di, t[q * 4]); t[q * 4 + 1] = fmaf(a4.y, di, t[q * 4 + 1]);
                t[q * 4 + 2] = fmaf(a4.z, di, t[q * 4 + 2]); t[q * 4 + 3] = fmaf(a4.w, di, t[q * 4 + 3]);
            }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) d[k] = t[k];                // dh4
    }
    // ---- layer 4: h4 = prelu(g1; slope4), g1 = A1 h1 + c1, h1 = prelu(g0; slope1) ----
    const float ds4 = prelu_bwd(p.slope4[c]);                    // d = dg1
    {
        const float4* gq = reinterpret_cast<const float4*>(p.g0 + crow * 32);
        const float s1 = p.slope1[c];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 w = gq[q];
            a[q * 4] = w.x; a[q * 4 + 1] = w.y; a[q * 4 + 2] = w.z; a[q * 4 + 3] = w.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; ++k) bufB[c * 32 + k] = a[k] > 0.f ? a[k] : a[k] * s1;     // h1
    }
    outer_reduce(part + 1056, part + 2080);                      // dA1, dc1
    {
        float t[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) t[k] = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float di = d[i];
            const float4* ar = reinterpret_cast<const float4*>(A1s + i * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a4 = ar[q];
                t[q * 4] = fmaf(a4.x, di, t[q * 4]); t[q * 4 + 1] = fmaf(a4.y, di, t[q * 4 + 1]);
                t[q * 4 + 2] = fmaf(a4.z, di, t[q * 4 + 2]); t[q * 4 + 3] = fmaf(a4.w, di, t[q * 4 + 3]);
            }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) d[k] = t[k];                // dh1
    }
    // ---- layer 1: h1 = prelu(g0; slope1), g0 = X W0a^T + inv_c (X T) + b0 ----
    const float ds1 = prelu_bwd(p.slope1[c]);                    // d = dg0
    p.dslope_part[((long long)n * 3 + 0) * 512 + c] = ds1;
    p.dslope_part[((long long)n * 3 + 1) * 512 + c] = ds4;
    p.dslope_part[((long long)n * 3 + 2) * 512 + c] = ds7;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) bufB[c * 32 + k] = d[k];        // dg0 for everyone
    __syncthreads();
    float* dW0 = part + 2144;                                    // [32][561]
    if (tid < 392) {          // (pixel hw, 4 consecutive k): dW0a[k][hw] = sum_c X[c][hw] dg0[c][k]; dT[hw][k] likewise with Xh
        const int hw = tid >> 3, k4 = (tid & 7) * 4;
        float wa0 = 0.f, wa1 = 0.f, wa2 = 0.f, wa3 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
        for (int r = 0; r < 512; ++r) {
            const float xv = xs[r * 49 + hw];
            const float xh = xv * invc[r];
            const float4 g4 = *reinterpret_cast<const float4*>(bufB + r * 32 + k4);
            wa0 = fmaf(xv, g4.x, wa0); wa1 = fmaf(xv, g4.y, wa1); wa2 = fmaf(xv, g4.z, wa2); wa3 = fmaf(xv, g4.w, wa3);
            t0 = fmaf(xh, g4.x, t0); t1 = fmaf(xh, g4.y, t1); t2 = fmaf(xh, g4.z, t2); t3 = fmaf(xh, g4.w, t3);
        }
        dW0[(k4 + 0) * 561 + hw] = wa0; dW0[(k4 + 1) * 561 + hw] = wa1;
        dW0[(k4 + 2) * 561 + hw] = wa2; dW0[(k4 + 3) * 561 + hw] = wa3;
        dT[hw * 32 + k4 + 0] = t0; dT[hw * 32 + k4 + 1] = t1; dT[hw * 32 + k4 + 2] = t2; dT[hw * 32 + k4 + 3] = t3;
    } else if (tid < 424) {   // db0[k] = sum_c dg0[c][k]
        const int k = tid - 392;
        float s = 0.f;
        for (int r = 0; r < 512; ++r) s += bufB[r * 32 + k];
        part[2112 + k] = s;
    }
    __syncthreads();
    {   // dW0b[k][c] = sum_hw Xh[c][hw] dT[hw][k]   (T = Xh^T W0b^T)
        float o[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) o[k] = 0.f;
        const float ic = invc[c];
        for (int hw = 0; hw < 49; ++hw) {
            const float xh = xs[c * 49 + hw] * ic;
            const float4* tr = reinterpret_cast<const float4*>(dT + hw * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 t4 = tr[q];
                o[q * 4] = fmaf(xh, t4.x, o[q * 4]); o[q * 4 + 1] = fmaf(xh, t4.y, o[q * 4 + 1]);
                o[q * 4 + 2] = fmaf(xh, t4.z, o[q * 4 + 2]); o[q * 4 + 3] = fmaf(xh, t4.w, o[q * 4 + 3]);
            }
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) dW0[k * 561 + 49 + c] = o[k];
    }
}

// out[i] (+)= sum_n part[n][i] in a fixed order; two segments: the CH_PART block and the slope block [3][512].
__global__ void __launch_bounds__(256) chan_reduce_kernel(const float* __restrict__ part, const float* __restrict__ dslope_part,
                                                          int n_img, float* __restrict__ tmp /*[2112]: dA2 dc2 dA1 dc1*/,
                                                          float* __restrict__ db0, float* __restrict__ dW0,
                                                          float* __restrict__ ds1, float* __restrict__ ds4,
                                                          float* __restrict__ ds7, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < CH_PART) {
        float s = 0.f;
        for (int n = 0; n < n_img; ++n) s += part[(long long)n * CH_PART + i];
        if (i < 2112) tmp[i] = s;
        else {
            float* o = (i < 2144) ? db0 + (i - 2112) : dW0 + (i - 2144);
            *o = accumulate ? *o + s : s;
        }
    } else if (i < CH_PART + 1536) {
        const int k = i - CH_PART;
        float s = 0.f;
        for (int n = 0; n < n_img; ++n) s += dslope_part[(long long)n * 1536 + k];
        float* o = (k < 512) ? ds1 + k : (k < 1024 ? ds4 + (k - 512) : ds7 + (k - 1024));
        *o = accumulate ? *o + s : s;
    }
}

}  // namespace ffr

using namespace ffr;
static inline cudaStream_t S_(ffr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline ActDst mk_dst(void* h, int ldh, int lo_off, void* b, int ldb) {
    ActDst d;
    d.h = reinterpret_cast<__half*>(h); d.ldh = ldh; d.lo_off = lo_off;
    d.b = reinterpret_cast<__nv_bfloat16*>(b); d.ldb = ldb;
    return d;
}

extern "C" {

FFR_API int ffr_recnet_prep_train(const ffr_prep_train_desc* d, int n, ffr_stream_t stream) {
    FFR_CHECK_ARG(d && d->x && d->w0 && d->b0 && d->slope1 && d->slope4 && d->slope7 && d->A1 && d->c1 && d->A2 && d->c2 &&
                  d->w8 && d->b8 && d->s0_h && d->cm_h && d->g0 && d->g1 && d->g2 && d->h7b && d->xk && d->mch2 && d->x3 &&
                  d->inv_c && d->tmat, "ffr_recnet_prep_train: null pointer");
    if (n == 0) return 0;
    PrepTrain p;
    p.x = d->x; p.w0 = d->w0; p.b0 = d->b0; p.slope1 = d->slope1; p.slope4 = d->slope4; p.slope7 = d->slope7;
    p.A1 = d->A1; p.c1 = d->c1; p.A2 = d->A2; p.c2 = d->c2; p.w8 = d->w8; p.b8 = d->b8;
    p.s0 = mk_dst(d->s0_h, d->s0_ld, d->s0_lo, d->s0_b, d->s0_ldb);
    p.cm = mk_dst(d->cm_h, d->cm_ld, d->cm_lo, d->cm_b, d->cm_ldb);
    p.g0 = d->g0; p.g1 = d->g1; p.g2 = d->g2;
    p.h7b = reinterpret_cast<__nv_bfloat16*>(d->h7b); p.xk = reinterpret_cast<__nv_bfloat16*>(d->xk);
    p.mch2 = reinterpret_cast<__half*>(d->mch2); p.x3 = reinterpret_cast<__half*>(d->x3);
    p.inv_c = d->inv_c; p.tmat = d->tmat; p.ss_space = d->ss_space;
    const int smem = (512 * XS + 512 + 64 + 49 * 49 + 3 + 49 * 32 * 2 + 2048 + 96 + 8 * 2401) * (int)sizeof(float);
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(recnet_prep_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    recnet_prep_train_kernel<<<n, 512, smem, S_(stream)>>>(p);
    return launch_status("recnet_prep_train_kernel");
}

FFR_API int ffr_fc_scatter(const float* fcraw, void* fm_h, int fm_ld, int fm_lo, void* fm_b, int fm_ldb, int n,
                           ffr_stream_t stream) {
    FFR_CHECK_ARG(fcraw && fm_h, "ffr_fc_scatter: null pointer");
    if (n == 0) return 0;
    fc_scatter_kernel<<<dim3(8, n), 256, 0, S_(stream)>>>(fcraw, mk_dst(fm_h, fm_ld, fm_lo, fm_b, fm_ldb));
    return launch_status("fc_scatter_kernel");
}

FFR_API int ffr_chan_compose(const float* w2, const float* b2, const float* w3, const float* b3, const float* w5,
                             const float* b5, const float* w6, const float* b6, float* A1, float* c1, float* A2,
                             float* c2, const float* w8, void* w8t, ffr_stream_t stream) {
    FFR_CHECK_ARG(w2 && b2 && w3 && b3 && w5 && b5 && w6 && b6 && A1 && c1 && A2 && c2, "ffr_chan_compose: null pointer");
    FFR_CHECK_ARG(!w8t || w8, "ffr_chan_compose: w8t without w8");
    ComposeArgs p;
    p.w8 = w8; p.w8t = reinterpret_cast<__nv_bfloat16*>(w8t);
    p.wa[0] = w2; p.ba[0] = b2; p.wb[0] = w3; p.bb[0] = b3; p.A[0] = A1; p.cv[0] = c1;
    p.wa[1] = w5; p.ba[1] = b5; p.wb[1] = w6; p.bb[1] = b6; p.A[1] = A2; p.cv[1] = c2;
    chan_compose_kernel<<<w8t ? 80 : 64, 264, 0, S_(stream)>>>(p);
    return launch_status("chan_compose_kernel");
}

FFR_API int ffr_chan_compose_bwd(const float* w2, const float* b2, const float* w3, const float* w5, const float* b5,
                                 const float* w6, const float* dA1, const float* dc1, const float* dA2, const float* dc2,
                                 float* dw2, float* db2, float* dw3, float* db3, float* dw5, float* db5, float* dw6,
                                 float* db6, int accumulate, ffr_stream_t stream) {
    FFR_CHECK_ARG(w2 && b2 && w3 && w5 && b5 && w6 && dA1 && dc1 && dA2 && dc2 && dw2 && db2 && dw3 && db3 && dw5 && db5 &&
                  dw6 && db6, "ffr_chan_compose_bwd: null pointer");
    ComposeBwdArgs p;
    p.wa[0] = w2; p.ba[0] = b2; p.wb[0] = w3; p.dA[0] = dA1; p.dc[0] = dc1;
    p.dwa[0] = dw2; p.dba[0] = db2; p.dwb[0] = dw3; p.dbb[0] = db3;
    p.wa[1] = w5; p.ba[1] = b5; p.wb[1] = w6; p.dA[1] = dA2; p.dc[1] = dc2;
    p.dwa[1] = dw5; p.dba[1] = db5; p.dwb[1] = dw6; p.dbb[1] = db6;
    p.accumulate = accumulate;
    chan_compose_bwd_kernel<<<2, 512, 0, S_(stream)>>>(p);
    return launch_status("chan_compose_bwd_kernel");
}

FFR_API int ffr_feat_space_train(const float* x, const float* mspace, void* cm_h, int cm_ld, int cm_lo, void* cm_b,
                                 int cm_ldb, float* fs_f32, int ldfs, int n, ffr_stream_t stream) {
    FFR_CHECK_ARG(x && mspace && cm_h, "ffr_feat_space_train: null pointer");
    if (n == 0) return 0;
    const int smem = 512 * 49 * (int)sizeof(float);
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(feat_space_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    feat_space_train_kernel<<<n, 256, smem, S_(stream)>>>(x, mspace, mk_dst(cm_h, cm_ld, cm_lo, cm_b, cm_ldb), fs_f32, ldfs);
    return launch_status("feat_space_train_kernel");
}

FFR_API int ffr_feat_space_bwd(const float* x, const float* mspace, const float* dcm, int lddcm, const float* dfs,
                               int lddfs, float* dmsp, int n, ffr_stream_t stream) {
    FFR_CHECK_ARG(x && mspace && (dcm || dfs) && dmsp, "ffr_feat_space_bwd: null pointer");
    if (n == 0) return 0;
    const int smem = (256 * 49 + 49 * 257 + 8 * 2401) * (int)sizeof(float);
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(feat_space_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    feat_space_bwd_kernel<<<n, 448, smem, S_(stream)>>>(x, mspace, dcm, lddcm, dfs, lddfs, dmsp);
    return launch_status("feat_space_bwd_kernel");
}

FFR_API int ffr_fc_bwd_gather(const float* dfm, int lddfm, void* dfc, int n, ffr_stream_t stream) {
    FFR_CHECK_ARG(dfm && dfc && lddfm >= 1024, "ffr_fc_bwd_gather: bad arguments");
    if (n == 0) return 0;
    fc_bwd_gather_kernel<<<dim3(8, n), 256, 0, S_(stream)>>>(dfm, lddfm, reinterpret_cast<__nv_bfloat16*>(dfc));
    return launch_status("fc_bwd_gather_kernel");
}

FFR_API int ffr_chan_bwd_part_floats(void) { return CH_PART; }

FFR_API int ffr_chan_bwd(const float* x, const float* dh7, const float* g0, const float* g1, const float* g2,
                         const float* inv_c, const float* tmat, const float* A1, const float* A2, const float* slope1,
                         const float* slope4, const float* slope7, float* part, float* dslope_part, float* tmp,
                         float* db0, float* dW0, float* dslope1, float* dslope4, float* dslope7, int accumulate, int n,
                         ffr_stream_t stream) {
    FFR_CHECK_ARG(x && dh7 && g0 && g1 && g2 && inv_c && tmat && A1 && A2 && slope1 && slope4 && slope7 && part &&
                  dslope_part && tmp && db0 && dW0 && dslope1 && dslope4 && dslope7, "ffr_chan_bwd: null pointer");
    if (n == 0) return 0;
    ChanBwd p{x, dh7, g0, g1, g2, inv_c, tmat, A1, A2, slope1, slope4, slope7, part, dslope_part};
    const int smem = (512 * 49 + 512 * 32 + 512 * 8 + 2048 + 1568 + 512) * (int)sizeof(float);
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(chan_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    chan_bwd_kernel<<<n, 512, smem, S_(stream)>>>(p);
    int rc = launch_status("chan_bwd_kernel");
    if (rc) return rc;
    const int total = CH_PART + 1536;
    chan_reduce_kernel<<<(total + 255) / 256, 256, 0, S_(stream)>>>(part, dslope_part, n, tmp, db0, dW0, dslope1, dslope4,
                                                                    dslope7, accumulate);
    return launch_status("chan_reduce_kernel");
}

}  // extern "C"
