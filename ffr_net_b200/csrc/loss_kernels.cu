// Loss side of the FFR-Net training step (models/trainer.py:31-43 TripletLoss, :154-178 Trainer.backward), forward AND
// gradient in hand-written kernels — no (N,512,512) tensor is ever materialised in fp32:
//   L1 = 1/2 (1/2 [MSE(ss_space(X), ss_space(space_non)) + MSE(ss_space(X), ss_space(space_ocl))]
//           + 1/2 [MSE(ss_chan(X),  ss_chan(channel_non)) + MSE(ss_chan(X),  ss_chan(channel_ocl))])     (:157-165)
//   L2 = mean relu((1 - cos(f_ocl, e_non)) - (1 - cos(f_ocl, e_ocl)) + 0.1)                               (:167-169)
//   L3 = 1/2 [MSE(f_non, e_non) + MSE(f_ocl, e_non)]                                                      (:171)
// X = feat_map_non (the frozen backbone's map of the unmasked image, no gradient), e_* = backbone embeddings.
// Batches: G groups (calls) of n samples, group 0 = unmasked ("non"), group 1 = masked ("ocl"); sample s of any group
// compares against X[s mod n].
//
// Channel Gram (512x512 per sample, contraction over the 49 pixels): on the tcgen05 GEMM. The pack kernel writes the
// row-normalised F^ and X^ as bf16 hi/lo splits, K-concatenated so that ONE accumulator holds D = F^F^T - X^X^T:
//   A = [ F^hi | F^lo | F^hi | -X^hi | -X^lo | -X^hi ],  B = [ F^hi | F^hi | F^lo | X^hi | X^hi | X^lo ]   (6 x 64 columns)
// The GEMM epilogue stores D (bf16) and deterministic per-tile sums of D^2 (the loss); a second GEMM forms D F^
// (the gradient w.r.t. F^ up to a factor) and selfsim_channel_bwd_kernel applies the normalisation Jacobian.
// Spatial Gram (49x49 per sample, contraction over 512 channels): one fused SIMT kernel per sample.
#include "../../include/ffr_sm100.h"
#include "host.h"
#include "ptx.cuh"

namespace ffr {

__device__ __forceinline__ float warp_sum_l(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int h9_row_l(int pix) { return (pix / 7 + 1) * 9 + (pix % 7 + 1); }

// block-wide sum (blockDim.x <= 1024, all threads call); result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* scratch /*[32]*/) {
    v = warp_sum_l(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        t = lane < nw ? scratch[lane] : 0.f;
        t = warp_sum_l(t);
    }
    return t;
}

// ----------------------------------------------------------------------------------------------------------
// Channel self-similarity: pack. CTA = (sample, 64-channel chunk); the chunk of F and of the target X goes through
// shared memory so that every global access is coalesced (rows of A6 / B6 are 768 bytes each, consecutive rows are
// contiguous).
//   f: fp32 H9 matrix [n_img*81][ldf], own rows valid (feat_channel of the RecNet call); x: [n][512][49] targets.
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) selfsim_channel_pack_kernel(const float* __restrict__ f, int ldf,
                                                                   const float* __restrict__ x, int n_per_group,
                                                                   __nv_bfloat16* __restrict__ A6,
                                                                   __nv_bfloat16* __restrict__ B6,
                                                                   __nv_bfloat16* __restrict__ FhT,
                                                                   float* __restrict__ inv_f) {
    __shared__ float fs[64][50], xs[64][50];     // [channel][pixel] (+1 pad column: conflict-free row walks)
    __shared__ float invf[64], invx[64];
    const int s = blockIdx.x, c0 = blockIdx.y * 64, tid = threadIdx.x;
    const int t = s % n_per_group;
    for (int i = tid; i < 49 * 64; i += 256) {
        const int hw = i >> 6, c = i & 63;
        fs[c][hw] = f[((long long)s * 81 + h9_row_l(hw)) * ldf + c0 + c];
    }
    const float* xc = x + ((long long)t * 512 + c0) * 49;
    for (int i = tid; i < 64 * 49; i += 256) xs[i / 49][i % 49] = __ldg(xc + i);
    __syncthreads();
    {   // row norms: 4 threads per channel row
        const int c = tid >> 2, q = tid & 3;
        float sf = 0.f, sx = 0.f;
        for (int hw = q; hw < 49; hw += 4) { sf = fmaf(fs[c][hw], fs[c][hw], sf); sx = fmaf(xs[c][hw], xs[c][hw], sx); }
        sf += __shfl_xor_sync(0xffffffffu, sf, 1); sf += __shfl_xor_sync(0xffffffffu, sf, 2);
        sx += __shfl_xor_sync(0xffffffffu, sx, 1); sx += __shfl_xor_sync(0xffffffffu, sx, 2);
        if (q == 0) {
            const float a = 1.0f / fmaxf(sqrtf(sf), 1e-12f);
            invf[c] = a;
            invx[c] = 1.0f / fmaxf(sqrtf(sx), 1e-12f);
            inv_f[(long long)s * 512 + c0 + c] = a;
        }
    }
    __syncthreads();
    uint32_t* a = reinterpret_cast<uint32_t*>(A6 + ((long long)s * 512 + c0) * 384);
    uint32_t* b = reinterpret_cast<uint32_t*>(B6 + ((long long)s * 512 + c0) * 384);
    for (int w = tid; w < 64 * 192; w += 256) {            // 192 words per row = 6 chunks x 32 words (64 columns each)
        const int c = w / 192, cw = w - c * 192;
        const int chunk = cw >> 5, k2 = (cw & 31) * 2;
        const bool is_x = chunk >= 3;
        const float iv = is_x ? invx[c] : invf[c];
        const float* src = is_x ? xs[c] : fs[c];
        const float v0 = (k2 < 49) ? src[k2] * iv : 0.f, v1 = (k2 + 1 < 49) ? src[k2 + 1] * iv : 0.f;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        const uint32_t HI = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        const uint32_t LO = pack_bf16x2(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
        // A = [Fh | Fl | Fh | -Xh | -Xl | -Xh],  B = [Fh | Fh | Fl | Xh | Xh | Xl]
        const int sub = is_x ? chunk - 3 : chunk;
        uint32_t av = (sub == 1) ? LO : HI;
        const uint32_t bv = (sub == 2) ? LO : HI;
        if (is_x) av ^= 0x80008000u;                        // sign flip of both halves
        a[w] = av;
        b[w] = bv;
    }
    // F^T (hi part): [s*64 + hw][c0 + c], rows 49..63 zero
    for (int i = tid; i < 64 * 64; i += 256) {
        const int hw = i >> 6, c = i & 63;
        FhT[((long long)s * 64 + hw) * 512 + c0 + c] = __float2bfloat16_rn(hw < 49 ? fs[c][hw] * invf[c] : 0.f);
    }
}

// Gradient finish: e = D F^ (fp32 [n_img*512][64], columns 0..48) ->
//   dF^ = coef * e;  dF[c] = inv_f[c] * (dF^[c] - F^[c] (F^[c] . dF^[c]))   written to df (fp32 H9 own rows, pitch lddf)
// coef = 4 * weight / (n_per_group * 512 * 512): d/dG of weight * mean((G - T)^2) is 2 weight D / numel, and
// dF^ = (dG + dG^T) F^ = 2 dG F^ because D is symmetric.
__global__ void __launch_bounds__(256) selfsim_channel_bwd_kernel(const float* __restrict__ e, const float* __restrict__ f,
                                                                  int ldf, const float* __restrict__ inv_f, float coef,
                                                                  float* __restrict__ df, int lddf) {
    const int s = blockIdx.x, c = blockIdx.y * 256 + threadIdx.x;
    const long long row = (long long)s * 512 + c;
    const float iv = inv_f[row];
    float fh[49], g[49];
    float dot = 0.f;
#pragma unroll
    for (int hw = 0; hw < 49; ++hw) {
        fh[hw] = f[((long long)s * 81 + h9_row_l(hw)) * ldf + c] * iv;
        g[hw] = e[row * 64 + hw] * coef;
        dot = fmaf(fh[hw], g[hw], dot);
    }
#pragma unroll
    for (int hw = 0; hw < 49; ++hw)
        df[((long long)s * 81 + h9_row_l(hw)) * lddf + c] = iv * (g[hw] - fh[hw] * dot);
}

// sums of the sum-of-squares halves of a stats_part buffer [R][2][C]: group g owns rows_per_group consecutive partial
// rows; CTA (g, chunk) adds its 1/64 of them in a fixed order -> out[g*64 + chunk] (loss_finalize adds the 64 chunks)
constexpr int SUMSQ_CHUNKS = 64;
__global__ void __launch_bounds__(256) sumsq_reduce_kernel(const float* __restrict__ part, int rows_per_group, int C,
                                                           float* __restrict__ out) {
    __shared__ float scratch[32];
    const int g = blockIdx.x, chunk = blockIdx.y;
    const int rows_per_chunk = (rows_per_group + SUMSQ_CHUNKS - 1) / SUMSQ_CHUNKS;
    const int r0 = chunk * rows_per_chunk, r1 = min(r0 + rows_per_chunk, rows_per_group);
    float acc = 0.f;
    for (int r = r0; r < r1; r += 8) {           // eight rows of loads in flight per thread (a fixed order all the same)
        const float* q = part + (((long long)g * rows_per_group + r) * 2 + 1) * C;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (r + u < r1) ? __ldg(q + (long long)u * 2 * C + c) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
    }
    const float t = block_sum(acc, scratch);
    if (threadIdx.x == 0) out[g * SUMSQ_CHUNKS + chunk] = t;
}

// ----------------------------------------------------------------------------------------------------------
// Spatial self-similarity loss, forward + gradient, one CTA per sample (512 threads).
//   fs: fp32 H9 matrix [n_img*81][ldfs] (feat_space, own rows), x targets [n][512][49].
//   G[i][j] = <F^_i, F^_j>, F^_i = pixel row i normalised over the 512 channels; T likewise from X.
//   loss_part[s] = sum_ij (G - T)^2;  dfs[i][c] = inv_i (dF^[i][c] - F^[i][c] (F^_i . dF^_i)),  dF^ = coef * D F^.
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
selfsim_space_loss_kernel(const float* __restrict__ fs, int ldfs, const float* __restrict__ x, int n_per_group, float coef,
                          float* __restrict__ loss_part, float* __restrict__ dfs, int lddfs) {
    extern __shared__ __align__(16) float lsm[];
    float* buf = lsm;                   // [49][516]  pixel rows (normalised), pitch 516 (16-byte aligned rows)
    float* inv = buf + 49 * 516;        // [64]
    float* D = inv + 64;                // [49][49] (+3): T, then D = G - T
    float* part = D + 2404;             // [8][2401]
    float* rowdot = part + 8 * 2401;    // [64]
    float* scratch = rowdot + 64;       // [32]
    const int s = blockIdx.x, t = s % n_per_group, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int item = tid % 49, split = tid / 49;
    const int bi = (item / 7) * 7, bj = (item % 7) * 7;

    for (int phase = 0; phase < 2; ++phase) {        // phase 0: target X, phase 1: F
        __syncthreads();
        if (phase == 0) {
            for (int i = tid; i < 512 * 49; i += 512) { const int c = i / 49, hw = i - c * 49; buf[hw * 516 + c] = x[(long long)t * 512 * 49 + i]; }
        } else {
            for (int i = tid; i < 49 * 512; i += 512) { const int hw = i >> 9, c = i & 511; buf[hw * 516 + c] = fs[((long long)s * 81 + h9_row_l(hw)) * ldfs + c]; }
        }
        __syncthreads();
        for (int hw = warp; hw < 49; hw += 16) {
            float ss = 0.f;
            for (int c = lane; c < 512; c += 32) { const float v = buf[hw * 516 + c]; ss = fmaf(v, v, ss); }
            ss = warp_sum_l(ss);
            if (lane == 0) inv[hw] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        }
        __syncthreads();
        for (int i = tid; i < 49 * 512; i += 512) { const int hw = i >> 9, c = i & 511; buf[hw * 516 + c] *= inv[hw]; }
        __syncthreads();
        if (split < 8) {                             // Gram: (7x7 block, 64-channel slice) per thread
            float acc[7][7];
#pragma unroll
            for (int a = 0; a < 7; ++a)
#pragma unroll
                for (int q = 0; q < 7; ++q) acc[a][q] = 0.f;
            for (int c = split * 64; c < split * 64 + 64; ++c) {
                float va[7], vb[7];
#pragma unroll
                for (int q = 0; q < 7; ++q) { va[q] = buf[(bi + q) * 516 + c]; vb[q] = buf[(bj + q) * 516 + c]; }
#pragma unroll
                for (int a = 0; a < 7; ++a)
#pragma unroll
                    for (int q = 0; q < 7; ++q) acc[a][q] = fmaf(va[a], vb[q], acc[a][q]);
            }
            float* gp = part + split * 2401;
#pragma unroll
            for (int a = 0; a < 7; ++a)
#pragma unroll
                for (int q = 0; q < 7; ++q) gp[(bi + a) * 49 + bj + q] = acc[a][q];
        }
        __syncthreads();
        for (int o = tid; o < 2401; o += 512) {
            float g = 0.f;
#pragma unroll
            for (int sp = 0; sp < 8; ++sp) g += part[sp * 2401 + o];
            D[o] = (phase == 0) ? g : g - D[o];
        }
    }
    __syncthreads();
    {   // loss partial
        float acc = 0.f;
        for (int o = tid; o < 2401; o += 512) acc = fmaf(D[o], D[o], acc);
        const float tot = block_sum(acc, scratch);
        if (tid == 0) loss_part[s] = tot;
    }
    if (dfs == nullptr) return;
    // dF^[i][c] = coef * sum_j D[i][j] F^[j][c]; thread = channel c; then the per-pixel normalisation Jacobian
    const int c = tid;
    float fcol[49];
#pragma unroll
    for (int j = 0; j < 49; ++j) fcol[j] = buf[j * 516 + c];
    // rowdot[i] = sum_c F^[i][c] dF^[i][c]: warp partials summed in a fixed order; dF^ is recomputed in the second pass
    // (49 x 49 FMAs per thread) instead of being kept in 49 more registers
    float* wpart = part;                 // reuse: [16 warps][49]
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 49; ++i) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 49; ++j) a = fmaf(D[i * 49 + j], fcol[j], a);
        const float v = warp_sum_l(fcol[i] * a * coef);
        if (lane == 0) wpart[warp * 49 + i] = v;
    }
    __syncthreads();
    if (tid < 49) {
        float a = 0.f;
        for (int w = 0; w < 16; ++w) a += wpart[w * 49 + tid];
        rowdot[tid] = a;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 49; ++i) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 49; ++j) a = fmaf(D[i * 49 + j], fcol[j], a);
        dfs[((long long)s * 81 + h9_row_l(i)) * lddfs + c] = inv[i] * (a * coef - fcol[i] * rowdot[i]);
    }
}

// ----------------------------------------------------------------------------------------------------------
// Triplet + identity losses with gradients (trainer.py:31-43, :167-171). One warp per sample.
//   pos = 1 - cos(f_ocl, e_non), neg = 1 - cos(f_ocl, e_ocl) (F.normalize, eps 1e-12);  trip = relu(pos - neg + 0.1)
//   ident = |f_non - e_non|^2 + |f_ocl - e_non|^2   (per sample sums; MSE = mean over n*512)
// row_part[s] = {trip, pos, neg, sq_non, sq_ocl}; df_non / df_ocl [n][512] receive (overwrite) the gradients of
//   w_trip * mean_s(trip) + w_id * (MSE_non + MSE_ocl) / 2.
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) triplet_identity_kernel(const float* __restrict__ f_non, const float* __restrict__ f_ocl,
                                                               const float* __restrict__ e_non, const float* __restrict__ e_ocl,
                                                               int n, float w_trip, float w_id, float margin,
                                                               float* __restrict__ row_part, float* __restrict__ df_non,
                                                               float* __restrict__ df_ocl) {
    const int s = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= n) return;
    float a[16], y[16], z[16], fn[16];
    float saa = 0.f, syy = 0.f, szz = 0.f, say = 0.f, saz = 0.f, qn = 0.f, qo = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int c = k * 32 + lane;
        a[k] = f_ocl[(long long)s * 512 + c]; y[k] = e_non[(long long)s * 512 + c]; z[k] = e_ocl[(long long)s * 512 + c];
        fn[k] = f_non[(long long)s * 512 + c];
        saa = fmaf(a[k], a[k], saa); syy = fmaf(y[k], y[k], syy); szz = fmaf(z[k], z[k], szz);
        say = fmaf(a[k], y[k], say); saz = fmaf(a[k], z[k], saz);
        qn = fmaf(fn[k] - y[k], fn[k] - y[k], qn); qo = fmaf(a[k] - y[k], a[k] - y[k], qo);
    }
    saa = warp_sum_l(saa); syy = warp_sum_l(syy); szz = warp_sum_l(szz); say = warp_sum_l(say); saz = warp_sum_l(saz);
    qn = warp_sum_l(qn); qo = warp_sum_l(qo);
    const float na = fmaxf(sqrtf(saa), 1e-12f), ny = fmaxf(sqrtf(syy), 1e-12f), nz = fmaxf(sqrtf(szz), 1e-12f);
    const float cy = say / (na * ny), cz = saz / (na * nz);
    const float pos = 1.f - cy, neg = 1.f - cz;
    const float tr = pos - neg + margin;
    if (lane == 0) {
        float* o = row_part + (long long)s * 5;
        o[0] = fmaxf(tr, 0.f); o[1] = pos; o[2] = neg; o[3] = qn; o[4] = qo;
    }
    // d trip / d a = -(d cy/da) + (d cz/da), active only where tr > 0; d cos(a,y)/da = (y^ - a^ cos) / |a|
    const float gt = (tr > 0.f) ? w_trip / (float)n : 0.f;
    const float gi = w_id / ((float)n * 512.f);          // d/df of w_id/2 * (sum sq)/(n*512) = w_id (f - e)/(n*512)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int c = k * 32 + lane;
        const float ah = a[k] / na;
        const float dcy = (y[k] / ny - ah * cy) / na, dcz = (z[k] / nz - ah * cz) / na;
        df_ocl[(long long)s * 512 + c] = gt * (dcz - dcy) + gi * (a[k] - y[k]);
        df_non[(long long)s * 512 + c] = gi * (fn[k] - y[k]);
    }
}

// Final scalars (one CTA, fixed order). out[0..3] = loss items weighted (SelfSimilarity, Triplet, Identity, Classifier),
// out[4] = mean pos, out[5] = mean neg, out[6] = total.
//   space_part [G*n] per-sample sums of squared spatial Gram differences; chan_sums [G][64] partial sums of squared
//   channel Gram differences (sumsq_reduce_kernel); row_part [n][5]; ce [2] device scalars (mean CE of the non / ocl call).
__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ space_part, const float* __restrict__ chan_sums,
                                                            const float* __restrict__ row_part, const float* __restrict__ ce,
                                                            int n, int G, float w0, float w1, float w2, float w3,
                                                            float* __restrict__ out) {
    __shared__ float scratch[32];
    float sp = 0.f, tr = 0.f, ps = 0.f, ng = 0.f, q = 0.f;
    for (int i = threadIdx.x; i < G * n; i += blockDim.x) sp += space_part[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        tr += row_part[i * 5]; ps += row_part[i * 5 + 1]; ng += row_part[i * 5 + 2];
        q += row_part[i * 5 + 3] + row_part[i * 5 + 4];
    }
    sp = block_sum(sp, scratch); tr = block_sum(tr, scratch); ps = block_sum(ps, scratch);
    ng = block_sum(ng, scratch); q = block_sum(q, scratch);
    if (threadIdx.x == 0) {
        float ch = 0.f;
        for (int g = 0; g < G * SUMSQ_CHUNKS; ++g) ch += chan_sums[g];
        // each MSE term is a mean over n*49*49 (space) / n*512*512 (channel); L1 = ((ms_non+ms_ocl)/2 + (mc_non+mc_ocl)/2)/2
        const float l_space = sp / ((float)n * 2401.f) * 0.5f, l_chan = ch / ((float)n * 262144.f) * 0.5f;
        out[0] = w0 * 0.5f * (l_space + l_chan);
        out[1] = w1 * tr / (float)n;
        out[2] = w2 * 0.5f * q / ((float)n * 512.f);
        out[3] = w3 * (ce[0] / (1e-8f + w3) + ce[1]);
        out[4] = ps / (float)n;
        out[5] = ng / (float)n;
        out[6] = out[0] + out[1] + out[2] + out[3];
    }
}

// out[i] = a[i] + b[i] (+ c[i]) — gradient of the pooled feature: head (CosFace) + triplet / identity contributions
__global__ void __launch_bounds__(256) add3_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                   const float* __restrict__ c, float* __restrict__ out, long long count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = a[i] + b[i] + (c ? c[i] : 0.f);
}

}  // namespace ffr

using namespace ffr;
static inline cudaStream_t S_(ffr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

FFR_API int ffr_selfsim_channel_pack(const float* f, int ldf, const float* x, int n_img, int n_per_group, void* A6,
                                     void* B6, void* FhT, float* inv_f, ffr_stream_t stream) {
    FFR_CHECK_ARG(f && x && A6 && B6 && FhT && inv_f && n_per_group > 0, "ffr_selfsim_channel_pack: bad arguments");
    if (n_img == 0) return 0;
    selfsim_channel_pack_kernel<<<dim3(n_img, 8), 256, 0, S_(stream)>>>(f, ldf, x, n_per_group,
                                                              reinterpret_cast<__nv_bfloat16*>(A6),
                                                              reinterpret_cast<__nv_bfloat16*>(B6),
                                                              reinterpret_cast<__nv_bfloat16*>(FhT), inv_f);
    return launch_status("selfsim_channel_pack_kernel");
}

FFR_API int ffr_selfsim_channel_bwd(const float* e, const float* f, int ldf, const float* inv_f, float coef, float* df,
                                    int lddf, int n_img, ffr_stream_t stream) {
    FFR_CHECK_ARG(e && f && inv_f && df, "ffr_selfsim_channel_bwd: null pointer");
    if (n_img == 0) return 0;
    selfsim_channel_bwd_kernel<<<dim3(n_img, 2), 256, 0, S_(stream)>>>(e, f, ldf, inv_f, coef, df, lddf);
    return launch_status("selfsim_channel_bwd_kernel");
}

FFR_API int ffr_sumsq_reduce(const float* part, int rows_per_group, int C, int groups, float* out, ffr_stream_t stream) {
    FFR_CHECK_ARG(part && out && groups > 0, "ffr_sumsq_reduce: bad arguments");
    sumsq_reduce_kernel<<<dim3(groups, SUMSQ_CHUNKS), 256, 0, S_(stream)>>>(part, rows_per_group, C, out);
    return launch_status("sumsq_reduce_kernel");
}

FFR_API int ffr_selfsim_space_loss(const float* fs, int ldfs, const float* x, int n_img, int n_per_group, float coef,
                                   float* loss_part, float* dfs, int lddfs, ffr_stream_t stream) {
    FFR_CHECK_ARG(fs && x && loss_part && n_per_group > 0, "ffr_selfsim_space_loss: bad arguments");
    if (n_img == 0) return 0;
    const int smem = (49 * 516 + 64 + 2404 + 8 * 2401 + 64 + 32) * (int)sizeof(float);
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(selfsim_space_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    selfsim_space_loss_kernel<<<n_img, 512, smem, S_(stream)>>>(fs, ldfs, x, n_per_group, coef, loss_part, dfs, lddfs);
    return launch_status("selfsim_space_loss_kernel");
}

FFR_API int ffr_triplet_identity(const float* f_non, const float* f_ocl, const float* e_non, const float* e_ocl, int n,
                                 float w_trip, float w_id, float margin, float* row_part, float* df_non, float* df_ocl,
                                 ffr_stream_t stream) {
    FFR_CHECK_ARG(f_non && f_ocl && e_non && e_ocl && row_part && df_non && df_ocl, "ffr_triplet_identity: null pointer");
    if (n == 0) return 0;
    triplet_identity_kernel<<<(n + 7) / 8, 256, 0, S_(stream)>>>(f_non, f_ocl, e_non, e_ocl, n, w_trip, w_id, margin, row_part,
                                                                df_non, df_ocl);
    return launch_status("triplet_identity_kernel");
}

FFR_API int ffr_loss_finalize(const float* space_part, const float* chan_sums, const float* row_part, const float* ce, int n,
                              int groups, float w0, float w1, float w2, float w3, float* out, ffr_stream_t stream) {
    FFR_CHECK_ARG(space_part && chan_sums && row_part && ce && out, "ffr_loss_finalize: null pointer");
    loss_finalize_kernel<<<1, 256, 0, S_(stream)>>>(space_part, chan_sums, row_part, ce, n, groups, w0, w1, w2, w3, out);
    return launch_status("loss_finalize_kernel");
}

FFR_API int ffr_add3_f32(const float* a, const float* b, const float* c, float* out, int64_t count, ffr_stream_t stream) {
    FFR_CHECK_ARG(a && b && out, "ffr_add3_f32: null pointer");
    if (count == 0) return 0;
    add3_kernel<<<(int)((count + 255) / 256), 256, 0, S_(stream)>>>(a, b, c, out, count);
    return launch_status("add3_kernel");
}

}  // extern "C"
