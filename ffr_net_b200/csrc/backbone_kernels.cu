// Memory-bound kernels of the IR-SE50 forward: stem conv, SE gate + residual, shortcut subsample,
// feature-map export (extra `bn` -> fp32 NCHW) and the embedding finish (bias + L2 normalise).
// All activations are bf16 in the halo-shared flat NHWC layout: image n, pixel (h, w) of an SxS map
// lives in row n*(S+1)^2 + h*(S+1) + w; rows with h == S or w == S are zero padding.
#include "host.h"
#include "ptx.cuh"

namespace ffr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ----------------------------------------------------------------------------------------------
// Stem: Conv3x3(3->64, s1, p1) + BN + PReLU on fp32 NCHW input  (model_ir_se50.py:118-120)
// w: [27][64] fp32 with the BN scale folded in (k = ci*9 + r*3 + s), b: [64] BN shift, a: [64] PReLU slope.
// One thread per output row of the flat layout (pad rows are written as zeros).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ b, const float* __restrict__ a,
                                                   __nv_bfloat16* __restrict__ out, int n_img, int S) {
    __shared__ __align__(16) float sw[27 * 64];
    __shared__ float sb[64], sa[64];
    for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 64) { sb[threadIdx.x] = b[threadIdx.x]; sa[threadIdx.x] = a[threadIdx.x]; }
    __syncthreads();
    const int G = S + 1;
    const long long total = (long long)n_img * G * G;
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= total) return;
    const int n = (int)(m / (G * G));
    const int rem = (int)(m - (long long)n * G * G);
    const int h = rem / G, wq = rem - h * G;
    uint4* o = reinterpret_cast<uint4*>(out + m * 64);
    if (h == S || wq == S) {
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = make_uint4(0, 0, 0, 0);
        return;
    }
    float in[27];
    const float* xi = x + (long long)n * 3 * S * S;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int hh = h + r - 1, ww = wq + s - 1;
                in[ci * 9 + r * 3 + s] =
                    (hh >= 0 && hh < S && ww >= 0 && ww < S) ? __ldg(xi + ((long long)ci * S + hh) * S + ww) : 0.f;
            }
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 16) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = sb[c0 + j];
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const float4* wr = reinterpret_cast<const float4*>(sw + k * 64 + c0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 wv = wr[q];
                acc[q * 4 + 0] = fmaf(in[k], wv.x, acc[q * 4 + 0]);
                acc[q * 4 + 1] = fmaf(in[k], wv.y, acc[q * 4 + 1]);
                acc[q * 4 + 2] = fmaf(in[k], wv.z, acc[q * 4 + 2]);
                acc[q * 4 + 3] = fmaf(in[k], wv.w, acc[q * 4 + 3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = acc[j] > 0.f ? acc[j] : acc[j] * sa[c0 + j];
#pragma unroll
        for (int q = 0; q < 2; ++q)
            o[c0 / 8 + q] = make_uint4(pack_bf16x2(acc[q * 8 + 0], acc[q * 8 + 1]), pack_bf16x2(acc[q * 8 + 2], acc[q * 8 + 3]),
                                       pack_bf16x2(acc[q * 8 + 4], acc[q * 8 + 5]), pack_bf16x2(acc[q * 8 + 6], acc[q * 8 + 7]));
    }
}

int stem_launch(const float* x, const float* w, const float* b, const float* a, void* out, int n_img, int S,
                cudaStream_t stream) {
    const long long total = (long long)n_img * (S + 1) * (S + 1);
    const int grid = (int)((total + 127) / 128);
    stem_kernel<<<grid, 128, 0, stream>>>(x, w, b, a, reinterpret_cast<__nv_bfloat16*>(out), n_img, S);
    return launch_status("stem_kernel");
}

// ----------------------------------------------------------------------------------------------
// SE gate + residual  (model_ir_se50.py:29-36, 73-76):
//   s = sigmoid(W2 relu(W1 mean_hw(u)));  y = u * s + shortcut
// pool holds per-(image, channel) SUMS of u over the SxS valid pixels (accumulated by the conv2 epilogue).
// shortcut_mode 0: x on the same grid (identity MaxPool(1,1)); 1: x on the (2S)x(2S) grid, subsampled
// (MaxPool(1,2)); 2: sc on the same grid (Conv1x1+BN shortcut, computed by a GEMM).
// grid = (chunks, n_img); every CTA recomputes the (tiny) gate of its image.
// ----------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) se_residual_kernel(const __nv_bfloat16* __restrict__ u,
                                                          const float* __restrict__ pool,
                                                          const float* __restrict__ w1, const float* __restrict__ w2,
                                                          const __nv_bfloat16* __restrict__ sc, int shortcut_mode,
                                                          __nv_bfloat16* __restrict__ y, int S) {
    constexpr int R = C / 16;
    constexpr int TPR = C / 8;          // threads per row (8 channels = 16 B each)
    constexpr int RPI = 256 / TPR;      // rows per iteration
    __shared__ float s_mean[C];
    __shared__ float s_hid[R];
    __shared__ float s_gate[C];
    const int n = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float inv = 1.0f / (float)(S * S);
    for (int c = tid; c < C; c += 256) s_mean[c] = pool[(long long)n * C + c] * inv;
    __syncthreads();
    for (int j = warp; j < R; j += 8) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a = fmaf(__ldg(w1 + j * C + c), s_mean[c], a);
        a = warp_sum(a);
        if (lane == 0) s_hid[j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int c = tid; c < C; c += 256) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < R; ++j) a = fmaf(__ldg(w2 + c * R + j), s_hid[j], a);
        s_gate[c] = 1.0f / (1.0f + __expf(-a));
    }
    __syncthreads();

    const int c8 = tid % TPR;
    float g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = s_gate[c8 * 8 + j];
    const int G = S + 1;
    const int rows = G * G;
    const int per = (rows + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * per;
    const int r1 = min(rows, r0 + per);
    const long long base = (long long)n * rows;
    const int G2 = 2 * S + 1;
    for (int r = r0 + tid / TPR; r < r1; r += RPI) {
        const uint4 uv = __ldg(reinterpret_cast<const uint4*>(u + (base + r) * C) + c8);
        long long srow = base + r;
        if (shortcut_mode == 1) {
            const int h = r / G, wq = r - h * G;
            srow = (long long)n * G2 * G2 + (long long)(2 * h) * G2 + 2 * wq;
        }
        const uint4 sv = __ldg(reinterpret_cast<const uint4*>(sc + srow * C) + c8);
        uint4 o;
        o.x = pack_bf16x2(fmaf(bf16lo(uv.x), g[0], bf16lo(sv.x)), fmaf(bf16hi(uv.x), g[1], bf16hi(sv.x)));
        o.y = pack_bf16x2(fmaf(bf16lo(uv.y), g[2], bf16lo(sv.y)), fmaf(bf16hi(uv.y), g[3], bf16hi(sv.y)));
        o.z = pack_bf16x2(fmaf(bf16lo(uv.z), g[4], bf16lo(sv.z)), fmaf(bf16hi(uv.z), g[5], bf16hi(sv.z)));
        o.w = pack_bf16x2(fmaf(bf16lo(uv.w), g[6], bf16lo(sv.w)), fmaf(bf16hi(uv.w), g[7], bf16hi(sv.w)));
        reinterpret_cast<uint4*>(y + (base + r) * C)[c8] = o;
    }
}

int se_residual_launch(const void* u, const float* pool, const float* w1, const float* w2, const void* sc,
                       int shortcut_mode, void* y, int n_img, int S, int C, cudaStream_t stream) {
    FFR_CHECK_ARG(shortcut_mode >= 0 && shortcut_mode <= 2, "se_residual: shortcut_mode=%d", shortcut_mode);
    const int rows = (S + 1) * (S + 1);
    int chunks = (num_sms() * 4 + n_img - 1) / n_img;
    const int max_chunks = (rows + 63) / 64;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, n_img);
    const __nv_bfloat16* up = reinterpret_cast<const __nv_bfloat16*>(u);
    const __nv_bfloat16* sp = reinterpret_cast<const __nv_bfloat16*>(sc);
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
    switch (C) {
        case 64:  se_residual_kernel<64><<<grid, 256, 0, stream>>>(up, pool, w1, w2, sp, shortcut_mode, yp, S); break;
        case 128: se_residual_kernel<128><<<grid, 256, 0, stream>>>(up, pool, w1, w2, sp, shortcut_mode, yp, S); break;
        case 256: se_residual_kernel<256><<<grid, 256, 0, stream>>>(up, pool, w1, w2, sp, shortcut_mode, yp, S); break;
        case 512: se_residual_kernel<512><<<grid, 256, 0, stream>>>(up, pool, w1, w2, sp, shortcut_mode, yp, S); break;
        default: return set_error(-1, "se_residual: unsupported C=%d", C);
    }
    return launch_status("se_residual_kernel");
}

// ----------------------------------------------------------------------------------------------
// Stride-2 subsample of a flat map (input of the Conv1x1(stride 2) shortcut, model_ir_se50.py:62-63):
// out(n,h,w) = x(n,2h,2w); the pad row/column of the output grid lands on the pad row/column of the input grid.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) subsample2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out,
                                                         int n_img, int So, int C8) {
    const int G = So + 1, G2 = 2 * So + 1;
    const long long total = (long long)n_img * G * G * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8);
        const long long row = i / C8;
        const int n = (int)(row / (G * G));
        const int rem = (int)(row - (long long)n * G * G);
        const int h = rem / G, w = rem - h * G;
        const long long srow = (long long)n * G2 * G2 + (long long)(2 * h) * G2 + 2 * w;
        out[i] = __ldg(x + srow * C8 + c);
    }
}

int subsample2_launch(const void* x, void* out, int n_img, int So, int C, cudaStream_t stream) {
    const long long total = (long long)n_img * (So + 1) * (So + 1) * (C / 8);
    int grid = (int)((total + 255) / 256);
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    subsample2_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), n_img,
                                                So, C / 8);
    return launch_status("subsample2_kernel");
}

// ----------------------------------------------------------------------------------------------
// Feature-map export: y = bn(h) as fp32 NCHW (model_ir_se50.py:126,139) from the flat bf16 map.
// grid = (C/64, n_img); a 64-channel slab of one image is transposed through shared memory.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) export_nchw_kernel(const __nv_bfloat16* __restrict__ h,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift, float* __restrict__ y,
                                                          int S, int C) {
    extern __shared__ float tile[];  // [S*S][65]
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const int G = S + 1, P = S * S;
    for (int i = threadIdx.x; i < P * 64; i += blockDim.x) {
        const int c = i & 63, pix = i >> 6;
        const int ph = pix / S, pw = pix - ph * S;
        const long long row = (long long)n * G * G + ph * G + pw;
        tile[pix * 65 + c] = __bfloat162float(h[row * C + c0 + c]) * scale[c0 + c] + shift[c0 + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * 64; i += blockDim.x) {
        const int pix = i % P, c = i / P;
        y[((long long)n * C + c0 + c) * P + pix] = tile[pix * 65 + c];
    }
}

int export_nchw_launch(const void* h, const float* scale, const float* shift, float* y, int n_img, int S, int C,
                       cudaStream_t stream) {
    FFR_CHECK_ARG(C % 64 == 0, "export_nchw: C=%d", C);
    dim3 grid(C / 64, n_img);
    const size_t smem = (size_t)S * S * 65 * sizeof(float);
    FFR_CHECK_ARG(smem <= 48 * 1024, "export_nchw: map %dx%d too large", S, S);
    export_nchw_kernel<<<grid, 256, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(h), scale, shift, y, S, C);
    return launch_status("export_nchw_kernel");
}

// ----------------------------------------------------------------------------------------------
// Embedding finish: f = l2_norm(acc + bias)  (model_ir_se50.py:13-16,141). One warp per row; D multiple of 32.
// `acc` is the split-K fp32 accumulator of the folded head GEMM (BN2d, Linear, BN1d folded into W', b').
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bias_l2norm_kernel(const float* __restrict__ acc, const float* __restrict__ bias,
                                                          float* __restrict__ f, int rows, int D) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float ss = 0.f;
    for (int c = lane; c < D; c += 32) {
        const float v = acc[(long long)row * D + c] + bias[c];
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / sqrtf(ss);
    for (int c = lane; c < D; c += 32) f[(long long)row * D + c] = (acc[(long long)row * D + c] + bias[c]) * inv;
}

int bias_l2norm_launch(const float* acc, const float* bias, float* f, int rows, int D, cudaStream_t stream) {
    bias_l2norm_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(acc, bias, f, rows, D);
    return launch_status("bias_l2norm_kernel");
}

}  // namespace ffr
