// CosFace head (AddMarginProduct, recnet.py:238-270) + CrossEntropy (trainer.py:173-176), training path.
//
//   cos = normalize(v) . normalize(W)^T          (N x 10575, K = 512)
//   z   = s * (cos - m * onehot(label)),  loss = mean_n( logsumexp_c z[n,c] - z[n,label_n] )
//
// The GEMM runs on the tcgen05 kernel of conv_gemm.cu with bf16 operands split into hi + lo parts
// (x = hi + lo, |lo| <= 2^-9 |x|): the K axis is laid out [hi | lo | hi] for the samples and [hi | hi | lo] for the
// classes, so one K = 1536 GEMM yields hi.hi + lo.hi + hi.lo — cosine error ~1e-5 instead of ~1e-3 for plain bf16,
// which matters because s = 30 multiplies it inside the softmax. Its epilogue (EPI_COSFACE) accumulates the softmax
// denominator per row with the fixed shift s (|cos| <= 1, so no running maximum is needed), and records z_label and
// the arg-max class; neither the logits nor the one-hot matrix are ever materialised (the cosines are kept in fp32
// for the backward pass). Backward: dcos = s * (softmax - onehot) * g / N in bf16 (plus its transpose), two more
// tcgen05 GEMMs (dv^ = dcos . W^, dW^ = dcos^T . v^), and the Jacobian of the row normalisation.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "host.h"
#include "kernels.h"

namespace ffr {

__device__ __forceinline__ float warp_sum_h(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- pack: rows of x (fp32 [rows x 512]) -> L2-normalised (eps 1e-12, F.normalize) bf16 hi/lo split --------------------
// packed [rows_pad x 1536]: mode 0 (samples) = [hi | lo | hi], mode 1 (classes) = [hi | hi | lo]; rows >= rows are zero.
// transposed (optional) [512 x t_ld]: hi part only, column = row index (the K-major operand of the backward GEMMs).
// 64 rows per CTA, 8 warps, each warp 8 rows; the transposed tile goes through shared memory.
constexpr int PACK_ROWS = 64;
constexpr int PACK_PITCH = 514;   // bf16 elements: 257 words, odd -> conflict-free column reads

__global__ void __launch_bounds__(256) cosface_pack_kernel(const float* __restrict__ x, int rows, int rows_pad, int mode,
                                                           __nv_bfloat16* __restrict__ packed,
                                                           __nv_bfloat16* __restrict__ transposed, int t_ld) {
    extern __shared__ __nv_bfloat16 tile[];   // [PACK_ROWS][PACK_PITCH], only when transposed != nullptr
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * PACK_ROWS;
    for (int i = 0; i < PACK_ROWS / 8; ++i) {
        const int rl = warp * (PACK_ROWS / 8) + i;
        const int r = r0 + rl;
        if (r >= rows_pad) break;
        float4 v[4];
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] = (r < rows) ? __ldg(reinterpret_cast<const float4*>(x + (long long)r * 512) + q * 32 + lane)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
            ss += v[q].x * v[q].x + v[q].y * v[q].y + v[q].z * v[q].z + v[q].w * v[q].w;
        }
        ss = warp_sum_h(ss);
        const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        __nv_bfloat16* prow = packed + (long long)r * 1536;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float f[4] = {v[q].x * inv, v[q].y * inv, v[q].z * inv, v[q].w * inv};
            __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                hi[e] = __float2bfloat16_rn(f[e]);
                lo[e] = __float2bfloat16_rn(f[e] - __bfloat162float(hi[e]));
            }
            const int k = (q * 32 + lane) * 4;
            const uint2 h2 = make_uint2((uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16),
                                        (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16));
            const uint2 l2 = make_uint2((uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16),
                                        (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16));
            *reinterpret_cast<uint2*>(prow + k) = h2;
            *reinterpret_cast<uint2*>(prow + 512 + k) = mode ? h2 : l2;
            *reinterpret_cast<uint2*>(prow + 1024 + k) = mode ? l2 : h2;
            if (transposed) {
#pragma unroll
                for (int e = 0; e < 4; ++e) tile[rl * PACK_PITCH + k + e] = hi[e];
            }
        }
    }
    if (!transposed) return;
    __syncthreads();
    // tile[rl][k] -> transposed[k][r0 + rl]: thread = (k sub-row, rl pair) -> 128-byte runs along the row index
    const int nr = min(PACK_ROWS, rows_pad - r0);
    for (int idx = threadIdx.x; idx < 512 * (PACK_ROWS / 2); idx += 256) {
        const int k = idx / (PACK_ROWS / 2), rp = (idx % (PACK_ROWS / 2)) * 2;
        if (rp < nr) {
            const uint32_t a = __bfloat16_as_ushort(tile[rp * PACK_PITCH + k]);
            const uint32_t b = (rp + 1 < nr) ? __bfloat16_as_ushort(tile[(rp + 1) * PACK_PITCH + k]) : 0u;
            *reinterpret_cast<uint32_t*>(transposed + (long long)k * t_ld + r0 + rp) = a | (b << 16);
        }
    }
}

int cosface_pack_launch(const float* x, int rows, int rows_pad, int mode, void* packed, void* transposed, int t_ld,
                        cudaStream_t stream) {
    if (rows_pad == 0) return 0;
    const int smem = transposed ? PACK_ROWS * PACK_PITCH * 2 : 0;
    static bool attr = false;
    if (!attr) {
        FFR_CUDA(cudaFuncSetAttribute(cosface_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PACK_ROWS * PACK_PITCH * 2));
        attr = true;
    }
    cosface_pack_kernel<<<(rows_pad + PACK_ROWS - 1) / PACK_ROWS, 256, smem, stream>>>(
        x, rows, rows_pad, mode, reinterpret_cast<__nv_bfloat16*>(packed), reinterpret_cast<__nv_bfloat16*>(transposed), t_ld);
    return launch_status("cosface_pack_kernel");
}

// ---- fixed-order sum of the per-(row, column tile) partial softmax denominators -------------------------------------
__global__ void __launch_bounds__(256) sumexp_reduce_kernel(const float* __restrict__ part, int n, int parts,
                                                            float* __restrict__ sumexp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < parts; ++k) s += part[(long long)i * parts + k];
    sumexp[i] = s;
}

int sumexp_reduce_launch(const float* part, int n, int parts, float* sumexp, cudaStream_t stream) {
    if (n == 0) return 0;
    sumexp_reduce_kernel<<<(n + 255) / 256, 256, 0, stream>>>(part, n, parts, sumexp);
    return launch_status("sumexp_reduce_kernel");
}

// ---- finish: loss = mean(log(sumexp) + s - z_label); pred = arg-max class ---------------------------------------------
__global__ void __launch_bounds__(1024) cosface_finish_kernel(const float* __restrict__ sumexp, const float* __restrict__ zlabel,
                                                              const unsigned long long* __restrict__ argkey, int n, float s,
                                                              float* __restrict__ loss, long long* __restrict__ pred) {
    __shared__ float part[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        acc += logf(sumexp[i]) + s - zlabel[i];
        if (pred) pred[i] = (long long)(0xFFFFFFFFu - (uint32_t)(argkey[i] & 0xFFFFFFFFull));
    }
    acc = warp_sum_h(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.f;
        v = warp_sum_h(v);
        if (threadIdx.x == 0) *loss = v / (float)n;
    }
}

int cosface_finish_launch(const float* sumexp, const float* zlabel, const unsigned long long* argkey, int n, float s,
                          float* loss, long long* pred, cudaStream_t stream) {
    cosface_finish_kernel<<<1, 1024, 0, stream>>>(sumexp, zlabel, argkey, n, s, loss, pred);
    return launch_status("cosface_finish_kernel");
}

// ---- backward of margin + softmax-CE: dcos[n,c] = s * (softmax[n,c] - onehot[n,c]) * g / N ------------------------------
// 64 x 64 tile per CTA; writes dcos [n_rows x c_pad] and its transpose dcosT [c_pad x n_pad] (both bf16, pads zero)
__global__ void __launch_bounds__(256) cosface_bwd_kernel(const float* __restrict__ cosv, int c_pad, int classes, int n, int n_pad,
                                                          const int* __restrict__ label, const float* __restrict__ sumexp,
                                                          const float* __restrict__ gloss, float s, float mrg,
                                                          __nv_bfloat16* __restrict__ dcos, __nv_bfloat16* __restrict__ dcosT,
                                                          int n_per_group) {
    __shared__ __nv_bfloat16 t[64][66];
    const int c0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // tx: class within tile, ty: 4 row groups
    // n_per_group > 0: the rows are several batches (RecNet calls) of n_per_group samples, each with its own mean CE
    // loss and upstream gradient gloss[group]
    for (int i = ty; i < 64; i += 4) {
        const int r = n0 + i, c = c0 + tx;
        float d = 0.f;
        if (r < n && c < classes) {
            const float gs = (n_per_group > 0) ? __ldg(gloss + r / n_per_group) * s / (float)n_per_group
                                               : __ldg(gloss) * s / (float)n;
            const int lab = __ldg(label + r);
            const float z = s * (__ldg(cosv + (long long)r * c_pad + c) - ((c == lab) ? mrg : 0.f));
            const float pr = __expf(z - s) / __ldg(sumexp + r);
            d = gs * (pr - ((c == lab) ? 1.f : 0.f));
        }
        const __nv_bfloat16 b = __float2bfloat16_rn(d);
        if (r < n) dcos[(long long)r * c_pad + c] = b;
        t[i][tx] = b;
    }
    __syncthreads();
    for (int i = ty; i < 64; i += 4) {                             // i: class within tile, tx: row within tile
        const int c = c0 + i, r = n0 + tx;
        if (r < n_pad) dcosT[(long long)c * n_pad + r] = t[tx][i];
    }
}

int cosface_bwd_launch_ex(const float* cosv, int c_pad, int classes, int n, int n_pad, const int* label, const float* sumexp,
                          const float* gloss, float s, float m, void* dcos, void* dcosT, int n_per_group, cudaStream_t stream) {
    if (n == 0) return 0;
    cosface_bwd_kernel<<<dim3(c_pad / 64, (n_pad + 63) / 64), 256, 0, stream>>>(
        cosv, c_pad, classes, n, n_pad, label, sumexp, gloss, s, m, reinterpret_cast<__nv_bfloat16*>(dcos),
        reinterpret_cast<__nv_bfloat16*>(dcosT), n_per_group);
    return launch_status("cosface_bwd_kernel");
}

int cosface_bwd_launch(const float* cosv, int c_pad, int classes, int n, int n_pad, const int* label, const float* sumexp,
                       const float* gloss, float s, float m, void* dcos, void* dcosT, cudaStream_t stream) {
    return cosface_bwd_launch_ex(cosv, c_pad, classes, n, n_pad, label, sumexp, gloss, s, m, dcos, dcosT, 0, stream);
}

// ---- Jacobian of F.normalize over rows of 512: dx = (dxh - xh (xh . dxh)) / max(||x||, eps) ---------------------------
__global__ void __launch_bounds__(256) normalize_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dxh, int rows,
                                                            float* __restrict__ dx) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + warp;
    if (r >= rows) return;
    float4 v[4], g[4];
    float ss = 0.f, dot = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[q] = __ldg(reinterpret_cast<const float4*>(x + (long long)r * 512) + q * 32 + lane);
        g[q] = __ldg(reinterpret_cast<const float4*>(dxh + (long long)r * 512) + q * 32 + lane);
        ss += v[q].x * v[q].x + v[q].y * v[q].y + v[q].z * v[q].z + v[q].w * v[q].w;
        dot += v[q].x * g[q].x + v[q].y * g[q].y + v[q].z * g[q].z + v[q].w * g[q].w;
    }
    ss = warp_sum_h(ss);
    dot = warp_sum_h(dot);
    const float nrm = sqrtf(ss);
    const bool clamped = nrm < 1e-12f;                 // F.normalize divides by max(norm, eps): constant when clamped
    const float inv = 1.0f / fmaxf(nrm, 1e-12f);
    const float k = clamped ? 0.f : dot * inv * inv;   // (xh . dxh) / norm, with xh = x * inv
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 o;
        o.x = (g[q].x - v[q].x * k) * inv; o.y = (g[q].y - v[q].y * k) * inv;
        o.z = (g[q].z - v[q].z * k) * inv; o.w = (g[q].w - v[q].w * k) * inv;
        reinterpret_cast<float4*>(dx + (long long)r * 512)[q * 32 + lane] = o;
    }
}

int normalize_bwd_launch(const float* x, const float* dxh, int rows, float* dx, cudaStream_t stream) {
    if (rows == 0) return 0;
    normalize_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, dxh, rows, dx);
    return launch_status("normalize_bwd_kernel");
}

}  // namespace ffr
