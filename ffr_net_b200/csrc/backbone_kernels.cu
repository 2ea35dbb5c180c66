// Memory-bound kernels of the IR-SE50 forward: stem conv, SE gate + residual, shortcut subsample,
// feature-map export (extra `bn` -> fp32 NCHW) and the embedding finish (bias + L2 normalise).
// All activations are bf16 in the halo-shared flat NHWC layout: image n, pixel (h, w) of an SxS map
// lives in row n*(S+1)^2 + h*(S+1) + w; rows with h == S or w == S are zero padding.
#include "host.h"
#include "kernels.h"
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
//
// K = 27 is far too small for a tcgen05/TMA pipeline (one 128-byte swizzle row would be mostly padding) and the
// layer is HBM-bound (reads 150 KB, writes 1.6 MB per image for 43 MFLOP): each warp builds the im2col fragment of
// 16 output rows straight from global memory (L1-resident neighbourhood), multiplies it with the register-resident
// weights using warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate, K padded 27 -> 32) and stores full 128-byte
// rows. Pad rows of the flat layout are written as zeros.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// U8 = true: the input is the decoded image itself, uint8 HWC (N,S,S,3), and the reference's host preprocessing
// (data/dataset.py:135-155 channel swap + horizontal flip, data/dataloader.py:15-19 ToTensor + Normalize(0.5, 0.5))
// is applied on the fly: value = (u/255 - 0.5)/0.5 through a 256-entry table computed with IEEE division (bit-equal to
// torchvision), source channel 2-ci when swap_rb, source column S-1-w for images whose flip flag is set.
template <bool U8>
__global__ void __launch_bounds__(128, 4) stem_kernel(const void* __restrict__ xin, const unsigned char* __restrict__ flip,
                                                   int swap_rb, const float* __restrict__ w,
                                                   const float* __restrict__ b, const float* __restrict__ a,
                                                   __nv_bfloat16* __restrict__ out, int n_img, int S) {
    pdl_sync();
    __shared__ __align__(16) uint32_t stage[4][16 * 32];   // per warp: 16 rows x 64 bf16
    __shared__ float lut[U8 ? 256 : 1];
    const float* x = reinterpret_cast<const float*>(xin);
    const unsigned char* xu = reinterpret_cast<const unsigned char*>(xin);
    if (U8) {
        for (int i = threadIdx.x; i < 256; i += 128) lut[i] = __fdiv_rn(__fdiv_rn((float)i, 255.0f) - 0.5f, 0.5f);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int G = S + 1;
    const long long total = (long long)n_img * G * G;
    const long long plane = (long long)S * S;

    // B fragments: weights as a [K=32][N=64] col-major operand, 8 n-tiles x 2 k-steps
    uint32_t bf[8][2][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = ks * 16 + h * 8 + tig * 2;
                const int n = nt * 8 + g;
                const float w0 = (k0 < 27) ? w[k0 * 64 + n] : 0.f;
                const float w1 = (k0 + 1 < 27) ? w[(k0 + 1) * 64 + n] : 0.f;
                bf[nt][ks][h] = pack_bf16x2(w0, w1);
            }
    // this thread's 8 K columns: k = ks*16 + h*8 + tig*2 + e  ->  (ci, r, s) offsets into the image
    int koff[8], kdr[8], kds[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = (j >> 2) * 16 + ((j >> 1) & 1) * 8 + tig * 2 + (j & 1);
        if (k < 27) {
            const int ci = k / 9, r = (k % 9) / 3, s = k % 3;
            kdr[j] = r - 1; kds[j] = s - 1;
            koff[j] = U8 ? (swap_rb ? 2 - ci : ci) : ci * (int)plane + (r - 1) * S + (s - 1);
        } else { kdr[j] = 1 << 20; kds[j] = 0; koff[j] = 0; }
    }
    // epilogue constants (BN shift, PReLU slope) live in shared memory: channels nt*8 + tig*2 + {0,1} per thread
    __shared__ float2 s_b[32], s_a[32];
    if (threadIdx.x < 32) {
        s_b[threadIdx.x] = make_float2(b[threadIdx.x * 2], b[threadIdx.x * 2 + 1]);
        s_a[threadIdx.x] = make_float2(a[threadIdx.x * 2], a[threadIdx.x * 2 + 1]);
    }
    __syncthreads();

    const long long num_tiles = (total + 15) / 16;
    const bool small = total < (1ll << 31) - 64;
    const long long warp_id = (long long)blockIdx.x * 4 + warp;
    const long long num_warps = (long long)gridDim.x * 4;
    // im2col fragments of one 16-row tile (rows m_base + g and m_base + g + 8 for this thread) and their validity
    auto gather = [&](long long tile, uint32_t (&af)[2][4], bool (&pv)[2]) {
        const long long m_base = tile * 16;
        const float* px[2];
        const unsigned char* pu[2];
        int ph[2], pw[2];
        bool pf[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const long long m = m_base + g + rr * 8;
            int n, rem;
            if (small) {                       // 32-bit division: the 64-bit one is a ~100-instruction routine
                n = (int)((unsigned)m / (unsigned)(G * G));
                rem = (int)m - n * G * G;
            } else {
                n = (int)(m / (G * G));
                rem = (int)(m - (long long)n * G * G);
            }
            ph[rr] = (int)((unsigned)rem / (unsigned)G);
            pw[rr] = rem - ph[rr] * G;
            pv[rr] = (m < total) && ph[rr] < S && pw[rr] < S;
            px[rr] = x + (long long)n * 3 * plane + (long long)ph[rr] * S + pw[rr];
            pu[rr] = xu + (long long)n * 3 * plane;
            pf[rr] = U8 && flip != nullptr && pv[rr] && flip[n] != 0;
        }
        // fast path (~80 % of the tiles at S = 112): all 16 rows are pixels at least one step away from the image
        // border, so no tap needs a bounds check (the kernel is instruction-issue bound, not HBM bound)
        const bool in0 = pv[0] && ph[0] >= 1 && ph[0] < S - 1 && pw[0] >= 1 && pw[0] < S - 1;
        const bool in1 = pv[1] && ph[1] >= 1 && ph[1] < S - 1 && pw[1] >= 1 && pw[1] < S - 1;
        if (__all_sync(0xffffffffu, in0 && in1)) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        float v[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int j = ks * 4 + h * 2 + e;
                            if (kdr[j] >= 2) {                     // k >= 27: zero column
                                v[e] = 0.f;
                            } else if (U8) {
                                const int ww = pw[rr] + kds[j];
                                const int ws = pf[rr] ? S - 1 - ww : ww;
                                v[e] = lut[__ldg(pu[rr] + ((ph[rr] + kdr[j]) * S + ws) * 3 + koff[j])];
                            } else {
                                v[e] = __ldg(px[rr] + koff[j]);
                            }
                        }
                        af[ks][h * 2 + rr] = pack_bf16x2(v[0], v[1]);
                    }
            return;
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    float v[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = ks * 4 + h * 2 + e;
                        const int hh = ph[rr] + kdr[j], ww = pw[rr] + kds[j];
                        const bool ok = pv[rr] && hh >= 0 && hh < S && ww >= 0 && ww < S;
                        if (U8) {
                            const int ws = pf[rr] ? S - 1 - ww : ww;
                            v[e] = ok ? lut[__ldg(pu[rr] + ((long long)hh * S + ws) * 3 + koff[j])] : 0.f;
                        } else {
                            v[e] = ok ? __ldg(px[rr] + koff[j]) : 0.f;
                        }
                    }
                    // A fragment order: a0:(g, k lo) a1:(g+8, k lo) a2:(g, k hi) a3:(g+8, k hi)
                    af[ks][h * 2 + rr] = pack_bf16x2(v[0], v[1]);
                }
    };
    // software pipeline: the gathers of the next tile are in flight while this tile's MMAs and stores run
    uint32_t af_next[2][4];
    bool pv_next[2];
    if (warp_id < num_tiles) gather(warp_id, af_next, pv_next);
    for (long long tile = warp_id; tile < num_tiles; tile += num_warps) {
        const long long m_base = tile * 16;
        uint32_t af[2][4];
        bool pv[2];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int q = 0; q < 4; ++q) af[ks][q] = af_next[ks][q];
        pv[0] = pv_next[0]; pv[1] = pv_next[1];
        if (tile + num_warps < num_tiles) gather(tile + num_warps, af_next, pv_next);
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            mma_16816_bf16(c, af[0], bf[nt][0][0], bf[nt][0][1]);
            mma_16816_bf16(c, af[1], bf[nt][1][0], bf[nt][1][1]);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const float2 bb = s_b[nt * 4 + tig], aa = s_a[nt * 4 + tig];
                float v0 = c[rr * 2 + 0] + bb.x, v1 = c[rr * 2 + 1] + bb.y;
                v0 = v0 > 0.f ? v0 : v0 * aa.x;
                v1 = v1 > 0.f ? v1 : v1 * aa.y;
                if (!pv[rr]) { v0 = 0.f; v1 = 0.f; }
                stage[warp][(g + rr * 8) * 32 + ((nt ^ g) & 7) * 4 + tig] = pack_bf16x2(v0, v1);  // XOR-swizzled chunks
            }
        }
        __syncwarp();
        // 16 rows x 128 B, written as 4 coalesced 512-byte stores
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = q * 32 + lane;          // 16-byte chunk index: row = idx / 8, chunk = idx % 8
            const long long m = m_base + (idx >> 3);
            if (m < total) {
                const uint4 v = reinterpret_cast<const uint4*>(stage[warp])[(idx & ~7) | ((idx ^ (idx >> 3)) & 7)];
                reinterpret_cast<uint4*>(out + m * 64)[idx & 7] = v;
            }
        }
        __syncwarp();
    }
}

// ----------------------------------------------------------------------------------------------
// Strip version of the stem for S % 16 == 0 (the 112x112 input of the model). The gather kernel above is bound by
// instruction issue and L1 wavefronts (ncu: 546 warp instructions per 16-pixel tile, 16 scattered 4-byte global loads
// per thread and tile, 64-bit address arithmetic, profiles/r02_ncu_membound_eval_kernels.csv). Here a CTA owns ST_R
// output rows of one image: the (ST_R + 2) x (S + 2) x 3 input patch is staged ONCE in shared memory as bf16 (coalesced
// 16-byte global loads; the uint8 variant applies channel swap / flip / normalisation while staging), zero halo
// included, so the im2col fragments are unconditional 2-byte shared loads at per-thread constant offsets; a warp tile
// is 16 consecutive pixels of one image row; BN shift and PReLU slopes stay in registers. Same K order and the same
// bf16 operands as the gather kernel: bit-identical activations.
// ----------------------------------------------------------------------------------------------
constexpr int ST_R = 8;
template <bool U8>
__global__ void __launch_bounds__(128, 4) stem_strip_kernel(const void* __restrict__ xin, const unsigned char* __restrict__ flip,
                                                         int swap_rb, const float* __restrict__ w,
                                                         const float* __restrict__ b, const float* __restrict__ a,
                                                         __nv_bfloat16* __restrict__ out, int n_img, int S) {
    pdl_sync();
    extern __shared__ __align__(16) uint8_t st_smem[];
    const int P = S + 2;                                    // patch row pitch (elements); even
    const int img_elems = 3 * (ST_R + 2) * P;
    __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(st_smem);
    uint32_t* stage = reinterpret_cast<uint32_t*>(st_smem + ((img_elems * 2 + 15) & ~15));   // [4][16 * 32]
    float* lut = reinterpret_cast<float*>(stage + 4 * 512);                                  // [256] (U8 only)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int G = S + 1;
    const int strips = S / ST_R;
    const int n = blockIdx.x / strips, h0 = (blockIdx.x - n * strips) * ST_R;

    if (U8) {
        for (int i = tid; i < 256; i += 128) lut[i] = __fdiv_rn(__fdiv_rn((float)i, 255.0f) - 0.5f, 0.5f);
        __syncthreads();
    }
    // ---- stage the input patch: rows h0-1 .. h0+ST_R, columns -1 .. S (zero outside the image) ----
    for (int i = tid; i < 3 * (ST_R + 2) * 2; i += 128) {                 // the two halo columns
        const int row = i >> 1;
        img[row * P + ((i & 1) ? S + 1 : 0)] = __float2bfloat16_rn(0.f);
    }
    if (U8) {
        const unsigned char* xu = reinterpret_cast<const unsigned char*>(xin) + (size_t)n * S * S * 3;
        const bool fl = flip != nullptr && flip[n] != 0;
        const int wpr = S * 3 / 4;                                        // 32-bit words per image row
        for (int i = tid; i < (ST_R + 2) * wpr; i += 128) {
            const int rr = i / wpr, wq = i - rr * wpr;
            const int h = h0 - 1 + rr;
            const bool ok = h >= 0 && h < S;
            const uint32_t word = ok ? __ldg(reinterpret_cast<const uint32_t*>(xu + (size_t)h * S * 3) + wq) : 0u;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int byte = wq * 4 + e;
                const int wsrc = byte / 3, c = byte - wsrc * 3;
                const int ci = swap_rb ? 2 - c : c;
                const int wd = fl ? S - 1 - wsrc : wsrc;
                const float v = ok ? lut[(word >> (8 * e)) & 255u] : 0.f;
                img[(ci * (ST_R + 2) + rr) * P + 1 + wd] = __float2bfloat16_rn(v);
            }
        }
    } else {
        const float* x = reinterpret_cast<const float*>(xin) + (size_t)n * 3 * S * S;
        const int qpr = S / 4;                                            // float4 per image row
        for (int i = tid; i < 3 * (ST_R + 2) * qpr; i += 128) {
            const int q = i % qpr, row = i / qpr;                         // row = ci * (ST_R + 2) + rr
            const int ci = row / (ST_R + 2), rr = row - ci * (ST_R + 2);
            const int h = h0 - 1 + rr;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h >= 0 && h < S) v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)ci * S + h) * S) + q);
            __nv_bfloat16* d = img + row * P + 1 + q * 4;
            d[0] = __float2bfloat16_rn(v.x); d[1] = __float2bfloat16_rn(v.y);
            d[2] = __float2bfloat16_rn(v.z); d[3] = __float2bfloat16_rn(v.w);
        }
    }

    // B fragments (weights [K = 32][N = 64] col-major, 8 n-tiles x 2 k-steps) and the epilogue constants of this
    // thread's channels nt*8 + tig*2 + {0, 1}
    uint32_t bf[8][2][2];
    float2 eb[8], ea[8];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = ks * 16 + h * 8 + tig * 2;
                const int nn = nt * 8 + g;
                const float w0 = (k0 < 27) ? w[k0 * 64 + nn] : 0.f;
                const float w1 = (k0 + 1 < 27) ? w[(k0 + 1) * 64 + nn] : 0.f;
                bf[nt][ks][h] = pack_bf16x2(w0, w1);
            }
        const int c = nt * 8 + tig * 2;
        eb[nt] = make_float2(b[c], b[c + 1]);
        ea[nt] = make_float2(a[c], a[c + 1]);
    }
    // this thread's 8 K columns -> offsets into the patch relative to (local row hl, column w): (ci, r, s) reads
    // patch row hl + r, patch column w + s
    int koff[8];
    bool kval[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = (j >> 2) * 16 + ((j >> 1) & 1) * 8 + tig * 2 + (j & 1);
        kval[j] = k < 27;
        const int ci = k / 9, r = (k % 9) / 3, sx = k % 3;
        koff[j] = kval[j] ? (ci * (ST_R + 2) + r) * P + sx : 0;
    }
    __syncthreads();

    const int tiles_per_row = S >> 4;
    const unsigned short* img16 = reinterpret_cast<const unsigned short*>(img);
    uint32_t* wstage = stage + warp * 512;
    for (int t = warp; t < ST_R * tiles_per_row; t += 4) {
        const int hl = t / tiles_per_row, w0 = (t - hl * tiles_per_row) << 4;
        const unsigned short* base = img16 + hl * P + w0 + g;
        uint32_t af[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int j = ks * 4 + h * 2;
                    uint32_t lo = base[koff[j] + rr * 8], hi = base[koff[j + 1] + rr * 8];
                    if (ks == 1) { lo = kval[j] ? lo : 0u; hi = kval[j + 1] ? hi : 0u; }    // k >= 27: zero columns
                    af[ks][h * 2 + rr] = lo | (hi << 16);
                }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            mma_16816_bf16(c, af[0], bf[nt][0][0], bf[nt][0][1]);
            mma_16816_bf16(c, af[1], bf[nt][1][0], bf[nt][1][1]);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                float v0 = c[rr * 2 + 0] + eb[nt].x, v1 = c[rr * 2 + 1] + eb[nt].y;
                v0 = v0 > 0.f ? v0 : v0 * ea[nt].x;
                v1 = v1 > 0.f ? v1 : v1 * ea[nt].y;
                wstage[(g + rr * 8) * 32 + ((nt ^ g) & 7) * 4 + tig] = pack_bf16x2(v0, v1);       // XOR-swizzled chunks
            }
        }
        __syncwarp();
        // 16 pixels x 128 B are contiguous in the flat map: 4 coalesced 512-byte stores
        __nv_bfloat16* orow = out + ((size_t)n * G * G + (size_t)(h0 + hl) * G + w0) * 64;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = q * 32 + lane;          // 16-byte chunk index: row = idx / 8, chunk = idx % 8
            const uint4 v = reinterpret_cast<const uint4*>(wstage)[(idx & ~7) | ((idx ^ (idx >> 3)) & 7)];
            reinterpret_cast<uint4*>(orow)[idx] = v;
        }
        __syncwarp();
    }
    // zero padding of the flat layout: column S of the strip's rows, and the whole row S after the last strip
    if (tid < ST_R * 8)
        reinterpret_cast<uint4*>(out + ((size_t)n * G * G + (size_t)(h0 + (tid >> 3)) * G + S) * 64)[tid & 7] = make_uint4(0, 0, 0, 0);
    if (h0 + ST_R == S) {
        uint4* z = reinterpret_cast<uint4*>(out + ((size_t)n * G * G + (size_t)S * G) * 64);
        for (int i = tid; i < G * 8; i += 128) z[i] = make_uint4(0, 0, 0, 0);
    }
}

static int stem_strip_smem(int S) { return ((3 * (ST_R + 2) * (S + 2) * 2 + 15) & ~15) + 4 * 512 * 4 + 256 * 4; }
static bool g_stem_strip = true;                 // ffr_debug_set_stem_strip(0): the gather kernel for every size (A/B, tests)
void set_stem_strip(bool on) { g_stem_strip = on; }
static bool stem_strip_ok(int n_img, int S) { return g_stem_strip && S % 16 == 0 && S >= 16 && S <= 512 && n_img > 0; }

int stem_launch(const float* x, const float* w, const float* b, const float* a, void* out, int n_img, int S,
                cudaStream_t stream) {
    if (stem_strip_ok(n_img, S)) {
        launch_ex(stem_strip_kernel<false>, dim3(n_img * (S / ST_R)), dim3(128), stem_strip_smem(S), stream, 1, PDL_SIMT, (const void*)x,
                  (const unsigned char*)nullptr, 0, w, b, a, reinterpret_cast<__nv_bfloat16*>(out), n_img, S);
        return launch_status("stem_strip_kernel");
    }
    const long long total = (long long)n_img * (S + 1) * (S + 1);
    const long long tiles = (total + 15) / 16;
    long long grid = (tiles + 3) / 4;
    const long long cap = (long long)num_sms() * 16;
    if (grid > cap) grid = cap;
    launch_ex(stem_kernel<false>, dim3((int)grid), dim3(128), 0, stream, 1, PDL_SIMT, x, nullptr, 0, w, b, a, reinterpret_cast<__nv_bfloat16*>(out), n_img, S);
    return launch_status("stem_kernel");
}

int stem_u8_launch(const unsigned char* img, const unsigned char* flip, int swap_rb, const float* w, const float* b,
                   const float* a, void* out, int n_img, int S, cudaStream_t stream) {
    if (stem_strip_ok(n_img, S)) {
        launch_ex(stem_strip_kernel<true>, dim3(n_img * (S / ST_R)), dim3(128), stem_strip_smem(S), stream, 1, PDL_SIMT, (const void*)img,
                  flip, swap_rb, w, b, a, reinterpret_cast<__nv_bfloat16*>(out), n_img, S);
        return launch_status("stem_strip_kernel<u8>");
    }
    const long long total = (long long)n_img * (S + 1) * (S + 1);
    const long long tiles = (total + 15) / 16;
    long long grid = (tiles + 3) / 4;
    const long long cap = (long long)num_sms() * 16;
    if (grid > cap) grid = cap;
    launch_ex(stem_kernel<true>, dim3((int)grid), dim3(128), 0, stream, 1, PDL_SIMT, img, flip, swap_rb, w, b, a, reinterpret_cast<__nv_bfloat16*>(out), n_img, S);
    return launch_status("stem_kernel<u8>");
}

// ----------------------------------------------------------------------------------------------
// SE gate (model_ir_se50.py:29-36):  s = sigmoid(W2 relu(W1 mean_hw(u)))  per image, from the per-32-row-block partial
// sums the conv2 epilogue stored (ConvGemmParams::pool_part). Image n owns rows [n*rpi, (n+1)*rpi) of the flat map; a
// block b covers rows [32b, 32b+31] and holds the sum of the rows of image floor(32b / rpi) in slot 0 and - when it
// straddles an image boundary (rpi >= 64 > 32: at most two images) - of the next image in slot 1. The blocks are added
// in a fixed order: no atomics, bit-reproducible. dense != 0: pool_part is [n_img][C] finished sums instead (the
// pixel-major experiment, whose epilogue adds per row with atomics). One CTA per image.
// ----------------------------------------------------------------------------------------------
// Gate of image n into s_gate[C] (shared): fixed-order sum of the partial blocks, FC1 + ReLU, FC2 + sigmoid. Called by
// all 256 threads of a CTA; ends with a __syncthreads(). Used by se_gate_kernel and by the fused se_gate_residual_kernel
// (identical arithmetic, so the two paths give bit-identical maps).
template <int C>
__device__ __forceinline__ void se_gate_compute(const float* __restrict__ pool_part, int dense,
                                                const float* __restrict__ w1, const float* __restrict__ w2, int n,
                                                int rpi, float inv, float* s_part /*[PARTS*C]*/, float* s_mean /*[C]*/,
                                                float* s_hid /*[C/16]*/, float* s_gate /*[C]*/) {
    constexpr int R = C / 16;
    constexpr int PARTS = (C >= 256) ? 1 : 256 / C;      // threads (c, part): part strides over the blocks
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (dense) {
        for (int c = tid; c < C; c += 256) s_mean[c] = pool_part[(long long)n * C + c];
    } else {
        const long long r0 = (long long)n * rpi, r1 = r0 + rpi - 1;
        const int b0 = (int)(r0 >> 5), b1 = (int)(r1 >> 5);
        for (int i = tid; i < PARTS * C; i += 256) {
            const int c = i % C, part = i / C;
            float a = 0.f;
            for (int b = b0 + part; b <= b1; b += PARTS) {
                const int slot = (((long long)b << 5) >= r0) ? 0 : 1;      // first row of the block belongs to image n?
                a += pool_part[((long long)b * 2 + slot) * C + c];
            }
            s_part[part * C + c] = a;
        }
        __syncthreads();
        for (int c = tid; c < C; c += 256) {
            float a = s_part[c];
#pragma unroll
            for (int q = 1; q < PARTS; ++q) a += s_part[q * C + c];
            s_mean[c] = a;
        }
    }
    __syncthreads();
    for (int j = warp; j < R; j += 8) {          // FC1: 16-byte loads of a weight row, 4 channels per lane and step
        float a = 0.f;
        const float4* wr = reinterpret_cast<const float4*>(w1 + j * C);
        for (int i = lane; i < C / 4; i += 32) {
            const float4 wv = __ldg(wr + i);
            a = fmaf(wv.x, s_mean[4 * i] * inv, a);
            a = fmaf(wv.y, s_mean[4 * i + 1] * inv, a);
            a = fmaf(wv.z, s_mean[4 * i + 2] * inv, a);
            a = fmaf(wv.w, s_mean[4 * i + 3] * inv, a);
        }
        a = warp_sum(a);
        if (lane == 0) s_hid[j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int c = tid; c < C; c += 256) {           // FC2: the R = C/16 weights of channel c are contiguous
        float a = 0.f;
        const float4* wr = reinterpret_cast<const float4*>(w2 + c * R);
#pragma unroll
        for (int j = 0; j < R / 4; ++j) {
            const float4 wv = __ldg(wr + j);
            a = fmaf(wv.x, s_hid[4 * j], a);
            a = fmaf(wv.y, s_hid[4 * j + 1], a);
            a = fmaf(wv.z, s_hid[4 * j + 2], a);
            a = fmaf(wv.w, s_hid[4 * j + 3], a);
        }
        s_gate[c] = 1.0f / (1.0f + __expf(-a));
    }
    __syncthreads();
}

template <int C>
__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ pool_part, int dense,
                                                      const float* __restrict__ w1, const float* __restrict__ w2,
                                                      float* __restrict__ gate, float* __restrict__ sums, int rpi, float inv) {
    pdl_sync();
    constexpr int PARTS = (C >= 256) ? 1 : 256 / C;
    __shared__ float s_part[PARTS * C];
    __shared__ float s_mean[C];
    __shared__ float s_hid[C / 16];
    __shared__ float s_gate[C];
    const int n = blockIdx.x;
    se_gate_compute<C>(pool_part, dense, w1, w2, n, rpi, inv, s_part, s_mean, s_hid, s_gate);
    for (int c = threadIdx.x; c < C; c += 256) {
        gate[(long long)n * C + c] = s_gate[c];
        if (sums != nullptr) sums[(long long)n * C + c] = s_mean[c];
    }
}

int se_gate_launch(const float* pool_part, int dense, const float* w1, const float* w2, float* gate, float* sums,
                   int n_img, int S, int C, cudaStream_t stream) {
    const int rpi = (S + 1) * (S + 1);
    const float inv = 1.0f / (float)(S * S);
    FFR_CHECK_ARG(rpi >= 64, "se_gate: map %dx%d too small for the 32-row block scheme", S, S);
    switch (C) {
        case 64:  launch_ex(se_gate_kernel<64>, dim3(n_img), dim3(256), 0, stream, 1, PDL_SIMT, pool_part, dense, w1, w2, gate, sums, rpi, inv); break;
        case 128: launch_ex(se_gate_kernel<128>, dim3(n_img), dim3(256), 0, stream, 1, PDL_SIMT, pool_part, dense, w1, w2, gate, sums, rpi, inv); break;
        case 256: launch_ex(se_gate_kernel<256>, dim3(n_img), dim3(256), 0, stream, 1, PDL_SIMT, pool_part, dense, w1, w2, gate, sums, rpi, inv); break;
        case 512: launch_ex(se_gate_kernel<512>, dim3(n_img), dim3(256), 0, stream, 1, PDL_SIMT, pool_part, dense, w1, w2, gate, sums, rpi, inv); break;
        default: return set_error(-1, "se_gate: unsupported C=%d", C);
    }
    return launch_status("se_gate_kernel");
}

// ----------------------------------------------------------------------------------------------
// SE scale + residual (model_ir_se50.py:36, 73-76):  y = u * gate[n] + shortcut, one pass over the flat map.
// shortcut_mode 0: x on the same grid (identity MaxPool(1,1)); 1: x on the (2S)x(2S) grid, subsampled
// (MaxPool(1,2)); 2: sc on the same grid (Conv1x1+BN shortcut, computed by a GEMM).
// A thread owns 8 channels (16 B) of a row; UNR rows are in flight per thread (2 x UNR 16-byte loads).
// ----------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) se_residual_kernel(const __nv_bfloat16* __restrict__ u,
                                                          const float* __restrict__ gate,
                                                          const __nv_bfloat16* __restrict__ sc, int shortcut_mode,
                                                          __nv_bfloat16* __restrict__ y, int S, long long total_rows) {
    pdl_sync();
    constexpr int TPR = C / 8;          // threads per row
    constexpr int RPB = 256 / TPR;      // rows per CTA pass
    constexpr int UNR = 4;
    const int c8 = threadIdx.x % TPR;
    const int rsub = threadIdx.x / TPR;
    const int G = S + 1, rows = G * G, G2 = 2 * S + 1;
    for (long long rb = (long long)blockIdx.x * (RPB * UNR); rb < total_rows; rb += (long long)gridDim.x * (RPB * UNR)) {
        uint4 uv[UNR], sv[UNR];
        float4 g0[UNR], g1[UNR];
        long long row[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            row[k] = rb + k * RPB + rsub;
            if (row[k] < total_rows) {
                const int n = (int)(row[k] / rows);
                long long srow = row[k];
                if (shortcut_mode == 1) {
                    const int r = (int)(row[k] - (long long)n * rows);
                    const int h = r / G, wq = r - h * G;
                    srow = (long long)n * G2 * G2 + (long long)(2 * h) * G2 + 2 * wq;
                }
                uv[k] = __ldg(reinterpret_cast<const uint4*>(u + row[k] * C) + c8);
                sv[k] = __ldg(reinterpret_cast<const uint4*>(sc + srow * C) + c8);
                const float4* gp = reinterpret_cast<const float4*>(gate + (long long)n * C + c8 * 8);
                g0[k] = __ldg(gp);
                g1[k] = __ldg(gp + 1);
            }
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            if (row[k] < total_rows) {
                uint4 o;
                o.x = pack_bf16x2(fmaf(bf16lo(uv[k].x), g0[k].x, bf16lo(sv[k].x)), fmaf(bf16hi(uv[k].x), g0[k].y, bf16hi(sv[k].x)));
                o.y = pack_bf16x2(fmaf(bf16lo(uv[k].y), g0[k].z, bf16lo(sv[k].y)), fmaf(bf16hi(uv[k].y), g0[k].w, bf16hi(sv[k].y)));
                o.z = pack_bf16x2(fmaf(bf16lo(uv[k].z), g1[k].x, bf16lo(sv[k].z)), fmaf(bf16hi(uv[k].z), g1[k].y, bf16hi(sv[k].z)));
                o.w = pack_bf16x2(fmaf(bf16lo(uv[k].w), g1[k].z, bf16lo(sv[k].w)), fmaf(bf16hi(uv[k].w), g1[k].w, bf16hi(sv[k].w)));
                reinterpret_cast<uint4*>(y + row[k] * C)[c8] = o;
            }
        }
    }
}

int se_residual_launch(const void* u, const float* gate, const void* sc, int shortcut_mode, void* y, int n_img, int S,
                       int C, cudaStream_t stream) {
    FFR_CHECK_ARG(shortcut_mode >= 0 && shortcut_mode <= 2, "se_residual: shortcut_mode=%d", shortcut_mode);
    const long long total_rows = (long long)n_img * (S + 1) * (S + 1);
    const int rows_per_pass = (256 / (C / 8)) * 4;
    long long grid = (total_rows + rows_per_pass - 1) / rows_per_pass;
    const long long cap = (long long)num_sms() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    const __nv_bfloat16* up = reinterpret_cast<const __nv_bfloat16*>(u);
    const __nv_bfloat16* sp = reinterpret_cast<const __nv_bfloat16*>(sc);
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
    switch (C) {
        case 64:  launch_ex(se_residual_kernel<64>, dim3((int)grid), dim3(256), 0, stream, 1, PDL_SIMT, up, gate, sp, shortcut_mode, yp, S, total_rows); break;
        case 128: launch_ex(se_residual_kernel<128>, dim3((int)grid), dim3(256), 0, stream, 1, PDL_SIMT, up, gate, sp, shortcut_mode, yp, S, total_rows); break;
        case 256: launch_ex(se_residual_kernel<256>, dim3((int)grid), dim3(256), 0, stream, 1, PDL_SIMT, up, gate, sp, shortcut_mode, yp, S, total_rows); break;
        case 512: launch_ex(se_residual_kernel<512>, dim3((int)grid), dim3(256), 0, stream, 1, PDL_SIMT, up, gate, sp, shortcut_mode, yp, S, total_rows); break;
        default: return set_error(-1, "se_residual: unsupported C=%d", C);
    }
    return launch_status("se_residual_kernel");
}

// ----------------------------------------------------------------------------------------------
// Fused gate + scale + residual: ONE CTA per image computes the image's gate in shared memory (se_gate_compute: 8-50 KB of
// partial sums and FC weights from L2, a few microseconds that the other resident CTAs of the SM hide) and then streams
// the image's rows exactly like se_residual_kernel. Saves the se_gate launch (~9 us of a ~13 us launch + drain
// per unit) for the 21 units with C <= 256; at C = 512 (7x7 maps: 64 KB of rows against 128 KB of FC weights per image)
// the separate gate kernel stays.
// ----------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) se_gate_residual_kernel(const __nv_bfloat16* __restrict__ u,
                                                               const float* __restrict__ pool_part, int dense,
                                                               const float* __restrict__ w1, const float* __restrict__ w2,
                                                               const __nv_bfloat16* __restrict__ sc, int shortcut_mode,
                                                               __nv_bfloat16* __restrict__ y, int S, float inv) {
    pdl_sync();
    constexpr int PARTS = (C >= 256) ? 1 : 256 / C;
    constexpr int TPR = C / 8;          // threads per row
    constexpr int RPB = 256 / TPR;      // rows per CTA pass
    constexpr int UNR = 4;
    __shared__ float s_part[PARTS * C];
    __shared__ float s_mean[C];
    __shared__ float s_hid[C / 16];
    __shared__ __align__(16) float s_gate[C];
    const int n = blockIdx.x;
    const int G = S + 1, rows = G * G, G2 = 2 * S + 1;
    se_gate_compute<C>(pool_part, dense, w1, w2, n, rows, inv, s_part, s_mean, s_hid, s_gate);
    const int c8 = threadIdx.x % TPR;
    const int rsub = threadIdx.x / TPR;
    const float4 g0 = reinterpret_cast<const float4*>(s_gate)[c8 * 2];
    const float4 g1 = reinterpret_cast<const float4*>(s_gate)[c8 * 2 + 1];
    const long long base = (long long)n * rows;
    for (int rb = 0; rb < rows; rb += RPB * UNR) {
        uint4 uv[UNR], sv[UNR];
        int r[UNR];
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            r[k] = rb + k * RPB + rsub;
            if (r[k] < rows) {
                long long srow = base + r[k];
                if (shortcut_mode == 1) {
                    const int h = r[k] / G, wq = r[k] - h * G;
                    srow = (long long)n * G2 * G2 + (long long)(2 * h) * G2 + 2 * wq;
                }
                uv[k] = __ldg(reinterpret_cast<const uint4*>(u + (base + r[k]) * C) + c8);
                sv[k] = __ldg(reinterpret_cast<const uint4*>(sc + srow * C) + c8);
            }
        }
#pragma unroll
        for (int k = 0; k < UNR; ++k) {
            if (r[k] < rows) {
                uint4 o;
                o.x = pack_bf16x2(fmaf(bf16lo(uv[k].x), g0.x, bf16lo(sv[k].x)), fmaf(bf16hi(uv[k].x), g0.y, bf16hi(sv[k].x)));
                o.y = pack_bf16x2(fmaf(bf16lo(uv[k].y), g0.z, bf16lo(sv[k].y)), fmaf(bf16hi(uv[k].y), g0.w, bf16hi(sv[k].y)));
                o.z = pack_bf16x2(fmaf(bf16lo(uv[k].z), g1.x, bf16lo(sv[k].z)), fmaf(bf16hi(uv[k].z), g1.y, bf16hi(sv[k].z)));
                o.w = pack_bf16x2(fmaf(bf16lo(uv[k].w), g1.z, bf16lo(sv[k].w)), fmaf(bf16hi(uv[k].w), g1.w, bf16hi(sv[k].w)));
                reinterpret_cast<uint4*>(y + (base + r[k]) * C)[c8] = o;
            }
        }
    }
}

int se_gate_residual_launch(const void* u, const float* pool_part, int dense, const float* w1, const float* w2,
                            const void* sc, int shortcut_mode, void* y, int n_img, int S, int C, cudaStream_t stream) {
    FFR_CHECK_ARG(shortcut_mode >= 0 && shortcut_mode <= 2, "se_gate_residual: shortcut_mode=%d", shortcut_mode);
    const int rpi = (S + 1) * (S + 1);
    const float inv = 1.0f / (float)(S * S);
    FFR_CHECK_ARG(rpi >= 64, "se_gate_residual: map %dx%d too small for the 32-row block scheme", S, S);
    const __nv_bfloat16* up = reinterpret_cast<const __nv_bfloat16*>(u);
    const __nv_bfloat16* sp = reinterpret_cast<const __nv_bfloat16*>(sc);
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
    switch (C) {
        case 64:  launch_ex(se_gate_residual_kernel<64>, dim3(n_img), dim3(256), 0, stream, 1, PDL_SIMT, up, pool_part, dense, w1, w2, sp, shortcut_mode, yp, S, inv); break;
        case 128: launch_ex(se_gate_residual_kernel<128>, dim3(n_img), dim3(256), 0, stream, 1, PDL_SIMT, up, pool_part, dense, w1, w2, sp, shortcut_mode, yp, S, inv); break;
        case 256: launch_ex(se_gate_residual_kernel<256>, dim3(n_img), dim3(256), 0, stream, 1, PDL_SIMT, up, pool_part, dense, w1, w2, sp, shortcut_mode, yp, S, inv); break;
        case 512: launch_ex(se_gate_residual_kernel<512>, dim3(n_img), dim3(256), 0, stream, 1, PDL_SIMT, up, pool_part, dense, w1, w2, sp, shortcut_mode, yp, S, inv); break;
        default: return set_error(-1, "se_gate_residual: unsupported C=%d", C);
    }
    return launch_status("se_gate_residual_kernel");
}

// ----------------------------------------------------------------------------------------------
// Stride-2 subsample of a flat map (input of the Conv1x1(stride 2) shortcut, model_ir_se50.py:62-63):
// out(n,h,w) = x(n,2h,2w); the pad row/column of the output grid lands on the pad row/column of the input grid.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) subsample2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out,
                                                         int n_img, int So, int C8) {
    pdl_sync();
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
    launch_ex(subsample2_kernel, dim3(grid), dim3(256), 0, stream, 1, PDL_SIMT, reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(out), n_img,
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
    pdl_sync();
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
    launch_ex(export_nchw_kernel, dim3(grid), dim3(256), smem, stream, 1, PDL_SIMT, reinterpret_cast<const __nv_bfloat16*>(h), scale, shift, y, S, C);
    return launch_status("export_nchw_kernel");
}

// ----------------------------------------------------------------------------------------------
// Embedding finish: f = l2_norm(sum_s acc[s] + bias)  (model_ir_se50.py:13-16,141). One warp per row; D = 512.
// `acc` holds the per-split fp32 partial products of the folded head GEMM (BN2d, Linear, BN1d folded into W', b'),
// [splits][rows][D], added here in split order (deterministic split-K).
// ----------------------------------------------------------------------------------------------
// One CTA of 128 threads per row: thread t owns columns 4t .. 4t+3 (float4 loads of every split's partial product, added
// in split order: deterministic), sum of squares by a fixed-order block reduction.
__global__ void __launch_bounds__(128) bias_l2norm_kernel(const float* __restrict__ acc, int splits, long long split_stride,
                                                          const float* __restrict__ bias, float* __restrict__ f, int rows) {
    pdl_sync();
    constexpr int D = 512;
    __shared__ float s_ss[4];
    const int row = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float4* src = reinterpret_cast<const float4*>(acc + (long long)row * D) + t;
    float4 a = __ldg(src);
    for (int s = 1; s < splits; ++s) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(acc + (long long)s * split_stride + (long long)row * D) + t);
        a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
    }
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias) + t);
    a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
    float ss = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, a.w * a.w)));
    ss = warp_sum(ss);
    if (lane == 0) s_ss[warp] = ss;
    __syncthreads();
    const float inv = 1.0f / sqrtf((s_ss[0] + s_ss[1]) + (s_ss[2] + s_ss[3]));
    reinterpret_cast<float4*>(f + (long long)row * D)[t] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}

int bias_l2norm_launch(const float* acc, int splits, long long split_stride, const float* bias, float* f, int rows, int D,
                       cudaStream_t stream) {
    FFR_CHECK_ARG(D == 512, "bias_l2norm: D=%d", D);
    launch_ex(bias_l2norm_kernel, dim3(rows), dim3(128), 0, stream, 1, PDL_SIMT, acc, splits, split_stride, bias, f, rows);
    return launch_status("bias_l2norm_kernel");
}

}  // namespace ffr
