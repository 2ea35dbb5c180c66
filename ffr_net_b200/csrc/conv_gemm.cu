#include "conv_gemm.cuh"
#include "host.h"
#include "kernels.h"
#include "ptx.cuh"

#include <cuda_fp16.h>

namespace ffr {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
// Warp roles. The two single-lane roles that everything else waits on (TMA producer, MMA issuer) get the highest
// warp ids, the eight instruction-heavy epilogue warps the lowest. Their loops are warp-uniform with the issue under
// elect.sync: inside `if (lane == 0)` the compiler wraps every UTCHMMA in an ELECT / BRA.U.ANY loop and each
// tcgen05.mma costs ~150 cycles regardless of N (profiles/r01_mma_issue_microbench.json).
constexpr int NUM_THREADS = 352;                  // warps 0-7 epilogue, warp 8 TMEM alloc, warp 9 TMA, warp 10 MMA
constexpr int EPI_THREADS = 256;                  // two warps per TMEM lane quadrant, each owning half of the columns
constexpr int WARP_ALLOC = 8, WARP_TMA = 9, WARP_MMA = 10;

template <int BN>
struct GemmCfg {
    static constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int STAGES_PAIR = (BN == 256) ? 6 : 8;       // CTA pairs: half of the weight tile per CTA
    static constexpr int TMEM_COLS = 2 * BN;      // double-buffered fp32 accumulator, 128 lanes x BN columns each
    static constexpr int PARAM_FLOATS = 10 * BN;  // 9 border-class biases (or 1) + PReLU slopes
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + 256 /*barriers*/ + PARAM_FLOATS * 4;
    static constexpr int SMEM_BYTES_PAIR = 1024 + STAGES_PAIR * (A_STAGE_BYTES + B_STAGE_BYTES / 2) + 256 + PARAM_FLOATS * 4;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// mbarrier wait that adds the time spent to a counter when profiling counters are enabled
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, bool timed, long long& acc) {
    if (!timed) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
}

// ex2.approx + rcp.approx (2 MUFU ops); an IEEE division here costs ~4x the whole epilogue of a K=64 GEMM
// (tools/mchannel_bench.py: 369 us with, 138 us without the sigmoid before this change)
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// Column sums across the 32 lanes of a warp for 32 per-lane values: after the call lane L holds, in x[0],
// sum over lanes of (their) x[L]. 31 shuffles (butterfly transpose-reduce).
__device__ __forceinline__ float warp_colsum32(float (&x)[32], int lane) {
#pragma unroll
    for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = up ? x[i] : x[i + n];
            const float keep = up ? x[i + n] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return x[0];
}

// q = m / d, r = m % d for 0 <= m < 2^24 via a float reciprocal and one correction step (exact in that range).
__device__ __forceinline__ void fast_divmod(int m, int d, float inv_d, int& q, int& r) {
    q = __float2int_rz(__int2float_rz(m) * inv_d);
    r = m - q * d;
    if (r >= d) { ++q; r -= d; }
    if (r < 0) { --q; r += d; }
}

// ===================== Epilogue: TMEM -> registers -> fused ops -> global =====================
// Runs on warps 0..7 (256 threads). Warp w reads TMEM lanes 32*(w%4).. (tile rows) and the column half w/4.
// A work item is SUB consecutive 128-row tiles (SUB accumulators side by side in the TMEM buffer).
// EPI_PIXMAJOR: M tile -> (output pixel, first image, mask of taps whose source pixel exists)
struct PixTile { int ho, wo, img0; uint32_t tapmask; };
// tile = (output pixel q, block ib of 128 images); ib >= pix_iblocks is an empty tile (CTA pairs with an odd block count)
__device__ __forceinline__ PixTile pix_tile_qi(const ConvGemmParams& p, int q, int ib) {
    PixTile t;
    t.img0 = ib * BLOCK_M;
    const int qh = q / p.pix_side;
    t.ho = qh + p.pix_off;
    t.wo = q - qh * p.pix_side + p.pix_off;
    t.tapmask = 0;
    if (t.ho >= p.pix_pad_from || t.wo >= p.pix_pad_from) return t;
    for (int tap = 0; tap < p.ntaps; ++tap) {
        const int r = (p.ntaps == 9) ? tap / 3 : 1, s = (p.ntaps == 9) ? tap % 3 : 1;
        const int hs = t.ho + r - 1, ws = t.wo + s - 1;
        if (hs >= p.pix_src_lo && hs <= p.pix_src_hi && ws >= p.pix_src_lo && ws <= p.pix_src_hi) t.tapmask |= 1u << tap;
    }
    return t;
}
__device__ __forceinline__ PixTile pix_tile(const ConvGemmParams& p, int m_tile) {
    const int q = m_tile / p.pix_iblocks;
    return pix_tile_qi(p, q, m_tile - q * p.pix_iblocks);
}

// 16-bit pair -> fp32 (bf16 by default, fp16 when f16 is set) and back
__device__ __forceinline__ float cvt16_lo(uint32_t u, bool f16) {
    return f16 ? __half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xFFFFu))) : bf16lo(u);
}
__device__ __forceinline__ float cvt16_hi(uint32_t u, bool f16) {
    return f16 ? __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16))) : bf16hi(u);
}
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, bool f16) {
    if (!f16) return pack_bf16x2(lo, hi);
    const __half2 t = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&t);
}

__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Which work items this CTA's epilogue walks and which M tile a (work item, sub-tile) is. Single CTAs: items
// blockIdx.x, +gridDim.x, ... and tile = m_group * SUB + sub. CTA pairs (conv_win2_kernel): the pair walks items
// cluster_id, +num_clusters, ...; an item is two M tiles, CTA rank r owns tile 2 * m_group + r; the non-leader CTA
// returns its accumulator buffers to the LEADER's tempty barrier (the MMA issuer lives there).
struct EpiSched {
    int work0, work_stride, num_work;
    int tile_mul, tile_add;
    int pix_ibp;                 // > 0 (pixel-major CTA pairs): image-block PAIRS per pixel; m_group = q * pix_ibp + pair,
                                 // this CTA owns image block 2 * pair + tile_add of pixel q
    uint32_t tempty_remote;      // 0, or the shared::cluster address of the leader's tempty_bar[0]
    // stream-K (conv_win2_kernel): work item whose accumulator is only DUMPED to this CTA's slot (-1: none), work item
    // that first adds the PEER pair's dumped partial (-1: none); see ConvGemmParams::sk_enable
    int sk_dump_work, sk_add_work;
    int n_pre, pre_work[4];      // work items walked BEFORE the arithmetic sequence work0, work0 + stride, ... (stream-K: the
                                 // partial items of the pair's step range go early, so that the fix-up of the item
                                 // shared with the next pair overlaps the MMAs of the whole items)
    float* sk_my_ws;             // this CTA's slot   [128][BN] fp32
    const float* sk_peer_ws;     // the slot of the same-rank CTA of the next pair
    int* sk_my_flag;
    int* sk_peer_flag;
};

template <int BN, int SUB>
__device__ __forceinline__ EpiSched epi_sched_single(const ConvGemmParams& p) {
    EpiSched es;
    es.work0 = blockIdx.x;
    es.work_stride = gridDim.x;
    es.num_work = ((p.num_m_tiles + SUB - 1) / SUB) * p.num_n_tiles * p.num_splits;
    es.tile_mul = 1;
    es.tile_add = 0;
    es.pix_ibp = 0;
    es.tempty_remote = 0;
    es.sk_dump_work = es.sk_add_work = -1;
    es.n_pre = 0;
    return es;
}

// FIXED != 0: the epilogue specialised at compile time for exactly this flag set (the kernels dispatch on p.flags among
// the sets the hot layers use, epilogue_dispatch below); every other feature is compiled out. The generic epilogue is
// ~10 k SASS instructions of runtime flag tests; on the 64- and 128-channel layers, which are epilogue-bound (the MMA
// warp waits for accumulator buffers 20-36 % of its time), ncu attributed 9 % of the epilogue warps' samples to
// instruction fetch and 5 % to branch resolution. Specialised: 64 -> 64 @112 583 -> 486 us, whole eval step -5 %.
template <int BN, int SUB, uint32_t FIXED = 0, bool SK = false>
__device__ __forceinline__ void epilogue_loop(const ConvGemmParams& p, const uint32_t tmem_base, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, float* sparam, const int warp, const int lane,
                                              const EpiSched es) {
    constexpr int HALF = BN / 2;                 // columns per epilogue warp
    const int num_work = es.num_work;
    const int quad = warp & 3;                   // TMEM lane quadrant == warp % 4
    const int chalf = warp >> 2;                    // which half of the accumulator columns
    const int row_in_tile = quad * 32 + lane;
    const int etid = threadIdx.x;
    const uint32_t flags = FIXED ? FIXED : (p.flags & ~EPI_GENERIC_ONLY);
    const bool border = (flags & EPI_BORDER_BIAS) != 0;
    const int nbias = border ? 9 : 1;
    const bool small_m = p.M < (1 << 24);
    // 32-byte aligned output rows (the S2D / plain addressing adds multiples of 64 channels): 256-bit stores
    const bool out32 = ((reinterpret_cast<uintptr_t>(p.out) | static_cast<uintptr_t>(p.ldo * 2)) & 31) == 0;
    const float inv_rpi = 1.0f / (float)max(p.rows_per_img, 1);
    const float inv_wp = 1.0f / (float)max(p.Wp, 1);
    int loaded_n_tile = -1;
    int it = 0;
    const bool timed = p.dbg != nullptr;
    long long w_e = 0;
    const long long t_begin = clock64();
    for (int widx = 0;; ++widx, ++it) {
        int work;
        if (SK && widx < es.n_pre) {
            work = (widx == 0) ? es.pre_work[0] : (widx == 1 ? es.pre_work[1] : (widx == 2 ? es.pre_work[2] : es.pre_work[3]));
        } else {
            work = es.work0 + (widx - (SK ? es.n_pre : 0)) * es.work_stride;
            if (work >= num_work) break;
        }
        const int t = work / p.num_splits;
        const int split = work - t * p.num_splits;
        const int n_tile = t % p.num_n_tiles;
        const int m_group = t / p.num_n_tiles;
        const int n0 = n_tile * BN;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const bool sk_dump = SK && work == es.sk_dump_work;
        const bool sk_add = SK && work == es.sk_add_work;
        if (sk_add) {                            // the peer dumped its part of this item at the START of its run
            if (etid == 0) {
                while (ld_acquire_gpu(es.sk_peer_flag) == 0) __nanosleep(64);
                *es.sk_peer_flag = 0;            // idle again for the next launch (stream order)
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }

        if (n_tile != loaded_n_tile) {           // stage per-channel epilogue parameters in smem
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (flags & (EPI_BIAS | EPI_BORDER_BIAS))
                for (int i = etid; i < nbias * BN; i += EPI_THREADS)
                    sparam[i] = p.bias[(i / BN) * p.Cout + n0 + (i % BN)];
            if (flags & EPI_PRELU)
                for (int i = etid; i < BN; i += EPI_THREADS) sparam[9 * BN + i] = p.slope[n0 + i] - 1.0f;   // slope - 1
            asm volatile("bar.sync 1, 256;" ::: "memory");
            loaded_n_tile = n_tile;
        }

#pragma unroll 1
        for (int sub = 0; sub < SUB; ++sub) {
        int m_tile = (m_group * SUB + sub) * es.tile_mul + es.tile_add;
        bool tile_ok = m_tile < p.num_m_tiles;
        int pq = 0, pib = 0;
        if (es.pix_ibp > 0) {
            pq = m_group / es.pix_ibp;
            pib = 2 * (m_group - pq * es.pix_ibp) + es.tile_add;
            tile_ok = pib < p.pix_iblocks;
            m_tile = pq * p.pix_iblocks + pib;
        }
        int m = m_tile * BLOCK_M + row_in_tile;
        int n_img = 0, h = 0, w = 0, r_local = 0;
        bool valid = m < p.M;
        int cls = 0;
        if (flags & EPI_PIXMAJOR) {              // row = image (img0 + row_in_tile) at the tile's pixel
            const PixTile pt = (es.pix_ibp > 0) ? pix_tile_qi(p, pq, pib) : pix_tile(p, m_tile);
            n_img = pt.img0 + row_in_tile;
            r_local = pt.ho * p.Wp + pt.wo;
            m = n_img * p.rows_per_img + r_local;
            valid = n_img < p.n_img;
            if (!valid) m = p.M;                 // keeps every "m < p.M" guard below false
            h = pt.ho - p.h0;
            w = pt.wo - p.h0;
            if ((flags & EPI_GEOM) && !(flags & EPI_SCATTER)) valid = valid && h >= 0 && h < p.S && w >= 0 && w < p.S;
            if (border) {
                const int ch = (h == 0) ? 0 : ((h == p.S - 1) ? 2 : 1);
                const int cw = (w == 0) ? 0 : ((w == p.S - 1) ? 2 : 1);
                cls = ch * 3 + cw;
            }
        } else if (flags & EPI_GEOM) {
            if (small_m) {
                fast_divmod(m, p.rows_per_img, inv_rpi, n_img, r_local);
                fast_divmod(r_local, p.Wp, inv_wp, h, w);
            } else {
                n_img = m / p.rows_per_img;
                r_local = m - n_img * p.rows_per_img;
                h = r_local / p.Wp;
                w = r_local - h * p.Wp;
            }
            h -= p.h0;
            w -= p.h0;
            if (!(flags & EPI_SCATTER)) valid = valid && h >= 0 && h < p.S && w >= 0 && w < p.S;
            if (border) {
                const int ch = (h == 0) ? 0 : ((h == p.S - 1) ? 2 : 1);
                const int cw = (w == 0) ? 0 : ((w == p.S - 1) ? 2 : 1);
                cls = ch * 3 + cw;
            }
        }
        // explicit ld.shared (a float* into dynamic shared memory compiles to generic LD.E: long-scoreboard latency)
        const uint32_t sbias_a = smem_u32(sparam) + static_cast<uint32_t>(cls * BN + chalf * HALF) * 4u;
        const uint32_t sslope_a = smem_u32(sparam) + static_cast<uint32_t>(9 * BN + chalf * HALF) * 4u;
        const int nc0 = n0 + chalf * HALF;       // first global output channel of this warp

        // output addressing
        __nv_bfloat16* orow = nullptr;
        bool do_store = false;
        __nv_bfloat16* sdst[8];
        if (flags & EPI_SCATTER) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                sdst[k] = nullptr;
                if (k < p.scatter_n && m < p.M) {
                    const int2 e = __ldg(p.scatter + r_local * p.scatter_n + k);
                    if (e.x >= 0)
                        sdst[k] = p.out + ((long long)n_img * p.out_rows_per_img + e.x) * p.ldo + e.y + nc0;
                }
            }
            valid = valid && sdst[0] != nullptr;
        } else if (p.out != nullptr) {
            if (flags & EPI_OUT_S2D) {
                const int g = p.s2d_So + 1;
                const long long r = (long long)n_img * g * g + (h >> 1) * g + (w >> 1);
                orow = p.out + r * p.ldo + ((h & 1) * 2 + (w & 1)) * p.Cout + nc0;
                do_store = valid;
            } else {
                orow = p.out + (long long)m * p.ldo + nc0;
                do_store = (m < p.M);
            }
        }
        const int n_lo = __shfl_sync(0xffffffffu, n_img, 0);
        const int n_hi = __shfl_sync(0xffffffffu, n_img, 31);

        if (sub == 0) {
            mbar_wait_timed(&tfull_bar[acc], acc_phase, timed, w_e);
            tc_fence_after();
        }
        const uint32_t t_row = tmem_base + (acc * SUB + sub) * BN + chalf * HALF + (static_cast<uint32_t>(quad * 32) << 16);

        // EPI_COSFACE per-row running values (one row per thread, HALF classes per tile)
        const int ce_lab = ((flags & EPI_COSFACE) && p.ce_label != nullptr && m < p.M) ? __ldg(p.ce_label + m) : -1;
        float ce_sum = 0.f, ce_zl = 0.f, ce_best = -3.0e38f;
        int ce_bestc = 0;

        uint32_t vbuf[2][32];
        tmem_ld_32x32(t_row, vbuf[0]);
#pragma unroll
        for (int ci = 0; ci < HALF / 32; ++ci) {
            const int c0 = ci * 32;
            tmem_ld_wait();
            if (ci + 1 < HALF / 32) tmem_ld_32x32(t_row + c0 + 32, vbuf[(ci + 1) & 1]);   // prefetch next chunk
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(vbuf[ci & 1][j]);
            if (SK && sk_dump) {                 // raw partial accumulator -> this CTA's stream-K slot, nothing else
                // slot layout [column / 4][row][4]: the 32 lanes (rows) of a warp store 512 contiguous bytes per instruction
                float4* o = reinterpret_cast<float4*>(es.sk_my_ws) + ((chalf * HALF + c0) / 4) * BLOCK_M + row_in_tile;
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q * BLOCK_M] = make_float4(x[q * 4], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
                continue;
            }
            if (SK && sk_add) {                  // all eight loads in flight before the first add (L2 round trips)
                const float* q0 = es.sk_peer_ws + ((long long)((chalf * HALF + c0) / 4) * BLOCK_M + row_in_tile) * 4;
                float4 t[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) t[q] = ld_cg_f32x4(q0 + q * (BLOCK_M * 4));
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    x[q * 4 + 0] += t[q].x; x[q * 4 + 1] += t[q].y; x[q * 4 + 2] += t[q].z; x[q * 4 + 3] += t[q].w;
                }
            }
            if (flags & (EPI_BIAS | EPI_BORDER_BIAS)) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b4 = lds_f32x4(sbias_a + (ci * 8 + q) * 16);
                    x[q * 4 + 0] += b4.x; x[q * 4 + 1] += b4.y; x[q * 4 + 2] += b4.z; x[q * 4 + 3] += b4.w;
                }
            }
            if (flags & EPI_PRELU) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 s4 = lds_f32x4(sslope_a + (ci * 8 + q) * 16);
                    // PReLU as x + (slope - 1) * min(x, 0): one FMNMX + one FFMA per element
                    x[q * 4 + 0] = fmaf(s4.x, fminf(x[q * 4 + 0], 0.f), x[q * 4 + 0]);
                    x[q * 4 + 1] = fmaf(s4.y, fminf(x[q * 4 + 1], 0.f), x[q * 4 + 1]);
                    x[q * 4 + 2] = fmaf(s4.z, fminf(x[q * 4 + 2], 0.f), x[q * 4 + 2]);
                    x[q * 4 + 3] = fmaf(s4.w, fminf(x[q * 4 + 3], 0.f), x[q * 4 + 3]);
                }
            }
            if (flags & (EPI_RESIDUAL | EPI_MUL_DSIG)) {
                if (valid) {
                    const bool rf16 = (flags & EPI_RES_F16) != 0;
                    const uint4* rp = reinterpret_cast<const uint4*>(p.res + (long long)m * p.ldres + nc0 + c0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 r = __ldg(rp + q);
                        const float rv[8] = {cvt16_lo(r.x, rf16), cvt16_hi(r.x, rf16), cvt16_lo(r.y, rf16), cvt16_hi(r.y, rf16),
                                             cvt16_lo(r.z, rf16), cvt16_hi(r.z, rf16), cvt16_lo(r.w, rf16), cvt16_hi(r.w, rf16)};
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            if (flags & EPI_MUL_DSIG) x[q * 8 + e] *= rv[e] * (1.0f - rv[e]);
                            else                      x[q * 8 + e] += rv[e];
                        }
                    }
                }
            }
            if (flags & EPI_SIGMOID) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = sigmoidf_fast(x[j]);
            }
            if (!valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = 0.f;
            }
            if (flags & EPI_COSFACE) {
                const int cbase = nc0 + c0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int c = cbase + j;
                    if (c < p.ce_classes) {
                        const float z = p.ce_s * (x[j] - ((c == ce_lab) ? p.ce_m : 0.f));
                        ce_sum += __expf(z - p.ce_s);
                        if (c == ce_lab) ce_zl = z;
                        if (x[j] > ce_best) { ce_best = x[j]; ce_bestc = c; }
                    }
                }
            }
            if (flags & EPI_OUT_F32_ATOMIC) {
                if (valid) {
                    float* o = p.out_f32 + (long long)m * p.Cout + nc0 + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(o + j, x[j]);
                }
            }
            if (flags & EPI_OUT_F32) {
                if (m < p.M) {
                    float4* o = reinterpret_cast<float4*>(p.out_f32 + (long long)split * p.out_f32_split_stride +
                                                          (long long)m * p.Cout + nc0 + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] = make_float4(x[q * 4], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
                }
            }
            if (do_store || (flags & EPI_SCATTER)) {
                uint4 pk[4];
                const bool of16 = (flags & EPI_OUT_F16) != 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    pk[q].x = pack16x2(x[q * 8 + 0], x[q * 8 + 1], of16);
                    pk[q].y = pack16x2(x[q * 8 + 2], x[q * 8 + 3], of16);
                    pk[q].z = pack16x2(x[q * 8 + 4], x[q * 8 + 5], of16);
                    pk[q].w = pack16x2(x[q * 8 + 6], x[q * 8 + 7], of16);
                }
                if (do_store) {
                    if (out32) {         // two 32-byte stores: each fills a whole sector of the thread's row
                        st_global_256(orow + c0, pk[0], pk[1]);
                        st_global_256(orow + c0 + 16, pk[2], pk[3]);
                    } else {
                        uint4* o = reinterpret_cast<uint4*>(orow + c0);
#pragma unroll
                        for (int q = 0; q < 4; ++q) o[q] = pk[q];
                    }
                }
                if (flags & EPI_SCATTER) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (sdst[k] != nullptr) {
                            uint4* o = reinterpret_cast<uint4*>(sdst[k] + c0);
#pragma unroll
                            for (int q = 0; q < 4; ++q) o[q] = pk[q];
                        }
                    }
                }
            }
            if (flags & EPI_STATS) {   // x is already zero on invalid rows
                float sq[32], xs[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) { sq[j] = x[j] * x[j]; xs[j] = x[j]; }
                const float s1 = warp_colsum32(xs, lane);
                const float s2 = warp_colsum32(sq, lane);
                if (p.stats_part != nullptr) {     // one plain store per (M tile, quadrant, channel): deterministic
                    if (tile_ok) {
                        float* o = p.stats_part + ((long long)(m_tile * 4 + quad) * 2) * p.Cout + nc0 + c0 + lane;
                        o[0] = s1;
                        o[p.Cout] = s2;
                    }
                } else {
                    atomicAdd(p.stats + nc0 + c0 + lane, s1);
                    atomicAdd(p.stats + p.Cout + nc0 + c0 + lane, s2);
                }
            }
            if ((flags & EPI_POOL) && (flags & EPI_PIXMAJOR)) {   // every row is a different image
                if (valid) {
                    float* o = p.pool + (long long)n_img * p.Cout + nc0 + c0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) red_add_f32x4(o + q * 4, x[q * 4], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
                }
            } else if (flags & EPI_POOL) {    // x is already zero on invalid rows
                // deterministic: one plain store per (32-row block, image slot, channel); else atomics into pool
                float* part = (p.pool_part != nullptr && tile_ok)
                    ? p.pool_part + ((long long)(m_tile * 4 + quad) * 2) * p.Cout + nc0 + c0 + lane : nullptr;
                if (n_lo == n_hi) {
                    const float s = warp_colsum32(x, lane);
                    if (n_lo < p.n_img) {
                        if (part) part[0] = s;
                        else if (p.pool_part == nullptr) atomicAdd(p.pool + (long long)n_lo * p.Cout + nc0 + c0 + lane, s);
                    }
                } else {               // the warp's 32 rows straddle two images
                    float xb[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        xb[j] = (n_img == n_lo) ? 0.f : x[j];
                        x[j] = (n_img == n_lo) ? x[j] : 0.f;
                    }
                    const float sa = warp_colsum32(x, lane);
                    const float sb = warp_colsum32(xb, lane);
                    if (n_lo < p.n_img) {
                        if (part) part[0] = sa;
                        else if (p.pool_part == nullptr) atomicAdd(p.pool + (long long)n_lo * p.Cout + nc0 + c0 + lane, sa);
                    }
                    if (n_hi < p.n_img) {
                        if (part) part[p.Cout] = sb;
                        else if (p.pool_part == nullptr) atomicAdd(p.pool + (long long)n_hi * p.Cout + nc0 + c0 + lane, sb);
                    }
                }
            }
        }
        if ((flags & EPI_COSFACE) && m < p.M) {
            if (p.ce_sumexp_part != nullptr)
                p.ce_sumexp_part[(long long)m * (2 * p.num_n_tiles) + n_tile * 2 + chalf] = ce_sum;
            else if (p.ce_sumexp != nullptr) atomicAdd(p.ce_sumexp + m, ce_sum);
            if (p.ce_zlabel != nullptr && ce_lab >= nc0 && ce_lab < nc0 + HALF) p.ce_zlabel[m] = ce_zl;
            if (nc0 < p.ce_classes) {
                uint32_t u = __float_as_uint(ce_best);
                u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);            // monotone map float -> uint
                atomicMax(p.ce_argkey + m, ((unsigned long long)u << 32) | (0xFFFFFFFFu - (uint32_t)ce_bestc));
            }
        }
        }   // sub
        tc_fence_before();
        if (es.tempty_remote) mbar_arrive_cluster(es.tempty_remote + acc * 8);
        else mbar_arrive(&tempty_bar[acc]);
        if (SK && sk_dump) {                     // every epilogue thread's stores are visible before the flag goes up
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (etid == 0) st_release_gpu(es.sk_my_flag, 1);
        }
    }
    if (timed && threadIdx.x == 0) {
        atomicAdd(p.dbg + DBG_EPI_WAIT, (unsigned long long)w_e);
        atomicAdd(p.dbg + DBG_EPI_TOTAL, (unsigned long long)(clock64() - t_begin));
    }
}

// Flag sets with a specialised epilogue. Backbone: conv1 of a unit (BatchNorm shift as border bias + PReLU, plain or
// space-to-depth store), conv2 (BatchNorm shift + SE squeeze sums), the 1x1 shortcut. RecNet eval (pixel-major tiles):
// ConvLayer and the second ConvLayer of a ResidualBlock. RecNet training: forward (fp32 z + BatchNorm partial sums),
// data gradient, plain fp32 GEMM.
constexpr uint32_t FS_CONV1 = EPI_GEOM | EPI_BORDER_BIAS | EPI_PRELU;
constexpr uint32_t FS_CONV1_S2D = FS_CONV1 | EPI_OUT_S2D;
constexpr uint32_t FS_CONV2 = EPI_GEOM | EPI_BIAS | EPI_POOL;
constexpr uint32_t FS_SHORTCUT = EPI_GEOM | EPI_BIAS;
constexpr uint32_t FS_REC = EPI_GEOM | EPI_BIAS | EPI_PRELU | EPI_SCATTER | EPI_PIXMAJOR;
constexpr uint32_t FS_REC_RES = FS_REC | EPI_RESIDUAL;
constexpr uint32_t FS_TRAIN_FWD = EPI_GEOM | EPI_STATS | EPI_OUT_F32 | EPI_PIXMAJOR;
constexpr uint32_t FS_TRAIN_DGRAD = EPI_OUT_F32 | EPI_PIXMAJOR | EPI_PIX_DGRAD;
constexpr uint32_t FS_F32 = EPI_OUT_F32;
constexpr uint32_t FS_MCHANNEL = EPI_BIAS | EPI_SIGMOID;        // M_channel = sigmoid(h W^T + b): a K = 64 GEMM, all epilogue
constexpr uint32_t FS_FEAT_CHANNEL = EPI_GEOM | EPI_SCATTER;    // M_channel @ X into the flip / cat slots

// WINDOW: the sliding-window kernels only ever see the backbone sets (and RecNet's row-major small-batch layers, generic)
template <int BN, int SUB, bool WINDOW, bool SK = false>
__device__ __forceinline__ void epilogue_dispatch(const ConvGemmParams& p, const uint32_t tmem_base, uint64_t* tfull_bar,
                                                  uint64_t* tempty_bar, float* sparam, const int warp, const int lane,
                                                  const EpiSched es) {
    switch (p.flags) {
        case FS_CONV1:     epilogue_loop<BN, SUB, FS_CONV1, SK>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
        case FS_CONV1_S2D: epilogue_loop<BN, SUB, FS_CONV1_S2D, SK>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
        case FS_CONV2:     epilogue_loop<BN, SUB, FS_CONV2, SK>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
        default: break;
    }
    if constexpr (!WINDOW) {
        switch (p.flags) {
            case FS_SHORTCUT:    epilogue_loop<BN, SUB, FS_SHORTCUT>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
            case FS_REC:         epilogue_loop<BN, SUB, FS_REC>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
            case FS_REC_RES:     epilogue_loop<BN, SUB, FS_REC_RES>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
            case FS_TRAIN_FWD:   epilogue_loop<BN, SUB, FS_TRAIN_FWD>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
            case FS_TRAIN_DGRAD: epilogue_loop<BN, SUB, FS_TRAIN_DGRAD>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
            case FS_F32:         epilogue_loop<BN, SUB, FS_F32>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
            case FS_MCHANNEL:    epilogue_loop<BN, SUB, FS_MCHANNEL>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
            case FS_FEAT_CHANNEL: epilogue_loop<BN, SUB, FS_FEAT_CHANNEL>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es); return;
            default: break;
        }
    }
    epilogue_loop<BN, SUB, 0, SK>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es);
}

// PAIR: CTA pairs (cluster of 2, tcgen05 cta_group::2). A work item is two M tiles x one N tile: CTA rank r loads the A
// tile of ITS M tile and rows [r * BN/2, (r+1) * BN/2) of the weight tile; the leader issues M = 256 MMAs that read both
// CTAs' shared memory. Per SM the weight bytes halve: the tile-per-tap kernel streams 16 KB (A) + 32 KB (B) per 512 MMA
// cycles at BN = 256 = 96 B/cycle/SM with single CTAs — more than twice what the L2 delivers when all SMs read
// (~43 B/cycle/SM) — and 64 B/cycle/SM in pairs. Pixel-major tiles pair two image blocks of the SAME pixel (same
// tap mask); with batched B both tiles must share a batch (b_mtile_div even).
template <int BN, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvGemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int B_BYTES = PAIR ? Cfg::B_STAGE_BYTES / 2 : Cfg::B_STAGE_BYTES;     // this CTA's part of a weight tile
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_BYTES;
    constexpr int STAGES = PAIR ? Cfg::STAGES_PAIR : Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;                   // [STAGES]  TMA -> MMA   (pairs: the leader's copy is used)
    uint64_t* empty_bar = bars + STAGES;         // [STAGES]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * STAGES;     // [2]       MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;// [2]       epilogue -> MMA (pairs: the leader's copy is used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    float* sparam = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const bool leader = rank == 0;

    if (warp == WARP_TMA && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == WARP_MMA && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], PAIR ? 2 * EPI_THREADS : EPI_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == WARP_ALLOC) {
        if constexpr (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
        else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();        // everything above touched no global memory; the predecessor's outputs are visible from here on

    // work items: single CTAs walk M tiles; pairs walk PAIRS of M tiles (pixel-major: pairs of image blocks per pixel)
    const int pix_ibp = (PAIR && (p.flags & EPI_PIXMAJOR)) ? (p.pix_iblocks + 1) / 2 : 0;
    const int m_items = !PAIR ? p.num_m_tiles
                              : (pix_ibp > 0 ? (p.num_m_tiles / p.pix_iblocks) * pix_ibp : (p.num_m_tiles + 1) / 2);
    const int num_work = m_items * p.num_n_tiles * p.num_splits;
    const int work0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int work_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int kb_total = p.ntaps * p.kpt_a;
    const int ibp_div = pix_ibp > 0 ? pix_ibp : 1;

    if (warp == WARP_TMA) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t full0 = PAIR ? mapa_shared(smem_u32(&full_bar[0]), 0) : 0u;   // the leader's full barriers
        for (int work = work0; work < num_work; work += work_stride) {
            const int split = work % p.num_splits;
            const int t = work / p.num_splits;
            const int n_tile = t % p.num_n_tiles;
            const int m_item = t / p.num_n_tiles;
            const int m_tile = PAIR ? 2 * m_item + rank : m_item;       // (not used by pixel-major pairs)
            const int m0 = m_tile * BLOCK_M;
            const int b_row = n_tile * BN + (m_tile / p.b_mtile_div) * p.b_rows_per_mtile + (PAIR ? rank * (BN / 2) : 0);
            if (p.flags & EPI_PIXMAJOR) {        // A tile = 128 images at the tap's source pixel; taps outside are skipped
                PixTile pt;
                if (pix_ibp > 0) {
                    const int q = m_item / ibp_div;
                    pt = pix_tile_qi(p, q, 2 * (m_item - q * ibp_div) + rank);
                } else {
                    pt = pix_tile(p, m_item);
                }
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    if (!((pt.tapmask >> tap) & 1u)) continue;
                    const int r = (p.ntaps == 9) ? tap / 3 : 1, s = (p.ntaps == 9) ? tap % 3 : 1;
                    for (int c = 0; c < p.kpt_a; ++c) {          // hi chunks, then (a_hilo) the lo chunks on the same weights
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        const int cb = (c < p.kb_per_tap) ? c : c - p.kb_per_tap;
                        const int a_col = p.tap_ch_off[tap] + cb * BLOCK_K + ((c < p.kb_per_tap) ? 0 : p.a_lo_off);
                        if (elect_one_sync()) {
                            if constexpr (PAIR) {
                                if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                                tma_load_4d_pair(sA + stage * A_STAGE_BYTES, &tmA, full0 + stage * 8, a_col,
                                                 pt.wo + s - 1, pt.ho + r - 1, pt.img0);
                                tma_load_2d_pair(sB + stage * B_BYTES, &tmB, full0 + stage * 8,
                                                 (tap * p.kb_per_tap + cb) * BLOCK_K, b_row);
                            } else {
                                mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                                tma_load_4d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], a_col,
                                            pt.wo + s - 1, pt.ho + r - 1, pt.img0);
                                tma_load_2d(sB + stage * B_BYTES, &tmB, &full_bar[stage],
                                            (tap * p.kb_per_tap + cb) * BLOCK_K, b_row);
                            }
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                continue;
            }
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(kb0 + p.kb_per_split, kb_total);
            int tap = kb0 / p.kpt_a;
            int c = kb0 - tap * p.kpt_a;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const int cb = (c < p.kb_per_tap) ? c : c - p.kb_per_tap;
                const int a_col = p.tap_ch_off[tap] + cb * BLOCK_K + ((c < p.kb_per_tap) ? 0 : p.a_lo_off);
                const int a_row = m0 + p.tap_row_shift[tap];
                if (elect_one_sync()) {
                    if constexpr (PAIR) {
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                        tma_load_2d_pair(sA + stage * A_STAGE_BYTES, &tmA, full0 + stage * 8, a_col, a_row);
                        tma_load_2d_pair(sB + stage * B_BYTES, &tmB, full0 + stage * 8,
                                         (tap * p.kb_per_tap + cb) * BLOCK_K, b_row);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                        tma_load_2d(sA + stage * A_STAGE_BYTES, &tmA, &full_bar[stage], a_col, a_row);
                        tma_load_2d(sB + stage * B_BYTES, &tmB, &full_bar[stage],
                                    (tap * p.kb_per_tap + cb) * BLOCK_K, b_row);
                    }
                }
                __syncwarp();
                if (++c == p.kpt_a) { c = 0; ++tap; }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == WARP_MMA && leader) {
        // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
        const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, BN) ^ p.idesc_xor;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int work = work0; work < num_work; work += work_stride, ++it) {
            const int split = work % p.num_splits;
            int kb0 = split * p.kb_per_split;
            int kb1 = min(kb0 + p.kb_per_split, kb_total);
            if (p.flags & EPI_PIXMAJOR) {        // the producer streams only the taps whose source pixel exists
                const int m_item = (work / p.num_splits) / p.num_n_tiles;
                const int q = (pix_ibp > 0) ? m_item / ibp_div : m_item / p.pix_iblocks;
                kb0 = 0;
                kb1 = __popc(pix_tile_qi(p, q, 0).tapmask) * p.kpt_a;
            }
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            if (kb1 <= kb0) {                    // pixel-major pad point: nothing to accumulate, the epilogue stores zeros
                if (elect_one_sync()) {
                    if constexpr (PAIR) umma_commit_pair(&tfull_bar[acc]);
                    else umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                continue;
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
                const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(sB + stage * B_BYTES));
                const uint32_t first = (kb > kb0) ? 1u : 0u;
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k) {  // +32 B per K step == +2 in the (addr >> 4) field
                        if constexpr (PAIR) umma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (k > 0) ? 1u : first);
                        else umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (k > 0) ? 1u : first);
                    }
                    if constexpr (PAIR) {
                        umma_commit_pair(&empty_bar[stage]);
                        if (kb == kb1 - 1) umma_commit_pair(&tfull_bar[acc]);
                    } else {
                        umma_commit(&empty_bar[stage]);          // frees the smem slot once these MMAs retire
                        if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete -> epilogue
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < 8) {
        EpiSched es;
        es.work0 = work0;
        es.work_stride = work_stride;
        es.num_work = num_work;
        es.tile_mul = PAIR ? 2 : 1;
        es.tile_add = rank;
        es.pix_ibp = pix_ibp;
        es.tempty_remote = (PAIR && !leader) ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : 0u;
        es.sk_dump_work = es.sk_add_work = -1;
        es.n_pre = 0;
        epilogue_dispatch<BN, 1, false>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es);
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();       // no CTA leaves (or frees TMEM) while its peer may still touch it
    if (warp == WARP_ALLOC) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
        else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}


// ----------------------------------------------------------------------------------------------------------
// Sliding-window variant for 3x3 / stride-1 convolutions on a flat map of row pitch G.
// For one 128-row tile and one 64-channel chunk all nine taps read rows inside the window
// [m0 - G - 1, m0 + 128 + G + 1): it is loaded ONCE by TMA (instead of nine shifted 128-row tiles) and tap (r,s)
// is just an operand descriptor that starts r*G + s rows (128 B each) into the window. SWIZZLE_128B is a function of
// the absolute shared-memory address, so a row-offset start address needs no base_offset (profiles/r01_probe_*).
// Weights stream through their own ring, one [BN x 64] tile per (chunk, tap).
// ----------------------------------------------------------------------------------------------------------
constexpr int WIN_MAX_A_STAGES = 4;
constexpr int WIN_MAX_B_STAGES = 9;

struct WinCfg {
    int G;            // row pitch of the flat map
    int box_rows;     // rows per TMA box (multiple of 8, <= 256)
    int nbox;         // 1 or 2 boxes per window
    int a_stages, b_stages;
    int b_resident;   // all 9 weight tiles of the (single) channel chunk stay in smem for the whole kernel
};

// SUB = 128-row sub-tiles per work item. With SUB = 2 every weight tile that lands in smem feeds two MMAs groups
// (halving the weight traffic through shared memory, which is what bounds the 64/128-channel layers) and the
// window halo is amortised over 256 rows. Needs 2*SUB*BN <= 512 TMEM columns.
// TB = taps per MMA issue batch. The tensor pipe buffers only ~one instruction, so whatever the issuing lane does
// between two batches (barrier wait, descriptor arithmetic, ~250 cycles) is exposed unless a batch is long: TB weight
// stages are awaited together and their TB*SUB*4 MMAs go out as straight-line code (b_stages is a multiple of TB).
template <int BN, int SUB, int TB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_win_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const ConvGemmParams p, const WinCfg wc) {
    constexpr int B_STAGE_BYTES = BN * BLOCK_K * 2;
    constexpr uint32_t TMEM_COLS = 2 * SUB * BN;
    static_assert(TMEM_COLS <= 512, "accumulators do not fit TMEM");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int win_bytes = wc.box_rows * wc.nbox * 128;            // multiple of 1024
    uint8_t* sA = smem;
    uint8_t* sB = smem + wc.a_stages * win_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + wc.b_stages * B_STAGE_BYTES);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + WIN_MAX_A_STAGES;
    uint64_t* b_full = bars + 2 * WIN_MAX_A_STAGES;
    uint64_t* b_empty = b_full + WIN_MAX_B_STAGES;
    uint64_t* tfull_bar = b_empty + WIN_MAX_B_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* sparam = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == WARP_TMA && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == WARP_MMA && lane == 0) {
        for (int s = 0; s < wc.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < wc.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], EPI_THREADS); }
        fence_mbar_init();
    }
    if (warp == WARP_ALLOC) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    const int num_work = ((p.num_m_tiles + SUB - 1) / SUB) * p.num_n_tiles;
    const int chunks = p.kb_per_tap;
    const bool resident = wc.b_resident != 0;

    if (warp == WARP_TMA) {
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        bool b_loaded = false;
        const bool timed = p.dbg != nullptr;
        long long w_a = 0, w_b = 0;
        const long long t_begin = clock64();
        for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
            const int n_tile = work % p.num_n_tiles;
            const int m_group = work / p.num_n_tiles;
            const int row0 = m_group * (SUB * BLOCK_M) - wc.G - 1;
            const int n0 = n_tile * BN;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait_timed(&a_empty[sa], pa ^ 1, timed, w_a);
                uint8_t* dst = sA + sa * win_bytes;
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&a_full[sa], win_bytes);
                    tma_load_2d(dst, &tmA, &a_full[sa], c * BLOCK_K, row0);
                    if (wc.nbox == 2)
                        tma_load_2d(dst + wc.box_rows * 128, &tmA, &a_full[sa], c * BLOCK_K, row0 + wc.box_rows);
                }
                __syncwarp();
                if (++sa == wc.a_stages) { sa = 0; pa ^= 1; }
                if (resident && b_loaded) continue;
#pragma unroll 1
                for (int t = 0; t < 9; ++t) {
                    if (!resident) mbar_wait_timed(&b_empty[sb], pb ^ 1, timed, w_b);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&b_full[sb], B_STAGE_BYTES);
                        tma_load_2d(sB + sb * B_STAGE_BYTES, &tmB, &b_full[sb], (t * chunks + c) * BLOCK_K, n0);
                    }
                    __syncwarp();
                    if (++sb == wc.b_stages) { sb = 0; pb ^= 1; }
                }
            }
            b_loaded = true;
        }
        if (timed && lane == 0) {
            atomicAdd(p.dbg + DBG_TMA_WAIT_A, (unsigned long long)w_a);
            atomicAdd(p.dbg + DBG_TMA_WAIT_B, (unsigned long long)w_b);
            atomicAdd(p.dbg + DBG_TMA_TOTAL, (unsigned long long)(clock64() - t_begin));
        }
    } else if (warp == WARP_MMA) {
        const uint32_t idesc = umma_idesc_bf16(BLOCK_M, BN) ^ p.idesc_xor;
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        int it = 0;
        const bool timed = p.dbg != nullptr;
        long long w_t = 0, w_a = 0, w_b = 0;
        const long long t_begin = clock64();
        for (int work = blockIdx.x; work < num_work; work += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait_timed(&tempty_bar[acc], acc_phase ^ 1, timed, w_t);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * (SUB * BN);
            for (int c = 0; c < chunks; ++c) {
                mbar_wait_timed(&a_full[sa], pa, timed, w_a);
                tc_fence_after();
                const uint32_t win_lo = static_cast<uint32_t>(umma_smem_desc_sw128(smem_u32(sA + sa * win_bytes)));
#pragma unroll 1
                for (int tb = 0; tb < 9 / TB; ++tb) {
                    if (!resident || it == 0) {
#pragma unroll
                        for (int j = 0; j < TB; ++j) mbar_wait_timed(&b_full[sb + j], pb, timed, w_b);
                    }
                    tc_fence_after();
                    const uint32_t b_lo = static_cast<uint32_t>(umma_smem_desc_sw128(smem_u32(sB + sb * B_STAGE_BYTES)));
                    // tap t = tb*TB + j = (r, sx): A rows start (r*G + sx) rows into the window, 8 (>>4 units) per row
                    const uint32_t a_lo =
                        win_lo + static_cast<uint32_t>((TB == 9 ? 0 : (TB == 3 ? tb * wc.G : (tb / 3) * wc.G + tb % 3)) * 8);
                    const uint32_t g8 = static_cast<uint32_t>(wc.G * 8);
                    const uint32_t first = (c > 0 || tb > 0) ? 1u : 0u;
                    if (elect_one_sync()) {
#pragma unroll
                        for (int j = 0; j < TB; ++j) {
                            const uint32_t a_tap = (TB == 9) ? a_lo + (j / 3) * g8 + (j % 3) * 8 : a_lo + j * 8;
#pragma unroll
                            for (int sub = 0; sub < SUB; ++sub)
#pragma unroll
                                for (int k = 0; k < BLOCK_K / 16; ++k)
                                    umma_bf16(d_tmem + sub * BN, UMMA_DESC_HI | (a_tap + sub * 1024 + 2 * k),
                                              UMMA_DESC_HI | (b_lo + j * (B_STAGE_BYTES >> 4) + 2 * k), idesc,
                                              (j > 0 || k > 0) ? 1u : first);
                            if (!resident) umma_commit(&b_empty[sb + j]);
                        }
                        if (tb == 9 / TB - 1) {
                            umma_commit(&a_empty[sa]);
                            if (c == chunks - 1) umma_commit(&tfull_bar[acc]);
                        }
                    }
                    __syncwarp();
                    sb += TB;
                    if (sb == wc.b_stages) { sb = 0; pb ^= 1; }
                }
                if (++sa == wc.a_stages) { sa = 0; pa ^= 1; }
            }
        }
        if (timed && lane == 0) {
            atomicAdd(p.dbg + DBG_MMA_WAIT_TMEM, (unsigned long long)w_t);
            atomicAdd(p.dbg + DBG_MMA_WAIT_A, (unsigned long long)w_a);
            atomicAdd(p.dbg + DBG_MMA_WAIT_B, (unsigned long long)w_b);
            atomicAdd(p.dbg + DBG_MMA_TOTAL, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.dbg + DBG_CTAS, 1ull);
        }
    } else if (warp < 8) {
        epilogue_dispatch<BN, SUB, true>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, epi_sched_single<BN, SUB>(p));
    }

    tc_fence_before();
    __syncthreads();
    if (warp == WARP_ALLOC) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_base);
    }
}

// ----------------------------------------------------------------------------------------------------------
// CTA-pair variant of the sliding-window kernel (cluster of 2, tcgen05 cta_group::2), N = 256 layers.
// With one CTA per tile every SM streams the WHOLE weight matrix through shared memory for each of its tiles: at
// 256 -> 256 that is 1.15 MB of weights + 80 KB of windows per 18.4 k MMA cycles = 67 B/cycle/SM, 9.9 KB/cycle over the
// chip — above what the L2 delivers to 148 SMs reading at once (~6.3 KB/cycle, B300_MICROARCH.md "LTS throughput cap"),
// so the single-CTA kernel is L2-feed bound, not tensor bound. A pair computes 256 rows x 256 channels per work item:
// each CTA loads the window of ITS 128 rows and HALF of every weight tile (128 of the 256 output channels); the
// leader issues one M = 256 MMA that reads both halves. Weight bytes per SM halve (36 B/cycle/SM).
// Barriers: full barriers live in the leader (both CTAs' TMA bytes are credited there), empty / tfull barriers are
// per CTA and signalled by multicast commits, tempty lives in the leader and counts the epilogue threads of both CTAs.
// ----------------------------------------------------------------------------------------------------------
template <int BN, int TB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv_win2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvGemmParams p, const WinCfg wc) {
    constexpr int B_HALF_BYTES = (BN / 2) * BLOCK_K * 2;          // this CTA's half of a weight tile
    constexpr uint32_t TMEM_COLS = 2 * BN;
    static_assert(TMEM_COLS <= 512, "accumulators do not fit TMEM");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int win_bytes = wc.box_rows * wc.nbox * 128;            // multiple of 1024
    uint8_t* sA = smem;
    uint8_t* sB = smem + wc.a_stages * win_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + wc.b_stages * B_HALF_BYTES);
    uint64_t* a_full = bars;
    uint64_t* a_empty = bars + WIN_MAX_A_STAGES;
    uint64_t* b_full = bars + 2 * WIN_MAX_A_STAGES;
    uint64_t* b_empty = b_full + WIN_MAX_B_STAGES;
    uint64_t* tfull_bar = b_empty + WIN_MAX_B_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* sparam = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (p.dbg != nullptr && threadIdx.x == 0) atomicMin(p.dbg + DBG_T_ENTRY, globaltimer_ns());
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (warp == WARP_TMA && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == WARP_MMA && lane == 0) {
        for (int s = 0; s < wc.a_stages; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < wc.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 2 * EPI_THREADS); }
        fence_mbar_init();
    }
    if (warp == WARP_ALLOC) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                   // the peer's barriers are initialised before anything is signalled there
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();

    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_work = ((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
    const int chunks = p.kb_per_tap;
    // Work = (item, chunk) steps. Without stream-K pair c walks whole items c, c + num_clusters, ...; with it the pair
    // owns the contiguous step range [s0, s1) of the item-major order, which may begin and end inside an item. Its steps
    // run as four ranges: [s0, hb) = the last chunks of an item the previous pair began (dumped to the scratch); up to
    // two whole items [hb, mb); [tb, s1) = the first chunks of an item the next pair finishes (this pair adds the
    // peer's part and runs the fused epilogue - the peer dumped that part as ITS first range, two items ago, and the
    // fix-up runs under the MMAs of what follows); then the remaining whole items [mb, tb).
    const bool sk = p.sk_enable != 0;
    const int steps_total = num_work * chunks;
    const int s0 = sk ? (int)((long long)cluster_id * steps_total / num_clusters) : 0;
    const int s1 = sk ? (int)((long long)(cluster_id + 1) * steps_total / num_clusters) : 0;
    const int hb = ((s0 + chunks - 1) / chunks) * chunks;
    const int tb = (s1 / chunks) * chunks;
    const int mb = min(hb + 2 * chunks, tb);
    const int n_rng = sk ? 4 : 1;

    if (warp == WARP_TMA) {
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        const uint32_t a_full0 = mapa_shared(smem_u32(&a_full[0]), 0);     // the leader's full barriers
        const uint32_t b_full0 = mapa_shared(smem_u32(&b_full[0]), 0);
        for (int rg = 0; rg < n_rng; ++rg) {
            int step = !sk ? 0 : (rg == 0 ? s0 : (rg == 1 ? hb : (rg == 2 ? tb : mb)));
            const int step_end = !sk ? 0 : (rg == 0 ? hb : (rg == 1 ? mb : (rg == 2 ? s1 : tb)));
            int work = sk ? step / chunks : cluster_id;
            int c = sk ? step - work * chunks : 0;
            while (sk ? (step < step_end) : (work < num_work)) {
                const int n_tile = work % p.num_n_tiles;
                const int m_pair = work / p.num_n_tiles;
                const int row0 = (m_pair * 2 + (int)rank) * BLOCK_M - wc.G - 1;
                const int n0 = n_tile * BN + (int)rank * (BN / 2);
                mbar_wait(&a_empty[sa], pa ^ 1);
                uint8_t* dst = sA + sa * win_bytes;
                if (elect_one_sync()) {
                    if (leader) mbar_arrive_expect_tx(&a_full[sa], 2 * win_bytes);
                    tma_load_2d_pair(dst, &tmA, a_full0 + sa * 8, c * BLOCK_K, row0);
                    if (wc.nbox == 2)
                        tma_load_2d_pair(dst + wc.box_rows * 128, &tmA, a_full0 + sa * 8, c * BLOCK_K, row0 + wc.box_rows);
                }
                __syncwarp();
                if (++sa == wc.a_stages) { sa = 0; pa ^= 1; }
#pragma unroll 1
                for (int t = 0; t < 9; ++t) {
                    mbar_wait(&b_empty[sb], pb ^ 1);
                    if (elect_one_sync()) {
                        if (leader) mbar_arrive_expect_tx(&b_full[sb], 2 * B_HALF_BYTES);
                        tma_load_2d_pair(sB + sb * B_HALF_BYTES, &tmB, b_full0 + sb * 8, (t * chunks + c) * BLOCK_K, n0);
                    }
                    __syncwarp();
                    if (++sb == wc.b_stages) { sb = 0; pb ^= 1; }
                }
                ++step;
                if (++c == chunks) { c = 0; work += sk ? 1 : num_clusters; }
            }
        }
    } else if (warp == WARP_MMA && leader) {
        const uint32_t idesc = umma_idesc_bf16(2 * BLOCK_M, BN) ^ p.idesc_xor;
        int sa = 0, sb = 0;
        uint32_t pa = 0, pb = 0;
        int it = 0;
        const bool timed = p.dbg != nullptr;
        long long w_t = 0, w_a = 0, w_b = 0;
        const long long t_begin = clock64();
        if (timed && lane == 0) atomicMin(p.dbg + DBG_T_MMA_BEGIN, globaltimer_ns());
        for (int rg = 0; rg < n_rng; ++rg) {
            int step = !sk ? 0 : (rg == 0 ? s0 : (rg == 1 ? hb : (rg == 2 ? tb : mb)));
            const int step_end = !sk ? 0 : (rg == 0 ? hb : (rg == 1 ? mb : (rg == 2 ? s1 : tb)));
            int work = sk ? step / chunks : cluster_id;
            int c = sk ? step - work * chunks : 0;
            bool seg_first = true;               // first chunk of an accumulator segment (an item, or its part in range)
            while (sk ? (step < step_end) : (work < num_work)) {
                const bool seg_last = (c == chunks - 1) || (sk && step == step_end - 1);
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                if (seg_first) {
                    mbar_wait_timed(&tempty_bar[acc], acc_phase ^ 1, timed, w_t);
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + acc * BN;
                mbar_wait_timed(&a_full[sa], pa, timed, w_a);
                tc_fence_after();
                const uint32_t win_lo = static_cast<uint32_t>(umma_smem_desc_sw128(smem_u32(sA + sa * win_bytes)));
#pragma unroll 1
                for (int tb_ = 0; tb_ < 9 / TB; ++tb_) {
#pragma unroll
                    for (int j = 0; j < TB; ++j) mbar_wait_timed(&b_full[sb + j], pb, timed, w_b);
                    tc_fence_after();
                    const uint32_t b_lo = static_cast<uint32_t>(umma_smem_desc_sw128(smem_u32(sB + sb * B_HALF_BYTES)));
                    const uint32_t a_lo =
                        win_lo + static_cast<uint32_t>((TB == 9 ? 0 : (TB == 3 ? tb_ * wc.G : (tb_ / 3) * wc.G + tb_ % 3)) * 8);
                    const uint32_t g8 = static_cast<uint32_t>(wc.G * 8);
                    const uint32_t first = (!seg_first || tb_ > 0) ? 1u : 0u;
                    if (elect_one_sync()) {
#pragma unroll
                        for (int j = 0; j < TB; ++j) {
                            const uint32_t a_tap = (TB == 9) ? a_lo + (j / 3) * g8 + (j % 3) * 8 : a_lo + j * 8;
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                umma_bf16_pair(d_tmem, UMMA_DESC_HI | (a_tap + 2 * k),
                                               UMMA_DESC_HI | (b_lo + j * (B_HALF_BYTES >> 4) + 2 * k), idesc,
                                               (j > 0 || k > 0) ? 1u : first);
                            umma_commit_pair(&b_empty[sb + j]);
                        }
                        if (tb_ == 9 / TB - 1) {
                            umma_commit_pair(&a_empty[sa]);
                            if (seg_last) umma_commit_pair(&tfull_bar[acc]);
                        }
                    }
                    __syncwarp();
                    sb += TB;
                    if (sb == wc.b_stages) { sb = 0; pb ^= 1; }
                }
                if (++sa == wc.a_stages) { sa = 0; pa ^= 1; }
                seg_first = seg_last;
                if (seg_last) ++it;
                ++step;
                if (++c == chunks) { c = 0; work += sk ? 1 : num_clusters; }
            }
        }
        if (timed && lane == 0) {
            atomicAdd(p.dbg + DBG_MMA_WAIT_TMEM, (unsigned long long)w_t);
            atomicAdd(p.dbg + DBG_MMA_WAIT_A, (unsigned long long)w_a);
            atomicAdd(p.dbg + DBG_MMA_WAIT_B, (unsigned long long)w_b);
            atomicAdd(p.dbg + DBG_MMA_TOTAL, (unsigned long long)(clock64() - t_begin));
            atomicAdd(p.dbg + DBG_CTAS, 1ull);
            atomicMax(p.dbg + DBG_T_MMA_END, globaltimer_ns());
            atomicMax(p.dbg + DBG_MMA_TOTAL_MAX, (unsigned long long)(clock64() - t_begin));
        }
    } else if (warp < 8) {
        EpiSched es;
        es.tile_mul = 2;
        es.tile_add = (int)rank;
        es.pix_ibp = 0;
        es.tempty_remote = leader ? 0u : mapa_shared(smem_u32(&tempty_bar[0]), 0);
        es.sk_dump_work = es.sk_add_work = -1;
        es.n_pre = 0;
        if (!sk) {
            es.work0 = cluster_id;
            es.work_stride = num_clusters;
            es.num_work = num_work;
            epilogue_dispatch<BN, 1, true, false>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es);
        } else {
            if (hb > s0) es.sk_dump_work = s0 / chunks;   // the item's first chunks belong to the previous pair: dump
            if (s1 > tb) es.sk_add_work = tb / chunks;    // ... its last chunks to the next pair: add the peer's part
            int pre[4] = {-1, -1, -1, -1}, np = 0;
            if (hb > s0) pre[np++] = es.sk_dump_work;
            for (int w = hb / chunks; w < mb / chunks; ++w) pre[np++] = w;
            if (s1 > tb) pre[np++] = es.sk_add_work;
            es.pre_work[0] = pre[0]; es.pre_work[1] = pre[1]; es.pre_work[2] = pre[2]; es.pre_work[3] = pre[3];
            es.n_pre = np;
            es.work0 = mb / chunks;               // the remaining whole items
            es.work_stride = 1;
            es.num_work = tb / chunks;
            const long long slot = (long long)BLOCK_M * BN;
            es.sk_my_ws = p.sk_ws + ((long long)cluster_id * 2 + rank) * slot;
            es.sk_peer_ws = p.sk_ws + ((long long)(cluster_id + 1) * 2 + rank) * slot;
            es.sk_my_flag = p.sk_flags + cluster_id * 2 + rank;
            es.sk_peer_flag = p.sk_flags + (cluster_id + 1) * 2 + rank;
            epilogue_dispatch<BN, 1, true, true>(p, tmem_base, tfull_bar, tempty_bar, sparam, warp, lane, es);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                   // no CTA leaves (or frees TMEM) while its peer may still touch it
    if (warp == WARP_ALLOC) {
        tc_fence_after();
        tmem_dealloc_pair<TMEM_COLS>(tmem_base);
    }
    if (p.dbg != nullptr && threadIdx.x == 0) atomicMax(p.dbg + DBG_T_EXIT, globaltimer_ns());
}

static bool g_lean = true;       // ffr_debug_set_lean_epilogue(0): always the generic epilogue (A/B runs, tests)
void set_lean_epilogue(bool on) { g_lean = on; }

template <int BN, int TB>
static int launch_win2(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, const WinCfg& wc,
                       int smem_bytes, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        FFR_CUDA(cudaFuncSetAttribute(conv_win2_kernel<BN, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_set = true;
    }
    FFR_CUDA(launch_ex(conv_win2_kernel<BN, TB>, dim3(grid), dim3(NUM_THREADS), smem_bytes, stream, 1, PDL_GEMM, tmA, tmB, p, wc));
    return launch_status("conv_win2_kernel");
}

template <int BN, int SUB, int TB>
static int launch_win(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, const WinCfg& wc,
                      int smem_bytes, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        FFR_CUDA(cudaFuncSetAttribute(conv_win_kernel<BN, SUB, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      232448));
        attr_set = true;
    }
    FFR_CUDA(launch_ex(conv_win_kernel<BN, SUB, TB>, dim3(grid), dim3(NUM_THREADS), smem_bytes, stream, 1, PDL_GEMM, tmA, tmB, p, wc));
    return launch_status("conv_win_kernel");
}

static bool g_use_window = true;
void set_use_window(bool on) { g_use_window = on; }
// CTA pairs: non-zero = on (default -1), 0 = off (tests, A/B runs)
static int g_pair_mode = -1;
void set_pair_mode(int mode) { g_pair_mode = mode; }
static unsigned long long* g_dbg = nullptr;
void set_debug_counters(unsigned long long* dptr) { g_dbg = dptr; }
// stream-K scratch of the CTA-pair window kernel (ConvGemmParams::sk_enable): registered by the caller for the launches
// that follow on this host thread (the library never allocates); layout [flags: 1024 B][pairs][2][128][256] fp32.
// The flag words must be zero when registered; every launch leaves them zero again.
static void* g_sk_scratch = nullptr;
static long long g_sk_bytes = 0;
static int g_streamk = 0;       // measured neutral inside a step (see DESIGN.md section 3 finding 10): opt-in
void set_conv_scratch(void* ptr, long long bytes) { g_sk_scratch = ptr; g_sk_bytes = bytes; }
long long conv_scratch_bytes() { return 1024 + (long long)(num_sms() / 2) * 2 * BLOCK_M * 256 * (long long)sizeof(float); }
void set_streamk(int on) { g_streamk = on; }
static int g_last_streamk = 0;          // did the most recent CTA-pair window launch run stream-K? (tests)
int last_streamk() { return g_last_streamk; }

// Is this launch a plain 3x3/stride-1 flat convolution (taps (r-1)*G + (s-1), no channel offsets)?
static bool window_eligible(const ConvGemmParams& p, int* G_out) {
    if (!g_use_window || p.ntaps != 9 || p.num_splits != 1 || p.b_rows_per_mtile != 0 || p.a_hilo) return false;
    const int G = p.tap_row_shift[7] - p.tap_row_shift[4];
    if (G < 4) return false;
    for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s)
            if (p.tap_row_shift[r * 3 + s] != (r - 1) * G + (s - 1) || p.tap_ch_off[r * 3 + s] != 0) return false;
    *G_out = G;
    return true;
}

template <int BN>
static int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, int grid,
                      cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        FFR_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES));
        attr_set = true;
    }
    FFR_CUDA(launch_ex(conv_gemm_kernel<BN, false>, dim3(grid), dim3(NUM_THREADS), Cfg::SMEM_BYTES, stream, 1, PDL_GEMM, tmA, tmB, p));
    return launch_status("conv_gemm_kernel");
}

// CTA pairs: `grid` = 2 x (number of pairs), launched as clusters of 2
template <int BN>
static int launch_cfg_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvGemmParams& p, int grid,
                           cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        FFR_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::SMEM_BYTES_PAIR));
        attr_set = true;
    }
    FFR_CUDA(launch_ex(conv_gemm_kernel<BN, true>, dim3(grid), dim3(NUM_THREADS), Cfg::SMEM_BYTES_PAIR, stream, 2, PDL_GEMM, tmA, tmB, p));
    return launch_status("conv_gemm_kernel<pair>");
}

// Host entry used by every C-ABI wrapper. `a`: activation matrix [a_rows, a_ld]; `wp`: packed weights
// [Cout, ntaps*Cin]. Fills the tiling fields of `p` (M, Cout, ntaps, kb_per_tap, taps, geometry and epilogue
// fields must be set by the caller).
// Pixel-major tiles pay off once 49 pixel tiles of 128 images undercut the row-major tile count (81 rows per image, of
// which 32 are halo) by more than what the sliding-window kernel wins back (~8 %): n >= ~96 except just above a
// multiple of 128. g_pix_mode: -1 = this rule, 0 / 1 = forced off / on (ffr_debug_set_pixmajor, tests and tuning).
static int g_pix_mode = -1;
// Experiment switch (default off): backbone 3x3/s1 convs on maps with S <= this use pixel-major tiles over the
// halo-shared flat layout. Measured at N=512 (tools/pix_backbone_bench.py, profiles/r01_pix_backbone_bench_512.json):
// slower than the sliding-window kernel (8.72 ms -> 8.90 ms for S <= 7, 10.1 ms for S <= 14): the 13-23 % fewer MMA
// rows do not pay for nine A tiles per k-chunk instead of one window and per-row pooling reductions.
static int g_pix_backbone_max_s = 0;
void set_pixmajor_mode(int mode) { g_pix_mode = mode; }
void set_pixmajor_backbone(int max_s) { g_pix_backbone_max_s = max_s; }
bool pixmajor_backbone(int S, int n_img) { return S <= g_pix_backbone_max_s && n_img >= 1; }
bool pixmajor_profitable(int n_img) {
    if (g_pix_mode >= 0) return g_pix_mode != 0;
    const long long pix = 49LL * ((n_img + BLOCK_M - 1) / BLOCK_M);
    const long long rowmajor = (81LL * n_img + BLOCK_M - 1) / BLOCK_M;
    return pix * 108 < rowmajor * 100;
}

// same question for a contraction over pixels in k-blocks of 64 rows (wgrad): 49 * ceil(n/64) against ceil(81 n / 64)
bool pixmajor_profitable_k64(int n_img) {
    if (g_pix_mode >= 0) return g_pix_mode != 0;
    return 49LL * ((n_img + 63) / 64) * 105 < ((81LL * n_img + 63) / 64) * 100;
}

int conv_gemm_launch(const void* a, long long a_rows, int a_cols, int a_ld, const void* wp, int Cin, ConvGemmParams p,
                     int num_splits, cudaStream_t stream) {
    p.dbg = g_dbg;
    if (!g_lean) p.flags |= EPI_GENERIC_ONLY;      // matches no specialised flag set; masked off by the generic epilogue
    FFR_CHECK_ARG(Cin % BLOCK_K == 0, "conv_gemm: Cin=%d not a multiple of 64", Cin);
    FFR_CHECK_ARG(p.Cout % 64 == 0, "conv_gemm: Cout=%d not a multiple of 64", p.Cout);
    FFR_CHECK_ARG(p.ntaps >= 1 && p.ntaps <= 9, "conv_gemm: ntaps=%d", p.ntaps);
    int BN = (p.Cout % 256 == 0) ? 256 : ((p.Cout % 128 == 0) ? 128 : 64);
    const bool pix = (p.flags & EPI_PIXMAJOR) != 0;
    if (pix) {
        FFR_CHECK_ARG(p.Wp >= 2 && p.rows_per_img == p.Wp * p.Wp && p.n_img > 0 && (p.ntaps == 9 || p.ntaps == 1) &&
                      num_splits <= 1 && p.b_rows_per_mtile == 0, "conv_gemm: pixel-major tiles need a square flat map");
        const bool dgrad = (p.flags & EPI_PIX_DGRAD) != 0;
        p.pix_iblocks = (p.n_img + BLOCK_M - 1) / BLOCK_M;
        if (p.h0 == 0) {          // halo-shared zero-padded map (backbone): outputs S x S from (0,0); taps that leave
            FFR_CHECK_ARG(!dgrad && p.S == p.Wp - 1, "conv_gemm: pixel-major zero-pad map needs S = Wp - 1");
            p.pix_side = p.Wp;    // the grid are zero-filled by TMA, the shared pad row/column is stored as zeros.
            p.pix_off = 0;        // The pad points are output tiles too (no taps: they only store the zeros the
            p.pix_src_lo = -1;    // next convolution's taps expect there)
            p.pix_src_hi = p.Wp;
            p.pix_pad_from = p.S;
        } else {
            p.pix_pad_from = 1 << 30;                  // H9 (reflection halo materialised)
            FFR_CHECK_ARG(p.Wp == 9 && p.S == 7 && p.h0 == 1, "conv_gemm: pixel-major H9 map expected");
            p.pix_side = dgrad ? 9 : 7;
            p.pix_off = dgrad ? 0 : 1;
            p.pix_src_lo = dgrad ? 1 : 0;
            p.pix_src_hi = dgrad ? 7 : 8;
        }
        p.M = p.n_img * p.rows_per_img;
        // N tile: fewest (waves x cycles per MMA) over the SMs; the measured issue rates are 128 / 64 / 48 cycles
        const int m_tiles = p.pix_side * p.pix_side * p.pix_iblocks;
        long long best = -1;
        const int cand[3] = {256, 128, 64}, rate[3] = {128, 64, 48};
        for (int i = 0; i < 3; ++i) {
            if (p.Cout % cand[i]) continue;
            const long long tiles = (long long)m_tiles * (p.Cout / cand[i]);
            const long long cost = ((tiles + num_sms() - 1) / num_sms()) * (rate[i] + 8);
            if (best < 0 || cost < best) { best = cost; BN = cand[i]; }
        }
    }
    p.kb_per_tap = Cin / BLOCK_K;
    p.a_hilo = p.a_hilo ? 1 : 0;
    p.kpt_a = p.kb_per_tap * (1 + p.a_hilo);
    if (p.b_mtile_div < 1) p.b_mtile_div = 1;
    FFR_CHECK_ARG(!p.a_hilo || (p.a_lo_off % 8 == 0 && p.a_lo_off + Cin <= a_cols), "conv_gemm: bad hi/lo layout");
    const int kb_total = p.ntaps * p.kpt_a;
    if (num_splits < 1) num_splits = 1;
    p.kb_per_split = (kb_total + num_splits - 1) / num_splits;
    p.num_splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.num_m_tiles = pix ? p.pix_side * p.pix_side * p.pix_iblocks : (p.M + BLOCK_M - 1) / BLOCK_M;
    int Gw = 0;
    const bool win_pair = !pix && BN == 256 && g_pair_mode != 0 && p.num_m_tiles >= 2 && window_eligible(p, &Gw);
    // (Measured and rejected: 128-wide N tiles for this kernel to halve the wave quantum — 900 half-size items = 6.5 tile
    // times instead of 7 at 256 -> 256 @14x14 — run 106 us against 91 us: every window is loaded twice and the N = 128
    // pair MMA does not reach the N = 256 rate. profiles/r02_role_counters_pair.json)
    p.num_n_tiles = p.Cout / BN;
    FFR_CHECK_ARG(p.num_splits == 1 || (p.flags & EPI_OUT_F32_ATOMIC) ||
                  ((p.flags & EPI_OUT_F32) && p.out_f32_split_stride >= (long long)p.M * p.Cout),
                  "conv_gemm: split-K needs the atomic epilogue or per-split fp32 outputs");
    if (p.flags & EPI_SCATTER)
        FFR_CHECK_ARG(p.scatter && p.scatter_n >= 1 && p.scatter_n <= 8 && p.out && p.out_rows_per_img > 0,
                      "conv_gemm: bad scatter table");
    if (p.flags & (EPI_POOL | EPI_OUT_S2D | EPI_BORDER_BIAS | EPI_SCATTER))
        FFR_CHECK_ARG(p.flags & EPI_GEOM, "conv_gemm: epilogue needs row geometry");
    if (p.flags & EPI_GEOM) FFR_CHECK_ARG(p.rows_per_img >= 32 && p.Wp > 0, "conv_gemm: bad geometry");

    const long long num_work = (long long)p.num_m_tiles * p.num_n_tiles * p.num_splits;
    const int grid = (int)((num_work < num_sms()) ? num_work : num_sms());
    if (grid == 0) return 0;

    CUtensorMap tmA, tmB;
    const uint64_t b_rows = (uint64_t)p.Cout + (uint64_t)((p.num_m_tiles - 1) / p.b_mtile_div) * (uint64_t)p.b_rows_per_mtile;
    int rc = make_tmap_2d_bf16(&tmB, wp, b_rows, (uint64_t)p.ntaps * Cin, (uint64_t)p.ntaps * Cin, BN);
    if (rc) return rc;

    int G = 0;
    // CTA pairs for the tile-per-tap kernel: two M tiles per work item, half of every weight tile per CTA
    const bool pair_ok = g_pair_mode != 0 && p.num_m_tiles >= 2 && (p.b_rows_per_mtile == 0 || p.b_mtile_div % 2 == 0);
    long long pair_items = pix ? (long long)(p.num_m_tiles / p.pix_iblocks) * ((p.pix_iblocks + 1) / 2)
                               : (long long)(p.num_m_tiles + 1) / 2;
    pair_items *= (long long)p.num_n_tiles * p.num_splits;
    const int max_pairs = num_sms() / 2;
    const int pair_grid = 2 * (int)((pair_items < max_pairs) ? pair_items : max_pairs);
    if (pix) {
        rc = make_tmap_pixel_bf16(&tmA, a, (uint64_t)p.n_img, (uint32_t)p.Wp, (uint64_t)a_cols, (uint64_t)a_ld, BLOCK_M);
        if (rc) return rc;
        if (pair_ok) {
            rc = make_tmap_2d_bf16(&tmB, wp, b_rows, (uint64_t)p.ntaps * Cin, (uint64_t)p.ntaps * Cin, BN / 2);
            if (rc) return rc;
            switch (BN) {
                case 256: return launch_cfg_pair<256>(tmA, tmB, p, pair_grid, stream);
                case 128: return launch_cfg_pair<128>(tmA, tmB, p, pair_grid, stream);
                default:  return launch_cfg_pair<64>(tmA, tmB, p, pair_grid, stream);
            }
        }
        switch (BN) {
            case 256: return launch_cfg<256>(tmA, tmB, p, grid, stream);
            case 128: return launch_cfg<128>(tmA, tmB, p, grid, stream);
            default:  return launch_cfg<64>(tmA, tmB, p, grid, stream);
        }
    }
    if (win_pair) {
        G = Gw;
        // CTA pairs: two windows of 128 rows + halves of the weight tiles per CTA
        WinCfg wc;
        wc.G = G;
        wc.b_resident = 0;
        const int need = BLOCK_M + 2 * G + 2;
        if (need <= 256) { wc.nbox = 1; wc.box_rows = (need + 7) & ~7; }
        else             { wc.nbox = 2; wc.box_rows = (((need + 1) / 2) + 7) & ~7; }
        const int win_bytes = wc.box_rows * wc.nbox * 128;
        const int b_half = (BN / 2) * BLOCK_K * 2;
        const int fixed = 1024 + 512 + 10 * BN * 4;
        const int budget = 226 * 1024 - fixed;
        wc.b_stages = 9;
        wc.a_stages = (budget - wc.b_stages * b_half) / win_bytes;
        if (wc.a_stages > WIN_MAX_A_STAGES) wc.a_stages = WIN_MAX_A_STAGES;
        if (wc.a_stages >= 2 && wc.box_rows <= 256) {
            rc = make_tmap_2d_bf16(&tmA, a, (uint64_t)a_rows, (uint64_t)a_cols, (uint64_t)a_ld, wc.box_rows);
            if (rc) return rc;
            rc = make_tmap_2d_bf16(&tmB, wp, b_rows, (uint64_t)p.ntaps * Cin, (uint64_t)p.ntaps * Cin, BN / 2);
            if (rc) return rc;
            const int smem_bytes = fixed + wc.a_stages * win_bytes + wc.b_stages * b_half;
            const long long work = (long long)((p.num_m_tiles + 1) / 2) * p.num_n_tiles;
            const int max_pairs = num_sms() / 2;
            const int wgrid = 2 * (int)((work < max_pairs) ? work : max_pairs);
            // stream-K when whole items leave the last wave mostly empty: e.g. 450 items on 74 pairs are 7 waves of
            // items but 6.25 waves of (item, chunk) steps
            p.sk_enable = 0;
            if (g_streamk && g_sk_scratch != nullptr && g_sk_bytes >= conv_scratch_bytes() && work > max_pairs &&
                work % max_pairs != 0) {
                const long long steps = work * p.kb_per_tap;
                const double waves_items = (double)((work + max_pairs - 1) / max_pairs);
                const double waves_steps = (double)((steps + max_pairs - 1) / max_pairs) / p.kb_per_tap;
                if (waves_steps + 0.15 < waves_items) {
                    p.sk_enable = 1;
                    p.sk_flags = reinterpret_cast<int*>(g_sk_scratch);
                    p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(g_sk_scratch) + 1024);
                }
            }
            g_last_streamk = p.sk_enable;
            return launch_win2<256, 3>(tmA, tmB, p, wc, smem_bytes, wgrid, stream);
        }
    }
    if (window_eligible(p, &G)) {
        WinCfg wc;
        wc.G = G;
        const int SUB = (BN <= 128) ? 2 : 1;
        const int need = SUB * BLOCK_M + 2 * G + 2;
        if (need <= 256) { wc.nbox = 1; wc.box_rows = (need + 7) & ~7; }
        else             { wc.nbox = 2; wc.box_rows = (((need + 1) / 2) + 7) & ~7; }
        const int win_bytes = wc.box_rows * wc.nbox * 128;
        const int b_stage = BN * BLOCK_K * 2;
        const int fixed = 1024 + 512 + 10 * BN * 4;
        const int budget = 226 * 1024 - fixed;
        // weights: resident if all 9 tiles of a single-chunk layer fit next to two windows, else a ring that takes
        // what two windows leave (deep enough to cover the TMA latency); windows take the rest.
        wc.b_resident = (p.kb_per_tap == 1 && p.num_n_tiles == 1 && 9 * b_stage + 2 * win_bytes <= budget) ? 1 : 0;
        const int TB = (BN == 256) ? 1 : (wc.b_resident ? 9 : 3);
        if (wc.b_resident) {
            wc.b_stages = 9;
        } else {
            wc.b_stages = (budget - 2 * win_bytes) / b_stage;
            if (wc.b_stages > 9) wc.b_stages = 9;
            wc.b_stages -= wc.b_stages % TB;
        }
        wc.a_stages = (budget - wc.b_stages * b_stage) / win_bytes;
        if (wc.a_stages > WIN_MAX_A_STAGES) wc.a_stages = WIN_MAX_A_STAGES;
        if (wc.a_stages >= 2 && wc.b_stages >= 3 && wc.b_stages % TB == 0 && wc.box_rows <= 256) {
            rc = make_tmap_2d_bf16(&tmA, a, (uint64_t)a_rows, (uint64_t)a_cols, (uint64_t)a_ld, wc.box_rows);
            if (rc) return rc;
            const int smem_bytes = fixed + wc.a_stages * win_bytes + wc.b_stages * b_stage;
            const long long work = (long long)((p.num_m_tiles + SUB - 1) / SUB) * p.num_n_tiles;
            const int wgrid = (int)((work < num_sms()) ? work : num_sms());
            switch (BN) {
                case 256: return launch_win<256, 1, 1>(tmA, tmB, p, wc, smem_bytes, wgrid, stream);
                case 128: return launch_win<128, 2, 3>(tmA, tmB, p, wc, smem_bytes, wgrid, stream);
                default:
                    if (wc.b_resident) return launch_win<64, 2, 9>(tmA, tmB, p, wc, smem_bytes, wgrid, stream);
                    return launch_win<64, 2, 3>(tmA, tmB, p, wc, smem_bytes, wgrid, stream);
            }
        }
    }
    rc = make_tmap_2d_bf16(&tmA, a, (uint64_t)a_rows, (uint64_t)a_cols, (uint64_t)a_ld, BLOCK_M);
    if (rc) return rc;
    if (pair_ok) {
        rc = make_tmap_2d_bf16(&tmB, wp, b_rows, (uint64_t)p.ntaps * Cin, (uint64_t)p.ntaps * Cin, BN / 2);
        if (rc) return rc;
        switch (BN) {
            case 256: return launch_cfg_pair<256>(tmA, tmB, p, pair_grid, stream);
            case 128: return launch_cfg_pair<128>(tmA, tmB, p, pair_grid, stream);
            default:  return launch_cfg_pair<64>(tmA, tmB, p, pair_grid, stream);
        }
    }
    switch (BN) {
        case 256: return launch_cfg<256>(tmA, tmB, p, grid, stream);
        case 128: return launch_cfg<128>(tmA, tmB, p, grid, stream);
        default:  return launch_cfg<64>(tmA, tmB, p, grid, stream);
    }
}

}  // namespace ffr
