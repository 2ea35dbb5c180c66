// Shifted-row implicit-GEMM on tcgen05 (sm_100a).
//
//   D[m, co] = sum_{tap} sum_{c < Cin}  A[m + row_shift[tap], ch_off[tap] + c] * Wp[co, tap*Cin + c]
//
// A is a row-major bf16 matrix whose rows are pixels of a "halo-shared flat NHWC" activation:
// a 3x3 convolution tap is then nothing but a row offset, so every A tile is a plain 2-D TMA box and the
// zero padding comes from rows that are stored as zeros (or from TMA out-of-bounds fill at the tensor ends).
// Wp is the packed K-major weight matrix [Cout, ntaps*Cin]. Accumulators live in TMEM (double buffered, so the
// epilogue of tile i overlaps the MMAs of tile i+1); one elected lane issues tcgen05.mma, one issues TMA, eight
// warps run the fused epilogue straight out of TMEM.
// Three tilings share this parameter block and epilogue: the sliding-window kernel (3x3 / stride 1, the nine taps are
// descriptor offsets into one TMA window), the tile-per-tap kernel (everything else), and on H9 maps the pixel-major
// variant of the latter (EPI_PIXMAJOR: an M tile is 128 images at one pixel, A tiles are 4-D TMA boxes).
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

namespace ffr {

enum EpiFlags : uint32_t {
    EPI_BIAS        = 1u << 0,   // + bias[co]
    EPI_BORDER_BIAS = 1u << 1,   // + bias9[border_class(h,w)][co]  (pre-conv BatchNorm shift under zero padding)
    EPI_PRELU       = 1u << 2,   // x > 0 ? x : slope[co] * x
    EPI_GEOM        = 1u << 3,   // rows carry (n,h,w); pad rows (h==S or w==S) are invalid
    EPI_POOL        = 1u << 4,   // atomically accumulate per-(image, co) sums of valid rows (SE squeeze)
    EPI_OUT_S2D     = 1u << 5,   // scatter valid rows into the space-to-depth layout of the next stride-2 conv
    EPI_OUT_F32_ATOMIC = 1u << 6,// split-K: atomicAdd fp32 partial sums into out_f32[m, co]
    EPI_SIGMOID     = 1u << 7,   // 1 / (1 + exp(-x)) after everything else
    EPI_SCATTER     = 1u << 8,   // bf16 rows go to up to scatter_n (row, channel-offset) destinations per image-local
                                 // row, looked up in a table (reflection-halo mirrors, W-flip, concatenation slots);
                                 // a row whose first entry is negative is invalid
    EPI_RESIDUAL    = 1u << 9,   // + res[m, co] (bf16, same row grid) after PReLU
    EPI_STATS       = 1u << 10,  // atomically accumulate per-co sum and sum of squares of valid rows (batch-stat BN)
    EPI_OUT_F32     = 1u << 11,  // plain fp32 store to out_f32[m, co] (no atomics)
    EPI_COSFACE     = 1u << 12,  // rows = samples, columns = classes, accumulator = cosine: per row accumulate
                                 // sum_c exp(z_c - s) with z_c = s*(cos_c - m*[c == label]) (|cos| <= 1, so the fixed
                                 // shift s replaces the running maximum), record z_label and the arg-max of cos
    EPI_PIXMAJOR    = 1u << 13,  // H9 maps only (rows_per_img 81, Wp 9): an M tile is 128 IMAGES at one pixel instead of
                                 // 128 consecutive rows, so only the 49 interior pixels are computed (the H9 halo rows
                                 // are 40 % of a row-major tile). A taps are 4-D TMA boxes (channel, w, h, image).
    EPI_PIX_DGRAD   = 1u << 14,  // with EPI_PIXMAJOR: outputs cover all 81 grid points, and a tap is skipped when its
                                 // source pixel is a halo point (the operand is zero there: dz of a reflection-padded conv)
    EPI_MUL_DSIG    = 1u << 15,  // x *= r * (1 - r) with r = res[m, co]: backward of a sigmoid whose OUTPUT r is stored
    EPI_RES_F16     = 1u << 16,  // res holds fp16 (default bf16)
    EPI_OUT_F16     = 1u << 17,  // 16-bit outputs (out / scatter) are fp16 (default bf16)
    EPI_GENERIC_ONLY = 1u << 31, // set by the launcher (ffr_debug_set_lean_epilogue(0)): run the generic epilogue even for a
                                 // flag set that has a compile-time specialisation
};

struct ConvGemmParams {
    // GEMM shape
    int M;              // rows of the output grid
    int Cout;           // total output channels (multiple of the N tile)
    int num_m_tiles, num_n_tiles, num_splits;
    int ntaps;          // 1..9
    int kb_per_tap;     // Cin / 64
    int kb_per_split;   // k-blocks per split (ntaps * kb_per_tap when num_splits == 1)
    int tap_row_shift[9];
    int tap_ch_off[9];
    // row geometry (EPI_GEOM): row m = n * rows_per_img + h * Wp + w ; valid iff h0 <= h < h0+S and w0 <= w < w0+S
    int rows_per_img, Wp, S, h0;
    int n_img;
    // epilogue
    uint32_t flags;
    const float* bias;     // [Cout] or [9][Cout]
    const float* slope;    // [Cout]
    __nv_bfloat16* out;    // bf16 output, row pitch ldo elements
    int ldo;
    int s2d_So;            // EPI_OUT_S2D: output grid is (So+1)x(So+1) rows per image, 4*Cout channels
    float* pool;           // [n_img, Cout] sums, atomic accumulation (pixel-major tiles, or pool_part == nullptr)
    float* pool_part;      // deterministic alternative for row-major tiles: [ceil(M/32)][2][Cout] per-32-row-block sums
                           // (slot 0: the image of the block's first row, slot 1: the next image when the block straddles
                           // two), plain stores; reduced per image in a fixed order by se_gate_kernel
    float* out_f32;        // [M, Cout]
    long long out_f32_split_stride;   // EPI_OUT_F32 with split-K: split s stores its partial product at
                                      // out_f32 + s * stride (plain stores, summed in a fixed order by the consumer)
    const __nv_bfloat16* res;  // residual, row pitch ldres
    int ldres;
    float* stats;          // [2, Cout] : sum, sum of squares (atomic accumulation), used when stats_part == nullptr
    float* stats_part;     // deterministic alternative: [4 * num_m_tiles][2][Cout] per-(M tile, TMEM quadrant) partial
                           // sums, every entry written exactly once (plain stores); reduced by bn_finalize in a fixed order
    const int2* scatter;   // [rows_per_img][scatter_n] : (destination row within the image, channel offset) or (-1, *)
    int scatter_n;         // 1..8
    int out_rows_per_img;  // rows per image of the destination matrix
    int b_rows_per_mtile;  // batched B: weight-matrix row offset added per group of b_mtile_div M tiles (0 = shared weights)
    int b_mtile_div;       // M tiles per B batch (>= 1)
    int a_hilo;            // 1: the A matrix holds [hi | lo] halves (lo at column a_lo_off); every tap / k-chunk is
    int a_lo_off;          //    accumulated twice, hi then lo, against the SAME weight tile (fp16 hi + lo activations)
    int kpt_a;             // A k-blocks per tap = kb_per_tap * (1 + a_hilo)  (filled in by conv_gemm_launch)
    uint32_t idesc_xor;    // XOR-ed into the bf16 instruction descriptor: selects fp16 operands (both A and B)
    // EPI_PIXMAJOR (filled in by conv_gemm_launch)
    int pix_iblocks;                // image blocks of 128 per pixel
    int pix_side, pix_off;          // output pixels: side x side, starting at (off, off) in H9 coordinates
    int pix_src_lo, pix_src_hi;     // a tap is skipped when its source pixel leaves [lo, hi]^2
    int pix_pad_from;               // output pixels with h or w >= this are pad points: no taps, zeros are stored
    // EPI_COSFACE (AddMarginProduct + CrossEntropy, recnet.py:257-270, trainer.py:173-176)
    const int* ce_label;            // [M]
    float* ce_sumexp;               // [M], zeroed by the caller (atomic accumulation), used when ce_sumexp_part == nullptr
    float* ce_sumexp_part;          // deterministic alternative: [M][2 * num_n_tiles] partial sums, plain stores
    float* ce_zlabel;               // [M]
    unsigned long long* ce_argkey;  // [M], zeroed by the caller: (orderable cos bits << 32) | (0xFFFFFFFF - class)
    int ce_classes;                 // real class count (columns >= ce_classes are padding)
    float ce_s, ce_m;
    // stream-K (conv_win2_kernel, filled in by conv_gemm_launch): the (work item, k-chunk) steps are split EVENLY over
    // the CTA pairs instead of whole work items, so the last wave is as full as the others. A pair whose range starts
    // inside a work item stores that partial accumulator (fp32) in its slot of sk_ws and raises its flag; the pair that
    // computed the item's first chunks adds it (fixed order: own + peer) and runs the fused epilogue.
    int sk_enable;
    float* sk_ws;             // [pairs][2 CTAs][128 rows][BN] fp32
    int* sk_flags;            // [pairs][2], zero when idle
    unsigned long long* dbg;  // optional: per-role wait-cycle counters (ffr_debug_set_counters), else nullptr
};

// dbg slots (cycles summed over CTAs): MMA warp waits, producer waits, epilogue waits, totals
enum DbgSlot { DBG_MMA_WAIT_TMEM = 0, DBG_MMA_WAIT_A = 1, DBG_MMA_WAIT_B = 2, DBG_MMA_TOTAL = 3,
               DBG_TMA_WAIT_A = 4, DBG_TMA_WAIT_B = 5, DBG_TMA_TOTAL = 6, DBG_EPI_WAIT = 8, DBG_EPI_TOTAL = 9,
               DBG_CTAS = 10,
               // conv_win2_kernel only, %globaltimer ns: earliest CTA entry (slot must be preset to ~0), earliest start
               // and latest end of an MMA issue loop, latest CTA exit; longest MMA issue loop in cycles
               DBG_T_ENTRY = 11, DBG_T_MMA_BEGIN = 12, DBG_T_MMA_END = 13, DBG_T_EXIT = 14, DBG_MMA_TOTAL_MAX = 15 };

}  // namespace ffr
