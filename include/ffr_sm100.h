/*
 * libffr_sm100 — C ABI of the B200-native (sm_100a) FFR-Net hot path.
 *
 * The reference (haoosz/FFR-Net) has no FFI: its operator surface is the PyTorch nn.Module layer
 * (pretrain/model_ir_se50.py, models/recnet.py, lfw/lfw_eval.py). Each entry point below replaces the library
 * kernels one reference call site dispatches to; the citation names that call site. The Python modules in
 * ffr_net_b200/ (same class names, constructor arguments and state_dict keys as the reference) bind these
 * functions with ctypes — see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (16-byte aligned); the library never allocates or frees
 *     device memory and never synchronises: work is enqueued on `stream`;
 *   - return 0 on success, <0 for an argument/shape/driver-lookup error, >0 for a cudaError_t;
 *     ffr_last_error() returns a thread-local message for the last non-zero return;
 *   - activations are bf16 in the "halo-shared flat NHWC" layout: an SxS map of C channels is a row-major matrix
 *     with n_img*(S+1)^2 rows of C elements, pixel (n,h,w) in row n*(S+1)^2 + h*(S+1) + w, rows with h==S or w==S
 *     being zero padding shared by neighbouring rows/images (DESIGN.md "Data layout");
 *   - packed conv weights are bf16 [Cout][ntaps*Cin], k = (r*3+s)*Cin + ci (BatchNorm scales folded in on the host).
 */
#ifndef FFR_SM100_H_
#define FFR_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ffr_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define FFR_API __attribute__((visibility("default")))
#else
#define FFR_API
#endif

/* Epilogue flags of the implicit-GEMM kernel (values mirror ffr::EpiFlags in csrc/conv_gemm.cuh).
 * Order of application: bias -> PReLU -> residual -> sigmoid -> zero on invalid rows -> outputs / reductions. */
#define FFR_EPI_BIAS           (1u << 0)  /* + bias[co] */
#define FFR_EPI_BORDER_BIAS    (1u << 1)  /* + bias9[border class of (h,w)][co] (BatchNorm before a zero-padded conv) */
#define FFR_EPI_PRELU          (1u << 2)  /* max(x,0) + slope[co]*min(x,0) */
#define FFR_EPI_GEOM           (1u << 3)  /* rows carry (image, h, w): row = n*rows_per_img + h*Wp + w; a row is valid iff
                                             h0 <= h < h0+S and h0 <= w < h0+S (or, with SCATTER, iff it has a destination) */
#define FFR_EPI_POOL           (1u << 4)  /* per-(image,co) sums of the valid rows: atomically into pool[n_img][Cout] */
#define FFR_EPI_OUT_S2D        (1u << 5)  /* bf16 rows go to the space-to-depth map of the following stride-2 conv */
#define FFR_EPI_OUT_F32_ATOMIC (1u << 6)  /* split-K: atomicAdd fp32 into out_f32[M][Cout] */
#define FFR_EPI_SIGMOID        (1u << 7)
#define FFR_EPI_SCATTER        (1u << 8)  /* bf16 rows go to up to scatter_n table destinations (reflection mirrors,
                                             concat slots, W-flip); see ffr_conv_gemm */
#define FFR_EPI_RESIDUAL       (1u << 9)  /* + res[m][co] (bf16, same row grid) */
#define FFR_EPI_STATS          (1u << 10) /* atomically add per-co sum and sum of squares of valid rows into stats[2][Cout] */
#define FFR_EPI_OUT_F32        (1u << 11) /* plain fp32 rows to out_f32[M][Cout] */
#define FFR_EPI_COSFACE        (1u << 12) /* internal to ffr_cosface_ce_fwd (softmax denominator / label logit / arg-max) */
#define FFR_EPI_PIXMAJOR       (1u << 13) /* H9 maps (rows_per_img 81, Wp 9, 9 taps or 1): an M tile is 128 IMAGES at one
                                             pixel, so only the 49 interior pixels are computed (n_img must be given);
                                             rows of the result are addressed exactly as in the row-major mode */
#define FFR_EPI_PIX_DGRAD      (1u << 14) /* with PIXMAJOR: outputs on all 81 grid points, taps whose source pixel is a halo
                                             point are skipped (operand zero there: dz of a reflection-padded conv) */

FFR_API int ffr_version(void);
FFR_API const char* ffr_last_error(void);
/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
FFR_API long long ffr_launch_count(void);

/* Generic shifted-row implicit GEMM on tcgen05/TMEM/TMA (ffr_net_b200/csrc/conv_gemm.cu):
 *   D[m,co] = sum_t sum_c A[m + tap_row_shift[t], tap_ch_off[t] + c] * Wp[co, t*Cin + c], then the fused epilogue.
 * Replaces every nn.Conv2d / nn.Linear GEMM of the path (model_ir_se50.py:63-69,118,124; recnet.py:65,373-385). */
FFR_API int ffr_conv_gemm(const void* a, int64_t a_rows, int a_cols, int a_ld, const void* wp, int Cin, int Cout, int ntaps,
                  const int* tap_row_shift, const int* tap_ch_off, int M, int rows_per_img, int Wp, int S, int h0,
                  int n_img, uint32_t flags, const float* bias, const float* slope, void* out, int ldo, int s2d_So,
                  float* pool, float* out_f32, const void* res, int ldres, float* stats, int num_splits,
                  const int* scatter, int scatter_n, int out_rows_per_img, int b_rows_per_mtile, ffr_stream_t stream);
/*   scatter: device table [rows_per_img][scatter_n][2] of (destination row within the image, channel offset), -1 = none
 *   (FFR_EPI_SCATTER); b_rows_per_mtile != 0 makes the weight operand batched: M tile i reads weight rows starting at
 *   i*b_rows_per_mtile (per-sample matrices such as M_channel). */

/* Extended form of ffr_conv_gemm (same kernel): operands may be fp16, the A matrix may carry fp16 hi + lo halves
 * (training forward: x = hi + lo keeps ~21 mantissa bits at twice the MMA count), BatchNorm statistics can be written
 * as deterministic per-tile partial sums, the weight operand can be batched per group of M tiles, and the epilogue can
 * apply the backward of a stored sigmoid. Fields not listed here mean what the ffr_conv_gemm arguments mean. */
#define FFR_EPI_MUL_DSIG       (1u << 15) /* x *= r*(1-r), r = res[m][co] (sigmoid backward from its stored output) */
#define FFR_EPI_RES_F16        (1u << 16) /* res is fp16 (default bf16) */
#define FFR_EPI_OUT_F16        (1u << 17) /* 16-bit outputs are fp16 (default bf16) */
typedef struct ffr_conv_gemm_desc {
    const void* a; int64_t a_rows; int a_cols; int a_ld;
    const void* wp; int Cin; int Cout; int ntaps;
    const int* tap_row_shift; const int* tap_ch_off;      /* host arrays [ntaps] or NULL (zeros) */
    int M, rows_per_img, Wp, S, h0, n_img;
    uint32_t flags;
    const float* bias; const float* slope;
    void* out; int ldo; int s2d_So;
    float* pool; float* out_f32;
    const void* res; int ldres;
    float* stats;                 /* [2][Cout], atomically accumulated (caller zeroes) when stats_part is NULL */
    float* stats_part;            /* [4*ceil(M/128) (row-major) or 4*49*ceil(n_img/128) (pixel-major)][2][Cout] partial sums,
                                     row = (M tile)*4 + 32-row quadrant; written, not accumulated */
    int num_splits;
    const int* scatter; int scatter_n; int out_rows_per_img;
    int b_rows_per_mtile; int b_mtile_div;                /* B rows advance by b_rows_per_mtile every b_mtile_div M tiles */
    int a_hilo; int a_lo_off;                             /* A = [hi | lo], lo at column a_lo_off */
    int f16;                                              /* 1: A and B hold fp16 (tcgen05 kind::f16 needs equal formats) */
} ffr_conv_gemm_desc;
FFR_API int ffr_conv_gemm_ex(const ffr_conv_gemm_desc* d, ffr_stream_t stream);

/* Conv2d(Cin,Cout,3,stride 1,pad 1) on a flat SxS map (model_ir_se50.py:67 with the BatchNorm of :66 folded:
 * scale into wp, shift into the 9-class border bias table `bias9` [9][Cout]) + PReLU (:68).
 * out_s2d != 0 writes the space-to-depth layout consumed by ffr_conv3x3_bn_pool_fwd(stride = 2). */
FFR_API int ffr_conv3x3_bnpre_prelu_fwd(const void* x, int n_img, int S, int Cin, const void* wp, int Cout,
                                const float* bias9, const float* slope, void* out, int out_s2d, ffr_stream_t stream);

/* Conv2d(C,Cout,3,stride,pad 1) + BatchNorm (model_ir_se50.py:69-70; scale folded into wp, shift = bias) and the
 * SE squeeze (:31): the epilogue stores the column sums of every 32-row block of the output map into pool_part
 * (ffr_se_pool_part_floats(n_img, S/stride, Cout) floats, every used entry written exactly once - no zeroing, no
 * atomics); ffr_se_gate_fwd adds them per image. pool_part may be NULL (no squeeze).
 * stride 1: x is flat SxS. stride 2: x is the space-to-depth map written by ffr_conv3x3_bnpre_prelu_fwd
 * (rows of the (S/2+1)^2 grid, 4*C channels); S is the INPUT size. Output is flat (S/stride)x(S/stride). */
FFR_API long long ffr_se_pool_part_floats(int n_img, int So, int Cout);
FFR_API int ffr_conv3x3_bn_pool_fwd(const void* x, int n_img, int S, int C, int stride, const void* wp, int Cout,
                            const float* bias, void* out, float* pool_part, ffr_stream_t stream);

/* Conv2d(Cin,Cout,1,stride 2)+BatchNorm shortcut (model_ir_se50.py:62-64) on an already subsampled flat map. */
FFR_API int ffr_conv1x1_bn_fwd(const void* xs, int n_img, int S, int Cin, const void* wp, int Cout, const float* bias,
                       void* out, ffr_stream_t stream);

/* out(n,h,w) = x(n,2h,2w) on flat maps; So = output size (MaxPool2d(1,2) / stride of the 1x1 conv, :60,63). */
FFR_API int ffr_subsample2(const void* x, void* out, int n_img, int So, int C, ffr_stream_t stream);

/* input_layer: Conv2d(3,64,3,1,1)+BatchNorm+PReLU on fp32 NCHW images (model_ir_se50.py:118-120).
 * w [27][64] fp32 (k = ci*9+r*3+s, BN scale folded), b [64] BN shift, a [64] PReLU slopes; out flat bf16 SxS x 64. */
FFR_API int ffr_stem_fwd(const float* x, const float* w, const float* b, const float* a, void* out, int n_img, int S,
                 ffr_stream_t stream);

/* The same layer fed by decoded images: img uint8 HWC (n,S,S,3). Applies the reference's host-side preprocessing on the
 * fly — channel swap (data/dataset.py:138-141, swap_rb != 0: source channel 2-c), horizontal flip for images with
 * flip[i] != 0 (dataset.py:149-152; flip may be NULL), ToTensor + Normalize(0.5,0.5) (data/dataloader.py:15-19):
 * (u/255 - 0.5)/0.5 with IEEE division, bit-equal to torchvision. 4x less host->device traffic than fp32 NCHW. */
FFR_API int ffr_stem_u8_fwd(const unsigned char* img, const unsigned char* flip, int swap_rb, const float* w,
                            const float* b, const float* a, void* out, int n_img, int S, ffr_stream_t stream);

/* SEModule (model_ir_se50.py:29-36,73-76): y = u*sigmoid(W2 relu(W1 mean(u))) + shortcut, in two launches.
 * ffr_se_gate_fwd: gate [n_img][C] fp32 from the partial sums ffr_conv3x3_bn_pool_fwd stored (added per image in a
 * fixed order: bit-reproducible); sums (optional) receives the per-(image,channel) sums of u. w1 [C/16][C], w2 [C][C/16].
 * ffr_se_residual_fwd: one pass over the map; shortcut_mode 0: same-grid x, 1: x on the 2Sx2S grid (MaxPool2d(1,2)),
 * 2: same-grid conv shortcut. */
FFR_API int ffr_se_gate_fwd(const float* pool_part, const float* w1, const float* w2, float* gate, float* sums, int n_img,
                            int S, int C, ffr_stream_t stream);
/* ffr_se_gate_residual_fwd: the two calls above in ONE launch (one CTA per image computes the image's gate from the
 * partial sums, then streams the image's rows); same arithmetic, bit-identical output. Backbone.forward uses it for the
 * units with C <= 256 (model_ir_se50.py:29-36, 73-76). */
FFR_API int ffr_se_gate_residual_fwd(const void* u, const float* pool_part, const float* w1, const float* w2,
                                     const void* shortcut, int shortcut_mode, void* y, int n_img, int S, int C,
                                     ffr_stream_t stream);
FFR_API int ffr_se_residual_fwd(const void* u, const float* gate, const void* shortcut, int shortcut_mode, void* y,
                                int n_img, int S, int C, ffr_stream_t stream);

/* y = bn(h) exported as fp32 NCHW (model_ir_se50.py:126,139). */
FFR_API int ffr_export_nchw_fwd(const void* h, const float* scale, const float* shift, float* y, int n_img, int S, int C,
                        ffr_stream_t stream);

/* output_layer (BatchNorm2d -> Dropout(eval) -> Flatten -> Linear(25088,512) -> BatchNorm1d) + l2_norm
 * (model_ir_se50.py:121-125,13-16,141). wp: folded bf16 weights [512][(S+1)^2*C] over the flat rows of one image
 * (zero columns at pad pixels), bias [512] folded; acc: fp32 scratch of ffr_head_workspace_floats() elements (one partial
 * product per K split, added in split order: bit-reproducible); f [n_img][512] fp32. */
FFR_API long long ffr_head_workspace_floats(int n_img, int S, int C);
FFR_API int ffr_head_fwd(const void* h, int n_img, int S, int C, const void* wp, const float* bias, float* acc, float* f,
                 ffr_stream_t stream);

/* ---- RecNet (models/recnet.py) ------------------------------------------------------------------------------
 * H9 layout: a 7x7 map with its reflection halo materialised, bf16 [n*81][C], pixel (h,w) at row (h+1)*9+(w+1). */

/* selfSimilarity (recnet.py:220-236), the cat() fan-out of X (:401-402,420) and the thin part of Conv4Channel
 * (:372-384, up to the input of its last Linear) for one batch. x fp32 NCHW (n,512,7,7).
 * Outputs: s0 H9 [n*81][576] (X | ss_space | zero pad), cm H9 [n*81][1536] slot [1024,1536) <- X,
 * xt [n*128][512] (X^T, 49 valid rows), h5 [n*512][64] (32 valid columns), ss_space optional fp32 [n][49][49]. */
FFR_API int ffr_recnet_prep(const float* x, int n, const float* w0aT, const float* w0bT, const float* b0,
                            const float* slope1, const float* A1, const float* c1, const float* slope4,
                            const float* A2, const float* c2, const float* slope7, void* s0, void* cm, void* xt,
                            void* h5, float* ss_space, ffr_stream_t stream);

/* ConvLayer.forward in eval mode (recnet.py:78-85): ReflectionPad2d(1) -> Conv2d 3x3 -> BatchNorm (folded: scale in
 * wp, shift = bias) -> PReLU, optionally + residual (ResidualBlock, :213-218) and sigmoid (:370), on H9 rows.
 * bf16 result rows are scattered through `scatter` (self + reflection mirrors, concat slot, W-flip); out_f32
 * (optional) receives plain fp32 rows [n*81][Cout]; pool (optional) per-(sample,channel) sums of valid rows. */
FFR_API int ffr_recnet_convlayer_fwd(const void* x, int n, int Cin, const void* wp, int Cout, const float* bias,
                                     const float* slope, const void* res, int ldres, int sigmoid, void* out, int ldo,
                                     const int* scatter, int scatter_n, int out_rows_per_img, float* out_f32,
                                     float* pool, ffr_stream_t stream);

/* selfSimilarity forward (recnet.py:226-236) in fp32: x (n,512,7,7) -> ss_space (n,49,49) [viewed (n,49,7,7) by the
 * caller] and/or ss_channel (n,512,512); either output may be NULL. */
FFR_API int ffr_self_similarity(const float* x, int n, float* ss_space, float* ss_channel, ffr_stream_t stream);

/* feat_space = X @ M_space (recnet.py:409) -> slot [0,512) of cm (H9 + mirrors); optional fp32 NCHW copy. */
FFR_API int ffr_feat_space(const float* x, const float* mspace, void* cm, float* out_nchw, int n, ffr_stream_t stream);
/* The same product (recnet.py:409) on warp-level tensor cores, fed by the bf16 X^T matrix ffr_recnet_prep wrote
 * (xt: [n*128][512], rows = pixels) instead of the fp32 map; M_space is rounded to bf16 for the contraction. */
FFR_API int ffr_feat_space_xt(const void* xt, const float* mspace, void* cm, float* out_nchw, int n, ffr_stream_t stream);

/* Rows of a haloed/flat grid -> fp32 NCHW (S x S valid pixels at row (h+off)*G + (w+off)), optional affine. */
FFR_API int ffr_rows_to_nchw(const void* rows, int is_f32, int ld, int ch0, const float* scale, const float* shift,
                             float* y, int n, int S, int G, int off, int rows_per_img, int C, ffr_stream_t stream);

/* out = in * scale (AvgPool2d(7) finish, recnet.py:423). */
FFR_API int ffr_scale_f32(const float* in, float* out, int64_t count, float scale, ffr_stream_t stream);

/* Packs one 3x3 conv weight (fp32 OIHW) for the forward GEMM and (optionally) for its dgrad in a single launch:
 * fwd [cout_p][9*cin_p], dgrad [cin_p][9*cout_p] (spatially flipped + transposed), zero padded. */
FFR_API int ffr_pack_conv3x3(const float* w, int cout, int cin, int cout_p, int cin_p, void* fwd, void* dgrad,
                             ffr_stream_t stream);

/* clip_grad_value_(clip) + Adam.step() for a whole parameter set in one launch (models/trainer.py:185-187,
 * torch.optim.Adam semantics). table: device array of {float* p, g, m, v; int64 n} per tensor; chunks: device array of
 * (tensor index, chunk index) pairs, one per 4096 elements; hyper: DEVICE float[2] = {learning rate, 1-based step count}
 * (device-resident so a captured CUDA graph of the training step survives LR-schedule and step changes). */
FFR_API int ffr_clip_adam(const void* table, const int* chunks, int n_chunks, const float* hyper, float beta1,
                          float beta2, float eps, float weight_decay, float clip, ffr_stream_t stream);

/* ---- RecNet training, batched over the two calls of an iteration (models/trainer.py:144-145) ----------------------
 * G = n_img / n_per_group "groups" (RecNet calls) run in one launch; BatchNorm statistics stay per group. Activations are
 * fp16 hi + lo halves ([rows][2*C]: hi at column c, lo at lo_off + c; lo_off 0 = hi only) for the forward GEMMs plus a
 * bf16 copy for the weight-gradient GEMMs; raw conv outputs z and all activation gradients are fp32; dz is bf16.
 * Every reduction is a fixed-order two-stage sum (no floating-point atomics): results are run-to-run deterministic. */

/* BatchNorm2d batch statistics (recnet.py:83 in training mode) from the per-tile partial sums the conv epilogue wrote
 * (ffr_conv_gemm_desc.stats_part, part_rows rows): mean_rstd [G][2][C] <- (mean, 1/sqrt(var + eps)) per group, and the
 * running statistics (momentum update, unbiased variance) applied group after group; num_batches_tracked += G. */
FFR_API int ffr_bn_finalize(const float* part, int part_rows, int pixmajor, int n_img, int n_per_group, int C, int C_real,
                            float momentum, float eps, float* running_mean, float* running_var,
                            long long* num_batches_tracked, float* mean_rstd, ffr_stream_t stream);

/* a = prelu(gamma (z - mean_g) rstd_g + beta) (+ res) (recnet.py:83-84, :217) from the fp32 conv output z [n_img*81][ldz]:
 * out_h / out_b receive a at every destination of the H9 scatter table (own row + reflection mirrors, recnet.py:64) as
 * fp16 hi/lo resp. bf16; out_f (optional) the fp32 value on the own row (sigmoid != 0: sigmoid(a), recnet.py:370). */
FFR_API int ffr_bn_act_fwd(const float* z, int ldz, const float* mean_rstd, const float* gamma, const float* beta,
                           const float* slope, const void* res, int ldres, int res_lo_off, void* out_h, int ldo,
                           int lo_off, void* out_b, int ldb, float* out_f, int ldf, int sigmoid, const int* scatter,
                           int scatter_n, int n_img, int n_per_group, int C, int C_real, ffr_stream_t stream);

/* Backward of the above. Output-gradient sources (fp32, any subset): da on the H9 grid (folded over the scatter table),
 * dadd on own rows, dv [n_img][lddv] * dv_scale broadcast over the pixels (AvgPool2d(7), recnet.py:423).
 * afold [n_img*81][ldaf] receives the folded gradient (own rows; it is also the residual-branch gradient of a
 * ResidualBlock); dgamma / dbeta / dslope the parameter gradients (overwritten, or += when accumulate);
 * dz [n_img*81][lddz] bf16 the gradient of the conv output (zeros on halo rows).
 * partial: fp32 workspace [ffr_bn_act_bwd_partial_rows(G, C)][3][C]; gsum: [G][2][C]. */
FFR_API int ffr_bn_act_bwd_partial_rows(int n_groups, int C);
FFR_API int ffr_bn_act_bwd(const float* da, int ldda, int da_ch0, const int* scatter, int scatter_n, const float* dadd,
                           int ldadd, int dadd_ch0, const float* dv, int lddv, float dv_scale, const float* z, int ldz,
                           const float* mean_rstd, const float* gamma, const float* beta, const float* slope,
                           float* afold, int ldaf, float* partial, float* gsum, float* dgamma, float* dbeta,
                           float* dslope, int accumulate, int C_real, void* dz, int lddz, int n_img, int n_per_group,
                           int C, ffr_stream_t stream);

/* fp32 NCHW (n,C,7,7) -> own rows of an fp32 H9 matrix [n*81][ld], channel slot ch0 (layout of the `dadd` source). */
FFR_API int ffr_nchw_to_h9_f32(const float* x, float* out, int ld, int ch0, int n, int C, ffr_stream_t stream);

/* v[n][c] = mean over the 49 valid rows of an fp32 H9 matrix (AvgPool2d(7), recnet.py:423). */
FFR_API int ffr_h9_avgpool(const float* a, int lda, float* v, int ldv, int n_img, int C, ffr_stream_t stream);
/* the same over a bf16 H9 matrix (the eval path pools its stored feature map: fixed order, bit-reproducible) */
FFR_API int ffr_h9_avgpool_bf16(const void* a, int lda, float* v, int ldv, int n_img, int C, ffr_stream_t stream);

/* Weight gradient, general form of ffr_wgrad3x3: dw[co][ci][t] (+)= sum_p dz[p][co] x[p + shift_t][x_ch0 + ci] over P rows;
 * ntaps 9 (3x3 on the H9 grid, P = n*81) or 1 (plain Y^T X, e.g. Linear weights); operands both bf16 (f16 = 0) or both
 * fp16; deterministic != 0: one staging slab per split of the row axis, added in a fixed order; bias_col >= 0: that
 * column of x is a column of ones and its result goes to db[co] instead of dw. ld_w = row pitch of dw in ci units.
 * workspace: ffr_wgrad_workspace_floats(...) fp32 elements. */
FFR_API int64_t ffr_wgrad_workspace_floats(int P, int Cout, int Cin, int ntaps, int deterministic);
FFR_API int ffr_wgrad(const void* dz, int ld_dz, const void* x, int ld_x, int x_ch0, int P, int Cout, int Cin, int ntaps,
                      int f16, int deterministic, int accumulate, int ld_w, int bias_col, float* dw, float* db,
                      float* workspace, ffr_stream_t stream);

/* Packs a 3x3 conv weight like ffr_pack_conv3x3 with the forward operand in fp16 (dgrad operand stays bf16). */
FFR_API int ffr_pack_conv3x3_f16(const float* w, int cout, int cin, int cout_p, int cin_p, void* fwd_f16, void* dgrad_bf16,
                                 ffr_stream_t stream);

/* RecNet.forward up to the convolution stacks for a training batch (recnet.py:399-402, :406, Conv4Channel :372-386):
 * selfSimilarity, the cat() fan-outs and the whole channel rectifier in fp32. M_channel = sigmoid(.) is written as fp16
 * [hi | lo] (mch2) and X^T as [hi | lo | hi] (x3): feat_channel = M_channel @ X (:410) is then ONE call of ffr_conv_gemm_ex
 * (f16; three "taps" of K = 512 with A column offsets 0, 0, 512 = hi.hi + hi.lo + lo.hi; batched weight operand x3; fp32
 * output [n*512][64]) followed by ffr_fc_scatter (flip / cat fan-out of :416-417 into the ChannelFlipMerge input). Saves
 * what the backward needs (pre-PReLU activations g0/g1/g2, h7 | 1, X rows, 1/|X_c|; m (1 - m) comes from the hi part of
 * mch2). */
typedef struct ffr_prep_train_desc {
    const float* x;                                  /* (n,512,7,7) fp32 NCHW */
    const float* w0; const float* b0;                /* Conv4Channel.0 [32][561], [32] */
    const float* slope1; const float* slope4; const float* slope7;    /* Conv4Channel.{1,4,7}.func.weight [512] */
    const float* A1; const float* c1; const float* A2; const float* c2;   /* ffr_chan_compose outputs */
    const float* w8; const float* b8;                /* Conv4Channel.8 [512][32], [512] */
    void* s0_h; int s0_ld; int s0_lo; void* s0_b; int s0_ldb;     /* Conv4Space input [n*81][..], 576 channels */
    void* cm_h; int cm_ld; int cm_lo; void* cm_b; int cm_ldb;     /* Conv4Merge input, 1536 channels (slot 1024.. <- X) */
    float* g0; float* g1; float* g2;                 /* [n*512][32] each */
    void* h7b; void* xk;                             /* bf16 [n*512][64] each */
    void* mch2; void* x3;                            /* fp16 [n*512][1024], [n*64][1536] */
    float* inv_c; float* tmat; float* ss_space;      /* [n*512], [n][49][32], optional [n][49][49] */
} ffr_prep_train_desc;
FFR_API int ffr_recnet_prep_train(const ffr_prep_train_desc* d, int n, ffr_stream_t stream);
/* fcraw fp32 [n*512][64] (row = channel, column = pixel) -> ChannelFlipMerge input (hi/lo + bf16 copy, 1024 channels):
 * slot [512,1024) <- feat_channel, slot [0,512) <- flip_W(feat_channel), each with its reflection mirrors. */
FFR_API int ffr_fc_scatter(const float* fcraw, void* fm_h, int fm_ld, int fm_lo, void* fm_b, int fm_ldb, int n,
                           ffr_stream_t stream);

/* Linear(32->512) directly followed by Linear(512->32) (recnet.py:375-376, :378-379) composed into 32x32 maps:
 * A1 = W3 W2, c1 = W3 b2 + b3, A2 = W6 W5, c2 = W6 b5 + b6; and the backward of the composition. */
FFR_API int ffr_chan_compose(const float* w2, const float* b2, const float* w3, const float* b3, const float* w5,
                             const float* b5, const float* w6, const float* b6, float* A1, float* c1, float* A2,
                             float* c2, const float* w8, void* w8t, ffr_stream_t stream);
/*   w8t (optional): bf16 [64][512] <- W8^T (Conv4Channel.8.weight [512][32]), rows 32..63 zero: K-major operand of the
 *   backward GEMM dh7 = dM_pre @ W8. */
FFR_API int ffr_chan_compose_bwd(const float* w2, const float* b2, const float* w3, const float* w5, const float* b5,
                                 const float* w6, const float* dA1, const float* dc1, const float* dA2, const float* dc2,
                                 float* dw2, float* db2, float* dw3, float* db3, float* dw5, float* db5, float* dw6,
                                 float* db6, int accumulate, ffr_stream_t stream);

/* feat_space = X @ M_space (recnet.py:409) into slot [0,512) of the Conv4Merge input (+ fp32 copy on own rows), and the
 * backward w.r.t. M_space through its sigmoid: dmsp [n*81][64] (own row of pixel j, column i). */
FFR_API int ffr_feat_space_train(const float* x, const float* mspace, void* cm_h, int cm_ld, int cm_lo, void* cm_b,
                                 int cm_ldb, float* fs_f32, int ldfs, int n, ffr_stream_t stream);
FFR_API int ffr_feat_space_bwd(const float* x, const float* mspace, const float* dcm, int lddcm, const float* dfs,
                               int lddfs, float* dmsp, int n, ffr_stream_t stream);

/* Gradient of the flip / cat / reflection fan-out of feat_channel (recnet.py:416-417): dfm fp32 [n*81][lddfm] on the H9
 * grid -> dfc bf16 [n*512][64] (row = channel, 49 valid columns). */
FFR_API int ffr_fc_bwd_gather(const float* dfm, int lddfm, void* dfc, int n, ffr_stream_t stream);

/* Backward of the thin Conv4Channel chain (recnet.py:373-384) given dh7 [n*512][64] fp32: gradients of
 * Conv4Channel.0 (dW0 [32][561], db0), the three PReLUs, and of the composed maps (tmp [2112] = dA2, dc2, dA1, dc1 for
 * ffr_chan_compose_bwd). part: workspace [n][ffr_chan_bwd_part_floats()], dslope_part [n][3][512]. */
FFR_API int ffr_chan_bwd_part_floats(void);
FFR_API int ffr_chan_bwd(const float* x, const float* dh7, const float* g0, const float* g1, const float* g2,
                         const float* inv_c, const float* tmat, const float* A1, const float* A2, const float* slope1,
                         const float* slope4, const float* slope7, float* part, float* dslope_part, float* tmp,
                         float* db0, float* dW0, float* dslope1, float* dslope4, float* dslope7, int accumulate, int n,
                         ffr_stream_t stream);

/* ---- losses of Trainer.backward (models/trainer.py:31-43, :154-178), forward + gradient ------------------------- */

/* Channel self-similarity (recnet.py:232 + trainer.py:158-165): operands of D = F^F^T - X^X^T as one K = 384 GEMM
 * (bf16 hi/lo splits, see csrc/loss_kernels.cu), F^T for the gradient GEMM and 1/|F_c|. f: fp32 H9 [n_img*81][ldf]. */
FFR_API int ffr_selfsim_channel_pack(const float* f, int ldf, const float* x, int n_img, int n_per_group, void* A6,
                                     void* B6, void* FhT, float* inv_f, ffr_stream_t stream);
/* e = D F^ [n_img*512][64] -> df (fp32 H9 own rows): normalisation Jacobian, coef = 4 w / (n * 512 * 512). */
FFR_API int ffr_selfsim_channel_bwd(const float* e, const float* f, int ldf, const float* inv_f, float coef, float* df,
                                    int lddf, int n_img, ffr_stream_t stream);
/* Partial sums of the sum-of-squares half of a stats_part buffer: out[g*64 + k]. */
FFR_API int ffr_sumsq_reduce(const float* part, int rows_per_group, int C, int groups, float* out, ffr_stream_t stream);
/* Spatial self-similarity (recnet.py:231 + trainer.py:158-164): loss_part[s] = sum (G - T)^2, dfs = gradient (optional). */
FFR_API int ffr_selfsim_space_loss(const float* fs, int ldfs, const float* x, int n_img, int n_per_group, float coef,
                                   float* loss_part, float* dfs, int lddfs, ffr_stream_t stream);
/* TripletLoss (trainer.py:38-43, :167-169) and the identity MSE (:171): row_part [n][5], gradients w.r.t. the pooled
 * RecNet features of the two calls. */
FFR_API int ffr_triplet_identity(const float* f_non, const float* f_ocl, const float* e_non, const float* e_ocl, int n,
                                 float w_trip, float w_id, float margin, float* row_part, float* df_non, float* df_ocl,
                                 ffr_stream_t stream);
/* out[0..3] weighted loss items (trainer.py:178), out[4] / out[5] mean pos / neg distance, out[6] total. */
FFR_API int ffr_loss_finalize(const float* space_part, const float* chan_sums, const float* row_part, const float* ce, int n,
                              int groups, float w0, float w1, float w2, float w3, float* out, ffr_stream_t stream);
FFR_API int ffr_add3_f32(const float* a, const float* b, const float* c, float* out, int64_t count, ffr_stream_t stream);

/* ---- 1:N gallery scoring: the paired scoring of lfw/lfw_eval.py:246-259 generalised to a similarity matrix ---- */

/* cos_out[p][g_pad] (fp32) = probe . gallery^T for operands packed by ffr_cosface_pack (probe: mode 0, gallery: mode 1;
 * L2-normalised, bf16 hi/lo split, K = 1536, cosine error ~1e-5). g_pad multiple of 256; columns >= G are padding.
 * argkey (may be NULL): per probe the arg-max gallery column as (orderable cosine bits << 32) | (0xFFFFFFFF - g), first
 * maximum wins — the rank-1 identification. */
FFR_API int ffr_gallery_cosine(const void* probe_packed, int P, const void* gallery_packed, int g_pad, int G,
                               float* cos_out, unsigned long long* argkey, ffr_stream_t stream);

/* Histograms for ROC / TAR@FAR: hist[2][T+1] (uint64; genuine = same identity, impostor), bin b of a score = number of
 * grid thresholds t with (double)score > t (the strict comparison of eval_acc, lfw_eval.py:141-147; thresholds
 * ascending, T <= 1024). Accepted-at-threshold counts are suffix sums of the bins. scores [P][ld] fp32, ids int32. */
FFR_API int ffr_roc_hist(const float* scores, int ld, int P, int G, const int* probe_id, const int* gallery_id,
                         const double* thresholds, int T, unsigned long long* hist, ffr_stream_t stream);

/* ---- CosFace head + CrossEntropy, training path (AddMarginProduct recnet.py:238-270, trainer.py:173-176) ------ */

/* Rows of x (fp32 [rows][512]) -> F.normalize'd (eps 1e-12) bf16 hi/lo split, packed [rows_pad][1536]:
 * mode 0 (samples) = [hi | lo | hi], mode 1 (classes) = [hi | hi | lo]; rows in [rows, rows_pad) are zero.
 * transposed (may be NULL): [512][t_ld] bf16, hi part, column index = row (K-major operand of the backward GEMMs). */
FFR_API int ffr_cosface_pack(const float* x, int rows, int rows_pad, int mode, void* packed, void* transposed, int t_ld,
                             ffr_stream_t stream);

/* cos[n][c_pad] (fp32) = v_packed . w_packed^T on the tcgen05 GEMM (K = 1536: hi.hi + lo.hi + hi.lo), with the fused
 * epilogue: sumexp[i] = sum_{c < classes} exp(z_ic - s), z_ic = s*(cos_ic - m*[c == label_i]); zlabel[i] = z at the
 * label; argkey[i] = arg-max_c cos_ic packed as (orderable bits << 32) | (0xFFFFFFFF - c) (first maximum wins, as
 * torch.argmax). c_pad must be a multiple of 256; label is int32. sumexp and argkey are zeroed here. */
FFR_API int ffr_cosface_ce_fwd(const void* v_packed, int n, const void* w_packed, int c_pad, int classes, const int* label,
                               float s, float m, float* cos_out, float* sumexp, float* zlabel,
                               unsigned long long* argkey, float* sumexp_part, ffr_stream_t stream);
/*   sumexp_part (optional): fp32 workspace [n][c_pad / 128]; when given, the softmax denominators are accumulated as
 *   per-tile partial sums added in a fixed order (bit-reproducible) instead of with fp32 atomics. */

/* loss = mean_i(log(sumexp[i]) + s - zlabel[i]) (device scalar); pred[i] = arg-max class (int64, may be NULL). */
FFR_API int ffr_cosface_ce_finish(const float* sumexp, const float* zlabel, const unsigned long long* argkey, int n, float s,
                                  float* loss, long long* pred, ffr_stream_t stream);

/* dcos[i][c] = s * (softmax(z_i)[c] - [c == label_i]) * gloss[0] / n as bf16 [n][c_pad] and transposed [c_pad][n_pad]
 * (pads zero). gloss: DEVICE scalar, the upstream gradient of the mean CE loss. The two backward contractions
 * (dv^ = dcos . W^, K = c_pad;  dW^ = dcos^T . v^, K = n_pad) are plain ffr_conv_gemm calls on these operands. */
FFR_API int ffr_cosface_ce_bwd(const float* cos_in, int c_pad, int classes, int n, int n_pad, const int* label,
                               const float* sumexp, const float* gloss, float s, float m, void* dcos, void* dcosT,
                               ffr_stream_t stream);

/* The same for several batches of n_per_group rows each (the two RecNet calls of an iteration share one GEMM): row r uses
 * the upstream gradient gloss[r / n_per_group] and the mean over n_per_group rows. */
FFR_API int ffr_cosface_ce_bwd_grouped(const float* cos_in, int c_pad, int classes, int n, int n_pad, const int* label,
                                       const float* sumexp, const float* gloss, int n_per_group, float s, float m, void* dcos,
                                       void* dcosT, ffr_stream_t stream);

/* Jacobian of F.normalize(x, dim=1) on rows of 512: dx = (dxh - xh (xh . dxh)) / max(|x|, 1e-12). */
FFR_API int ffr_normalize_bwd(const float* x, const float* dxh, int rows, float* dx, ffr_stream_t stream);

/* ---- LFW-style verification scoring (lfw/lfw_eval.py) ------------------------------------------------------ */

/* score[i] = sum(f1[i]*f2[i]) / (|f1[i]|*|f2[i]| + 1e-8)  (lfw_eval.py:246,248); f1, f2 fp32 [pairs][D]. */
FFR_API int ffr_pair_cosine(const float* f1, const float* f2, float* score, int pairs, int D, ffr_stream_t stream);

/* K-fold threshold sweep (lfw_eval.py:110-118 KFold, :137-153 eval_acc, :155-162 find_best_threshold,
 * :255-259 get_fold_accuracy): contiguous folds; "same" iff (double)score > thresholds[t]; per fold the LAST threshold
 * with maximal training accuracy and the held-out accuracy at it. thresholds: fp64 [T] (upload np.arange(-1,1,.005)
 * bit-for-bit); outputs per fold: best_idx, best_thr (fp64), test_correct, train_correct (integer counts). */
FFR_API int ffr_threshold_sweep(const float* score, const int* label, const double* thresholds, int n, int T, int folds,
                                int* best_idx, double* best_thr, int* test_correct, int* train_correct,
                                ffr_stream_t stream);

/* Debug/tuning: 1 (default) lets 3x3 stride-1 convolutions use the sliding-window kernel, 0 forces the
 * tile-per-tap kernel (the two must agree; tests run both). */
FFR_API int ffr_debug_set_window(int enable);

/* Debug/tuning: CTA pairs (cluster of 2, tcgen05 cta_group::2, half of every weight tile per CTA) for the sliding-window
 * layers with 256-wide N tiles: -1 / 1 = on (default), 0 = off (single-CTA kernel; the two must agree, tests run both). */
FFR_API void ffr_debug_set_pair(int mode);

/* Debug/tuning: programmatic dependent launch, a bit mask: 1 = the tcgen05 GEMM kernels, 2 = the memory-bound kernels of
 * the backbone chain are launched with the programmatic-stream-serialization attribute, so a kernel's CTAs are scheduled,
 * and its prologue runs, while the previous kernel drains; every such kernel executes griddepcontrol.wait before touching
 * global memory. 0: plain stream order; negative: the library default. Results are identical in every mode (A/B timing
 * and tests). */
FFR_API void ffr_debug_set_pdl(int mask);

/* Debug/tuning: 1 (default) lets the sliding-window kernels run the epilogue instantiation that is compiled for the
 * backbone's flag set only (bias / border bias / PReLU / geometry / squeeze sums / space-to-depth store); 0 forces the
 * generic epilogue. Same arithmetic, bit-identical results. */
FFR_API void ffr_debug_set_lean_epilogue(int enable);

/* Debug/tuning: 1 (default) runs the shared-memory strip kernel of the stem (ffr_stem_fwd / ffr_stem_u8_fwd) when
 * S % 16 == 0; 0 forces the gather kernel that serves every other size. Same arithmetic, bit-identical results. */
FFR_API void ffr_debug_set_stem_strip(int enable);

/* Stream-K scratch of the 3x3 / stride-1 convolutions with 256-wide N tiles (ffr_conv3x3_bnpre_prelu_fwd,
 * ffr_conv3x3_bn_pool_fwd at Cout % 256 == 0): when whole (256-row, 256-channel) work items would leave the last wave of
 * the 74 CTA pairs mostly empty (e.g. 450 items = 7 waves at 14x14 x 512 images), the (item, 64-channel k-chunk) steps are
 * split evenly over the pairs instead (6.25 waves); a pair that starts inside an item parks its fp32 partial
 * accumulator in this scratch and the pair that began the item adds it before the fused epilogue (own + peer, a fixed
 * order: bit-reproducible). The library never allocates: register ffr_conv_scratch_bytes() bytes of ZEROED device memory
 * (the first 1024 bytes are flag words that every launch leaves zero again) before the launches that should use it, on
 * the launching host thread; one scratch per stream that runs such convolutions concurrently. NULL unregisters (whole
 * items, as without scratch). The schedule is OPT-IN: ffr_debug_set_streamk(1) enables it for launches with a scratch
 * registered (default 0). Measured on B200 (tools/streamk_ab.py, tools/ab_bench.py): the last MMA of the 256->256@14x14
 * layer retires 7 % earlier, but the whole eval step does not get faster — with whole items the 68 of 74 pairs that
 * finish early already run the NEXT kernel's prologue under programmatic dependent launch, and the step is power-capped.
 * Replaces nothing in the reference: scheduling of model_ir_se50.py:67,69 on 148 SMs. */
FFR_API long long ffr_conv_scratch_bytes(void);
FFR_API int ffr_set_conv_scratch(void* scratch, long long bytes);
FFR_API void ffr_debug_set_streamk(int enable);
FFR_API int ffr_debug_last_streamk(void);   /* 1 if the most recent such launch was scheduled stream-K */

/* Debug/tuning: 1 (default) runs ffr_recnet_prep on the warp-MMA kernel (bf16 / tf32 tensor-core contractions of the
 * staged bf16 X); 0 forces the fp32 SIMT kernel. Same algebra; the results agree within the bf16 noise of the eval path. */
FFR_API void ffr_debug_set_prep_mma(int enable);

/* Debug/tuning: device buffer of 16 uint64 that the sliding-window kernel fills with per-role barrier-wait cycle
 * counts (summed over CTAs; slots in csrc/conv_gemm.cuh DbgSlot); NULL (default) disables the counters. */
FFR_API int ffr_debug_set_counters(void* counters);

/* 1 when H9 convolutions over n images run with pixel-major tiles (FFR_EPI_PIXMAJOR): the rule the library applies in
 * ffr_recnet_convlayer_fwd and that callers of ffr_conv_gemm / ffr_wgrad3x3 on H9 maps should follow. */
FFR_API int ffr_pixmajor_profitable(int n);
/* Tests / tuning: -1 = the rule above, 0 = never, 1 = always. */
FFR_API void ffr_debug_set_pixmajor(int mode);
/* Experiment switch: backbone 3x3 stride-1 convolutions on maps with S <= max_s (and >= 96 images) use pixel-major
 * tiles over the halo-shared flat layout (0 = off, the default; measured no gain). */
FFR_API void ffr_debug_set_pixmajor_backbone(int max_s);

/* Tuning only: splits > 0 overrides the split-count heuristic of ffr_wgrad3x3 (0 restores it). */
FFR_API void ffr_debug_set_wgrad_splits(int splits);

#ifdef __cplusplus
}
#endif
#endif /* FFR_SM100_H_ */
