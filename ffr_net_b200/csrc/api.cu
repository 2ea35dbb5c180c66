// extern "C" entry points declared in include/ffr_sm100.h.
#include "../../include/ffr_sm100.h"

#include <cstring>

#include "conv_gemm.cuh"
#include "host.h"
#include "kernels.h"



using namespace ffr;

static inline cudaStream_t S_(ffr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// 3x3 taps on a flat grid with row pitch G: tap (r,s) reads row m + (r-1)*G + (s-1).
static void taps_3x3_flat(ConvGemmParams& p, int G) {
    p.ntaps = 9;
    for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
            p.tap_row_shift[r * 3 + s] = (r - 1) * G + (s - 1);
            p.tap_ch_off[r * 3 + s] = 0;
        }
}

// 3x3 stride-2 taps on the space-to-depth grid (pitch G = So+1, 4*C channels): input pixel 2*ho + r - 1 lives in
// block ho + dh with parity ph where r=0 -> (dh=-1, ph=1), r=1 -> (0,0), r=2 -> (0,1).
static void taps_3x3_s2d(ConvGemmParams& p, int G, int C) {
    p.ntaps = 9;
    for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
            const int dh = (r == 0) ? -1 : 0, ph = (r == 1) ? 0 : 1;
            const int dw = (s == 0) ? -1 : 0, pw = (s == 1) ? 0 : 1;
            p.tap_row_shift[r * 3 + s] = dh * G + dw;
            p.tap_ch_off[r * 3 + s] = (ph * 2 + pw) * C;
        }
}

extern "C" {

FFR_API int ffr_version(void) { return 100; }

FFR_API const char* ffr_last_error(void) { return last_error_buf(); }

FFR_API long long ffr_launch_count(void) { return launch_count(); }

FFR_API int ffr_debug_set_window(int enable) { set_use_window(enable != 0); return 0; }
FFR_API void ffr_debug_set_pair(int mode) { set_pair_mode(mode); }
FFR_API void ffr_debug_set_pdl(int mask) { set_pdl_mask(mask); }
FFR_API void ffr_debug_set_lean_epilogue(int enable) { set_lean_epilogue(enable != 0); }
FFR_API void ffr_debug_set_stem_strip(int enable) { set_stem_strip(enable != 0); }
FFR_API void ffr_debug_set_streamk(int enable) { set_streamk(enable); }
FFR_API void ffr_debug_set_prep_mma(int enable) { set_prep_mma(enable); }
FFR_API int ffr_debug_last_streamk(void) { return last_streamk(); }
FFR_API long long ffr_conv_scratch_bytes(void) { return conv_scratch_bytes(); }
FFR_API int ffr_set_conv_scratch(void* scratch, long long bytes) {
    FFR_CHECK_ARG(scratch == nullptr || (bytes >= 1024 && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0),
                  "ffr_set_conv_scratch: scratch must be 16-byte aligned and hold the flag words");
    set_conv_scratch(scratch, scratch ? bytes : 0);
    return 0;
}

FFR_API int ffr_debug_set_counters(void* counters) {
    set_debug_counters(reinterpret_cast<unsigned long long*>(counters));
    return 0;
}

FFR_API int ffr_conv_gemm(const void* a, int64_t a_rows, int a_cols, int a_ld, const void* wp, int Cin, int Cout, int ntaps,
                  const int* tap_row_shift, const int* tap_ch_off, int M, int rows_per_img, int Wp, int S, int h0,
                  int n_img, uint32_t flags, const float* bias, const float* slope, void* out, int ldo, int s2d_So,
                  float* pool, float* out_f32, const void* res, int ldres, float* stats, int num_splits,
                  const int* scatter, int scatter_n, int out_rows_per_img, int b_rows_per_mtile, ffr_stream_t stream) {
    FFR_CHECK_ARG(a && wp, "ffr_conv_gemm: null operand");
    FFR_CHECK_ARG(ntaps >= 1 && ntaps <= 9, "ffr_conv_gemm: ntaps=%d", ntaps);
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = M;
    p.Cout = Cout;
    p.ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t) {
        p.tap_row_shift[t] = tap_row_shift ? tap_row_shift[t] : 0;
        p.tap_ch_off[t] = tap_ch_off ? tap_ch_off[t] : 0;
    }
    p.rows_per_img = rows_per_img; p.Wp = Wp; p.S = S; p.h0 = h0; p.n_img = n_img;
    p.flags = flags;
    p.bias = bias; p.slope = slope;
    p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo; p.s2d_So = s2d_So;
    p.pool = pool; p.out_f32 = out_f32;
    p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.ldres = ldres;
    p.stats = stats;
    p.scatter = reinterpret_cast<const int2*>(scatter); p.scatter_n = scatter_n; p.out_rows_per_img = out_rows_per_img;
    p.b_rows_per_mtile = b_rows_per_mtile;
    if (flags & (EPI_BIAS | EPI_BORDER_BIAS)) FFR_CHECK_ARG(bias, "ffr_conv_gemm: bias flag without bias");
    if (flags & EPI_PRELU) FFR_CHECK_ARG(slope, "ffr_conv_gemm: PReLU flag without slopes");
    if (flags & EPI_POOL) FFR_CHECK_ARG(pool, "ffr_conv_gemm: pool flag without buffer");
    if (flags & (EPI_OUT_F32_ATOMIC | EPI_OUT_F32)) FFR_CHECK_ARG(out_f32, "ffr_conv_gemm: fp32 output missing");
    if (flags & EPI_RESIDUAL) FFR_CHECK_ARG(res, "ffr_conv_gemm: residual missing");
    if (flags & EPI_STATS) FFR_CHECK_ARG(stats, "ffr_conv_gemm: stats buffer missing");
    return conv_gemm_launch(a, a_rows, a_cols, a_ld, wp, Cin, p, num_splits, S_(stream));
}

FFR_API int ffr_conv_gemm_ex(const ffr_conv_gemm_desc* d, ffr_stream_t stream) {
    FFR_CHECK_ARG(d && d->a && d->wp, "ffr_conv_gemm_ex: null operand");
    FFR_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= 9, "ffr_conv_gemm_ex: ntaps=%d", d->ntaps);
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = d->M;
    p.Cout = d->Cout;
    p.ntaps = d->ntaps;
    for (int t = 0; t < d->ntaps; ++t) {
        p.tap_row_shift[t] = d->tap_row_shift ? d->tap_row_shift[t] : 0;
        p.tap_ch_off[t] = d->tap_ch_off ? d->tap_ch_off[t] : 0;
    }
    p.rows_per_img = d->rows_per_img; p.Wp = d->Wp; p.S = d->S; p.h0 = d->h0; p.n_img = d->n_img;
    p.flags = d->flags;
    p.bias = d->bias; p.slope = d->slope;
    p.out = reinterpret_cast<__nv_bfloat16*>(d->out); p.ldo = d->ldo; p.s2d_So = d->s2d_So;
    p.pool = d->pool; p.out_f32 = d->out_f32;
    p.res = reinterpret_cast<const __nv_bfloat16*>(d->res); p.ldres = d->ldres;
    p.stats = d->stats; p.stats_part = d->stats_part;
    p.scatter = reinterpret_cast<const int2*>(d->scatter); p.scatter_n = d->scatter_n;
    p.out_rows_per_img = d->out_rows_per_img;
    p.b_rows_per_mtile = d->b_rows_per_mtile; p.b_mtile_div = d->b_mtile_div;
    p.a_hilo = d->a_hilo; p.a_lo_off = d->a_lo_off;
    p.idesc_xor = d->f16 ? ((1u << 7) | (1u << 10)) : 0u;     // a_format / b_format: BF16 (1) -> F16 (0)
    const uint32_t flags = d->flags;
    if (flags & (EPI_BIAS | EPI_BORDER_BIAS)) FFR_CHECK_ARG(d->bias, "ffr_conv_gemm_ex: bias flag without bias");
    if (flags & EPI_PRELU) FFR_CHECK_ARG(d->slope, "ffr_conv_gemm_ex: PReLU flag without slopes");
    if (flags & EPI_POOL) FFR_CHECK_ARG(d->pool, "ffr_conv_gemm_ex: pool flag without buffer");
    if (flags & (EPI_OUT_F32_ATOMIC | EPI_OUT_F32)) FFR_CHECK_ARG(d->out_f32, "ffr_conv_gemm_ex: fp32 output missing");
    if (flags & (EPI_RESIDUAL | EPI_MUL_DSIG)) FFR_CHECK_ARG(d->res, "ffr_conv_gemm_ex: residual operand missing");
    if (flags & EPI_STATS) FFR_CHECK_ARG(d->stats || d->stats_part, "ffr_conv_gemm_ex: stats buffer missing");
    return conv_gemm_launch(d->a, d->a_rows, d->a_cols, d->a_ld, d->wp, d->Cin, p, d->num_splits, S_(stream));
}

FFR_API int ffr_conv3x3_bnpre_prelu_fwd(const void* x, int n_img, int S, int Cin, const void* wp, int Cout,
                                const float* bias9, const float* slope, void* out, int out_s2d, ffr_stream_t stream) {
    FFR_CHECK_ARG(x && wp && bias9 && slope && out, "ffr_conv3x3_bnpre_prelu_fwd: null pointer");
    FFR_CHECK_ARG(!out_s2d || S % 2 == 0, "ffr_conv3x3_bnpre_prelu_fwd: s2d output needs even S (got %d)", S);
    const int G = S + 1;
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = n_img * G * G;
    p.Cout = Cout;
    taps_3x3_flat(p, G);
    p.rows_per_img = G * G; p.Wp = G; p.S = S; p.h0 = 0; p.n_img = n_img;
    p.flags = EPI_GEOM | EPI_BORDER_BIAS | EPI_PRELU | (out_s2d ? EPI_OUT_S2D : 0u) |
              (pixmajor_backbone(S, n_img) ? EPI_PIXMAJOR : 0u);
    p.bias = bias9; p.slope = slope;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ldo = out_s2d ? 4 * Cout : Cout;
    p.s2d_So = S / 2;
    return conv_gemm_launch(x, (long long)p.M, Cin, Cin, wp, Cin, p, 1, S_(stream));
}

FFR_API long long ffr_se_pool_part_floats(int n_img, int So, int Cout) {
    const long long M = (long long)n_img * (So + 1) * (So + 1);
    return ((M + 127) / 128) * 4 * 2 * Cout;
}

FFR_API int ffr_conv3x3_bn_pool_fwd(const void* x, int n_img, int S, int C, int stride, const void* wp, int Cout,
                            const float* bias, void* out, float* pool_part, ffr_stream_t stream) {
    FFR_CHECK_ARG(x && wp && bias && out, "ffr_conv3x3_bn_pool_fwd: null pointer");
    FFR_CHECK_ARG(stride == 1 || (stride == 2 && S % 2 == 0), "ffr_conv3x3_bn_pool_fwd: stride=%d S=%d", stride, S);
    const int So = S / stride, G = So + 1;
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = n_img * G * G;
    p.Cout = Cout;
    if (stride == 1) taps_3x3_flat(p, G); else taps_3x3_s2d(p, G, C);
    p.rows_per_img = G * G; p.Wp = G; p.S = So; p.h0 = 0; p.n_img = n_img;
    const bool pix = stride == 1 && pixmajor_backbone(So, n_img);
    p.flags = EPI_GEOM | EPI_BIAS | (pool_part ? EPI_POOL : 0u) | (pix ? EPI_PIXMAJOR : 0u);
    p.bias = bias;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ldo = Cout;
    if (pixmajor_backbone(So, n_img)) {   // experiment switch (all strides, so that ffr_se_gate_fwd can tell from S and
                                          // n_img alone): atomics into dense [n_img][Cout] sums
        p.pool = pool_part;
        if (pool_part) FFR_CUDA(cudaMemsetAsync(pool_part, 0, sizeof(float) * (size_t)n_img * Cout, S_(stream)));
    } else {
        p.pool_part = pool_part;
    }
    const int a_cols = (stride == 1) ? C : 4 * C;
    return conv_gemm_launch(x, (long long)p.M, a_cols, a_cols, wp, C, p, 1, S_(stream));
}

FFR_API int ffr_conv1x1_bn_fwd(const void* xs, int n_img, int S, int Cin, const void* wp, int Cout, const float* bias,
                       void* out, ffr_stream_t stream) {
    FFR_CHECK_ARG(xs && wp && bias && out, "ffr_conv1x1_bn_fwd: null pointer");
    const int G = S + 1;
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = n_img * G * G;
    p.Cout = Cout;
    p.ntaps = 1;
    p.rows_per_img = G * G; p.Wp = G; p.S = S; p.h0 = 0; p.n_img = n_img;
    p.flags = EPI_GEOM | EPI_BIAS;
    p.bias = bias;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ldo = Cout;
    return conv_gemm_launch(xs, (long long)p.M, Cin, Cin, wp, Cin, p, 1, S_(stream));
}

FFR_API int ffr_subsample2(const void* x, void* out, int n_img, int So, int C, ffr_stream_t stream) {
    FFR_CHECK_ARG(x && out && C % 8 == 0, "ffr_subsample2: bad arguments");
    return subsample2_launch(x, out, n_img, So, C, S_(stream));
}

FFR_API int ffr_stem_fwd(const float* x, const float* w, const float* b, const float* a, void* out, int n_img, int S,
                 ffr_stream_t stream) {
    FFR_CHECK_ARG(x && w && b && a && out, "ffr_stem_fwd: null pointer");
    return stem_launch(x, w, b, a, out, n_img, S, S_(stream));
}

FFR_API int ffr_stem_u8_fwd(const unsigned char* img, const unsigned char* flip, int swap_rb, const float* w,
                            const float* b, const float* a, void* out, int n_img, int S, ffr_stream_t stream) {
    FFR_CHECK_ARG(img && w && b && a && out, "ffr_stem_u8_fwd: null pointer");
    return stem_u8_launch(img, flip, swap_rb, w, b, a, out, n_img, S, S_(stream));
}

FFR_API int ffr_se_gate_fwd(const float* pool_part, const float* w1, const float* w2, float* gate, float* sums, int n_img,
                            int S, int C, ffr_stream_t stream) {
    FFR_CHECK_ARG(pool_part && w1 && w2 && gate, "ffr_se_gate_fwd: null pointer");
    return se_gate_launch(pool_part, pixmajor_backbone(S, n_img) ? 1 : 0, w1, w2, gate, sums, n_img, S, C, S_(stream));
}

FFR_API int ffr_se_gate_residual_fwd(const void* u, const float* pool_part, const float* w1, const float* w2,
                                     const void* shortcut, int shortcut_mode, void* y, int n_img, int S, int C,
                                     ffr_stream_t stream) {
    FFR_CHECK_ARG(u && pool_part && w1 && w2 && shortcut && y, "ffr_se_gate_residual_fwd: null pointer");
    return se_gate_residual_launch(u, pool_part, pixmajor_backbone(S, n_img) ? 1 : 0, w1, w2, shortcut, shortcut_mode, y,
                                   n_img, S, C, S_(stream));
}

FFR_API int ffr_se_residual_fwd(const void* u, const float* gate, const void* shortcut, int shortcut_mode, void* y,
                                int n_img, int S, int C, ffr_stream_t stream) {
    FFR_CHECK_ARG(u && gate && shortcut && y, "ffr_se_residual_fwd: null pointer");
    return se_residual_launch(u, gate, shortcut, shortcut_mode, y, n_img, S, C, S_(stream));
}

FFR_API int ffr_export_nchw_fwd(const void* h, const float* scale, const float* shift, float* y, int n_img, int S, int C,
                        ffr_stream_t stream) {
    FFR_CHECK_ARG(h && scale && shift && y, "ffr_export_nchw_fwd: null pointer");
    return export_nchw_launch(h, scale, shift, y, n_img, S, C, S_(stream));
}

// split count of the head GEMM: enough to put ~one wave of CTAs on the machine
static int head_splits(int n_img, int K) {
    const int tiles = ((n_img + 127) / 128) * (512 / 256);
    int splits = (num_sms() + tiles - 1) / tiles;
    const int kb = K / 64;
    if (splits > kb / 4) splits = kb / 4;
    if (splits < 1) splits = 1;
    const int per = (kb + splits - 1) / splits;
    return (kb + per - 1) / per;
}

FFR_API long long ffr_head_workspace_floats(int n_img, int S, int C) {
    return (long long)head_splits(n_img, (S + 1) * (S + 1) * C) * n_img * 512;
}

FFR_API int ffr_head_fwd(const void* h, int n_img, int S, int C, const void* wp, const float* bias, float* acc, float* f,
                 ffr_stream_t stream) {
    FFR_CHECK_ARG(h && wp && bias && acc && f, "ffr_head_fwd: null pointer");
    const int D = 512;
    const int K = (S + 1) * (S + 1) * C;       // one image's flat rows, viewed as a single GEMM row
    FFR_CHECK_ARG(K % 64 == 0, "ffr_head_fwd: K=%d", K);
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = n_img;
    p.Cout = D;
    p.ntaps = 1;
    p.flags = EPI_OUT_F32;                     // deterministic split-K: one fp32 partial product per split, plain stores
    p.out_f32 = acc;
    p.out_f32_split_stride = (long long)n_img * D;
    const int splits = head_splits(n_img, K);
    int rc = conv_gemm_launch(h, n_img, K, K, wp, K, p, splits, S_(stream));
    if (rc) return rc;
    return bias_l2norm_launch(acc, splits, p.out_f32_split_stride, bias, f, n_img, D, S_(stream));
}

FFR_API int ffr_recnet_prep(const float* x, int n, const float* w0aT, const float* w0bT, const float* b0,
                            const float* slope1, const float* A1, const float* c1, const float* slope4,
                            const float* A2, const float* c2, const float* slope7, void* s0, void* cm, void* xt,
                            void* h5, float* ss_space, ffr_stream_t stream) {
    FFR_CHECK_ARG(x && w0aT && w0bT && b0 && slope1 && A1 && c1 && slope4 && A2 && c2 && slope7 && s0 && cm && xt && h5,
                  "ffr_recnet_prep: null pointer");
    PrepParams p{x, w0aT, w0bT, b0, slope1, A1, c1, slope4, A2, c2, slope7,
                 reinterpret_cast<__nv_bfloat16*>(s0), reinterpret_cast<__nv_bfloat16*>(cm),
                 reinterpret_cast<__nv_bfloat16*>(xt), reinterpret_cast<__nv_bfloat16*>(h5), ss_space};
    return recnet_prep_launch(p, n, S_(stream));
}

FFR_API int ffr_recnet_convlayer_fwd(const void* x, int n, int Cin, const void* wp, int Cout, const float* bias,
                                     const float* slope, const void* res, int ldres, int sigmoid, void* out, int ldo,
                                     const int* scatter, int scatter_n, int out_rows_per_img, float* out_f32,
                                     float* pool, ffr_stream_t stream) {
    FFR_CHECK_ARG(x && wp && bias, "ffr_recnet_convlayer_fwd: null pointer");
    FFR_CHECK_ARG(out || out_f32, "ffr_recnet_convlayer_fwd: no output");
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = n * 81;
    p.Cout = Cout;
    taps_3x3_flat(p, 9);
    p.rows_per_img = 81; p.Wp = 9; p.S = 7; p.h0 = 1; p.n_img = n;
    p.flags = EPI_GEOM | EPI_BIAS | (slope ? EPI_PRELU : 0u) | (res ? EPI_RESIDUAL : 0u) |
              (sigmoid ? EPI_SIGMOID : 0u) | (out ? EPI_SCATTER : 0u) | (out_f32 ? EPI_OUT_F32 : 0u) |
              (pool ? EPI_POOL : 0u);
    p.bias = bias; p.slope = slope;
    p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.ldres = ldres;
    p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo;
    p.scatter = reinterpret_cast<const int2*>(scatter); p.scatter_n = scatter_n; p.out_rows_per_img = out_rows_per_img;
    p.out_f32 = out_f32;
    p.pool = pool;
    if (pixmajor_profitable(n)) p.flags |= EPI_PIXMAJOR;
    if (pool) FFR_CUDA(cudaMemsetAsync(pool, 0, sizeof(float) * (size_t)n * Cout, S_(stream)));
    return conv_gemm_launch(x, (long long)p.M, Cin, Cin, wp, Cin, p, 1, S_(stream));
}

FFR_API int ffr_pixmajor_profitable(int n) { return pixmajor_profitable(n) ? 1 : 0; }
FFR_API void ffr_debug_set_pixmajor(int mode) { set_pixmajor_mode(mode); }
FFR_API void ffr_debug_set_pixmajor_backbone(int max_s) { set_pixmajor_backbone(max_s); }

FFR_API int ffr_self_similarity(const float* x, int n, float* ss_space, float* ss_channel, ffr_stream_t stream) {
    FFR_CHECK_ARG(n == 0 || x, "ffr_self_similarity: null input");
    FFR_CHECK_ARG(ss_space || ss_channel, "ffr_self_similarity: no output requested");
    return self_similarity_launch(x, n, ss_space, ss_channel, S_(stream));
}

FFR_API int ffr_feat_space(const float* x, const float* mspace, void* cm, float* out_nchw, int n, ffr_stream_t stream) {
    FFR_CHECK_ARG(x && mspace && cm, "ffr_feat_space: null pointer");
    return feat_space_launch(x, mspace, cm, out_nchw, n, S_(stream));
}

FFR_API int ffr_feat_space_xt(const void* xt, const float* mspace, void* cm, float* out_nchw, int n, ffr_stream_t stream) {
    FFR_CHECK_ARG(xt && mspace && cm, "ffr_feat_space_xt: null pointer");
    if (n == 0) return 0;
    return feat_space_xt_launch(xt, mspace, cm, out_nchw, n, S_(stream));
}

FFR_API int ffr_rows_to_nchw(const void* rows, int is_f32, int ld, int ch0, const float* scale, const float* shift,
                             float* y, int n, int S, int G, int off, int rows_per_img, int C, ffr_stream_t stream) {
    FFR_CHECK_ARG(rows && y, "ffr_rows_to_nchw: null pointer");
    return rows_to_nchw_launch(rows, is_f32, ld, ch0, scale, shift, y, n, S, G, off, rows_per_img, C, S_(stream));
}

FFR_API int ffr_scale_f32(const float* in, float* out, int64_t count, float scale, ffr_stream_t stream) {
    FFR_CHECK_ARG(in && out, "ffr_scale_f32: null pointer");
    return scale_f32_launch(in, out, count, scale, S_(stream));
}

FFR_API void ffr_debug_set_wgrad_splits(int splits) { set_wgrad_splits(splits); }

FFR_API int64_t ffr_wgrad_workspace_floats(int P, int Cout, int Cin, int ntaps, int deterministic) {
    return wgrad_workspace_floats(P, Cout, Cin, ntaps, 9, deterministic, nullptr);
}

FFR_API int ffr_wgrad(const void* dz, int ld_dz, const void* x, int ld_x, int x_ch0, int P, int Cout, int Cin, int ntaps,
                      int f16, int deterministic, int accumulate, int ld_w, int bias_col, float* dw, float* db,
                      float* workspace, ffr_stream_t stream) {
    FFR_CHECK_ARG(dz && x && dw && workspace, "ffr_wgrad: null pointer");
    FFR_CHECK_ARG(ntaps == 9 || ntaps == 1, "ffr_wgrad: ntaps=%d", ntaps);
    FFR_CHECK_ARG(ld_dz % 64 == 0 && ld_x % 8 == 0 && x_ch0 % 8 == 0, "ffr_wgrad: bad pitches");
    FFR_CHECK_ARG(bias_col < 0 || db, "ffr_wgrad: bias column without db");
    return wgrad_launch_ex(dz, ld_dz, x, ld_x, x_ch0, P, Cout, Cin, 9, ntaps, f16, deterministic, accumulate, ld_w, bias_col,
                           dw, db, workspace, S_(stream));
}

FFR_API int ffr_pack_conv3x3_f16(const float* w, int cout, int cin, int cout_p, int cin_p, void* fwd_f16, void* dgrad_bf16,
                                 ffr_stream_t stream) {
    FFR_CHECK_ARG(w && fwd_f16, "ffr_pack_conv3x3_f16: null pointer");
    return pack_conv3x3_launch_ex(w, cout, cin, cout_p, cin_p, fwd_f16, dgrad_bf16, 1, S_(stream));
}

FFR_API int ffr_pack_conv3x3(const float* w, int cout, int cin, int cout_p, int cin_p, void* fwd, void* dgrad,
                             ffr_stream_t stream) {
    FFR_CHECK_ARG(w && fwd, "ffr_pack_conv3x3: null pointer");
    return pack_conv3x3_launch(w, cout, cin, cout_p, cin_p, fwd, dgrad, S_(stream));
}

FFR_API int ffr_cosface_pack(const float* x, int rows, int rows_pad, int mode, void* packed, void* transposed, int t_ld,
                             ffr_stream_t stream) {
    FFR_CHECK_ARG((rows == 0 || x) && packed, "ffr_cosface_pack: null pointer");
    FFR_CHECK_ARG(rows >= 0 && rows_pad >= rows && rows_pad % 64 == 0, "ffr_cosface_pack: rows=%d rows_pad=%d", rows, rows_pad);
    FFR_CHECK_ARG(!transposed || (t_ld >= rows_pad && t_ld % 8 == 0), "ffr_cosface_pack: t_ld=%d", t_ld);
    return cosface_pack_launch(x, rows, rows_pad, mode, packed, transposed, t_ld, S_(stream));
}

FFR_API int ffr_cosface_ce_fwd(const void* v_packed, int n, const void* w_packed, int c_pad, int classes, const int* label,
                               float s, float m, float* cos_out, float* sumexp, float* zlabel,
                               unsigned long long* argkey, float* sumexp_part, ffr_stream_t stream) {
    FFR_CHECK_ARG(v_packed && w_packed && label && cos_out && sumexp && zlabel && argkey, "ffr_cosface_ce_fwd: null pointer");
    FFR_CHECK_ARG(n > 0 && c_pad % 256 == 0 && classes > 0 && classes <= c_pad, "ffr_cosface_ce_fwd: n=%d c_pad=%d classes=%d",
                  n, c_pad, classes);
    if (!sumexp_part) FFR_CUDA(cudaMemsetAsync(sumexp, 0, sizeof(float) * (size_t)n, S_(stream)));
    FFR_CUDA(cudaMemsetAsync(argkey, 0, sizeof(unsigned long long) * (size_t)n, S_(stream)));
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = n;
    p.Cout = c_pad;
    p.ntaps = 1;
    p.flags = EPI_OUT_F32 | EPI_COSFACE;
    p.out_f32 = cos_out;
    p.ce_label = label; p.ce_sumexp = sumexp; p.ce_zlabel = zlabel; p.ce_argkey = argkey;
    p.ce_sumexp_part = sumexp_part;
    p.ce_classes = classes; p.ce_s = s; p.ce_m = m;
    int rc = conv_gemm_launch(v_packed, (long long)n, 1536, 1536, w_packed, 1536, p, 1, S_(stream));
    if (rc || !sumexp_part) return rc;
    return sumexp_reduce_launch(sumexp_part, n, c_pad / 128, sumexp, S_(stream));   // N tile 256, two column halves each
}

FFR_API int ffr_cosface_ce_finish(const float* sumexp, const float* zlabel, const unsigned long long* argkey, int n, float s,
                                  float* loss, long long* pred, ffr_stream_t stream) {
    FFR_CHECK_ARG(sumexp && zlabel && argkey && loss && n > 0, "ffr_cosface_ce_finish: bad argument");
    return cosface_finish_launch(sumexp, zlabel, argkey, n, s, loss, pred, S_(stream));
}

FFR_API int ffr_cosface_ce_bwd(const float* cos_in, int c_pad, int classes, int n, int n_pad, const int* label,
                               const float* sumexp, const float* gloss, float s, float m, void* dcos, void* dcosT,
                               ffr_stream_t stream) {
    FFR_CHECK_ARG(cos_in && label && sumexp && gloss && dcos && dcosT, "ffr_cosface_ce_bwd: null pointer");
    FFR_CHECK_ARG(c_pad % 64 == 0 && n_pad % 64 == 0 && n_pad >= n && classes <= c_pad, "ffr_cosface_ce_bwd: bad shape");
    return cosface_bwd_launch(cos_in, c_pad, classes, n, n_pad, label, sumexp, gloss, s, m, dcos, dcosT, S_(stream));
}

FFR_API int ffr_cosface_ce_bwd_grouped(const float* cos_in, int c_pad, int classes, int n, int n_pad, const int* label,
                                       const float* sumexp, const float* gloss, int n_per_group, float s, float m, void* dcos,
                                       void* dcosT, ffr_stream_t stream) {
    FFR_CHECK_ARG(cos_in && label && sumexp && gloss && dcos && dcosT, "ffr_cosface_ce_bwd_grouped: null pointer");
    FFR_CHECK_ARG(c_pad % 64 == 0 && n_pad % 64 == 0 && n_pad >= n && classes <= c_pad && n_per_group > 0 &&
                  n % n_per_group == 0, "ffr_cosface_ce_bwd_grouped: bad shape");
    return cosface_bwd_launch_ex(cos_in, c_pad, classes, n, n_pad, label, sumexp, gloss, s, m, dcos, dcosT, n_per_group,
                                 S_(stream));
}

FFR_API int ffr_normalize_bwd(const float* x, const float* dxh, int rows, float* dx, ffr_stream_t stream) {
    FFR_CHECK_ARG(rows == 0 || (x && dxh && dx), "ffr_normalize_bwd: null pointer");
    return normalize_bwd_launch(x, dxh, rows, dx, S_(stream));
}

FFR_API int ffr_gallery_cosine(const void* probe_packed, int P, const void* gallery_packed, int g_pad, int G,
                               float* cos_out, unsigned long long* argkey, ffr_stream_t stream) {
    FFR_CHECK_ARG(probe_packed && gallery_packed && cos_out, "ffr_gallery_cosine: null pointer");
    FFR_CHECK_ARG(P > 0 && g_pad % 256 == 0 && G > 0 && G <= g_pad, "ffr_gallery_cosine: P=%d g_pad=%d G=%d", P, g_pad, G);
    ConvGemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = P;
    p.Cout = g_pad;
    p.ntaps = 1;
    p.flags = EPI_OUT_F32;
    p.out_f32 = cos_out;
    if (argkey) {                                  // rank-1 match per probe through the arg-max part of EPI_COSFACE
        FFR_CUDA(cudaMemsetAsync(argkey, 0, sizeof(unsigned long long) * (size_t)P, S_(stream)));
        p.flags |= EPI_COSFACE;
        p.ce_label = nullptr; p.ce_sumexp = nullptr; p.ce_zlabel = nullptr; p.ce_argkey = argkey;
        p.ce_classes = G; p.ce_s = 0.f; p.ce_m = 0.f;
    }
    return conv_gemm_launch(probe_packed, (long long)P, 1536, 1536, gallery_packed, 1536, p, 1, S_(stream));
}

FFR_API int ffr_roc_hist(const float* scores, int ld, int P, int G, const int* probe_id, const int* gallery_id,
                         const double* thresholds, int T, unsigned long long* hist, ffr_stream_t stream) {
    FFR_CHECK_ARG(hist && thresholds && ((P == 0 || G == 0) || (scores && probe_id && gallery_id)), "ffr_roc_hist: null pointer");
    return roc_hist_launch(scores, ld, P, G, probe_id, gallery_id, thresholds, T, hist, S_(stream));
}

FFR_API int ffr_clip_adam(const void* table, const int* chunks, int n_chunks, const float* hyper, float beta1,
                          float beta2, float eps, float weight_decay, float clip, ffr_stream_t stream) {
    FFR_CHECK_ARG(n_chunks == 0 || (table && chunks && hyper), "ffr_clip_adam: null pointer");
    return clip_adam_launch(table, chunks, n_chunks, hyper, beta1, beta2, eps, weight_decay, clip, S_(stream));
}

FFR_API int ffr_pair_cosine(const float* f1, const float* f2, float* score, int pairs, int D, ffr_stream_t stream) {
    FFR_CHECK_ARG(pairs == 0 || (f1 && f2 && score), "ffr_pair_cosine: null pointer");
    return pair_cosine_launch(f1, f2, score, pairs, D, S_(stream));
}

FFR_API int ffr_threshold_sweep(const float* score, const int* label, const double* thresholds, int n, int T, int folds,
                                int* best_idx, double* best_thr, int* test_correct, int* train_correct,
                                ffr_stream_t stream) {
    FFR_CHECK_ARG(score && label && thresholds && best_idx && best_thr && test_correct && train_correct,
                  "ffr_threshold_sweep: null pointer");
    return threshold_sweep_launch(score, label, thresholds, n, T, folds, best_idx, best_thr, test_correct,
                                  train_correct, S_(stream));
}

}  // extern "C"
