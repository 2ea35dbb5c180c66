// Host-side launch functions of every kernel file (implemented in *_kernels.cu / conv_gemm.cu), shared with api.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_gemm.cuh"

namespace ffr {
int conv_gemm_launch(const void* a, long long a_rows, int a_cols, int a_ld, const void* wp, int Cin, ConvGemmParams p,
                     int num_splits, cudaStream_t stream);
int stem_launch(const float* x, const float* w, const float* b, const float* a, void* out, int n_img, int S,
                cudaStream_t stream);
int stem_u8_launch(const unsigned char* img, const unsigned char* flip, int swap_rb, const float* w, const float* b,
                   const float* a, void* out, int n_img, int S, cudaStream_t stream);
int se_gate_launch(const float* pool_part, int dense, const float* w1, const float* w2, float* gate, float* sums,
                   int n_img, int S, int C, cudaStream_t stream);
int se_gate_residual_launch(const void* u, const float* pool_part, int dense, const float* w1, const float* w2,
                            const void* sc, int shortcut_mode, void* y, int n_img, int S, int C, cudaStream_t stream);
int se_residual_launch(const void* u, const float* gate, const void* sc, int shortcut_mode, void* y, int n_img, int S,
                       int C, cudaStream_t stream);
int subsample2_launch(const void* x, void* out, int n_img, int So, int C, cudaStream_t stream);
int export_nchw_launch(const void* h, const float* scale, const float* shift, float* y, int n_img, int S, int C,
                       cudaStream_t stream);
int bias_l2norm_launch(const float* acc, int splits, long long split_stride, const float* bias, float* f, int rows, int D,
                       cudaStream_t stream);
void set_use_window(bool on);
void set_pair_mode(int mode);
void set_lean_epilogue(bool on);
void set_stem_strip(bool on);
void set_conv_scratch(void* ptr, long long bytes);
long long conv_scratch_bytes();
void set_streamk(int on);
void set_prep_mma(int on);
int last_streamk();
void set_debug_counters(unsigned long long* dptr);
struct PrepParams {
    const float* x; const float* w0aT; const float* w0bT; const float* b0; const float* slope1; const float* A1;
    const float* c1; const float* slope4; const float* A2; const float* c2; const float* slope7;
    __nv_bfloat16* s0; __nv_bfloat16* cm; __nv_bfloat16* xt; __nv_bfloat16* h5; float* ss_space;
};
int recnet_prep_launch(const PrepParams& p, int n, cudaStream_t stream);
int feat_space_launch(const float* x, const float* mspace, void* cm, float* out_nchw, int n, cudaStream_t stream);
int feat_space_xt_launch(const void* xt, const float* mspace, void* cm, float* out_nchw, int n, cudaStream_t stream);
int rows_to_nchw_launch(const void* rows, int is_f32, int ld, int ch0, const float* scale, const float* shift, float* y,
                        int n, int S, int G, int off, int rows_per_img, int C, cudaStream_t stream);
int cosface_pack_launch(const float* x, int rows, int rows_pad, int mode, void* packed, void* transposed, int t_ld,
                        cudaStream_t stream);
int sumexp_reduce_launch(const float* part, int n, int parts, float* sumexp, cudaStream_t stream);
int cosface_finish_launch(const float* sumexp, const float* zlabel, const unsigned long long* argkey, int n, float s,
                          float* loss, long long* pred, cudaStream_t stream);
int cosface_bwd_launch(const float* cosv, int c_pad, int classes, int n, int n_pad, const int* label, const float* sumexp,
                       const float* gloss, float s, float m, void* dcos, void* dcosT, cudaStream_t stream);
int cosface_bwd_launch_ex(const float* cosv, int c_pad, int classes, int n, int n_pad, const int* label, const float* sumexp,
                          const float* gloss, float s, float m, void* dcos, void* dcosT, int n_per_group, cudaStream_t stream);
int normalize_bwd_launch(const float* x, const float* dxh, int rows, float* dx, cudaStream_t stream);
int roc_hist_launch(const float* scores, int ld, int P, int G, const int* probe_id, const int* gallery_id,
                    const double* thresholds, int T, unsigned long long* hist, cudaStream_t stream);
int self_similarity_launch(const float* x, int n, float* ss_space, float* ss_channel, cudaStream_t stream);
int scale_f32_launch(const float* in, float* out, long long count, float scale, cudaStream_t stream);
long long wgrad_workspace_floats(int P, int Cout, int Cin, int ntaps, int G, int deterministic, int* splits_out);
int wgrad_launch_ex(const void* dz, int ld_dz, const void* x, int ld_x, int x_ch0, int P, int Cout, int Cin, int G,
                    int ntaps, int f16, int deterministic, int accumulate, int ld_w, int bias_col, float* dw, float* db,
                    float* ws, cudaStream_t stream);
void set_wgrad_splits(int s);
void set_pixmajor_mode(int mode);
bool pixmajor_profitable(int n_img);
void set_pixmajor_backbone(int max_s);
bool pixmajor_backbone(int S, int n_img);
bool pixmajor_profitable_k64(int n_img);
int pack_conv3x3_launch(const float* w, int cout, int cin, int cout_p, int cin_p, void* fwd, void* dgrad,
                        cudaStream_t stream);
int pack_conv3x3_launch_ex(const float* w, int cout, int cin, int cout_p, int cin_p, void* fwd, void* dgrad, int fwd_f16,
                           cudaStream_t stream);
int clip_adam_launch(const void* table, const int* chunks, int n_chunks, const float* hyper, float b1, float b2,
                     float eps, float wd, float clip, cudaStream_t stream);
int pair_cosine_launch(const float* f1, const float* f2, float* score, int pairs, int D, cudaStream_t stream);
int threshold_sweep_launch(const float* score, const int* label, const double* thresholds, int n, int T, int folds,
                           int* best_idx, double* best_thr, int* test_correct, int* train_correct,
                           cudaStream_t stream);
}  // namespace ffr
