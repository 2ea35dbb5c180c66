/*
 * libffr_sm100_probe — hardware-semantics probes and micro-benchmarks (csrc/probe.cu). NOT part of the product
 * library: built as a separate shared object, loaded only by tests/test_probe_gpu.py and the developer tools.
 * They established the facts the kernels of libffr_sm100 rely on (row-offset SWIZZLE_128B descriptors, MN-major
 * operands, fp16 operand format, tcgen05.mma issue rates); results are recorded under profiles/.
 */
#ifndef FFR_SM100_PROBE_H_
#define FFR_SM100_PROBE_H_

#include "ffr_sm100.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Debug: hardware-semantics probe for row-offset UMMA descriptors (csrc/probe.cu); not on the product path.
 * a [256][64] bf16, w [64][64] bf16, out [128][64] fp32 = a[row_off : row_off+128] @ w^T. */
FFR_API int ffr_debug_rowshift_probe(const void* a, const void* w, float* out, int row_off, int variant,
                                     ffr_stream_t stream);

/* Debug: MN-major UMMA operand probe (csrc/probe.cu): a [96][128], b [96][64] bf16 (k rows);
 * out[128][64] = sum_{k<64} a[k][m] * b[r0+k][n]. */
FFR_API int ffr_debug_mn_probe(const void* a, const void* b, float* out, int r0, int variant, ffr_stream_t stream);

/* Debug: tcgen05.mma issue-rate micro-benchmark (csrc/probe.cu). out_cycles[grid] = cycles for `iters` MMAs of shape
 * M x N x 16 alternating between n_acc accumulators. */
FFR_API int ffr_debug_mma_bench(long long* out_cycles, int M, int N, int n_acc, int iters, int grid,
                                ffr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FFR_SM100_PROBE_H_ */
