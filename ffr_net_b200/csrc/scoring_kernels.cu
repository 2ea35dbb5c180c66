// LFW-style verification scoring (lfw/lfw_eval.py): row-wise cosine of embedding pairs and the 10-fold threshold
// sweep, both on the device.
#include "host.h"
#include "kernels.h"
#include "ptx.cuh"

namespace ffr {

// score[i] = sum(f1[i]*f2[i]) / (|f1[i]| * |f2[i]| + 1e-8)     (lfw_eval.py:246,248). One warp per pair, fp32.
__global__ void __launch_bounds__(256) pair_cosine_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                          float* __restrict__ score, int pairs, int D) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= pairs) return;
    const float4* a = reinterpret_cast<const float4*>(f1 + (long long)i * D);
    const float4* b = reinterpret_cast<const float4*>(f2 + (long long)i * D);
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int k = lane; k < D / 4; k += 32) {
        const float4 x = __ldg(a + k), y = __ldg(b + k);
        dot = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, dot))));
        na = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, na))));
        nb = fmaf(y.x, y.x, fmaf(y.y, y.y, fmaf(y.z, y.z, fmaf(y.w, y.w, nb))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
        na += __shfl_xor_sync(0xffffffffu, na, o);
        nb += __shfl_xor_sync(0xffffffffu, nb, o);
    }
    if (lane == 0) score[i] = dot / (sqrtf(na) * sqrtf(nb) + 1e-8f);
}

int pair_cosine_launch(const float* f1, const float* f2, float* score, int pairs, int D, cudaStream_t stream) {
    FFR_CHECK_ARG(D % 4 == 0, "pair_cosine: D=%d not a multiple of 4", D);
    if (pairs == 0) return 0;
    pair_cosine_kernel<<<(pairs + 7) / 8, 256, 0, stream>>>(f1, f2, score, pairs, D);
    return launch_status("pair_cosine_kernel");
}

// K-fold threshold sweep (lfw_eval.py:110-118,137-162,255-259). Folds are contiguous: fold f tests pairs
// [f*n/k, (f+1)*n/k) and trains on the rest. A pair is predicted "same" iff (double)score > threshold (strict).
// For every fold: best threshold = the LAST threshold reaching the maximal training accuracy (the reference updates
// on >=), then the accuracy on the held-out pairs at that threshold. Counts are integers, so ties are exact.
// One CTA; thread t owns threshold t (T <= 1024); per-fold correct counts stay in registers.
constexpr int SWEEP_MAX_FOLDS = 16;
__global__ void __launch_bounds__(1024) threshold_sweep_kernel(const float* __restrict__ score,
                                                               const int* __restrict__ label,
                                                               const double* __restrict__ thresholds, int n, int T,
                                                               int folds, int* __restrict__ best_idx,
                                                               double* __restrict__ best_thr,
                                                               int* __restrict__ test_correct,
                                                               int* __restrict__ train_correct) {
    __shared__ float s_score[1024];
    __shared__ int s_label[1024];
    __shared__ unsigned long long s_key[32];
    __shared__ int s_best;
    const int t = threadIdx.x;
    const double thr = (t < T) ? thresholds[t] : 0.0;
    int cnt[SWEEP_MAX_FOLDS];
#pragma unroll
    for (int f = 0; f < SWEEP_MAX_FOLDS; ++f) cnt[f] = 0;
#pragma unroll
    for (int f = 0; f < SWEEP_MAX_FOLDS; ++f) {
        if (f < folds) {
            const int lo = (int)((long long)f * n / folds), hi = (int)((long long)(f + 1) * n / folds);
            for (int base = lo; base < hi; base += 1024) {
                __syncthreads();
                if (base + t < hi) { s_score[t] = score[base + t]; s_label[t] = label[base + t]; }
                __syncthreads();
                const int m = min(1024, hi - base);
                int c = 0;
                for (int i = 0; i < m; ++i) {
                    const int same = ((double)s_score[i] > thr) ? 1 : 0;
                    c += (same == s_label[i]) ? 1 : 0;
                }
                cnt[f] += c;
            }
        }
    }
    int total = 0;
#pragma unroll
    for (int f = 0; f < SWEEP_MAX_FOLDS; ++f) total += cnt[f];
#pragma unroll
    for (int f = 0; f < SWEEP_MAX_FOLDS; ++f) {
        if (f < folds) {
            // argmax over thresholds of the training count, last index wins on ties
            unsigned long long key = (t < T) ? (((unsigned long long)(unsigned)(total - cnt[f]) << 32) | (unsigned)t) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            __syncthreads();
            if ((t & 31) == 0) s_key[t >> 5] = key;
            __syncthreads();
            if (t < 32) {
                key = s_key[t];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other > key ? other : key;
                }
                if (t == 0) {
                    s_best = (int)(key & 0xffffffffu);
                    best_idx[f] = s_best;
                    train_correct[f] = (int)(key >> 32);
                }
            }
            __syncthreads();
            if (t == s_best) {
                best_thr[f] = thr;
                test_correct[f] = cnt[f];
            }
        }
    }
}

int threshold_sweep_launch(const float* score, const int* label, const double* thresholds, int n, int T, int folds,
                           int* best_idx, double* best_thr, int* test_correct, int* train_correct,
                           cudaStream_t stream) {
    FFR_CHECK_ARG(T >= 1 && T <= 1024, "threshold_sweep: T=%d (max 1024)", T);
    FFR_CHECK_ARG(folds >= 1 && folds <= SWEEP_MAX_FOLDS, "threshold_sweep: folds=%d (max %d)", folds, SWEEP_MAX_FOLDS);
    FFR_CHECK_ARG(n >= folds, "threshold_sweep: n=%d < folds=%d", n, folds);
    threshold_sweep_kernel<<<1, 1024, 0, stream>>>(score, label, thresholds, n, T, folds, best_idx, best_thr,
                                                  test_correct, train_correct);
    return launch_status("threshold_sweep_kernel");
}

// ----------------------------------------------------------------------------------------------------------
// 1:N gallery scoring (generalises the paired scoring of lfw_eval.py:246-259 to a full similarity matrix):
// accept/reject counts of a (P x G) cosine matrix at every threshold of the grid, split into genuine (same identity)
// and impostor pairs. A score falls into bin b = number of thresholds t with (double)score > t (strict, exactly the
// comparison of eval_acc, lfw_eval.py:141-147; the grid is ascending so b comes from a binary search); the counts
// of "accepted at threshold t" are the suffix sums of the two histograms, taken on the host in integers.
// hist: [2][T+1] unsigned long long (genuine, impostor), zeroed by the launcher.
// ----------------------------------------------------------------------------------------------------------
constexpr int ROC_MAX_T = 1024;
__global__ void __launch_bounds__(256) roc_hist_kernel(const float* __restrict__ scores, int ld, int P, int G,
                                                       const int* __restrict__ probe_id, const int* __restrict__ gallery_id,
                                                       const double* __restrict__ thresholds, int T,
                                                       unsigned long long* __restrict__ hist) {
    __shared__ double s_thr[ROC_MAX_T];
    __shared__ unsigned int s_hist[2][ROC_MAX_T + 1];
    for (int i = threadIdx.x; i < T; i += 256) s_thr[i] = thresholds[i];
    for (int i = threadIdx.x; i < 2 * (ROC_MAX_T + 1); i += 256) (&s_hist[0][0])[i] = 0u;
    __syncthreads();
    for (int p = blockIdx.x; p < P; p += gridDim.x) {
        const int pid = probe_id[p];
        const float* row = scores + (long long)p * ld;
        for (int g = threadIdx.x; g < G; g += 256) {
            const double sc = (double)__ldg(row + g);
            int lo = 0, hi = T;                    // b = |{t : sc > thr[t]}| = first index with !(sc > thr[index])
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sc > s_thr[mid]) lo = mid + 1; else hi = mid;
            }
            atomicAdd(&s_hist[(gallery_id[g] == pid) ? 0 : 1][lo], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * (T + 1); i += 256) {
        const int k = i / (T + 1), b = i - k * (T + 1);
        const unsigned int c = s_hist[k][b];
        if (c) atomicAdd(hist + (long long)k * (T + 1) + b, (unsigned long long)c);
    }
}

int roc_hist_launch(const float* scores, int ld, int P, int G, const int* probe_id, const int* gallery_id,
                    const double* thresholds, int T, unsigned long long* hist, cudaStream_t stream) {
    FFR_CHECK_ARG(T >= 1 && T <= ROC_MAX_T, "roc_hist: T=%d (max %d)", T, ROC_MAX_T);
    FFR_CHECK_ARG(ld >= G && P >= 0 && G >= 0, "roc_hist: P=%d G=%d ld=%d", P, G, ld);
    FFR_CUDA(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * 2 * (size_t)(T + 1), stream));
    if (P == 0 || G == 0) return 0;
    const int grid = P < num_sms() * 4 ? P : num_sms() * 4;
    roc_hist_kernel<<<grid, 256, 0, stream>>>(scores, ld, P, G, probe_id, gallery_id, thresholds, T, hist);
    return launch_status("roc_hist_kernel");
}

}  // namespace ffr
