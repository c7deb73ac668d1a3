// On-device confusion matrices for label_accuracy_score (utils.py:104-154): the trainers call it after EVERY
// iteration on host numpy arrays, which forces a D2H copy of two (n,h,w) label maps and a numpy bincount
// (trainer_fcn.py:164,223,248).  Here the three histograms ("all", "seen", "unseen" targets of _fast_hist) are built
// in one pass over the device-resident labels; only n_class^2 counters travel to the host, where the four scores are
// formed exactly as _hist_to_metrics does.  (the step immediately after the hot path.)
#include "szn_internal.h"

namespace szn {

// hist[kind][t * n_class + p] += 1 for every pixel with 0 <= t < n_class (and 0 <= p < n_class);
// kind 0 = all, 1 = target class is seen, 2 = target class is unseen.  is_unseen: n_class flags or null (kind 0 only).
template <bool SMEM>
__global__ void __launch_bounds__(256) confusion_kernel(const long long* __restrict__ lt, const long long* __restrict__ lp,
                                                        long long n, int n_class, const unsigned char* __restrict__ is_unseen,
                                                        unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int sh[];  // [kinds][n_class^2] when SMEM
  const int kinds = is_unseen ? 3 : 1;
  const int cells = n_class * n_class;
  if (SMEM) {
    for (int i = threadIdx.x; i < kinds * cells; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long t = lt[i], q = lp[i];
    if (t < 0 || t >= n_class || q < 0 || q >= n_class) continue;
    const int cell = (int)t * n_class + (int)q;
    if (SMEM) {
      atomicAdd(&sh[cell], 1u);
      if (is_unseen) atomicAdd(&sh[(is_unseen[t] ? 2 : 1) * cells + cell], 1u);
    } else {
      atomicAdd(&hist[cell], 1ull);
      if (is_unseen) atomicAdd(&hist[(size_t)(is_unseen[t] ? 2 : 1) * cells + cell], 1ull);
    }
  }
  if (SMEM) {
    __syncthreads();
    for (int i = threadIdx.x; i < kinds * cells; i += blockDim.x)
      if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
  }
}

}  // namespace szn
using namespace szn;

// hist: int64 [kinds][n_class][n_class] (kinds = 3 when is_unseen is given, else 1), ACCUMULATED into (zero it first)
extern "C" int szn_confusion_hist(const long long* label_true, const long long* label_pred, long long n, int n_class,
                                  const unsigned char* is_unseen, long long* hist, void* stream) {
  if (n_class < 1) return set_error(SZN_ERR_ARG, "szn_confusion_hist: n_class");
  const int kinds = is_unseen ? 3 : 1;
  const size_t smem = (size_t)kinds * n_class * n_class * sizeof(unsigned int);
  long long blocks = (n + 256 * 16 - 1) / (256 * 16);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (smem <= 48 * 1024)
    confusion_kernel<true><<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(label_true, label_pred, n, n_class, is_unseen,
                                                                                  (unsigned long long*)hist);
  else
    confusion_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(label_true, label_pred, n, n_class, is_unseen,
                                                                                 (unsigned long long*)hist);
  return check_launch("szn_confusion_hist");
}
