// Gradient scaling of the fp16 deltas (shared by the tensor-core chains and the CUDA-core kernels).
#pragma once
#include <cuda_fp16.h>

namespace npp {

// Power-of-two gradient scale so that fp16 deltas sit in the middle of the half range:
// amax * scale == 2^10 (rounded down to a power of two).  Exact to undo in fp32.
__device__ __forceinline__ float npp_grad_scale(float amax) {
  if (!(amax > 0.0f) || !isfinite(amax)) return 1.0f;
  int e;
  frexpf(amax, &e);  // amax = m * 2^e, m in [0.5, 1)
  int k = 10 - e;
  k = max(-60, min(60, k));
  return ldexpf(1.0f, k);
}

// Power-of-two scale of the fp16 deltas of a fused train step.  It comes from the previous step's max |dL/dlogit|
// (relative to inv_count, so that a change of the batch size between steps does not matter): every CTA knows it
// before its head runs, which is what lets forward, head and backward of a stripe share one launch.  A power of two
// leaves the fp16 rounding of every delta unchanged, so the result is the one the same-step maximum would give
// unless the maximum moves by more than the 2^6 of headroom between two steps.  Unknown history (first step):
// the bound |dL/dlogit| <= inv_count / 2 (|d| <= 1, weight <= 1, yh (1 - yh) <= 1/4).
// A sudden jump of the maximum between two steps (a converged fit meeting an outlier batch) must not overflow fp16: the
// relative maximum is floored at 2^-6, so that even the largest possible gradient (0.5 relative) lands at 2^15 < 65504 at
// the head; a floor that high costs nothing (deltas of a converged fit then peak near 2^4 instead of 2^10, still ten
// binades above fp16's normal minimum).  Deeper layers, where deltas may grow, pack with saturation (pack_h2_sat).
__device__ __forceinline__ float npp_step_amax(const unsigned int* amax_prev, float inv_count) {
  float a = amax_prev != nullptr ? __uint_as_float(__ldcg(amax_prev)) : 0.f;
  if (!(a > 0.f) || !isfinite(a)) a = 0.5f;
  a = fmaxf(a, 0.015625f);
  return a * inv_count;
}

}  // namespace npp
