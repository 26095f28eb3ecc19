// Fused cluster tail (tail_fused.cu): the slide tail of MCAT (concat fusion) as three cluster kernels plus one
// grouped weight-gradient kernel instead of ~180 dependent launches.  tail.cu dispatches here when eligible().
#pragma once
#include <cuda_runtime.h>
#include "../../include/mpo_b200.h"
#include "tail_ws.h"

namespace mpo {
namespace fused {

enum : int { F_FWD = 1, F_LOSS = 2, F_BWD = 4, F_HEAD = 8 };   // F_HEAD: fusion + survival head (internal)

struct LossArgs {              // models/loss.py:5-43 through mpo_surv_loss's argument meaning
  int kind;                    // MPO_LOSS_NLL / MPO_LOSS_CES
  const int64_t* label;        // [B]
  const float* censor;         // [B]
  float alpha, eps, grad_scale;
  float* loss;                 // [B]
  float* dhaz; float* dS;      // [B][n_classes] (written for API parity with mpo_surv_loss; may be NULL)
};

// MCAT + concat fusion, omic widths multiples of 4, n_classes <= 8, not disabled through MPO_TAIL_FUSED=0
bool eligible(const mpo_model* m, const mpo_tail_io* io);

int pre_fwd(const mpo_model* m, const mpo_tail_io* io, const tailws::Ws& w, cudaStream_t st);
// flags: F_FWD (mcat.py:97-138), F_LOSS (needs F_FWD; `loss` non-null), F_BWD (gradients from the in-kernel loss, or
// from dhaz/dS/dY when F_LOSS is not set); F_BWD also launches the grouped weight-gradient kernel of the post stage
// side_wgrad: launch that kernel on an internal side stream (joined into `st` at the end of pre_bwd) so that it
// overlaps the bag backward pass; only for callers that always finish the step with pre_bwd (mpo_tail_post_step)
int post(const mpo_model* m, const mpo_tail_io* io, const tailws::Ws& w, int flags, const LossArgs* loss,
         const float* dhaz, const float* dS, const float* dY, cudaStream_t st, bool side_wgrad = false);
int pre_bwd(const mpo_model* m, const mpo_tail_io* io, const tailws::Ws& w, cudaStream_t st);
// Adam over a bucket completed by the pending side-stream weight gradients (see mpo_tail_side_adam)
int side_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
              float eps, float weight_decay, int32_t* step_dev, bool zero_grad, cudaStream_t st);

}  // namespace fused
}  // namespace mpo
