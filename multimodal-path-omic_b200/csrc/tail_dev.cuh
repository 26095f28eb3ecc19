// Device-side helpers shared by the per-op tail kernels (tail_kernels.cuh) and the fused cluster tail
// (tail_fused.cu): activations, the stateless dropout of the tail, programmatic-dependent-launch helpers.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "mpo_ptx.cuh"
#include "launchers.h"

namespace mpo {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_ELU = 2, ACT_TANH = 3, ACT_SIGMOID = 4 };

__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_ELU: return x > 0.f ? x : expm1f(x);
    case ACT_TANH: return tanhf(x);
    case ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
// ------------------------------------------------------------------------------------------------
// Train-mode dropout sites of the tail (SNN AlphaDropout mcat.py:38,42; encoder layers mcat.py:51-53; pooling heads
// blocks.py:34-36; rho mcat.py:57; bilinear fusion fusion.py:58-76).  Masks come from the stateless hash RNG keyed by
// (seed, site, element index): the backward pass regenerates them instead of storing them.
// ------------------------------------------------------------------------------------------------
// dropout sites of the tail (the bag stage uses sites 0 and 1)
enum : uint32_t { SITE_SNN = 16, SITE_ENC = 32, SITE_POOL = 48, SITE_RHO = 52, SITE_BIL = 56 };
struct DropSpec {
  uint32_t thr;               // drop when 8 random bits < thr ; 0 = no dropout at this site
  float scale;                // regular: 1/(1-p).  alpha: a
  float shift;                // alpha: b (regular: 0)
  uint32_t site;
  uint32_t seed;
  const uint32_t* seed_dev;   // optional device-side seed word (CUDA-graph replays)
  int alpha;                  // 1: nn.AlphaDropout
};
constexpr float kAlphaPrime = -1.7580993408473766f;      // -selu_lambda * selu_alpha
__device__ __forceinline__ uint32_t drop_seed(const DropSpec& d) {
  return d.seed_dev != nullptr ? (d.seed ^ __ldg(d.seed_dev)) : d.seed;
}
__device__ __forceinline__ bool drop_keep(const DropSpec& d, uint32_t seedv, uint32_t idx) {
  return (rng_u32(seedv, d.site, idx) & 0xFFu) >= d.thr;
}
// forward: value after the dropout layer
__device__ __forceinline__ float drop_fwd(float v, const DropSpec& d, uint32_t seedv, uint32_t idx) {
  const bool keep = drop_keep(d, seedv, idx);
  if (d.alpha) return (keep ? v : kAlphaPrime) * d.scale + d.shift;
  return keep ? v * d.scale : 0.f;
}
// backward: d(out)/d(in) of the dropout layer at this element
__device__ __forceinline__ float drop_grad(const DropSpec& d, uint32_t seedv, uint32_t idx) {
  return drop_keep(d, seedv, idx) ? d.scale : 0.f;
}
// value BEFORE the dropout layer, recovered from the stored (post-dropout) output of a kept element
__device__ __forceinline__ float drop_invert(float y, const DropSpec& d) {
  return d.alpha ? (y - d.shift) / d.scale : y / d.scale;
}

// derivative expressed through the activation OUTPUT y
__device__ __forceinline__ float act_bwd_from_out(float y, int act) {
  switch (act) {
    case ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case ACT_TANH: return 1.f - y * y;
    case ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: the tail is a chain of ~100 small dependent kernels, so the launch latency
// between them is the critical path.  Every tail kernel is launched with the programmatic-stream-serialization
// attribute and starts with pdl_enter(): wait until the predecessor grid has completed (its writes are visible),
// then let the successor be scheduled so that its launch -- and, for the GEMMs, the loads of its weight
// operand, which no kernel of the pass writes -- overlaps this kernel's execution.
// ------------------------------------------------------------------------------------------------
// (pdl_wait / pdl_launch_dependents / pdl_enter: mpo_ptx.cuh)

inline int& pdl_kind() { static int k = 0; return k; }   // bisect aid: 1 while a GEMM is being launched
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MPO_TAIL_PDL"); on = e ? atoi(e) : 1; }
  if (on == 2) return pdl_kind() >= 1;      // GEMMs only
  if (on == 3) return pdl_kind() == 0;      // everything but the GEMMs
  if (on == 4) return pdl_kind() == 2;      // full-K GEMMs only
  return on != 0;
}
// A stream that has just been made to wait on another stream's event launches its next kernel fully serialized:
// the programmatic relaxation is only meant for the kernel -> kernel edge inside one stream.
struct PdlBars { cudaStream_t s[8]; bool used[8]; };
inline PdlBars& pdl_bars() { static PdlBars b = {}; return b; }
inline void pdl_bar_next(cudaStream_t st) {       // (a null handle is the legacy default stream: a valid key)
  PdlBars& b = pdl_bars();
  for (int i = 0; i < 8; ++i) if (b.used[i] && b.s[i] == st) return;
  for (int i = 0; i < 8; ++i) if (!b.used[i]) { b.used[i] = true; b.s[i] = st; return; }
}
inline bool pdl_take_bar(cudaStream_t st) {
  PdlBars& b = pdl_bars();
  for (int i = 0; i < 8; ++i) if (b.used[i] && b.s[i] == st) { b.used[i] = false; return true; }
  return false;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl_enabled() && !pdl_take_bar(st)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// The kernels of the fused train step (bag kernels, cluster tail, Adam) form a chain of ~17 dependent launches per step;
// MPO_STEP_PDL=1 launches them with the same attribute (every one of them starts with pdl_enter()).  Measured neutral to
// slightly negative inside the step graph -- MCAT 1.074 / 1.076 ms with, 1.068 / 1.067 ms without, NaCAGaT 2.025 vs
// 2.035 ms, all 67 GPU tests green either way -- so it is OFF by default: the ~70 us between the step and the sum of its
// kernels' isolated times are not launch latency.
inline bool step_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MPO_STEP_PDL"); on = e ? atoi(e) : 0; }
  return on != 0;
}
// appends the attribute to `at` (returns the new count) unless the stream has just been made to wait on an event
inline int step_pdl_attr(cudaLaunchAttribute* at, int n, cudaStream_t st) {
  if (!step_pdl_enabled() || pdl_take_bar(st)) return n;
  at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[n].val.programmaticStreamSerializationAllowed = 1;
  return n + 1;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_step(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  cfg.attrs = at;
  cfg.numAttrs = step_pdl_attr(at, 0, st);
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace mpo
