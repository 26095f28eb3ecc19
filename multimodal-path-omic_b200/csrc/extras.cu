// Small entry points around the returned attention map and the remaining loss / regulariser branches of the reference
// (SURVEY 8f N4): map dot products, the attention-norm regulariser of CrossEntropySurvivalAttnRegLoss
// (models/loss.py:88-101), SurvivalClassificationTobitLoss (models/loss.py:62-85) and l1_reg (models/utils.py:33-40).
#include <cstdint>
#include "../../include/mpo_b200.h"
#include "mpo_ptx.cuh"
#include "mpo_common.cuh"
#include "launchers.h"

namespace mpo {

// one block per tile, one thread per patch row: per-query partial dots, warp-reduced, then one atomic per warp and query
__global__ void __launch_bounds__(kTileM)
attn_map_dot_kernel(const TileInfo* __restrict__ tile_info, const float* __restrict__ amap, const float* __restrict__ other,
                    float* __restrict__ dot, int total_rows) {
  const TileInfo ti = tile_info[blockIdx.x];
  const int r = threadIdx.x;
  const bool valid = r < ti.nvalid;
#pragma unroll
  for (int i = 0; i < kQ; ++i) {
    const size_t o = static_cast<size_t>(i) * total_rows + ti.row0 + r;
    float v = valid ? amap[o] * other[o] : 0.f;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((r & 31) == 0 && v != 0.f) atomicAdd(dot + ti.slide * kQ + i, v);
  }
}

__global__ void __launch_bounds__(kTileM)
cesar_reg_kernel(const TileInfo* __restrict__ tile_info, const float* __restrict__ amap, const float* __restrict__ sumsq,
                 float lambda_reg, float grad_scale, float* __restrict__ reg, float* __restrict__ d_amap, int total_rows) {
  const TileInfo ti = tile_info[blockIdx.x];
  const int r = threadIdx.x;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kQ; ++i) ss += sumsq[ti.slide * kQ + i];
  const float norm = sqrtf(ss);
  if (ti.tile_in_slide == 0 && r == 0) reg[ti.slide] = lambda_reg * norm;
  if (r >= ti.nvalid) return;
  const float c = norm > 0.f ? grad_scale * lambda_reg / norm : 0.f;      // torch.norm backward: x / ||x||, 0 at the origin
#pragma unroll
  for (int i = 0; i < kQ; ++i) {
    const size_t o = static_cast<size_t>(i) * total_rows + ti.row0 + r;
    d_amap[o] = c * amap[o];
  }
}

__global__ void sct_loss_kernel(const float* __restrict__ Y, const int64_t* __restrict__ label, const float* __restrict__ censor,
                                float eps, float grad_scale, float* __restrict__ loss, float* __restrict__ dY, int B, int K) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int y = static_cast<int>(label[b]);
  for (int j = 0; j < K; ++j) dY[b * K + j] = 0.f;
  if (y < 0 || y >= K) { loss[b] = NAN; return; }       // out-of-range label: poison the loss, touch nothing else
  if (censor[b] == 0.f) {                               // loss.py:76-78: uncensored -> cross-entropy on the class
    const float pr = Y[b * K + y] + eps;
    loss[b] = -logf(pr);
    dY[b * K + y] = -grad_scale / pr;
  } else {                                              // loss.py:79-82: censored -> survival at least up to the label
    float cum = 0.f;
    for (int j = y; j < K; ++j) cum += Y[b * K + j];
    cum += eps;
    loss[b] = -logf(cum);
    for (int j = y; j < K; ++j) dY[b * K + j] = -grad_scale / cum;
  }
}

__global__ void __launch_bounds__(256)
l1_sum_kernel(const float* __restrict__ p, int64_t n, float* __restrict__ out) {
  float s = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    s += fabsf(p[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256)
l1_grad_kernel(const float* __restrict__ p, float* __restrict__ g, int64_t n, float scale) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = p[i];
    g[i] += v > 0.f ? scale : (v < 0.f ? -scale : 0.f);       // torch.abs backward: sign(x), 0 at 0
  }
}

}  // namespace mpo

using namespace mpo;

extern "C" {

int mpo_attn_map_dot(const mpo_bag* bag, const float* amap, const float* other, float* dot, void* stream) {
  if (!bag || !amap || !other || !dot) return fail(MPO_E_ARG, "%s", "mpo_attn_map_dot: null pointer");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_attn_map_dot: no CUDA device (this library has no CPU fallback)");
  if (bag->num_slides <= 0) return MPO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cuda(cudaMemsetAsync(dot, 0, sizeof(float) * kQ * bag->num_slides, st), "memset dot");
  if (rc || bag->num_tiles == 0) return rc;
  attn_map_dot_kernel<<<bag->num_tiles, kTileM, 0, st>>>(reinterpret_cast<const TileInfo*>(bag->tile_info), amap, other, dot,
                                                         static_cast<int>(bag->total_rows));
  count_launch();
  return check_cuda(cudaGetLastError(), "attn_map_dot_kernel");
}

int mpo_cesar_reg(const mpo_bag* bag, const float* amap, const float* sumsq, float lambda_reg, float grad_scale,
                  float* reg, float* d_amap, void* stream) {
  if (!bag || !amap || !sumsq || !reg || !d_amap) return fail(MPO_E_ARG, "%s", "mpo_cesar_reg: null pointer");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_cesar_reg: no CUDA device (this library has no CPU fallback)");
  if (bag->num_tiles == 0) return MPO_OK;
  cesar_reg_kernel<<<bag->num_tiles, kTileM, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const TileInfo*>(bag->tile_info), amap, sumsq, lambda_reg, grad_scale, reg, d_amap,
      static_cast<int>(bag->total_rows));
  count_launch();
  return check_cuda(cudaGetLastError(), "cesar_reg_kernel");
}

int mpo_sct_loss(const float* Y, const int64_t* label, const float* censor, float eps, float grad_scale, float* loss,
                 float* dY, int32_t B, int32_t n_classes, void* stream) {
  if (!Y || !label || !censor || !loss || !dY || B <= 0 || n_classes <= 0)
    return fail(MPO_E_ARG, "%s", "mpo_sct_loss: bad arguments");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_sct_loss: no CUDA device (this library has no CPU fallback)");
  sct_loss_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(Y, label, censor, eps, grad_scale, loss, dY,
                                                                                  B, n_classes);
  count_launch();
  return check_cuda(cudaGetLastError(), "sct_loss_kernel");
}

int mpo_l1_sum(const float* p, int64_t n, float* sum_out, void* stream) {
  if (!p || !sum_out || n < 0) return fail(MPO_E_ARG, "%s", "mpo_l1_sum: bad arguments");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_l1_sum: no CUDA device (this library has no CPU fallback)");
  if (n == 0) return MPO_OK;
  int blocks = static_cast<int>((n + 1023) / 1024);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  l1_sum_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, n, sum_out);
  count_launch();
  return check_cuda(cudaGetLastError(), "l1_sum_kernel");
}

int mpo_l1_grad(const float* p, float* grad, int64_t n, float scale, void* stream) {
  if (!p || !grad || n < 0) return fail(MPO_E_ARG, "%s", "mpo_l1_grad: bad arguments");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_l1_grad: no CUDA device (this library has no CPU fallback)");
  if (n == 0) return MPO_OK;
  int blocks = static_cast<int>((n + 1023) / 1024);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  l1_grad_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, grad, n, scale);
  count_launch();
  return check_cuda(cudaGetLastError(), "l1_grad_kernel");
}

}  // extern "C"
