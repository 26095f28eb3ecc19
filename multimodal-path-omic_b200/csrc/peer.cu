// Peer-memory collectives of the slide path over NVLink 5 / NVSwitch (one process per GPU).
//
// The two exchanges of this path are small and latency-bound, so they are written as kernels over peer-mapped memory
// (CUDA IPC: every rank maps every other rank's exchange buffer) instead of library collectives.  They are ordinary
// kernel launches, so a whole data-parallel step -- or a sharded inference call -- replays as ONE CUDA graph:
//
//  * mpo_peer_lse_combine : patch-range sharded bag (BASELINE config 5).  Each rank stores its (lse[6], pooled[6][256])
//    into a slot of every peer's exchange buffer (st over NVLink), signals, waits for the W - 1 other signals and merges
//    the W partial soft-max states locally: one launch, no gather + combine round trip.
//  * mpo_peer_adam_step   : data-parallel training.  Reduce-scatter + optimizer + all-gather as one sharded step: rank
//    r sums slice r of the flat fp32 gradient straight out of the peers' gradient buffers (ld over NVLink, fixed rank
//    order: bit-identical on every rank), applies Adam to slice r only (rank r touches 1/W of exp_avg / exp_avg_sq)
//    and stores the updated parameters into every peer's parameter buffer.  A signal/wait kernel in front
//    (gradients complete everywhere) and one behind (parameters delivered, gradients consumed) bracket it; the
//    gradient buffer is zeroed behind the second one.
//
// Synchronisation: flags[p][slot * 8 + r] lives in rank p's memory and is written only by rank r with a monotonically
// increasing epoch (st.release.sys); rank p spins on its own flags (ld.acquire.sys).  Epoch counters live on the device,
// so graph replays need no host involvement.  A spin that does not complete within ~2^31 clocks traps instead of hanging.
#include <cstdint>
#include <cstring>
#include "../../include/mpo_b200.h"
#include "mpo_ptx.cuh"
#include "mpo_common.cuh"
#include "launchers.h"

namespace mpo {

constexpr int kPeerMax = MPO_PEER_MAX;
constexpr int kStatFloats = kQ * (kD + 1);          // lse + pooled of one rank: 6 x 257
constexpr int kPeerSlots = 8;                       // independent barrier slots (epoch counters)

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {      // never served from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_f4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// signal every peer with this slot's next epoch, wait for every peer's signal of the same epoch (threads 0..W-1)
__device__ __forceinline__ void peer_signal_wait(const mpo_peer_group& g, int slot, uint32_t epoch) {
  const int t = threadIdx.x;
  if (t < g.world) {
    __threadfence_system();
    st_release_sys(g.flags[t] + slot * kPeerMax + g.rank, epoch);
    const uint32_t* mine = g.flags[g.rank] + slot * kPeerMax + t;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t0 > (1ll << 31)) __trap();      // a peer never arrived: fail loudly, do not hang the GPU
    }
  }
}

__global__ void peer_barrier_kernel(const mpo_peer_group g, int slot) {
  __shared__ uint32_t ep;
  if (threadIdx.x == 0) ep = ++g.epochs[slot];
  __syncthreads();
  peer_signal_wait(g, slot, ep);
}

// one block, 256 threads (thread = feature): publish the local soft-max state to every peer, barrier, merge
__global__ void __launch_bounds__(256)
peer_lse_combine_kernel(const mpo_peer_group g, int slot, const float* lse_local, const float* pooled_local,
                        float* lse_out, float* pooled_out) {      // outputs may alias the inputs (in-place merge)
  __shared__ uint32_t ep;
  const int d = threadIdx.x;
  if (d == 0) ep = ++g.epochs[slot];
  __syncthreads();
  const uint32_t epoch = ep;
  // two data buffers by epoch parity: a rank that runs ahead writes call k + 1 while a slow peer still reads call k
  const size_t base = static_cast<size_t>(epoch & 1u) * kPeerMax * kStatFloats + static_cast<size_t>(g.rank) * kStatFloats;
  for (int p = 0; p < g.world; ++p) {
    float* dst = static_cast<float*>(g.data[p]) + base;
#pragma unroll
    for (int i = 0; i < kQ; ++i) dst[kQ + i * kD + d] = pooled_local[i * kD + d];
    if (d < kQ) dst[d] = lse_local[d];
  }
  __syncthreads();                 // every thread's remote stores are issued before the signalling threads fence + release
  peer_signal_wait(g, slot, epoch);
  __syncthreads();
  const float* src = static_cast<const float*>(g.data[g.rank]) + static_cast<size_t>(epoch & 1u) * kPeerMax * kStatFloats;
#pragma unroll
  for (int i = 0; i < kQ; ++i) {
    float M = -INFINITY;
    for (int s = 0; s < g.world; ++s) M = fmaxf(M, __ldcg(src + s * kStatFloats + i));
    float L = 0.f, acc = 0.f;
    for (int s = 0; s < g.world; ++s) {
      const float w = __expf(__ldcg(src + s * kStatFloats + i) - M);      // exp(-inf) = 0: an empty shard drops out
      L += w;
      acc = fmaf(__ldcg(src + s * kStatFloats + kQ + i * kD + d), w, acc);
    }
    pooled_out[i * kD + d] = acc / L;
    if (d == 0) lse_out[i] = M + __logf(L);
  }
}

// slice r of [lo, hi): sum of every rank's gradients (fixed order), Adam (torch.optim.Adam with L2 weight decay, same
// arithmetic as adam_step_kernel in api_bag.cu), updated parameters stored into every rank's parameter buffer
__global__ void __launch_bounds__(256)
peer_adam_kernel(const mpo_peer_group g, int64_t s0, int64_t s1, int64_t m_off, float* __restrict__ exp_avg,
                 float* __restrict__ exp_avg_sq, float lr, float b1, float b2, float eps, float wd, float grad_scale,
                 const int32_t* __restrict__ step_dev) {
  const float t = static_cast<float>(*step_dev + 1);
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  for (int64_t i = s0 + (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < s1;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x * 4) {
    float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < g.world; ++p) {
      const float4 v = ld_peer_f4(static_cast<const float*>(g.grad[p]) + i);
      gs.x += v.x; gs.y += v.y; gs.z += v.z; gs.w += v.w;
    }
    float4 pv = *reinterpret_cast<const float4*>(static_cast<const float*>(g.param[g.rank]) + i);
    float4 mv = *reinterpret_cast<float4*>(exp_avg + (i - m_off)), vv = *reinterpret_cast<float4*>(exp_avg_sq + (i - m_off));
    float* pp = &pv.x; float* gp = &gs.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ge = gp[e] * grad_scale + wd * pp[e];
      mp[e] = mp[e] + (ge - mp[e]) * (1.f - b1);
      vp[e] = b2 * vp[e] + (1.f - b2) * ge * ge;
      pp[e] -= step_size * mp[e] / (sqrtf(vp[e]) / bc2_sqrt + eps);
    }
    *reinterpret_cast<float4*>(exp_avg + (i - m_off)) = mv;
    *reinterpret_cast<float4*>(exp_avg_sq + (i - m_off)) = vv;
    for (int p = 0; p < g.world; ++p) st_peer_f4(static_cast<float*>(g.param[p]) + i, pv);
  }
}

__global__ void peer_bump_step_kernel(int32_t* s) { *s += 1; }

static int check_group(const mpo_peer_group* g, const char* who) {
  if (!g) return fail(MPO_E_ARG, "%s: group is NULL", who);
  if (g->world < 1 || g->world > kPeerMax || g->rank < 0 || g->rank >= g->world) return fail(MPO_E_ARG, "%s: bad world/rank", who);
  if (!g->epochs) return fail(MPO_E_ARG, "%s: epoch counters are NULL", who);
  for (int p = 0; p < g->world; ++p)
    if (!g->flags[p] || !g->data[p]) return fail(MPO_E_ARG, "%s: unmapped peer buffer", who);
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s: no CUDA device (this library has no CPU fallback)", who);
  return MPO_OK;
}

}  // namespace mpo

using namespace mpo;

extern "C" {

int64_t mpo_peer_exchange_bytes(void) {
  // [2 parities][W_max ranks][6 x 257] floats of soft-max state, then [slots][W_max] flags
  return static_cast<int64_t>(2) * kPeerMax * kStatFloats * 4 + static_cast<int64_t>(kPeerSlots) * kPeerMax * 4;
}
int64_t mpo_peer_flags_offset(void) { return static_cast<int64_t>(2) * kPeerMax * kStatFloats * 4; }

int mpo_peer_alloc(int64_t bytes, void** ptr) {
  if (!ptr || bytes <= 0) return fail(MPO_E_ARG, "%s", "mpo_peer_alloc: bad arguments");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_peer_alloc: no CUDA device (this library has no CPU fallback)");
  int rc = check_cuda(cudaMalloc(ptr, static_cast<size_t>(bytes)), "cudaMalloc (peer buffer)");   // its own allocation: IPC offset 0
  if (rc) return rc;
  return check_cuda(cudaMemset(*ptr, 0, static_cast<size_t>(bytes)), "cudaMemset (peer buffer)");
}
int mpo_peer_free(void* ptr) { return ptr ? check_cuda(cudaFree(ptr), "cudaFree (peer buffer)") : MPO_OK; }

int mpo_peer_export(const void* ptr, void* handle_out) {
  if (!ptr || !handle_out) return fail(MPO_E_ARG, "%s", "mpo_peer_export: null pointer");
  cudaIpcMemHandle_t h;
  int rc = check_cuda(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)), "cudaIpcGetMemHandle");
  if (rc) return rc;
  static_assert(sizeof(h) == MPO_PEER_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  return MPO_OK;
}
int mpo_peer_open(const void* handle, void** ptr_out) {
  if (!handle || !ptr_out) return fail(MPO_E_ARG, "%s", "mpo_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  return check_cuda(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}
int mpo_peer_close(void* ptr) { return ptr ? check_cuda(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle") : MPO_OK; }

int mpo_peer_warmup(void) {
  // loads the kernels of this file (lazy module loading must not happen for the first time inside a stream capture)
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_peer_warmup: no CUDA device (this library has no CPU fallback)");
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, peer_barrier_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, peer_lse_combine_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, peer_adam_kernel);
  if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, peer_bump_step_kernel);
  return check_cuda(e, "mpo_peer_warmup");
}

int mpo_peer_barrier(const mpo_peer_group* g, int32_t slot, void* stream) {
  int rc = check_group(g, "mpo_peer_barrier");
  if (rc) return rc;
  if (slot < 0 || slot >= kPeerSlots) return fail(MPO_E_ARG, "%s", "mpo_peer_barrier: slot out of range");
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(*g, slot);
  count_launch();
  return check_cuda(cudaGetLastError(), "peer_barrier_kernel");
}

int mpo_peer_lse_combine(const mpo_peer_group* g, int32_t slot, const float* lse_local, const float* pooled_local,
                         float* lse_out, float* pooled_out, void* stream) {
  int rc = check_group(g, "mpo_peer_lse_combine");
  if (rc) return rc;
  if (slot < 0 || slot >= kPeerSlots || !lse_local || !pooled_local || !lse_out || !pooled_out)
    return fail(MPO_E_ARG, "%s", "mpo_peer_lse_combine: bad arguments");
  peer_lse_combine_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(*g, slot, lse_local, pooled_local, lse_out, pooled_out);
  count_launch();
  return check_cuda(cudaGetLastError(), "peer_lse_combine_kernel");
}

int mpo_peer_adam_step(const mpo_peer_group* g, int32_t slot, int64_t lo, int64_t hi, float* exp_avg, float* exp_avg_sq,
                       int64_t state_offset, float lr, float beta1, float beta2, float eps, float weight_decay,
                       float grad_scale, int32_t* step_dev, int32_t bump_step, void* stream) {
  int rc = check_group(g, "mpo_peer_adam_step");
  if (rc) return rc;
  if (slot < 0 || slot + 1 >= kPeerSlots || lo < 0 || hi < lo || (lo & 3) || (hi & 3) || !exp_avg || !exp_avg_sq || !step_dev)
    return fail(MPO_E_ARG, "%s", "mpo_peer_adam_step: bad arguments (lo / hi must be multiples of 4)");
  for (int p = 0; p < g->world; ++p)
    if (!g->grad[p] || !g->param[p]) return fail(MPO_E_ARG, "%s", "mpo_peer_adam_step: unmapped gradient / parameter buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t s0, s1;
  mpo_peer_slice(lo, hi, g->world, g->rank, &s0, &s1);
  peer_barrier_kernel<<<1, 32, 0, st>>>(*g, slot);                 // every rank's gradients of [lo, hi) are complete
  if (s1 > s0) {
    const int64_t groups = (s1 - s0) / 4;
    int blocks = static_cast<int>((groups + 255) / 256);
    if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
    peer_adam_kernel<<<blocks, 256, 0, st>>>(*g, s0, s1, state_offset, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay,
                                             grad_scale, step_dev);
  }
  peer_barrier_kernel<<<1, 32, 0, st>>>(*g, slot + 1);             // parameters delivered everywhere, gradients consumed
  rc = check_cuda(cudaMemsetAsync(static_cast<float*>(g->grad[g->rank]) + lo, 0, static_cast<size_t>(hi - lo) * 4, st),
                  "memset gradients");
  if (rc) return rc;
  if (bump_step) peer_bump_step_kernel<<<1, 1, 0, st>>>(step_dev);
  count_launch(3 + (bump_step ? 1 : 0));
  return check_cuda(cudaGetLastError(), "peer_adam_kernel");
}

void mpo_peer_slice(int64_t lo, int64_t hi, int32_t world, int32_t rank, int64_t* s0, int64_t* s1) {
  // contiguous slices of [lo, hi) in units of 4 floats, the first (n4 % world) ranks one unit longer
  const int64_t n4 = (hi - lo) / 4;
  const int64_t base = n4 / world, rem = n4 % world;
  const int64_t a = rank * base + (rank < rem ? rank : rem);
  const int64_t b = a + base + (rank < rem ? 1 : 0);
  *s0 = lo + 4 * a;
  *s1 = lo + 4 * b;
}

}  // extern "C"
