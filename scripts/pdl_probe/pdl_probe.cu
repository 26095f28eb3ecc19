// Probe of programmatic dependent launch semantics on sm_100a (chain of dependent kernels, eager and captured).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void step_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int mode) {
  if (mode == 1) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
  if (mode == 2) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] + 1.f;
}
__global__ void step_ldg_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __ldg(in + i) + 1.f;
}
static void launch(bool pdl, bool ldg, const float* in, float* out, int n, int mode, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3((n + 255) / 256); cfg.blockDim = dim3(256); cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  if (ldg) cudaLaunchKernelEx(&cfg, step_ldg_kernel, in, out, n);
  else cudaLaunchKernelEx(&cfg, step_kernel, in, out, n, mode);
}
int main() {
  const int n = 192 * 256, K = 200;
  float *a, *b; cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4);
  float* h = new float[n];
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  for (int variant = 0; variant < 5; ++variant) {
    for (int use_default = 0; use_default < 2; ++use_default) {
      cudaStream_t s = use_default ? 0 : st;
      cudaMemset(a, 0, n * 4); cudaMemset(b, 0, n * 4); cudaDeviceSynchronize();
      const bool pdl = variant > 0; const bool ldg = variant == 4; const int mode = variant == 1 ? 1 : variant == 2 ? 2 : variant == 3 ? 0 : 1;
      for (int k = 0; k < K; ++k) launch(pdl, ldg, (k & 1) ? b : a, (k & 1) ? a : b, n, mode, s);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, a, n * 4, cudaMemcpyDeviceToHost);
      int bad = 0; for (int i = 0; i < n; ++i) if (h[i] != (float)K) ++bad;
      printf("variant %d (pdl %d mode %d ldg %d) stream %s: %s bad=%d first=%g\n", variant, pdl, mode, ldg, use_default ? "default" : "nonblocking",
             cudaGetErrorString(e), bad, h[0]);
    }
  }
  return 0;
}
