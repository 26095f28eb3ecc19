// fp32 building blocks of the slide "tail": everything that runs on the 6 omic tokens per slide
// (SNN encoders, query projection, transformer encoders, gated attention pooling, fusion, survival head, losses).
// All kernels are batched over the B slides of a step: token rows are laid out [slide][token] (R = 6B rows of 256).
// The work is latency-bound: plain CUDA-core kernels, fp32 throughout.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "mpo_ptx.cuh"
#include "launchers.h"
#include "tail_dev.cuh"

namespace mpo {

// ------------------------------------------------------------------------------------------------
// generic strided SIMT GEMM:  C[m][n] (+)= act(alpha * sum_k A(m,k) B(k,n) + bias[n])
//   A(m,k) = A[m*sa_m + k*sa_k],  B(k,n) = B[k*sb_k + n*sb_n];  64x64x16 tiles, 256 threads, 4x4 per thread
// ------------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A; long long sa_m, sa_k;
  const float* B; long long sb_k, sb_n;
  float* C; long long ldc;
  const float* bias;
  int M, N, K;
  float alpha;
  int accumulate;
  int act;
  float* rowsum;      // optional: rowsum[m] += alpha * sum_k A(m,k)   (fused bias gradient)
  int b_static;       // 1: B is a parameter (never written inside the pass): its loads may run ahead of the PDL wait
  DropSpec drop;      // dropout on the output, after the activation (element index m * N + n)
  const float* addend; long long ld_add;    // optional: + addend[m][n] (a residual-branch gradient), before `act`
  const float* bwd_y; long long ld_bwd;     // optional backward epilogue: * act'(y[m][n]) of activation `bwd_act`, where
  int bwd_act;                              //   y was stored after the dropout layer `bwd_drop` (also undone here)
  DropSpec bwd_drop;
};

// The tail is a long chain of small dependent GEMMs, so the kernel is built for latency: 32-deep K steps,
// global loads of step k+1 issued into registers before the math of step k (double-buffered shared memory, one
// barrier per step), and a narrow 32-row tile variant so that few-row problems still spread over many SMs.
// Optional fused bias gradient: rowsum[m] += sum_k A(m,k) (used by the weight-gradient GEMMs, where A = dz^T).
template <int BM, bool A_KC, bool B_NC>
__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs g) {
  pdl_enter();
  constexpr int BN = 64, BK = 32;
  constexpr int TM = BM / 16;                 // rows per thread (4 or 2)
  constexpr int A_PER = BM * BK / 256;        // elements of the A tile each thread stages (8 or 4)
  constexpr int B_PER = BN * BK / 256;        // 8
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = t & 15, ty = t >> 4;
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[A_PER], rb[B_PER];
  float rsum = 0.f;                            // fused row sum of A (threads t < BM own row m0 + t)

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      int mm, kk;
      if (A_KC) { kk = t & 31; mm = (t >> 5) + 8 * j; } else { mm = t % BM; kk = t / BM + (256 / BM) * j; }
      const int m = m0 + mm, k = k0 + kk;
      ra[j] = (m < g.M && k < g.K) ? __ldg(g.A + m * g.sa_m + k * g.sa_k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      int nn, kk;
      if (B_NC) { nn = t & 63; kk = (t >> 6) + 4 * j; } else { kk = t & 31; nn = (t >> 5) + 8 * j; }
      const int n = n0 + nn, k = k0 + kk;
      rb[j] = (n < g.N && k < g.K) ? __ldg(g.B + k * g.sb_k + n * g.sb_n) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      int mm, kk;
      if (A_KC) { kk = t & 31; mm = (t >> 5) + 8 * j; } else { mm = t % BM; kk = t / BM + (256 / BM) * j; }
      As[buf][kk][mm] = ra[j];
    }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      int nn, kk;
      if (B_NC) { nn = t & 63; kk = (t >> 6) + 4 * j; } else { kk = t & 31; nn = (t >> 5) + 8 * j; }
      Bs[buf][kk][nn] = rb[j];
    }
  };

  const int nsteps = (g.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    if (s + 1 < nsteps) load_tile((s + 1) * BK);       // in flight during the math below
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM];
      if (TM == 4) {
        const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
        av[0] = a.x; av[1] = a.y; av[TM > 2 ? 2 : 0] = a.z; av[TM > 2 ? 3 : 1] = a.w;
      } else {
        const float2 a = *reinterpret_cast<const float2*>(&As[buf][kk][ty * 2]);
        av[0] = a.x; av[1] = a.y;
      }
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (g.rowsum != nullptr && blockIdx.x == 0 && t < BM) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) rsum += As[buf][kk][t];
    }
    if (s + 1 < nsteps) store_tile(buf ^ 1);
    __syncthreads();
  }
  if (g.rowsum != nullptr && blockIdx.x == 0 && t < BM && m0 + t < g.M) g.rowsum[m0 + t] += g.alpha * rsum;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias != nullptr) v += g.bias[n];
      if (g.addend != nullptr) v += g.addend[m * g.ld_add + n];
      if (g.bwd_y != nullptr) {
        float yv = g.bwd_y[m * g.ld_bwd + n];
        if (g.bwd_drop.thr != 0) {
          v *= drop_grad(g.bwd_drop, drop_seed(g.bwd_drop), static_cast<uint32_t>(m) * static_cast<uint32_t>(g.N) + n);
          yv = drop_invert(yv, g.bwd_drop);
        }
        v *= act_bwd_from_out(yv, g.bwd_act);
      }
      v = act_fwd(v, g.act);
      if (g.drop.thr != 0) v = drop_fwd(v, g.drop, drop_seed(g.drop), static_cast<uint32_t>(m) * static_cast<uint32_t>(g.N) + n);
      float* c = g.C + m * g.ldc + n;
      *c = g.accumulate ? (*c + v) : v;
    }
  }
}

template <int BM>
inline void launch_gemm_bm(const GemmArgs& g, cudaStream_t st) {
  dim3 grid((g.N + 63) / 64, (g.M + BM - 1) / BM);
  const bool akc = (g.sa_k == 1), bnc = (g.sb_n == 1);
  if (akc && bnc) launch_k(gemm_kernel<BM, true, true>, dim3(grid), dim3(256), 0, st, g);
  else if (akc && !bnc) launch_k(gemm_kernel<BM, true, false>, dim3(grid), dim3(256), 0, st, g);
  else if (!akc && bnc) launch_k(gemm_kernel<BM, false, true>, dim3(grid), dim3(256), 0, st, g);
  else launch_k(gemm_kernel<BM, false, false>, dim3(grid), dim3(256), 0, st, g);
}

// Latency-first variant for the chain of small dependent GEMMs: a 32x32 output tile per CTA, and the WHOLE K
// extent of the tile (in chunks of 256) fetched with one wave of independent loads before any math, so a GEMM costs
// one L2 round trip instead of one per 32-deep K step.  Same contract as gemm_kernel.
constexpr int kFkBM = 32, kFkBN = 32, kFkBK = 256;
template <bool A_KC, bool B_NC>
__global__ void __launch_bounds__(256) gemm_fullk_kernel(const GemmArgs g) {
  extern __shared__ __align__(16) float fk_smem[];
  float (*As)[kFkBM + 2] = reinterpret_cast<float (*)[kFkBM + 2]>(fk_smem);
  float (*Bs)[kFkBN + 2] = reinterpret_cast<float (*)[kFkBN + 2]>(fk_smem + kFkBK * (kFkBM + 2));
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * kFkBM, n0 = blockIdx.x * kFkBN;
  const int tx = t & 15, ty = t >> 4;          // thread -> rows ty*2..+1, cols tx*2..+1
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float rsum = 0.f;
  constexpr int PER = kFkBM * kFkBK / 256;     // 32 elements of each operand tile per thread and chunk
  for (int k0 = 0; k0 < g.K; k0 += kFkBK) {
    float ra[PER], rb[PER];
    if (k0 == 0 && !g.b_static) pdl_enter();      // B may come from the preceding kernel: nothing is read before the wait
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      int nn, kk;
      if (B_NC) { nn = t & 31; kk = (t >> 5) + 8 * j; } else { kk = (t & 31) + 32 * (j & 7); nn = (t >> 5) + 8 * (j >> 3); }
      const int n = n0 + nn, k = k0 + kk;
      rb[j] = 0.f;
      if (n < g.N && k < g.K) {
        if (g.b_static) rb[j] = __ldg(g.B + k * g.sb_k + n * g.sb_n);
        else asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(rb[j]) : "l"(g.B + k * g.sb_k + n * g.sb_n));
      }
    }
    // when B is a parameter (b_static) its loads (above) run ahead of the grid-dependency wait.  A is loaded with
    // volatile asm: __ldg loads are "pure" to the compiler and would be scheduled above the wait together with B's.
    if (k0 == 0 && g.b_static) pdl_enter();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      int mm, kk;
      if (A_KC) { kk = (t & 31) + 32 * (j & 7); mm = (t >> 5) + 8 * (j >> 3); } else { mm = t & 31; kk = (t >> 5) + 8 * j; }
      const int m = m0 + mm, k = k0 + kk;
      ra[j] = 0.f;
      if (m < g.M && k < g.K) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(ra[j]) : "l"(g.A + m * g.sa_m + k * g.sa_k));
    }
    if (k0 > 0) __syncthreads();               // previous chunk's math is done with the tiles
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      int mm, kk;
      if (A_KC) { kk = (t & 31) + 32 * (j & 7); mm = (t >> 5) + 8 * (j >> 3); } else { mm = t & 31; kk = (t >> 5) + 8 * j; }
      As[kk][mm] = ra[j];
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      int nn, kk;
      if (B_NC) { nn = t & 31; kk = (t >> 5) + 8 * j; } else { kk = (t & 31) + 32 * (j & 7); nn = (t >> 5) + 8 * (j >> 3); }
      Bs[kk][nn] = rb[j];
    }
    __syncthreads();
    const int kmax = min(kFkBK, g.K - k0);
#pragma unroll 8
    for (int kk = 0; kk < kmax; ++kk) {
      const float2 a = *reinterpret_cast<const float2*>(&As[kk][ty * 2]);
      const float2 b = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
      acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
      acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
    }
    if (g.rowsum != nullptr && blockIdx.x == 0 && t < kFkBM) {
      for (int kk = 0; kk < kmax; ++kk) rsum += As[kk][t];
    }
  }
  if (g.rowsum != nullptr && blockIdx.x == 0 && t < kFkBM && m0 + t < g.M) g.rowsum[m0 + t] += g.alpha * rsum;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + ty * 2 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = n0 + tx * 2 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias != nullptr) v += g.bias[n];
      if (g.addend != nullptr) v += g.addend[m * g.ld_add + n];
      if (g.bwd_y != nullptr) {
        float yv = g.bwd_y[m * g.ld_bwd + n];
        if (g.bwd_drop.thr != 0) {
          v *= drop_grad(g.bwd_drop, drop_seed(g.bwd_drop), static_cast<uint32_t>(m) * static_cast<uint32_t>(g.N) + n);
          yv = drop_invert(yv, g.bwd_drop);
        }
        v *= act_bwd_from_out(yv, g.bwd_act);
      }
      v = act_fwd(v, g.act);
      if (g.drop.thr != 0) v = drop_fwd(v, g.drop, drop_seed(g.drop), static_cast<uint32_t>(m) * static_cast<uint32_t>(g.N) + n);
      float* c = g.C + m * g.ldc + n;
      *c = g.accumulate ? (*c + v) : v;
    }
  }
}
inline void launch_gemm_fullk(const GemmArgs& g, cudaStream_t st) {
  dim3 grid((g.N + kFkBN - 1) / kFkBN, (g.M + kFkBM - 1) / kFkBM);
  const bool akc = (g.sa_k == 1), bnc = (g.sb_n == 1);
  constexpr int smem = kFkBK * (kFkBM + 2 + kFkBN + 2) * 4;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_fullk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gemm_fullk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gemm_fullk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(gemm_fullk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = true;
  }
  if (akc && bnc) launch_k(gemm_fullk_kernel<true, true>, dim3(grid), dim3(256), smem, st, g);
  else if (akc && !bnc) launch_k(gemm_fullk_kernel<true, false>, dim3(grid), dim3(256), smem, st, g);
  else if (!akc && bnc) launch_k(gemm_fullk_kernel<false, true>, dim3(grid), dim3(256), smem, st, g);
  else launch_k(gemm_fullk_kernel<false, false>, dim3(grid), dim3(256), smem, st, g);
}

// ------------------------------------------------------------------------------------------------
// Throughput variant for the bag-scale fp32 GEMMs of GE-NaCAGaT (N x N attention over the patches): 128 x 128 tiles
// (or 256 x 32 when the output is only a head wide), 16-deep K steps, 8 x 8 (8 x 4) outputs per thread accumulated
// with packed FFMA2 (fma.rn.f32x2: two fp32 FMAs per instruction), double-buffered shared memory, global loads of
// step k + 1 in flight during the math of step k.  Same contract as gemm_kernel for the epilogues it supports
// (alpha, bias, accumulate, activation).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gb_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long gb_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <int BN, bool A_KC, bool B_NC>
__global__ void __launch_bounds__(256) gemm_big_kernel(const GemmArgs g) {
  constexpr int BK = 16;
  constexpr int TX = BN == 128 ? 16 : 8;        // thread columns
  constexpr int TY = 256 / TX;                  // thread rows (16 or 32)
  constexpr int BM = TY * 8;                    // 128 or 256
  constexpr int CN = BN / TX;                   // output columns per thread: 8 (two float4 halves) or 4
  constexpr int A_PER = BM * BK / 256, B_PER = BN * BK / 256;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int t = threadIdx.x, tx = t % TX, ty = t / TX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  unsigned long long acc[8][CN / 2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CN / 2; ++j) acc[i][j] = 0ull;
  float ra[A_PER], rb[B_PER];
  auto a_pos = [&](int j, int& mm, int& kk) {
    if (A_KC) { kk = t & 15; mm = (t >> 4) + 16 * j; } else { mm = t % BM; kk = t / BM + (256 / BM) * j; }
  };
  auto b_pos = [&](int j, int& nn, int& kk) {
    if (B_NC) { nn = t % BN; kk = t / BN + (256 / BN) * j; } else { kk = t & 15; nn = (t >> 4) + 16 * j; }
  };
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      int mm, kk; a_pos(j, mm, kk);
      const long long m = m0 + mm, k = k0 + kk;
      ra[j] = (m < g.M && k < g.K) ? __ldg(g.A + m * g.sa_m + k * g.sa_k) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      int nn, kk; b_pos(j, nn, kk);
      const long long n = n0 + nn, k = k0 + kk;
      rb[j] = (n < g.N && k < g.K) ? __ldg(g.B + k * g.sb_k + n * g.sb_n) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int j = 0; j < A_PER; ++j) { int mm, kk; a_pos(j, mm, kk); As[buf][kk][mm] = ra[j]; }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) { int nn, kk; b_pos(j, nn, kk); Bs[buf][kk][nn] = rb[j]; }
  };
  pdl_enter();
  // split-K (gridDim.z > 1): this CTA reduces steps [s_begin, s_end) and adds its partial with atomics (C pre-zeroed)
  const int nsteps_all = (g.K + BK - 1) / BK;
  const int per_z = (nsteps_all + static_cast<int>(gridDim.z) - 1) / static_cast<int>(gridDim.z);
  const int s_begin = static_cast<int>(blockIdx.z) * per_z, s_end = min(nsteps_all, s_begin + per_z);
  if (s_begin >= s_end) return;
  const int nsteps = s_end - s_begin, kbase = s_begin * BK;
  load_tile(kbase);
  store_tile(0);
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    if (s + 1 < nsteps) load_tile(kbase + (s + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      // rows ty*4 .. +3 and BM/2 + ty*4 .. +3; columns tx*4 .. +3 (and BN/2 + tx*4 .. +3): conflict-free float4 reads
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][BM / 2 + ty * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      unsigned long long bv[CN / 2];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        bv[0] = gb_pack2(b0.x, b0.y); bv[1] = gb_pack2(b0.z, b0.w);
        if (CN == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][BN / 2 + tx * 4]);
          bv[CN / 2 - 2] = gb_pack2(b1.x, b1.y); bv[CN / 2 - 1] = gb_pack2(b1.z, b1.w);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const unsigned long long aa = gb_pack2(av[i], av[i]);
#pragma unroll
        for (int j = 0; j < CN / 2; ++j) acc[i][j] = gb_fma2(aa, bv[j], acc[i][j]);
      }
    }
    if (s + 1 < nsteps) store_tile(buf ^ 1);
    __syncthreads();
  }
  // epilogue: each thread owns groups of 4 consecutive columns -> one 16-byte store per row and group when aligned
  const bool vec_ok = (g.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int grp = 0; grp < CN / 4; ++grp) {
      const long long nb = n0 + (grp == 0 ? tx * 4 : BN / 2 + tx * 4);
      float v[4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i][grp * 2 + j]));
        v[2 * j] = lo; v[2 * j + 1] = hi;
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        v[h] *= g.alpha;
        if (g.bias != nullptr && nb + h < g.N) v[h] += g.bias[nb + h];
        v[h] = act_fwd(v[h], g.act);
      }
      float* c = g.C + m * g.ldc + nb;
      if (gridDim.z > 1) {
#pragma unroll
        for (int h = 0; h < 4; ++h)
          if (nb + h < g.N) atomicAdd(c + h, v[h]);
      } else if (vec_ok && nb + 3 < g.N) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (g.accumulate) { const float4 p = *reinterpret_cast<const float4*>(c); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
        *reinterpret_cast<float4*>(c) = o;
      } else {
#pragma unroll
        for (int h = 0; h < 4; ++h)
          if (nb + h < g.N) c[h] = g.accumulate ? (c[h] + v[h]) : v[h];
      }
    }
  }
}
// C[m][n] = 0 over an [M, N] window of a strided matrix (split-K GEMMs accumulate into it with atomics)
__global__ void zero_window_kernel(float* __restrict__ C, long long ldc, int M, int N) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < static_cast<long long>(M) * N) C[(i / N) * ldc + (i % N)] = 0.f;
}
template <int BN>
inline void launch_gemm_big(const GemmArgs& g, cudaStream_t st) {
  constexpr int BM = BN == 128 ? 128 : 256;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
  // tall-skinny outputs with a bag-long reduction (P V, P^T dctx, dS K, dS^T Q of GE-NaCAGaT): too few tiles to fill
  // the GPU and thousands of dependent K steps each -> split K
  if (g.K >= 512 && g.bias == nullptr && g.act == ACT_NONE && grid.x * grid.y < 296) {
    int splits = g.K / 256;
    const int cap = static_cast<int>(296u / (grid.x * grid.y)) > 16 ? static_cast<int>(296u / (grid.x * grid.y)) : 16;
    if (splits > cap) splits = cap;
    if (splits > 1) {
      grid.z = splits;
      if (!g.accumulate) {
        const long long n = static_cast<long long>(g.M) * g.N;
        launch_k(zero_window_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, g.C, g.ldc, g.M, g.N);
        count_launch();
      }
    }
  }
  const bool akc = (g.sa_k == 1), bnc = (g.sb_n == 1);
  if (akc && bnc) launch_k(gemm_big_kernel<BN, true, true>, dim3(grid), dim3(256), 0, st, g);
  else if (akc && !bnc) launch_k(gemm_big_kernel<BN, true, false>, dim3(grid), dim3(256), 0, st, g);
  else if (!akc && bnc) launch_k(gemm_big_kernel<BN, false, true>, dim3(grid), dim3(256), 0, st, g);
  else launch_k(gemm_big_kernel<BN, false, false>, dim3(grid), dim3(256), 0, st, g);
}

inline cudaError_t launch_gemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("MPO_TAIL_GEMM"); mode = e ? atoi(e) : 1; }
  const long long blocks32 = static_cast<long long>((g.N + 31) / 32) * ((g.M + 31) / 32);
  pdl_kind() = 1;
  const bool plain = g.rowsum == nullptr && g.drop.thr == 0 && g.addend == nullptr && g.bwd_y == nullptr;
  static int big = -1;
  if (big < 0) { const char* e = getenv("MPO_GEMM_BIG"); big = e ? atoi(e) : 1; }
  if (big && plain && g.M >= 256 && static_cast<long long>(g.M) * g.N * g.K >= (1LL << 24)) {
    if (g.N >= 96) launch_gemm_big<128>(g, st); else launch_gemm_big<32>(g, st);      // bag-scale GEMMs (GE-NaCAGaT)
  } else if (mode == 1 && blocks32 <= 1184 && g.K <= 4 * kFkBK) {
    pdl_kind() = 2;
    launch_gemm_fullk(g, st);                  // the latency-bound regime of the slide tail
  } else {
    // narrow tiles when 64-row tiles would leave most SMs idle
    const long long blocks64 = static_cast<long long>((g.N + 63) / 64) * ((g.M + 63) / 64);
    if (blocks64 < 96) launch_gemm_bm<32>(g, st);
    else launch_gemm_bm<64>(g, st);
  }
  pdl_kind() = 0;
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------------
// dz[r][c] = dy[r][c] * act'(y[r][c])     (row strides allow views into wider buffers)
__global__ void act_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ y,
                               long long ldy, float* __restrict__ dz, long long lddz, int rows, int cols, int act,
                               const DropSpec drop) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * cols) return;
  const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
  float yv = y[r * ldy + c];
  float g = dy[r * lddy + c];
  if (drop.thr != 0) {     // y is stored after the dropout layer: undo it for the activation derivative
    g *= drop_grad(drop, drop_seed(drop), static_cast<uint32_t>(i));
    yv = drop_invert(yv, drop);
  }
  dz[r * lddz + c] = g * act_bwd_from_out(yv, act);
}

// out[r][c] = a[r][c] + b[r][c]
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                           long long n) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                           long long n) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * b[i];
}
// y = act(x) elementwise, in or out of place
__global__ void act_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int act) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = act_fwd(x[i], act);
}
__global__ void fill_kernel(float* __restrict__ x, long long n, float v) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

// g[c] += sum_r x[r][c]   (and optionally of x*y): bias / LayerNorm parameter gradients.
// block = 32 columns x 8 row groups; rows are strided over the groups, then reduced through shared memory.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ y, long long ldy,
              float* __restrict__ g, int rows, int cols) {
  pdl_enter();
  __shared__ float part[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  // gridDim.y > 1: the rows are split over blockIdx.y as well (bag-long reductions) and the sums are added atomically
  const int rstep = 8 * static_cast<int>(gridDim.y);
  if (c < cols) {
    if (y == nullptr) {
#pragma unroll 4
      for (int r = ry + 8 * static_cast<int>(blockIdx.y); r < rows; r += rstep) s += x[r * ldx + c];
    } else {
#pragma unroll 4
      for (int r = ry + 8 * static_cast<int>(blockIdx.y); r < rows; r += rstep) s = fmaf(x[r * ldx + c], y[r * ldy + c], s);
    }
  }
  part[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < cols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += part[i][cx];
    if (gridDim.y > 1) atomicAdd(&g[c], v);
    else g[c] += v;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over 256 features, one warp per row:  y = LN(a + b) * gamma + beta   (b may be null)
// saves xhat and rstd for the backward pass           (torch.nn.LayerNorm, eps 1e-5, biased variance)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ xhat,
                     float* __restrict__ rstd, int rows) {
  pdl_enter();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane + 32 * j;
    v[j] = a[row * 256 + c] + (b ? b[row * 256 + c] : 0.f);
    s += v[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { const float d = v[j] - mu; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rs = rsqrtf(q * (1.f / 256.f) + 1e-5f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane + 32 * j;
    const float xh = (v[j] - mu) * rs;
    xhat[row * 256 + c] = xh;
    y[row * 256 + c] = fmaf(xh, gamma[c], beta[c]);
  }
  if (lane == 0) rstd[row] = rs;
}

// dx = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)),  dxh = dy * gamma;  also writes t = dy * xhat
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ gamma, const float* __restrict__ xhat,
                     const float* __restrict__ rstd, float* __restrict__ dx, int rows, float* __restrict__ dx_drop,
                     const DropSpec drop) {   // dx_drop (optional) = dx through the dropout layer in front of the residual
  pdl_enter();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float dxh[8], xh[8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane + 32 * j;
    xh[j] = xhat[row * 256 + c];
    dxh[j] = dy[row * 256 + c] * gamma[c];
    s1 += dxh[j];
    s2 = fmaf(dxh[j], xh[j], s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  s1 *= (1.f / 256.f);
  s2 *= (1.f / 256.f);
  const float rs = rstd[row];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int idx = row * 256 + lane + 32 * j;
    const float v = rs * (dxh[j] - s1 - xh[j] * s2);
    dx[idx] = v;
    if (dx_drop != nullptr) dx_drop[idx] = v * drop_grad(drop, drop_seed(drop), static_cast<uint32_t>(idx));
  }
}

// ------------------------------------------------------------------------------------------------
// 8-head self-attention over the 6 tokens of a slide (nn.TransformerEncoderLayer's self_attn, head_dim 32)
// one warp per (slide, head); lane = head-dim index.  qkv rows: [q(256) | k(256) | v(256)]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mha6_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ probs, float* __restrict__ ctx, int B,
                const DropSpec drop) {
  pdl_enter();
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= B * 8) return;
  const int b = w >> 3, h = w & 7;
  float q[6], k[6], v[6];
#pragma unroll
  for (int l = 0; l < 6; ++l) {
    const float* row = qkv + static_cast<size_t>(b * 6 + l) * 768 + h * 32 + lane;
    q[l] = row[0]; k[l] = row[256]; v[l] = row[512];
  }
  const float scale = 0.17677669529663687f;   // 1/sqrt(32)
#pragma unroll
  for (int l1 = 0; l1 < 6; ++l1) {
    float s[6];
#pragma unroll
    for (int l2 = 0; l2 < 6; ++l2) {
      float d = q[l1] * k[l2];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      s[l2] = d * scale;
    }
    float m = s[0];
#pragma unroll
    for (int l2 = 1; l2 < 6; ++l2) m = fmaxf(m, s[l2]);
    float sum = 0.f;
#pragma unroll
    for (int l2 = 0; l2 < 6; ++l2) { s[l2] = expf(s[l2] - m); sum += s[l2]; }
    const float inv = 1.f / sum;
    float c = 0.f;
#pragma unroll
    for (int l2 = 0; l2 < 6; ++l2) {
      s[l2] *= inv;
      const uint32_t pi = (static_cast<uint32_t>(w) * 6 + l1) * 6 + l2;
      if (lane == 0) probs[pi] = s[l2];                        // kept before dropout (softmax backward needs it)
      const float sd = drop.thr != 0 ? drop_fwd(s[l2], drop, drop_seed(drop), pi) : s[l2];
      c = fmaf(sd, v[l2], c);
    }
    ctx[static_cast<size_t>(b * 6 + l1) * 256 + h * 32 + lane] = c;
  }
}

__global__ void __launch_bounds__(256)
mha6_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ probs, const float* __restrict__ dctx,
                float* __restrict__ dqkv, int B, const DropSpec drop) {
  pdl_enter();
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= B * 8) return;
  const int b = w >> 3, h = w & 7;
  float q[6], k[6], v[6], dc[6];
#pragma unroll
  for (int l = 0; l < 6; ++l) {
    const float* row = qkv + static_cast<size_t>(b * 6 + l) * 768 + h * 32 + lane;
    q[l] = row[0]; k[l] = row[256]; v[l] = row[512];
    dc[l] = dctx[static_cast<size_t>(b * 6 + l) * 256 + h * 32 + lane];
  }
  float dq[6], dk[6], dv[6];
#pragma unroll
  for (int l = 0; l < 6; ++l) { dq[l] = 0.f; dk[l] = 0.f; dv[l] = 0.f; }
  const float scale = 0.17677669529663687f;
#pragma unroll
  for (int l1 = 0; l1 < 6; ++l1) {
    float a[6], da[6];
    float dot = 0.f;
#pragma unroll
    for (int l2 = 0; l2 < 6; ++l2) {
      const uint32_t pi = (static_cast<uint32_t>(w) * 6 + l1) * 6 + l2;
      a[l2] = probs[pi];
      const float mg = drop.thr != 0 ? drop_grad(drop, drop_seed(drop), pi) : 1.f;   // d(dropped prob)/d(prob)
      float d = dc[l1] * v[l2];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      da[l2] = d * mg;
      dot = fmaf(da[l2], a[l2], dot);
      dv[l2] = fmaf(a[l2] * mg, dc[l1], dv[l2]);
    }
#pragma unroll
    for (int l2 = 0; l2 < 6; ++l2) {
      const float ds = a[l2] * (da[l2] - dot) * scale;
      dq[l1] = fmaf(ds, k[l2], dq[l1]);
      dk[l2] = fmaf(ds, q[l1], dk[l2]);
    }
  }
#pragma unroll
  for (int l = 0; l < 6; ++l) {
    float* row = dqkv + static_cast<size_t>(b * 6 + l) * 768 + h * 32 + lane;
    row[0] = dq[l]; row[256] = dk[l]; row[512] = dv[l];
  }
}

// ------------------------------------------------------------------------------------------------
// gated attention pooling core (reference: models/blocks.py:42-48 + models/mcat/mcat.py:105-108)
// one block (256 threads = feature index) per slide
//   A[l] = sum_d a[l][d] b[l][d] wc[d] + bc ;  w = softmax_l(A) ;  hp[d] = sum_l w[l] x[l][d]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += sh[i];
  return r;
}

__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ bgate,
                const float* __restrict__ wc, const float* __restrict__ bc, float* __restrict__ logits,
                float* __restrict__ w, float* __restrict__ hp) {
  pdl_enter();
  __shared__ float sh[8];
  const int b = blockIdx.x, d = threadIdx.x;
  float A[6];
#pragma unroll
  for (int l = 0; l < 6; ++l) {
    const size_t o = static_cast<size_t>(b * 6 + l) * 256 + d;
    A[l] = block_sum_256(a[o] * bgate[o] * wc[d], sh) + bc[0];
  }
  float m = A[0];
#pragma unroll
  for (int l = 1; l < 6; ++l) m = fmaxf(m, A[l]);
  float e[6], sum = 0.f;
#pragma unroll
  for (int l = 0; l < 6; ++l) { e[l] = expf(A[l] - m); sum += e[l]; }
  float acc = 0.f;
#pragma unroll
  for (int l = 0; l < 6; ++l) {
    e[l] /= sum;
    acc = fmaf(e[l], x[static_cast<size_t>(b * 6 + l) * 256 + d], acc);
  }
  hp[static_cast<size_t>(b) * 256 + d] = acc;
  if (d < 6) { logits[b * 6 + d] = A[d]; w[b * 6 + d] = e[d]; }
}

// dx (overwritten) gets the value-path gradient w[l]*dhp; da_pre/db_pre are the gradients at the pre-activations of
// the tanh / sigmoid branches; gwc/gbc are accumulated with atomics (one add per slide and feature)
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ bgate,
                const float* __restrict__ wc, const float* __restrict__ w, const float* __restrict__ dhp,
                float* __restrict__ dx, float* __restrict__ da_pre, float* __restrict__ db_pre,
                float* __restrict__ gwc, float* __restrict__ gbc, const DropSpec drop_a, const DropSpec drop_b) {
  pdl_enter();
  __shared__ float sh[8];
  const int b = blockIdx.x, d = threadIdx.x;
  const float g = dhp[static_cast<size_t>(b) * 256 + d];
  float wl[6], dw[6];
  float dot = 0.f;
#pragma unroll
  for (int l = 0; l < 6; ++l) {
    wl[l] = w[b * 6 + l];
    dw[l] = block_sum_256(g * x[static_cast<size_t>(b * 6 + l) * 256 + d], sh);
    dot = fmaf(dw[l], wl[l], dot);
  }
  float gw = 0.f, gb = 0.f;
#pragma unroll
  for (int l = 0; l < 6; ++l) {
    const size_t o = static_cast<size_t>(b * 6 + l) * 256 + d;
    const float dA = wl[l] * (dw[l] - dot);
    const float av = a[o], bv = bgate[o];       // as stored: after their dropout layers in train mode
    const float dab = dA * wc[d];
    dx[o] = wl[l] * g;
    float ga = 1.f, gbm = 1.f, a0 = av, b0 = bv;  // dropout derivatives and the pre-dropout activations
    if (drop_a.thr != 0) { ga = drop_grad(drop_a, drop_seed(drop_a), static_cast<uint32_t>(o)); a0 = drop_invert(av, drop_a); }
    if (drop_b.thr != 0) { gbm = drop_grad(drop_b, drop_seed(drop_b), static_cast<uint32_t>(o)); b0 = drop_invert(bv, drop_b); }
    da_pre[o] = dab * bv * ga * (1.f - a0 * a0);
    db_pre[o] = dab * av * gbm * b0 * (1.f - b0);
    gw = fmaf(dA, av * bv, gw);
    gb += dA;
  }
  if (gwc != nullptr) atomicAdd(gwc + d, gw);
  if (gbc != nullptr && d == 0) atomicAdd(gbc, gb);
}

// ------------------------------------------------------------------------------------------------
// survival head (models/mcat/mcat.py:126-138) and losses (models/loss.py:5-43), one thread per slide
// ------------------------------------------------------------------------------------------------
__global__ void surv_head_fwd_kernel(const float* __restrict__ logits, float* __restrict__ hazards,
                                     float* __restrict__ S, float* __restrict__ Y, int B, int K) {
  pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float m = -INFINITY;
  for (int j = 0; j < K; ++j) m = fmaxf(m, logits[b * K + j]);
  float sum = 0.f, s = 1.f;
  for (int j = 0; j < K; ++j) {
    const float z = logits[b * K + j];
    const float hz = 1.f / (1.f + expf(-z));
    hazards[b * K + j] = hz;
    s *= (1.f - hz);
    S[b * K + j] = s;
    sum += expf(z - m);
  }
  for (int j = 0; j < K; ++j) Y[b * K + j] = expf(logits[b * K + j] - m) / sum;
}

// dlogits from upstream gradients of hazards, S and Y (any may be null)
__global__ void surv_head_bwd_kernel(const float* __restrict__ hazards, const float* __restrict__ S,
                                     const float* __restrict__ Y, const float* __restrict__ dhaz,
                                     const float* __restrict__ dS, const float* __restrict__ dY,
                                     float* __restrict__ dlogits, int B, int K) {
  pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float dotY = 0.f;
  if (dY) for (int j = 0; j < K; ++j) dotY = fmaf(dY[b * K + j], Y[b * K + j], dotY);
  float tail = 0.f;   // sum_{j >= t} dS_j S_j, built from the last interval backwards
  for (int t = K - 1; t >= 0; --t) {
    const float hz = hazards[b * K + t];
    if (dS) tail = fmaf(dS[b * K + t], S[b * K + t], tail);
    float dh = dhaz ? dhaz[b * K + t] : 0.f;
    dh -= tail / (1.f - hz);
    float dl = dh * hz * (1.f - hz);
    if (dY) dl += Y[b * K + t] * (dY[b * K + t] - dotY);
    dlogits[b * K + t] = dl;
  }
}

// kind 0 = NLL (alpha .15), 1 = CES (alpha .75); grad_scale multiplies the gradients (e.g. 1/grad_acc_step)
__global__ void surv_loss_kernel(int kind, const float* __restrict__ hazards, const float* __restrict__ S,
                                 const int64_t* __restrict__ label, const float* __restrict__ censor, float alpha,
                                 float eps, float grad_scale, float* __restrict__ loss, float* __restrict__ dhaz,
                                 float* __restrict__ dS, int B, int K) {
  pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int y = static_cast<int>(label[b]);
  const float c = censor[b];
  for (int j = 0; j < K; ++j) { dhaz[b * K + j] = 0.f; dS[b * K + j] = 0.f; }
  if (y < 0 || y >= K) { loss[b] = NAN; return; }   // out-of-range label: poison the loss, touch nothing else
  const float s_prev = y == 0 ? 1.f : S[b * K + y - 1];
  const float h_y = hazards[b * K + y];
  const float unc = -(1.f - c) * (logf(fmaxf(s_prev, eps)) + logf(fmaxf(h_y, eps)));
  float w_unc;
  float l;
  if (kind == 0) {
    const float s_y = S[b * K + y];
    const float cen = -c * logf(fmaxf(s_y, eps));
    l = (1.f - alpha) * (cen + unc) + alpha * unc;
    w_unc = 1.f;
    if (s_y > eps) dS[b * K + y] += grad_scale * (1.f - alpha) * (-c / s_y);
  } else {
    const float s_y = fmaxf(S[b * K + y], eps);
    const float ce = -(c * logf(s_y) + (1.f - c) * logf(1.f - s_y));
    l = (1.f - alpha) * ce + alpha * unc;
    w_unc = alpha;
    if (S[b * K + y] > eps) dS[b * K + y] += grad_scale * (1.f - alpha) * (-(c / s_y) + (1.f - c) / (1.f - s_y));
  }
  if (y >= 1 && s_prev > eps) dS[b * K + y - 1] += grad_scale * w_unc * (-(1.f - c) / s_prev);
  if (h_y > eps) dhaz[b * K + y] += grad_scale * w_unc * (-(1.f - c) / h_y);
  loss[b] = l;
}

// ------------------------------------------------------------------------------------------------
// GatedConcatFusion gates (models/fusion.py:25-27, 35-40): item_p * sigmoid(w_p . item_p + b_p), p = path / omic.
// One block per slide, 256 threads (one feature of both halves each); cat / catg / dcat are [B][512].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gate_concat_fwd_kernel(const float* __restrict__ cat, const float* __restrict__ w0, const float* __restrict__ b0,
                       const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ catg,
                       float* __restrict__ gateg) {
  pdl_enter();
  __shared__ float red[8];
  const int b = blockIdx.x, f = threadIdx.x;
  const float x0 = cat[b * 512 + f], x1 = cat[b * 512 + 256 + f];
  const float g0 = 1.f / (1.f + __expf(-(block_sum_256(x0 * w0[f], red) + b0[0])));
  const float g1 = 1.f / (1.f + __expf(-(block_sum_256(x1 * w1[f], red) + b1[0])));
  catg[b * 512 + f] = x0 * g0;
  catg[b * 512 + 256 + f] = x1 * g1;
  if (f == 0) { gateg[b * 2] = g0; gateg[b * 2 + 1] = g1; }
}
// dcat: in = gradient at the gated items, out = gradient at the items (in place); gw / gb (may be null) accumulate the
// gradients of the gate layers themselves
__global__ void __launch_bounds__(256)
gate_concat_bwd_kernel(const float* __restrict__ cat, const float* __restrict__ gateg, const float* __restrict__ w0,
                       const float* __restrict__ w1, float* __restrict__ dcat, float* __restrict__ gw0,
                       float* __restrict__ gb0, float* __restrict__ gw1, float* __restrict__ gb1) {
  pdl_enter();
  __shared__ float red[8];
  const int b = blockIdx.x, f = threadIdx.x;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float x = cat[b * 512 + p * 256 + f], d = dcat[b * 512 + p * 256 + f], g = gateg[b * 2 + p];
    const float ds = block_sum_256(d * x, red) * g * (1.f - g);      // gradient at the gate's pre-activation
    dcat[b * 512 + p * 256 + f] = d * g + ds * (p == 0 ? w0[f] : w1[f]);
    float* gw = p == 0 ? gw0 : gw1;
    float* gb = p == 0 ? gb0 : gb1;
    if (gw != nullptr) atomicAdd(gw + f, ds * x);
    if (gb != nullptr && f == 0) atomicAdd(gb, ds);
  }
}

// ------------------------------------------------------------------------------------------------
// bilinear-fusion helpers (models/fusion.py:81-113)
// ------------------------------------------------------------------------------------------------
// z[b][k] = sum_i x1[b][i] U[b][k*256+i] + bias[k];  g = sigmoid(z);  gh = g * h          (U = x2 W_z^T, by GEMM)
__global__ void __launch_bounds__(256)
bil_gate_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ U, const float* __restrict__ bias,
                    const float* __restrict__ h, float* __restrict__ g, float* __restrict__ gh) {
  pdl_enter();
  const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = wid; k < 32; k += 8) {
    float s = 0.f;
    for (int i = lane; i < 256; i += 32) s = fmaf(x1[b * 256 + i], U[(static_cast<size_t>(b) * 32 + k) * 256 + i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const float gg = 1.f / (1.f + expf(-(s + bias[k])));
      g[b * 32 + k] = gg;
      gh[b * 32 + k] = gg * h[b * 32 + k];
    }
  }
}
// from dgh: dh_pre = dgh*g*1[h>0], dz = dgh*h*g(1-g);  V[b][k*256+i] = dz[b][k] x1[b][i];  dx1[b][i] (+)= sum_k dz U
__global__ void __launch_bounds__(256)
bil_gate_bwd_kernel(const float* __restrict__ x1, const float* __restrict__ U, const float* __restrict__ h,
                    const float* __restrict__ g, const float* __restrict__ dgh, float* __restrict__ dh_pre,
                    float* __restrict__ dz, float* __restrict__ V, float* __restrict__ dx1, int accumulate_dx1) {
  pdl_enter();
  __shared__ float dz_s[32];
  const int b = blockIdx.x, i = threadIdx.x;
  if (i < 32) {
    const float gg = g[b * 32 + i], hh = h[b * 32 + i], d = dgh[b * 32 + i];
    dh_pre[b * 32 + i] = hh > 0.f ? d * gg : 0.f;
    const float z = d * hh * gg * (1.f - gg);
    dz[b * 32 + i] = z;
    dz_s[i] = z;
  }
  __syncthreads();
  const float xi = x1[b * 256 + i];
  float acc = 0.f;
#pragma unroll 4
  for (int k = 0; k < 32; ++k) {
    const size_t o = (static_cast<size_t>(b) * 32 + k) * 256 + i;
    V[o] = dz_s[k] * xi;
    acc = fmaf(dz_s[k], U[o], acc);
  }
  float* dst = dx1 + b * 256 + i;
  *dst = accumulate_dx1 ? (*dst + acc) : acc;
}
// kp[b][i*33+j] = o1e[i]*o2e[j] with o?e = [o?, 1];  cat tail = [o1e, o2e] written at cat[b][64..130)
__global__ void bil_kron_fwd_kernel(const float* __restrict__ o1, const float* __restrict__ o2,
                                    float* __restrict__ kp, float* __restrict__ cat, const DropSpec drop) {
  pdl_enter();
  const int b = blockIdx.x;
  for (int e = threadIdx.x; e < 33 * 33; e += blockDim.x) {
    const int i = e / 33, j = e % 33;
    const float a = i < 32 ? o1[b * 32 + i] : 1.f;
    const float c = j < 32 ? o2[b * 32 + j] : 1.f;
    float v = a * c;
    if (drop.thr != 0) v = drop_fwd(v, drop, drop_seed(drop), static_cast<uint32_t>(b) * 1089u + e);   // post_fusion_dropout
    kp[static_cast<size_t>(b) * 1089 + e] = v;
  }
  for (int e = threadIdx.x; e < 66; e += blockDim.x) {
    const int i = e % 33;
    const float v = e < 33 ? (i < 32 ? o1[b * 32 + i] : 1.f) : (i < 32 ? o2[b * 32 + i] : 1.f);
    cat[static_cast<size_t>(b) * 130 + 64 + e] = v;
  }
}
// do1[b][i] = dcat[b][64+i] + sum_j dkp[b][i*33+j] o2e[j];  do2[b][j] = dcat[b][97+j] + sum_i dkp[b][i*33+j] o1e[i]
__global__ void bil_kron_bwd_kernel(const float* __restrict__ o1, const float* __restrict__ o2,
                                    const float* __restrict__ dkp, const float* __restrict__ dcat,
                                    float* __restrict__ do1, float* __restrict__ do2, const DropSpec drop) {
  pdl_enter();
  const int b = blockIdx.x, t = threadIdx.x;   // 64 threads
  const uint32_t sd = drop.thr != 0 ? drop_seed(drop) : 0u;
  auto dk = [&](int e) {
    const float g = dkp[static_cast<size_t>(b) * 1089 + e];
    return drop.thr != 0 ? g * drop_grad(drop, sd, static_cast<uint32_t>(b) * 1089u + e) : g;
  };
  if (t < 32) {
    float s = dcat[static_cast<size_t>(b) * 130 + 64 + t];
    for (int j = 0; j < 33; ++j) s = fmaf(dk(t * 33 + j), j < 32 ? o2[b * 32 + j] : 1.f, s);
    do1[b * 32 + t] = s;
  } else if (t < 64) {
    const int j = t - 32;
    float s = dcat[static_cast<size_t>(b) * 130 + 97 + j];
    for (int i = 0; i < 33; ++i) s = fmaf(dk(i * 33 + j), i < 32 ? o1[b * 32 + i] : 1.f, s);
    do2[b * 32 + j] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// GE-NaCAGaT helpers (models/ge_nacagat/ge_nacagat.py): soft-max over rows of N columns, forward and backward
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce_256(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}
// x[row][:] = softmax(x[row][:]) in place; one 256-thread block per row
__global__ void __launch_bounds__(256) row_softmax_kernel(float* __restrict__ x, long long ld, int cols) {
  pdl_enter();
  __shared__ float red[8];
  float* row = x + static_cast<long long>(blockIdx.x) * ld;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, row[c]);
  m = block_reduce_256(m, true, red);
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) { const float e = __expf(row[c] - m); row[c] = e; s += e; }
  s = block_reduce_256(s, false, red);
  const float inv = 1.f / s;
  for (int c = threadIdx.x; c < cols; c += 256) row[c] *= inv;
}
// ds[row][c] = a[row][c] (da[row][c] - sum_c a da) * scale, in place over da
__global__ void __launch_bounds__(256)
row_softmax_bwd_kernel(const float* __restrict__ a, long long lda, float* __restrict__ da, long long ldd, int cols, float scale) {
  pdl_enter();
  __shared__ float red[8];
  const float* ar = a + static_cast<long long>(blockIdx.x) * lda;
  float* dr = da + static_cast<long long>(blockIdx.x) * ldd;
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) s = fmaf(ar[c], dr[c], s);
  s = block_reduce_256(s, false, red);
  for (int c = threadIdx.x; c < cols; c += 256) dr[c] = ar[c] * (dr[c] - s) * scale;
}
// Tensor-core path of GE-NaCAGaT: the soft-max backward of one attention row written straight as the bf16 (hi, lo)
// operand pair of the dQ / dK products (tc_gemm.cu) -- ds itself is never stored in fp32 and never re-read by a split
// pass.  DROP: da is first multiplied by the regenerated dropout factor of the probabilities (element base + row * cols + c).
template <bool DROP>
__global__ void __launch_bounds__(256)
row_softmax_bwd_pair_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ da, long long ldd, int cols,
                            float scale, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int pitch,
                            uint32_t base, DropSpec drop) {
  pdl_enter();
  __shared__ float red[8];
  uint32_t seedv = 0;
  if (DROP) seedv = drop_seed(drop);
  const float* ar = a + static_cast<long long>(blockIdx.x) * lda;
  const float* dr = da + static_cast<long long>(blockIdx.x) * ldd;
  const uint32_t rb = base + static_cast<uint32_t>(blockIdx.x) * static_cast<uint32_t>(cols);
  const bool vec = ((lda | ldd) & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(da)) & 15u) == 0;
  float s = 0.f;
  for (int c0 = threadIdx.x * 8; c0 < cols; c0 += 256 * 8) {
    float av[8], dv[8];
    if (vec && c0 + 8 <= cols) {
      const float4 a0 = *reinterpret_cast<const float4*>(ar + c0), a1 = *reinterpret_cast<const float4*>(ar + c0 + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(dr + c0), d1 = *reinterpret_cast<const float4*>(dr + c0 + 4);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
      dv[0] = d0.x; dv[1] = d0.y; dv[2] = d0.z; dv[3] = d0.w; dv[4] = d1.x; dv[5] = d1.y; dv[6] = d1.z; dv[7] = d1.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) { av[e] = c0 + e < cols ? ar[c0 + e] : 0.f; dv[e] = c0 + e < cols ? dr[c0 + e] : 0.f; }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (DROP) dv[e] *= drop_grad(drop, seedv, rb + c0 + e);
      s = fmaf(av[e], dv[e], s);
    }
  }
  s = block_reduce_256(s, false, red);
  __nv_bfloat16* hr = hi + static_cast<long long>(blockIdx.x) * pitch;
  __nv_bfloat16* lr = lo + static_cast<long long>(blockIdx.x) * pitch;
  for (int c0 = threadIdx.x * 8; c0 < pitch; c0 += 256 * 8) {
    float v[8];
    if (vec && c0 + 8 <= cols) {
      const float4 a0 = *reinterpret_cast<const float4*>(ar + c0), a1 = *reinterpret_cast<const float4*>(ar + c0 + 4);
      const float4 d0 = *reinterpret_cast<const float4*>(dr + c0), d1 = *reinterpret_cast<const float4*>(dr + c0 + 4);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (DROP) dv[e] *= drop_grad(drop, seedv, rb + c0 + e);
        v[e] = av[e] * (dv[e] - s) * scale;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[e] = 0.f;
        if (c0 + e < cols) {
          float d = dr[c0 + e];
          if (DROP) d *= drop_grad(drop, seedv, rb + c0 + e);
          v[e] = ar[c0 + e] * (d - s) * scale;
        }
      }
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * e] - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * e + 1] - __bfloat162float(h1));
      h[e] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      l[e] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
    *reinterpret_cast<uint4*>(hr + c0) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lr + c0) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
// Stand-alone operators (mpo_op_*): LayerNorm over rows of any width (one warp per row, biased variance), and the
// product of every row with its own scalar (the gates of GatedConcatFusion, models/fusion.py:34-38)
__global__ void __launch_bounds__(256)
op_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ y, int rows, int cols, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * cols;
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(cols);
  float v = 0.f;
  for (int c = lane; c < cols; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / static_cast<float>(cols) + eps);
  for (int c = lane; c < cols; c += 32)
    y[static_cast<long long>(row) * cols + c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}
__global__ void __launch_bounds__(256)
op_rowscale_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ y, int rows, int cols) {
  const long long n = static_cast<long long>(rows) * cols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = x[i] * g[i / cols];
}
// GE-NaCAGaT train mode: attention-probability dropout of the N-token encoder layers (nn.TransformerEncoderLayer's
// self_attn carries the layer's dropout rate).  out = dropout(p), element index = base + row * cols + col.
__global__ void __launch_bounds__(256)
probs_dropout_kernel(const float* __restrict__ p, float* __restrict__ out, long long n, uint32_t base, DropSpec drop) {
  pdl_enter();
  const uint32_t seedv = drop_seed(drop);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = drop_fwd(p[i], drop, seedv, base + static_cast<uint32_t>(i));
}
// ds = a (m da - sum_c a m da) * scale with m the regenerated dropout factor of the probabilities, in place over da
__global__ void __launch_bounds__(256)
row_softmax_bwd_drop_kernel(const float* __restrict__ a, long long lda, float* __restrict__ da, long long ldd, int cols,
                            float scale, uint32_t base, DropSpec drop) {
  pdl_enter();
  __shared__ float red[8];
  const uint32_t seedv = drop_seed(drop);
  const float* ar = a + static_cast<long long>(blockIdx.x) * lda;
  float* dr = da + static_cast<long long>(blockIdx.x) * ldd;
  const uint32_t rb = base + static_cast<uint32_t>(blockIdx.x) * static_cast<uint32_t>(cols);
  float s = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float d = dr[c] * drop_grad(drop, seedv, rb + c);
    dr[c] = d;
    s = fmaf(ar[c], d, s);
  }
  s = block_reduce_256(s, false, red);
  for (int c = threadIdx.x; c < cols; c += 256) dr[c] = ar[c] * (dr[c] - s) * scale;
}
// H = fp16 hi + fp16 lo (the projection pass keeps both: bag_fwd.cu)
__global__ void ge_h_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo, float* __restrict__ H, long long n) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) H[i] = __half2float(hi[i]) + __half2float(lo[i]);
}
// dz = dH * 1[h > 0] * keep_scale: fp32 (bias gradient) and bf16 (operand of the dW_H kernel)
__global__ void ge_dz_kernel(const float* __restrict__ dH, const __half* __restrict__ hi, float keep, float* __restrict__ dzf,
                             __nv_bfloat16* __restrict__ dz, long long n) {
  pdl_enter();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = __half2float(hi[i]) > 0.f ? dH[i] * keep : 0.f;
    dzf[i] = v;
    dz[i] = __float2bfloat16_rn(v);
  }
}
// nn.CrossEntropyLoss on the already soft-maxed Y (models/ge_nacagat/main.py:29,33): loss = -log softmax(Y)[label]
__global__ void ge_ce_loss_kernel(const float* __restrict__ Y, const long long* __restrict__ label, int ncls, float scale,
                                  float* __restrict__ loss, float* __restrict__ dY) {
  pdl_enter();
  if (threadIdx.x != 0) return;
  float m = -INFINITY;
  for (int j = 0; j < ncls; ++j) m = fmaxf(m, Y[j]);
  float s = 0.f;
  for (int j = 0; j < ncls; ++j) s += expf(Y[j] - m);
  const int y = static_cast<int>(label[0]);
  *loss = -(Y[y] - m - logf(s));
  for (int j = 0; j < ncls; ++j) dY[j] = (expf(Y[j] - m) / s - (j == y ? 1.f : 0.f)) * scale;
}

}  // namespace mpo
