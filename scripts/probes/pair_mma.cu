// Probe for the pair-tile forward kernel: one cluster of two CTAs computes H[256 x 256] = X[256 x 1024] W^T
// with tcgen05.mma.cta_group::2 (M = 256 split over the two CTAs, each CTA stages HALF of every W block), checks it
// against a host reference, and then times back-to-back MMAs on resident operands for cta_group::1 and ::2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multimodal-path-omic_b200/csrc \
//        scripts/probes/pair_mma.cu -o scripts/probes/pair_mma && timeout 60 scripts/probes/pair_mma
// Barrier protocol under test:
//   full[s]  (used in rank 0 only): one arrive.expect_tx by rank 0's producer for the bytes of BOTH CTAs; rank 1 issues
//            its loads with the .cta_group::2 TMA form on the barrier address with the peer bit (24) cleared
//   empty[s] (both CTAs): tcgen05.commit.cta_group::2 ... multicast::cluster from rank 0's MMA thread
//   done     (both CTAs): same commit after the last MMA
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "mpo_ptx.cuh"

using namespace mpo;

constexpr int kM = 256, kN = 256, kK = 1024, kBKp = 64, kKB = kK / kBKp, kSt = 2;
constexpr int kAB = 128 * kBKp * 2;          // 16 KB: this CTA's X block, and its half of the W block
constexpr int kStage = 2 * kAB;
constexpr int kSmem = kSt * kStage + 4096 + 256 + 1024;   // + the 4 KB soft-max weight operand of mode 3

__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity, int what) {
  for (long long i = 0; i < (1ll << 24); ++i)
    if (mbar_try_wait(bar, parity)) return;
  printf("probe: wait %d timed out (block %d thread %d)\n", what, blockIdx.x, threadIdx.x);
  __trap();
}
// (the cta_group::2 instruction forms are in mpo_ptx.cuh: tmem_alloc2, umma2_f16, umma2_commit_mcast, tma2_load_2d, ...)

// mode 0: correctness (TMA pipeline, H written out).  mode 1 / 2: `reps` x 64 MMAs on whatever is resident, cta_group 1 / 2
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
pair_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
            const __grid_constant__ CUtensorMap tm_h, float* H, int mode, int reps, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSt * kStage + 4096);
  uint64_t* full = bars;            // [kSt]
  uint64_t* empty = bars + kSt;     // [kSt]
  uint64_t* done = bars + 2 * kSt;
  uint64_t* pready = bars + 2 * kSt + 1;   // rank 0: both CTAs' operands of the pooled pair MMA are in place (count 2)
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2 * kSt + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool pair = mode != 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kSt; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    mbar_init(pready, 2);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (pair) { tmem_alloc2(slot, 256); tmem_relinquish2(); }
    else { tmem_alloc(slot, 256); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;

  if (mode == 0) {
    if (warp == 0 && lane == 0) {
      const uint64_t pol = policy_evict_first();
      for (int kb = 0; kb < kKB; ++kb) {
        const int s = kb % kSt;
        const uint32_t ph = (kb / kSt) & 1;
        wait_or_trap(&empty[s], ph ^ 1, 1);
        uint8_t* sa = smem + s * kStage;
        if (rank == 0) mbar_expect_tx(&full[s], 2 * kStage);      // the bytes of both CTAs
        tma2_load_2d(sa, &tm_x, &full[s], kb * kBKp, static_cast<int>(rank) * 128, pol);        // own 128 rows of X
        tma2_load_2d(sa + kAB, &tm_w, &full[s], kb * kBKp, static_cast<int>(rank) * 128, pol);  // own half of W (N rows)
      }
    } else if (warp == 1 && lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kM, kN, 0, 0);
      for (int kb = 0; kb < kKB; ++kb) {
        const int s = kb % kSt;
        wait_or_trap(&full[s], (kb / kSt) & 1, 2);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + s * kStage), b = a + kAB;
#pragma unroll
        for (int k = 0; k < kBKp / 16; ++k)
          umma2_f16(tmem, umma_desc_sw128(a + k * 32, 16, 1024), umma_desc_sw128(b + k * 32, 16, 1024), idesc,
                     (kb | k) != 0 ? 1u : 0u);
        umma2_commit_mcast(&empty[s], 3);
      }
      umma2_commit_mcast(done, 3);
    } else if (warp >= 2) {
      wait_or_trap(done, 0, 3);
      tc_fence_after();
      const int qd = warp & 3, row = qd * 32 + lane;
      for (int c = 0; c < kN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(qd * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) H[(static_cast<size_t>(rank) * 128 + row) * kN + c + j] = __uint_as_float(v[j]);
      }
    }
  } else if (mode == 3) {
    // The pooled product of the forward kernel as ONE pair MMA: D_r[feature, 16 r + i] = sum_patch H_r[patch, feature] P_r[i, patch].
    // A = each CTA's staged fp16 tile read M-major (M = 256 over the pair = 128 features per CTA and pass), B = [P_0 ; P_1]
    // (N = 32: 16 rows per CTA, K = the CTA's own 128 patches), hand-off through a count-2 barrier in rank 0 that rank 1
    // reaches with a remote arrive (peer bit cleared).
    uint8_t* staging = smem;
    uint8_t* Pb = smem + kSt * kStage;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(Pb)[i] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(&full[0], 65536);
      for (int cb = 0; cb < 4; ++cb) tma_load_2d(staging + cb * 16384, &tm_h, &full[0], cb * 64, static_cast<int>(rank) * 128, policy_evict_first());
    }
    if (threadIdx.x < 128) {
      const int r = threadIdx.x;
      for (int i = 0; i < 6; ++i) {
        const float pv = static_cast<float>((r * 7 + i * 13 + static_cast<int>(rank) * 5) % 17 - 8) / 16.f;
        uint8_t* pcol = Pb + (r >> 6) * 2048 + (r & 7) * 2;
        *reinterpret_cast<__half*>(pcol + i * 128 + ((((r & 63) >> 3) ^ i) << 4)) = __float2half_rn(pv);
      }
    }
    wait_or_trap(&full[0], 0, 21);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0)
      mbar_arrive_pair0(pready);
    if (warp == 1 && lane == 0 && rank == 0) {
      wait_or_trap(pready, 0, 22);
      tc_fence_after();
      const uint32_t idesc_d = umma_idesc(256, 32, 0, 0, 1, 0);
      const uint32_t a0 = smem_u32(staging), b0 = smem_u32(Pb);
      for (int mh = 0; mh < 2; ++mh)
        for (int kk = 0; kk < 8; ++kk)
          umma2_f16(tmem + mh * 32, umma_desc_sw128(a0 + mh * 2 * 16384 + kk * 2048, 128 * 128, 1024),
                     umma_desc_sw128(b0 + (kk >> 2) * 2048 + (kk & 3) * 32, 16, 1024), idesc_d, kk != 0 ? 1u : 0u);
      umma2_commit_mcast(done, 3);
    }
    if (warp >= 2) {
      wait_or_trap(done, 0, 23);
      tc_fence_after();
      const int qd = warp & 3;
      for (int mh = 0; mh < 2; ++mh) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(qd * 32) << 16) + mh * 32 + 16 * rank, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) H[(static_cast<size_t>(rank) * 256 + mh * 128 + qd * 32 + lane) * 16 + i] = __uint_as_float(v[i]);
      }
    }
  } else {
    // MMA pacing on resident (uninitialised) operands: cta_group::1 -> every CTA issues M128 N256, B = 32 KB at the stage
    // base; cta_group::2 -> rank 0 issues M256 N256, each CTA contributes A 16 KB and B 16 KB
    if (warp == 1 && lane == 0 && (!pair || rank == 0)) {
      const uint32_t a = smem_u32(smem), b = a + kAB;
      const uint32_t idesc = pair ? umma_idesc_bf16(256, 256, 0, 0) : umma_idesc_bf16(128, 256, 0, 0);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r)
#pragma unroll 1
        for (int i = 0; i < 64; ++i) {
          const uint64_t da = umma_desc_sw128(a + (i & 3) * 32, 16, 1024), db = umma_desc_sw128(b + (i & 3) * 32, 16, 1024);
          if (pair) umma2_f16(tmem, da, db, idesc, 1u);
          else umma_bf16(tmem, da, db, idesc, 1u);
        }
      if (pair) umma2_commit_mcast(done, 3);
      else umma_commit(done);
      wait_or_trap(done, 0, 4);
      cycles[blockIdx.x] = clock64() - t0;
    } else if (pair && rank == 1 && warp == 1 && lane == 0) {
      wait_or_trap(done, 0, 5);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (pair) tmem_dealloc2(tmem, 256);
    else tmem_dealloc(tmem, 256);
  }
}


// Load + MMA loop of the forward bag kernel alone, persistent over `num_tiles` 128-row tiles of a large X, no epilogue:
// PAIR = false: one CTA per tile, full W_H block per stage (48 KB stages)  -- the structure of bag_fwd_kernel today
// PAIR = true : one CTA pair per tile pair, half of the W_H block per CTA (32 KB stages), cta_group::2 MMAs
template <bool PAIR>
__global__ void __launch_bounds__(192, 1)
stream_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
              const __grid_constant__ CUtensorMap tm_h, int num_tiles, int stages, int store) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int kStB = PAIR ? 2 * kAB : 3 * kAB;
  uint8_t* staging = smem + stages * kStB;       // 64 KB: stands in for the fp16 H tile the epilogue would have written
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 65536);
  uint64_t* full = bars;
  uint64_t* empty = bars + 8;
  uint64_t* done = bars + 16;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 17);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc2(slot, 256); tmem_relinquish2(); }
    else { tmem_alloc(slot, 256); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *slot;
  // units: tile pairs (PAIR) or tiles; the persistent grid walks consecutive ranges
  const int nunits = PAIR ? num_tiles / 2 : num_tiles;
  const int nworkers = PAIR ? gridDim.x / 2 : gridDim.x;
  const int me = PAIR ? blockIdx.x / 2 : blockIdx.x;
  const int per = (nunits + nworkers - 1) / nworkers;
  const int u0 = me * per, u1 = min(nunits, u0 + per);
  if (warp == 0 && lane == 0) {
    const uint64_t pol_x = policy_evict_first(), pol_w = policy_evict_last();
    int s = 0; uint32_t ph = 0;
    for (int u = u0; u < u1; ++u) {
      const int row0 = (PAIR ? 2 * u + static_cast<int>(rank) : u) * 128;
      for (int kb = 0; kb < kKB; ++kb) {
        wait_or_trap(&empty[s], ph ^ 1, 11);
        uint8_t* sa = smem + s * kStB;
        if (PAIR) {
          if (rank == 0) mbar_expect_tx(&full[s], 2 * kStB);
          tma2_load_2d(sa, &tm_x, &full[s], kb * kBKp, row0, pol_x);
          tma2_load_2d(sa + kAB, &tm_w, &full[s], kb * kBKp, static_cast<int>(rank) * 128, pol_w);
        } else {
          mbar_expect_tx(&full[s], kStB);
          tma_load_2d(sa, &tm_x, &full[s], kb * kBKp, row0, pol_x);
          tma_load_2d(sa + kAB, &tm_w, &full[s], kb * kBKp, 0, pol_w);
          tma_load_2d(sa + 2 * kAB, &tm_w, &full[s], kb * kBKp, 128, pol_w);
        }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      if (store) {       // the h_saved traffic of the product kernel: 64 KB per tile, TMA store, evict-first
        tma_store_wait_read();
        const uint64_t pol_s = policy_evict_first();
        for (int cb = 0; cb < 4; ++cb) tma_store_2d_hint(&tm_h, staging + cb * 16384, cb * 64, row0, pol_s);
        tma_store_commit();
      }
    }
    if (store) tma_store_wait_read();
  } else if (warp == 1 && lane == 0 && rank == 0) {
    const uint32_t idesc = PAIR ? umma_idesc_bf16(256, 256, 0, 0) : umma_idesc_bf16(128, 256, 0, 0);
    int s = 0; uint32_t ph = 0;
    for (int u = u0; u < u1; ++u)
      for (int kb = 0; kb < kKB; ++kb) {
        wait_or_trap(&full[s], ph, 12);
        tc_fence_after();
        const uint32_t a = smem_u32(smem + s * kStB), b = a + kAB;
#pragma unroll
        for (int k = 0; k < kBKp / 16; ++k) {
          const uint64_t da = umma_desc_sw128(a + k * 32, 16, 1024), db = umma_desc_sw128(b + k * 32, 16, 1024);
          if (PAIR) umma2_f16(tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          else umma_bf16(tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if (PAIR) umma2_commit_mcast(&empty[s], 3);
        else umma_commit(&empty[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    if (PAIR) umma2_commit_mcast(done, 3);
    else umma_commit(done);
    wait_or_trap(done, 0, 13);
  } else if (PAIR && warp == 1 && lane == 0) {
    wait_or_trap(done, 0, 14);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem, 256);
    else tmem_dealloc(tmem, 256);
  }
}

template <bool PAIR>
static int run_stream(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& th, int num_tiles, int stages, int sms, int store) {
  const int stb = PAIR ? 2 * kAB : 3 * kAB;
  const int smem = stages * stb + 65536 + 256 + 1024;
  auto kern = stream_kernel<PAIR>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { printf("smem %d too large\n", smem); return 1; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(sms / 2 * 2); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) if (cudaLaunchKernelEx(&cfg, kern, tx, tw, th, num_tiles, stages, store) != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) cudaLaunchKernelEx(&cfg, kern, tx, tw, th, num_tiles, stages, store);
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("stream kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  printf("load+MMA loop%s, %s, %d stages of %d KB: %.2f us per 16384-patch slide (%.0f GB/s of X)\n", store ? " + 64 KB store per tile" : "", PAIR ? "pair tiles (cta_group::2, half W_H per CTA)" : "single tiles (cta_group::1, full W_H per CTA)",
         stages, stb / 1024, ms * 1e3 / (num_tiles / 128.0), num_tiles * 128.0 * 2048 / (ms * 1e-3) / 1e9);
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_map(CUtensorMap* out, void* base, uint64_t rows, uint64_t cols, bool f16 = false) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
  cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  return reinterpret_cast<EncodeTiledFn>(p)(out, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main() {
  std::vector<__nv_bfloat16> hx(kM * kK), hw(kN * kK);
  std::vector<float> fx(kM * kK), fw(kN * kK);
  srand(1);
  for (size_t i = 0; i < hx.size(); ++i) { hx[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fx[i] = __bfloat162float(hx[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { hw[i] = __float2bfloat16((rand() % 2001 - 1000) / 8000.f); fw[i] = __bfloat162float(hw[i]); }
  __nv_bfloat16 *dx, *dw; float* dH; long long* dc;
  CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2)); CK(cudaMalloc(&dH, kM * kN * 4)); CK(cudaMalloc(&dc, 16 * 8));
  CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dH, 0, kM * kN * 4));
  CUtensorMap tx, tw;
  if (make_map(&tx, dx, kM, kK) || make_map(&tw, dw, kN, kK)) { printf("tensor map failed\n"); return 1; }
  CK(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  std::vector<__half> hh(256 * 256);
  for (size_t i = 0; i < hh.size(); ++i) hh[i] = __float2half((rand() % 2001 - 1000) / 1000.f);
  __half* dh;
  CK(cudaMalloc(&dh, hh.size() * 2));
  CK(cudaMemcpy(dh, hh.data(), hh.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap th0;
  pair_kernel<<<2, 192, kSmem>>>(tx, tw, tx, dH, 0, 0, dc);
  CK(cudaDeviceSynchronize());
  std::vector<float> H(kM * kN);
  CK(cudaMemcpy(H.data(), dH, H.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double acc = 0;
      for (int k = 0; k < kK; ++k) acc += static_cast<double>(fx[m * kK + k]) * fw[n * kK + k];
      worst = fmax(worst, fabs(acc - H[m * kN + n]));
    }
  printf("pair MMA (cta_group::2, M256 N256, half of W per CTA): max |err| = %.3g  %s\n", worst, worst < 2e-3 ? "OK" : "MISMATCH");
  for (int mode = 1; mode <= 2; ++mode) {
    const int reps = 50;
    pair_kernel<<<2, 192, kSmem>>>(tx, tw, tx, dH, mode, reps, dc);
    CK(cudaDeviceSynchronize());
    long long c[2];
    CK(cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost));
    printf("cta_group::%d: %.1f cycles per MMA instruction (%s per CTA), %d instructions\n", mode, static_cast<double>(c[0]) / (64.0 * reps),
           mode == 1 ? "M128 N256 K16" : "M256 N256 K16 over the pair = M128 N256 K16", 64 * reps);
  }
  if (worst >= 2e-3) return 2;
  {
    if (make_map(&th0, dh, 256, 256, true)) { printf("tensor map failed\n"); return 1; }
    CK(cudaMemset(dH, 0, kM * kN * 4));
    pair_kernel<<<2, 192, kSmem>>>(tx, tw, th0, dH, 3, 0, dc);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(2 * 256 * 16);
    CK(cudaMemcpy(D.data(), dH, D.size() * 4, cudaMemcpyDeviceToHost));
    double w3 = 0;
    for (int r = 0; r < 2; ++r)
      for (int f = 0; f < 256; ++f)
        for (int i = 0; i < 16; ++i) {
          double acc = 0;
          if (i < 6)
            for (int pt = 0; pt < 128; ++pt)
              acc += static_cast<double>(__half2float(hh[(r * 128 + pt) * 256 + f])) * (static_cast<double>((pt * 7 + i * 13 + r * 5) % 17 - 8) / 16.0);
          w3 = fmax(w3, fabs(acc - D[(r * 256 + f) * 16 + i]));
        }
    printf("pooled pair MMA (M-major A, N = 32 = 16 weight rows per CTA, count-2 barrier with a remote arrive): max |err| = %.3g  %s\n", w3,
           w3 < 1e-3 ? "OK" : "MISMATCH");
    if (w3 >= 1e-3) return 4;
  }
  // the load + MMA loop over a bag-sized X (32 slides x 16384 patches), by pipeline depth
  const int num_tiles = 4096;
  __nv_bfloat16* bx;
  CK(cudaMalloc(&bx, static_cast<size_t>(num_tiles) * 128 * kK * 2));
  CK(cudaMemset(bx, 0x11, static_cast<size_t>(num_tiles) * 128 * kK * 2));
  CUtensorMap tbx;
  if (make_map(&tbx, bx, static_cast<uint64_t>(num_tiles) * 128, kK)) { printf("tensor map failed\n"); return 1; }
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  __half* bh;
  CK(cudaMalloc(&bh, static_cast<size_t>(num_tiles) * 128 * 256 * 2));
  CUtensorMap tbh;
  if (make_map(&tbh, bh, static_cast<uint64_t>(num_tiles) * 128, 256, true)) { printf("tensor map failed\n"); return 1; }
  for (int store = 0; store <= 1; ++store) {
    for (int st = 3; st <= 3; ++st) if (run_stream<false>(tbx, tw, tbh, num_tiles, st, sms, store)) return 3;
    for (int st = 3; st <= 5; ++st) if (run_stream<true>(tbx, tw, tbh, num_tiles, st, sms, store)) return 3;
  }
  return 0;
}
