// Bag-pass backward for MCAT (autograd of models/mcat/mcat.py:87,97 in the reference).
//
//  bag_bwd_dz_kernel  (CUDA cores, streaming): from the saved fp16 activations h_n, the saved raw scores and
//      the upstream gradient dPooled[6,256] it forms, per patch,
//          a_in  = exp(s_in - lse_i)
//          ds_in = a_in (dPooled_i . h_n - delta_i),      delta_i = dPooled_i . pooled_i
//          dh_n  = sum_i a_in dPooled_i + ds_in qk_i
//          dz_n  = dh_n * 1[h_n > 0] * keep_scale          (ReLU + inverted dropout in one mask)
//      writes dz (bf16) and per-tile partials of dqk_i = sum_n ds_in h_n and db_H = sum_n dz_n.
//  bag_bwd_dw_kernel  (tcgen05): dW_H[256,1024] += dz^T X over every packed row of the batch; both operands
//      are M/N-major straight out of their row-major global layout (no transposes), split-K over patch rows,
//      fp32 accumulators in TMEM, reduced into the gradient buffer with red.global.add.
#include "mpo_ptx.cuh"
#include "mpo_common.cuh"
#include "launchers.h"

namespace mpo {


constexpr int kDzWarps = 8;
constexpr int kDzSmemBytes = kDzWarps * 7 * kD * 4;   // cross-warp reduction buffer

__global__ void __launch_bounds__(kDzWarps * 32, 1) bag_bwd_dz_kernel(const BagBwdDzParams p) {
  extern __shared__ float red[];   // [8 warps][7][256]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x;
  const TileInfo ti = p.tile_info[t];
  const size_t sb = static_cast<size_t>(ti.slide) * kQ * kD;

  // per-lane slices (8 features) of dPooled and qk; delta_i via a warp reduction
  float dP[kQ][8], qk[kQ][8];
  float lse_r[kQ];
#pragma unroll
  for (int i = 0; i < kQ; ++i) {
    const float4 a0 = *reinterpret_cast<const float4*>(p.dpooled + sb + i * kD + lane * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(p.dpooled + sb + i * kD + lane * 8 + 4);
    dP[i][0] = a0.x; dP[i][1] = a0.y; dP[i][2] = a0.z; dP[i][3] = a0.w;
    dP[i][4] = a1.x; dP[i][5] = a1.y; dP[i][6] = a1.z; dP[i][7] = a1.w;
    const float4 q0 = *reinterpret_cast<const float4*>(p.qk + sb + i * kD + lane * 8);
    const float4 q1 = *reinterpret_cast<const float4*>(p.qk + sb + i * kD + lane * 8 + 4);
    qk[i][0] = q0.x; qk[i][1] = q0.y; qk[i][2] = q0.z; qk[i][3] = q0.w;
    qk[i][4] = q1.x; qk[i][5] = q1.y; qk[i][6] = q1.z; qk[i][7] = q1.w;
    lse_r[i] = p.lse[ti.slide * kQ + i];
  }
  float delta[kQ];
#pragma unroll
  for (int i = 0; i < kQ; ++i) {
    const float4 c0 = *reinterpret_cast<const float4*>(p.pooled + sb + i * kD + lane * 8);
    const float4 c1 = *reinterpret_cast<const float4*>(p.pooled + sb + i * kD + lane * 8 + 4);
    float v = dP[i][0] * c0.x + dP[i][1] * c0.y + dP[i][2] * c0.z + dP[i][3] * c0.w + dP[i][4] * c1.x +
              dP[i][5] * c1.y + dP[i][6] * c1.z + dP[i][7] * c1.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    delta[i] = v;
  }

  float dqk[kQ][8], db[8];
#pragma unroll
  for (int i = 0; i < kQ; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) dqk[i][e] = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) db[e] = 0.f;

  // each warp owns 16 consecutive rows of the tile
#pragma unroll 1
  for (int rr = 0; rr < kTileM / kDzWarps; ++rr) {
    const int n = warp * (kTileM / kDzWarps) + rr;
    if (n >= ti.nvalid) break;   // warp-uniform
    const size_t grow = static_cast<size_t>(ti.row0 + n);
    const uint4 hv = *reinterpret_cast<const uint4*>(p.h + grow * kD + lane * 8);
    float h[8];
    { const float2 t0 = unpack_f16x2(hv.x), t1 = unpack_f16x2(hv.y), t2 = unpack_f16x2(hv.z), t3 = unpack_f16x2(hv.w);
      h[0] = t0.x; h[1] = t0.y; h[2] = t1.x; h[3] = t1.y; h[4] = t2.x; h[5] = t2.y; h[6] = t3.x; h[7] = t3.y; }
    float a[kQ], ds[kQ];
#pragma unroll
    for (int i = 0; i < kQ; ++i) {
      float g = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) g = fmaf(dP[i][e], h[e], g);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      a[i] = __expf(__ldg(p.scores + static_cast<size_t>(i) * p.total_rows + grow) - lse_r[i]);
      ds[i] = a[i] * (g - delta[i]);
    }
    float dzv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < kQ; ++i) v = fmaf(a[i], dP[i][e], fmaf(ds[i], qk[i][e], v));
      dzv[e] = h[e] > 0.f ? v * p.keep_scale : 0.f;
      db[e] += dzv[e];
    }
#pragma unroll
    for (int i = 0; i < kQ; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) dqk[i][e] = fmaf(ds[i], h[e], dqk[i][e]);
    uint4 pk;
    pk.x = pack_bf16x2(dzv[0], dzv[1]);
    pk.y = pack_bf16x2(dzv[2], dzv[3]);
    pk.z = pack_bf16x2(dzv[4], dzv[5]);
    pk.w = pack_bf16x2(dzv[6], dzv[7]);
    *reinterpret_cast<uint4*>(p.dz + grow * kD + lane * 8) = pk;
  }

  // cross-warp reduction of the per-lane partial sums
  float* mine = red + warp * (7 * kD);
#pragma unroll
  for (int i = 0; i < kQ; ++i) {
    *reinterpret_cast<float4*>(mine + i * kD + lane * 8) = make_float4(dqk[i][0], dqk[i][1], dqk[i][2], dqk[i][3]);
    *reinterpret_cast<float4*>(mine + i * kD + lane * 8 + 4) = make_float4(dqk[i][4], dqk[i][5], dqk[i][6], dqk[i][7]);
  }
  *reinterpret_cast<float4*>(mine + 6 * kD + lane * 8) = make_float4(db[0], db[1], db[2], db[3]);
  *reinterpret_cast<float4*>(mine + 6 * kD + lane * 8 + 4) = make_float4(db[4], db[5], db[6], db[7]);
  __syncthreads();
  for (int e = threadIdx.x; e < 7 * kD; e += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kDzWarps; ++w) v += red[w * (7 * kD) + e];
    if (e < kQ * kD) p.part_dqk[static_cast<size_t>(t) * (kQ * kD) + e] = v;
    else p.part_db[static_cast<size_t>(t) * kD + (e - kQ * kD)] = v;
  }
}

// dqk[b][i][d] = sum over the slide's tiles;  db_H[d] += sum over all tiles (gradient accumulation)
// 4 tile groups x 64 float4 columns per block so that many independent loads are in flight
__global__ void __launch_bounds__(256)
bag_bwd_reduce_kernel(const int* __restrict__ tile_prefix, const float* __restrict__ part_dqk,
                      const float* __restrict__ part_db, float* __restrict__ dqk, float* __restrict__ grad_bias,
                      int B, int num_tiles) {
  __shared__ float4 acc_s[4][64];
  const int tid = threadIdx.x, tg = tid >> 6, dq = tid & 63;
  const bool is_bias = static_cast<int>(blockIdx.x) >= B;
  const int b = blockIdx.x, i = blockIdx.y;
  // bias blocks: (gridDim.x - B) * 6 chunks share the tiles round-robin and finish with atomics
  const int nchunk = (static_cast<int>(gridDim.x) - B) * kQ;
  const int chunk = (static_cast<int>(blockIdx.x) - B) * kQ + i;
  const int t0 = is_bias ? chunk * 4 : tile_prefix[b];
  const int t1 = is_bias ? num_tiles : tile_prefix[b + 1];
  const int tstep = is_bias ? nchunk * 4 : 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int t = t0 + tg; t < t1; t += tstep) {
    const float* src = is_bias ? part_db + static_cast<size_t>(t) * kD : part_dqk + (static_cast<size_t>(t) * kQ + i) * kD;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + dq);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  acc_s[tg][dq] = acc;
  __syncthreads();
  if (tid < 64) {
    float4 r = acc_s[0][tid];
#pragma unroll
    for (int g = 1; g < 4; ++g) { r.x += acc_s[g][tid].x; r.y += acc_s[g][tid].y; r.z += acc_s[g][tid].z; r.w += acc_s[g][tid].w; }
    if (is_bias) {
      float* g4 = grad_bias + tid * 4;
      atomicAdd(g4 + 0, r.x); atomicAdd(g4 + 1, r.y); atomicAdd(g4 + 2, r.z); atomicAdd(g4 + 3, r.w);
    } else {
      reinterpret_cast<float4*>(dqk + (static_cast<size_t>(b) * kQ + i) * kD)[tid] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dW_H += dz^T X        M = 256 features (two UMMA M=128 halves), N = 256 input columns per CTA, K = patch rows
// ------------------------------------------------------------------------------------------------
constexpr int kDwStages = 3;
constexpr int kDwBK = 64;                               // patch rows per stage
constexpr int kDwBox = 64 * kDwBK * 2;                  // one TMA box: 64 contiguous elements x 64 rows = 8 KB
constexpr int kDwABytes = 4 * kDwBox;                   // dz  [64 rows x 256 features]
constexpr int kDwBBytes = 4 * kDwBox;                   // X   [64 rows x 256 columns]
constexpr int kDwStageBytes = kDwABytes + kDwBBytes;    // 64 KB
constexpr int kDwThreads = 64 + 128;
constexpr int kDwSmemBytes = kDwStages * kDwStageBytes + 256 + 1024;

__global__ void __launch_bounds__(kDwThreads, 1)
bag_bwd_dw_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_x,
                  float* __restrict__ grad_w,   // [256][1024] fp32, accumulated
                  int total_rows, int num_splits) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms, by pointer arithmetic so the compiler keeps the shared state space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * kDwStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kDwStages;
  uint64_t* done_bar = bars + 2 * kDwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDwStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb = blockIdx.x & 3;          // which 256-column block of X / dW_H
  const int sp = blockIdx.x >> 2;         // split-K index over patch rows
  const int chunks_total = (total_rows + kDwBK - 1) / kDwBK;
  const int per = (chunks_total + num_splits - 1) / num_splits;
  const int c_begin = sp * per;
  const int c_end = min(chunks_total, c_begin + per);
  const int n_chunks = max(0, c_end - c_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_dz);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < kDwStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      const uint64_t pol_keep = policy_evict_last();     // dz is re-read by the four column-block CTAs
      int stage = 0; uint32_t phase = 0;
      for (int c = 0; c < n_chunks; ++c) {
        const int row = (c_begin + c) * kDwBK;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kDwStageBytes;
        mbar_expect_tx(&full_bar[stage], kDwStageBytes);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tma_load_2d(sa + j * kDwBox, &tm_dz, &full_bar[stage], j * 64, row, pol_keep);
          tma_load_2d(sa + kDwABytes + j * kDwBox, &tm_x, &full_bar[stage], cb * 256 + j * 64, row, pol_stream);
        }
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 1, 1);   // A and B both M/N-major
      int stage = 0; uint32_t phase = 0;
      for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * kDwStageBytes);
        const uint32_t b_addr = a_addr + kDwABytes;
#pragma unroll
        for (int kk = 0; kk < kDwBK / 16; ++kk) {
          // 16 patch rows = two 8-row swizzle atoms = 2048 B down each box
          const uint64_t db = umma_desc_sw128(b_addr + kk * 2048, kDwBox, 1024);
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {
            const uint64_t da = umma_desc_sw128(a_addr + mh * 2 * kDwBox + kk * 2048, kDwBox, 1024);
            umma_bf16(tmem_base + mh * 256, da, db, idesc, (c | kk) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else {
    // epilogue: 4 warps, thread = one feature row per M half; reduce into the fp32 gradient
    const int qd = warp & 3;
    if (n_chunks > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int mh = 0; mh < 2; ++mh) {
        const int f = mh * 128 + qd * 32 + lane;
        float* dst = grad_w + static_cast<size_t>(f) * kDIn + cb * 256;
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + mh * 256 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + c0 + j, __uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
cudaError_t launch_bag_bwd_dz(const BagBwdDzParams& prm, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bag_bwd_dz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDzSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (prm.num_tiles <= 0) return cudaSuccess;
  bag_bwd_dz_kernel<<<prm.num_tiles, kDzWarps * 32, kDzSmemBytes, stream>>>(prm);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bag_bwd_reduce(const int* tile_prefix, const float* part_dqk, const float* part_db, float* dqk,
                                  float* grad_bias, int B, int num_tiles, cudaStream_t stream) {
  bag_bwd_reduce_kernel<<<dim3(B + 16, kQ), 256, 0, stream>>>(tile_prefix, part_dqk, part_db, dqk, grad_bias, B,
                                                            num_tiles);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bag_bwd_dw(const CUtensorMap& tm_dz, const CUtensorMap& tm_x, float* grad_w, int total_rows,
                              int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bag_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (total_rows <= 0) return cudaSuccess;
  const int chunks = (total_rows + kDwBK - 1) / kDwBK;
  int splits = num_sms / 4;
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  bag_bwd_dw_kernel<<<4 * splits, kDwThreads, kDwSmemBytes, stream>>>(tm_dz, tm_x, grad_w, total_rows, splits);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mpo
