// NaCAGaT bag stage (reference: models/blocks.py:156-192, PreGatingContextualAttention's attention core).
//
//   k_n    = W_k h_n + b_k                                  (blocks.py:156, key in-projection)
//   s_in   = q_i . k_n / 16 = h_n . qk_i + kc_i             (blocks.py:180,184; folded, from the forward bag pass)
//   P_in   = (tanh(q_i) . tanh(k_n) + 1) / 2                (blocks.py:185-186)
//   s'_in  = s_in P_in ;  a = softmax_n(s') ;  a' = dropout(a)   (blocks.py:187-190)
//   out_i  = sum_n a'_in v_n = W_v (sum_n a'_in h_n) + b_v sum_n a'_in      (blocks.py:192; W_v folded into the tail)
//
// bag_gate_kernel (forward): per 128-patch tile the saved fp16 activation tile is TMA-loaded once and used twice:
// as the A operand of K = H W_k^T (tcgen05; H as fp16 hi + fp16 lo tiles, W_k streamed as fp16 K blocks, fp32
// accumulators in TMEM) and,
// read M-major, as the A operand of the pooled contraction sum_n p_n h_n.  The epilogue threads turn the K tile into
// tanh(k) (kept as fp16 for the backward pass), the six gate dots, the gated scores and the tile-local softmax
// statistics; per-tile partials are merged by bag_merge_kernel (bag_fwd.cu).
//
// bag_dhk_kernel (backward): dz = dz_part + (dkg W_k) * 1[h > 0] * keep_scale -- the key-projection path of the
// gradient at the bag activations -- with the same tile/resident-weight structure, W_k read N-major.
#include "mpo_ptx.cuh"
#include "mpo_common.cuh"
#include "launchers.h"
#include <cstdlib>

namespace mpo {

constexpr int kGateThreads = 64 + 256;        // bag_dhk_kernel
constexpr int kGateFwdThreads = 64 + 512;     // bag_gate_kernel: 16 epilogue warps (a patch row is shared by four threads)
struct GateSmem {
  static constexpr int A = 0;                          // fp16 tile hi [4 blocks][128 rows][64] SW128   64 KB
  static constexpr int Alo = 65536;                    // fp16 tile lo (h - fp16(h)), same layout        64 KB
  static constexpr int W = 131072;                     // fp16 W_k ring: 2 x [256 rows][64] SW128        64 KB
  static constexpr int Pb = W + 65536;                 // fp16 [2][16][64] softmax weights (B operand)  4 KB
  static constexpr int tq = Pb + 4096;                 // fp32 [6][256] tanh(q_i) of the slide          6 KB
  static constexpr int bk = tq + 6144;                 // fp32 [256] key bias                           1 KB
  static constexpr int spart = bk + 1024;              // fp32 [3][128][8] gate partials of column quarters 1..3  12 KB
  static constexpr int wred = spart + 3 * 4096;        // fp32 [3][4][8] warp partials (max, sum, dropped sum)
  static constexpr int misc = wred + 384;              // fp32 [8] kc_i
  static constexpr int bars = misc + 64;
  static constexpr int tmem_slot = bars + 128;
  static constexpr int total = tmem_slot + 16;
};
constexpr int kGateSmemBytes = GateSmem::total + 1024;

// batch-wide power-of-two scale of the gate-path gradients (same formula as gate_scale in bag_bwd.cu)
__device__ __forceinline__ float pow2_scale_gate(const uint32_t* dg_max, float* inv) {
  const float m = 8.f * __uint_as_float(*dg_max);
  const uint32_t e = (__float_as_uint(m) >> 23) & 0xFFu;
  if (e == 0u || e >= 254u) { *inv = 1.f; return 1.f; }
  *inv = __uint_as_float(e << 23);
  return __uint_as_float((254u - e) << 23);
}

// tanh from ex2 + rcp (2 MUFU): absolute error ~1e-7 -- tanh.approx (2^-11) is too coarse for a 256-term gate dot
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(2.f * x);
  return 1.f - __fdividef(2.f, e + 1.f);
}

__global__ void __launch_bounds__(kGateFwdThreads, 1)
bag_gate_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_hlo,
                const __grid_constant__ CUtensorMap tm_w, const BagGateParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GateSmem::bars);
  uint64_t* a_full = bars + 12;       // [4] K block kb of the tile (hi and lo, 32 KB) landed: the key MMAs of block 0
                                      //     start while blocks 1..3 are still in flight
  uint64_t* w_full = bars + 8;        // [2] W_k block landed
  uint64_t* w_empty = bars + 10;      // [2] W_k block consumed
  uint64_t* a_empty = bars + 2;       // tile buffer free (pooled MMA retired)
  uint64_t* acc_full = bars + 3;      // K accumulators complete
  uint64_t* acc_empty = bars + 4;     // epilogue has drained TMEM (16 warp arrivals)
  uint64_t* p_ready = bars + 5;       // softmax weights written (16 warp arrivals)
  uint64_t* d_bar = bars + 6;         // pooled MMA retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + GateSmem::tmem_slot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (p.num_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int t_begin = min(p.num_tiles, static_cast<int>(blockIdx.x) * per);
  const int t_end = min(p.num_tiles, t_begin + per);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_h);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_hlo);
    for (int s = 0; s < 2; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int kb = 0; kb < 4; ++kb) mbar_init(&a_full[kb], 1);
    mbar_init(a_empty, 1); mbar_init(acc_full, 1);
    mbar_init(acc_empty, 16); mbar_init(p_ready, 16); mbar_init(d_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColK = 0, kColP = 256, kColPlo = 288;

  if (warp == 0) {
    if (lane == 0 && t_begin < t_end) {
      const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
      int it = 0, ws = 0;
      uint32_t wph = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        mbar_wait(a_empty, (it & 1) ^ 1);
        const int row0 = p.tile_info[t].row0;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          mbar_expect_tx(&a_full[cb], 32768);
          tma_load_2d(smem + GateSmem::A + cb * 16384, &tm_h, &a_full[cb], cb * 64, row0, pol_stream);
          tma_load_2d(smem + GateSmem::Alo + cb * 16384, &tm_hlo, &a_full[cb], cb * 64, row0, pol_stream);
        }
        // The one tile buffer (128 KB of (hi, lo)) is busy until this tile's pooled MMAs have retired, so the next tile's
        // load cannot overlap anything on the SM.  Its DRAM half can: MPO_GATE_L2PF=1 prefetches the next tile into L2
        // here (19 MB for the whole GPU).  Measured neutral (NaCAGaT step 1.996 / 1.998 ms with, 1.985 / 1.992 ms without):
        // the tile load is not what the serial tile chain (32 key MMAs -> tanh epilogue -> 32 pooled MMAs) waits for.
        if (p.l2_prefetch && t + 1 < t_end) {
          const int row1 = p.tile_info[t + 1].row0;
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) {
            tma_prefetch_l2_2d(&tm_h, cb * 64, row1);
            tma_prefetch_l2_2d(&tm_hlo, cb * 64, row1);
          }
        }
        for (int kb = 0; kb < 4; ++kb) {          // W_k is re-streamed from L2 for every tile, one 64-wide K block at a time
          mbar_wait(&w_empty[ws], wph ^ 1);
          mbar_expect_tx(&w_full[ws], 32768);
          tma_load_2d(smem + GateSmem::W + ws * 32768, &tm_w, &w_full[ws], kb * 64, 0, pol_keep);
          if (++ws == 2) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && t_begin < t_end) {
      constexpr uint32_t id_k = umma_idesc(128, 256, 0, 0, 0, 0);     // fp16 x fp16, both K-major
      constexpr uint32_t id_p = umma_idesc(128, 16, 0, 0, 1, 0);      // pooled: A = H^T (M-major), B K-major
      const uint32_t aW = smem_u32(smem + GateSmem::W), aA = smem_u32(smem + GateSmem::A);
      const uint32_t aP = smem_u32(smem + GateSmem::Pb);
      const uint32_t aL = smem_u32(smem + GateSmem::Alo);
      int it = 0, ws = 0;
      uint32_t wph = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(acc_empty, ph ^ 1);
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&a_full[kb], ph);
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          const uint32_t wb = aW + ws * 32768;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t db = umma_desc_sw128(wb + k * 32, 16, 1024);
            umma_bf16(tmem_base + kColK, umma_desc_sw128(aA + kb * 16384 + k * 32, 16, 1024), db, id_k, (kb | k) != 0 ? 1u : 0u);
            umma_bf16(tmem_base + kColK, umma_desc_sw128(aL + kb * 16384 + k * 32, 16, 1024), db, id_k, 1u);
          }
          umma_commit(&w_empty[ws]);
          if (++ws == 2) { ws = 0; wph ^= 1; }
        }
        umma_commit(acc_full);
        mbar_wait(p_ready, ph);
        tc_fence_after();
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tmem_base + kColP + mh * 16, umma_desc_sw128(aA + mh * 2 * 16384 + kk * 2048, 16384, 1024),
                      umma_desc_sw128(aP + (kk >> 2) * 2048 + (kk & 3) * 32, 16, 1024), id_p, kk != 0 ? 1u : 0u);
        // + p^T H_lo: NaCAGaT's gated softmax is often sharp (one patch can carry > 0.9 of a query's weight), so the fp16
        // rounding of that patch's activations (2^-11) does not average out of the pooled vector; measured on the GPU it
        // flipped ReLU units of the slide tail against the reference (8-13 % gradient error on 2 of 13 slides).  The
        // remainder tile is already staged for the key projection: 16 more N = 16 MMAs make pooled ~22-bit exact.
        // The remainder part goes to its own accumulator columns and (when the backward pass follows) to its own
        // partial buffer: the backward forms g = dP . fp16(h) from the saved hi tile alone, so its softmax-gradient
        // offset delta = dP . pooled must use the hi-only pooled vector as well, or a (g - delta) cancels to the wrong
        // value on a sharp softmax (measured: 2 % on the query gradients of the sharp fixture).
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tmem_base + kColPlo + mh * 16, umma_desc_sw128(aL + mh * 2 * 16384 + kk * 2048, 16384, 1024),
                      umma_desc_sw128(aP + (kk >> 2) * 2048 + (kk & 3) * 32, 16, 1024), id_p, kk != 0 ? 1u : 0u);
        umma_commit(d_bar);
        umma_commit(a_empty);
      }
    }
  } else {
    const int et = threadIdx.x - 64;           // 0..511
    const int qd = warp & 3;                    // TMEM lane quadrant
    const int ch = (warp - 2) >> 2;             // column quarter 0..3 (64 key features each)
    const int r = qd * 32 + lane;               // patch row of the tile
    float* tq_s = reinterpret_cast<float*>(smem + GateSmem::tq);
    float* bk_s = reinterpret_cast<float*>(smem + GateSmem::bk);
    float* spart_s = reinterpret_cast<float*>(smem + GateSmem::spart);
    float* wmax_s = reinterpret_cast<float*>(smem + GateSmem::wred);
    float* wsum_s = wmax_s + 32;
    float* wdrp_s = wmax_s + 64;
    float* kc_s = reinterpret_cast<float*>(smem + GateSmem::misc);
    uint8_t* Pb = smem + GateSmem::Pb;
    if (et < kD) bk_s[et] = p.bias_k[et];
    for (int o = et * 16; o < 4096; o += 512 * 16) *reinterpret_cast<uint4*>(Pb + o) = make_uint4(0, 0, 0, 0);
    const uint32_t seed = p.seed_dev != nullptr ? (p.seed ^ __ldg(p.seed_dev)) : p.seed;
    int cur_slide = -1;
    int it = 0;
    TileInfo ti_next = t_begin < t_end ? p.tile_info[t_begin] : TileInfo{};
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const TileInfo ti = ti_next;                 // the tile table is read one tile ahead
      if (t + 1 < t_end) ti_next = p.tile_info[t + 1];
      const uint32_t ph = it & 1;
      // raw scores of this tile's rows (written by the projection pass): loaded before the wait for the K accumulators,
      // so their DRAM latency overlaps the tile load and the key MMAs instead of sitting in front of the soft-max
      float sraw_pre[kQ];
      if (ch == 0) {
        const bool v = r < ti.nvalid;
#pragma unroll
        for (int i = 0; i < kQ; ++i)
          sraw_pre[i] = v ? p.scores[static_cast<size_t>(i) * p.total_rows + static_cast<size_t>(ti.row0 + r)] : 0.f;
      }
      if (ti.slide != cur_slide) {
        cur_slide = ti.slide;
        named_bar_sync(1, 512);     // nobody still reads the previous slide's operands
        const float* q = p.qp + static_cast<size_t>(ti.slide) * kQ * kD;
#pragma unroll
        for (int j = 0; j < kQ * kD / 512; ++j) tq_s[et + j * 512] = tanhf(q[et + j * 512]);
        if (et < kQ) kc_s[et] = p.kc[ti.slide * kQ + et];
      }
      named_bar_sync(1, 512);

      mbar_wait(acc_full, ph);
      tc_fence_after();
      float g[kQ];
#pragma unroll
      for (int i = 0; i < kQ; ++i) g[i] = 0.f;
      const size_t grow = static_cast<size_t>(ti.row0 + r);
      const bool valid = r < ti.nvalid;
#pragma unroll 1
      for (int c4 = 0; c4 < 2; ++c4) {
        const int col0 = ch * 64 + c4 * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + kColK + (static_cast<uint32_t>(qd * 32) << 16) + col0, v);
        tmem_ld_wait();
        uint4 pk_even = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float tt[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) tt[e] = tanh_fast(__uint_as_float(v[j + e]) + bk_s[col0 + j + e]);
#pragma unroll
          for (int i = 0; i < kQ; ++i) {
            const float4 q0 = *reinterpret_cast<const float4*>(tq_s + i * kD + col0 + j);
            const float4 q1 = *reinterpret_cast<const float4*>(tq_s + i * kD + col0 + j + 4);
            g[i] = fmaf(tt[0], q0.x, g[i]); g[i] = fmaf(tt[1], q0.y, g[i]);
            g[i] = fmaf(tt[2], q0.z, g[i]); g[i] = fmaf(tt[3], q0.w, g[i]);
            g[i] = fmaf(tt[4], q1.x, g[i]); g[i] = fmaf(tt[5], q1.y, g[i]);
            g[i] = fmaf(tt[6], q1.z, g[i]); g[i] = fmaf(tt[7], q1.w, g[i]);
          }
          if (p.t_out != nullptr && valid) {
            uint4 pk;
            pk.x = pack_f16x2(tt[0], tt[1]); pk.y = pack_f16x2(tt[2], tt[3]);
            pk.z = pack_f16x2(tt[4], tt[5]); pk.w = pack_f16x2(tt[6], tt[7]);
            // the thread owns the row: two chunks = one full 32-byte sector per store
            if ((j & 8) == 0) pk_even = pk;
            else st_global_256(p.t_out + grow * kD + col0 + j - 8, pk_even, pk);
          }
        }
      }
      // K accumulators drained
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      if (ch != 0) {
        float* sp = spart_s + (ch - 1) * 1024 + r * 8;
        *reinterpret_cast<float4*>(sp) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float2*>(sp + 4) = make_float2(g[4], g[5]);
      }
      named_bar_sync(1, 512);
      if (ch == 0) {
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {
          const float4 o0 = *reinterpret_cast<const float4*>(spart_s + qq * 1024 + r * 8);
          const float2 o1 = *reinterpret_cast<const float2*>(spart_s + qq * 1024 + r * 8 + 4);
          g[0] += o0.x; g[1] += o0.y; g[2] += o0.z; g[3] += o0.w; g[4] += o1.x; g[5] += o1.y;
        }
        float sg[kQ];
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          const size_t o = static_cast<size_t>(i) * p.total_rows + grow;
          const float sraw = valid ? sraw_pre[i] + kc_s[i] : 0.f;
          const float P = 0.5f * (g[i] + 1.f);
          sg[i] = sraw * P;
          if (valid) {
            p.scores_g[o] = sg[i];
            if (p.pgate != nullptr) { p.pgate[o] = P; p.scores[o] = sraw; }
          }
          float m = valid ? sg[i] : -INFINITY;
#pragma unroll
          for (int o2 = 16; o2 > 0; o2 >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o2));
          if (lane == 0) wmax_s[qd * 8 + i] = m;
        }
        named_bar_sync(2, 128);
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          const float m = fmaxf(fmaxf(wmax_s[i], wmax_s[8 + i]), fmaxf(wmax_s[16 + i], wmax_s[24 + i]));
          const float pr = valid ? __expf(sg[i] - m) : 0.f;
          float prd = pr;
          if (p.drop_thr != 0) {   // attention dropout (blocks.py:189-190): the normaliser keeps every weight
            const uint32_t rb = rng_u32(seed, 1u, static_cast<uint32_t>(i) * static_cast<uint32_t>(p.total_rows) +
                                                      static_cast<uint32_t>(grow)) & 0xFFu;
            prd = rb < p.drop_thr ? 0.f : pr * p.drop_scale;
          }
          const __half ph_ = __float2half_rn(prd);
          const __half pl_ = __float2half_rn(prd - __half2float(ph_));
          uint8_t* pcol = Pb + (r >> 6) * 2048 + (r & 7) * 2;
          *reinterpret_cast<__half*>(pcol + i * 128 + ((((r & 63) >> 3) ^ i) << 4)) = ph_;
          *reinterpret_cast<__half*>(pcol + (i + 6) * 128 + ((((r & 63) >> 3) ^ ((i + 6) & 7)) << 4)) = pl_;
          float l = pr, ld = __half2float(ph_) + __half2float(pl_);
#pragma unroll
          for (int o2 = 16; o2 > 0; o2 >>= 1) {
            l += __shfl_xor_sync(0xffffffffu, l, o2);
            ld += __shfl_xor_sync(0xffffffffu, ld, o2);
          }
          if (lane == 0) { wsum_s[qd * 8 + i] = l; wdrp_s[qd * 8 + i] = ld; }
          if (et == 0) p.part_ml[static_cast<size_t>(t) * 18 + i] = m;
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      named_bar_sync(1, 512);
      if (et < kQ) {
        p.part_ml[static_cast<size_t>(t) * 18 + 6 + et] = wsum_s[et] + wsum_s[8 + et] + wsum_s[16 + et] + wsum_s[24 + et];
        p.part_ml[static_cast<size_t>(t) * 18 + 12 + et] = wdrp_s[et] + wdrp_s[8 + et] + wdrp_s[16 + et] + wdrp_s[24 + et];
      }
      mbar_wait(d_bar, ph);
      tc_fence_after();
      if (ch < 2) {                  // pooled read-back: feature half ch, 128 features x 12 (hi | lo) query columns
        uint32_t dv[16], dl[16];
        tmem_ld_32x32b_x16(tmem_base + kColP + ch * 16 + (static_cast<uint32_t>(qd * 32) << 16), dv);
        tmem_ld_32x32b_x16(tmem_base + kColPlo + ch * 16 + (static_cast<uint32_t>(qd * 32) << 16), dl);
        tmem_ld_wait();
        const size_t po = static_cast<size_t>(t) * (kQ * kD) + ch * 128 + qd * 32 + lane;
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          const float lo = __uint_as_float(dl[i]) + __uint_as_float(dl[i + 6]);
          p.part_pool[po + i * kD] = __uint_as_float(dv[i]) + __uint_as_float(dv[i + 6]) + lo;
          if (p.part_pool_lo != nullptr) p.part_pool_lo[po + i * kD] = lo;
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// bag_dhk_kernel (backward): the key-projection path of the gradient at the bag activations,
//     dz = dz_part + (dkg W_k) 1[h > 0] keep_scale,      dkg W_k on tcgen05 with W_k read N-major,
// then db_H = column sums of the final dz tile (a ones-operand MMA on the staged tile) and a TMA store of dz.
// dkg arrives as fp16 scaled by the batch-wide power of two gs (bag_bwd.cu); un-scaled here.
// ------------------------------------------------------------------------------------------------
struct DhkSmem {
  static constexpr int A = 0;                          // 2 x fp16 dkg tile [4][128][64] SW128; later the bf16 dz tile  128 KB
  static constexpr int W = 2 * 65536;                  // fp16 W_k ring: 2 x ([4 N blocks][64 k rows][64]) SW128     64 KB
  static constexpr int ones = W + 65536;               // bf16 [2][16][64] row 0 = 1                                 4 KB
  static constexpr int bars = ones + 4096;
  static constexpr int tmem_slot = bars + 128;
  static constexpr int total = tmem_slot + 16;
};
constexpr int kDhkSmemBytes = DhkSmem::total + 1024;

// Round 2: the dkg / output tile is double-buffered (the load of tile t + 1 no longer waits for the store of tile t) and
// every epilogue thread issues ALL of its dz / h loads of a tile before it waits for the accumulators, so their DRAM
// latency overlaps the tile load and the MMAs (before: plain loads issued at use, long-scoreboard 77 %, 3.45 TB/s).
__global__ void __launch_bounds__(kGateThreads, 1)
bag_dhk_kernel(const __grid_constant__ CUtensorMap tm_dkg, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_dz, const BagDhkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DhkSmem::bars);
  uint64_t* a_full = bars;            // [2] dkg tile landed
  uint64_t* a_empty = bars + 2;       // [2] tile buffer free (2 arrivals: MMA-db retired, store has read it)
  uint64_t* acc_full = bars + 4;      // dkg W_k complete
  uint64_t* acc_empty = bars + 5;     // accumulators drained (8 warp arrivals)
  uint64_t* w_ready = bars + 6;       // dz tile written in place (8 warp arrivals)
  uint64_t* b_bar = bars + 7;         // MMA-db retired
  uint64_t* w_full = bars + 8;        // [2]
  uint64_t* w_empty = bars + 10;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DhkSmem::tmem_slot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (p.num_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int t_begin = min(p.num_tiles, static_cast<int>(blockIdx.x) * per);
  const int t_end = min(p.num_tiles, t_begin + per);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_dkg); tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_dz);
    mbar_init(acc_full, 1); mbar_init(acc_empty, 8); mbar_init(w_ready, 8); mbar_init(b_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 2);
      mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  for (int o = threadIdx.x * 16; o < 4096; o += kGateThreads * 16)
    *reinterpret_cast<uint4*>(smem + DhkSmem::ones + o) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (threadIdx.x < 128) {
    const int k = threadIdx.x;
    *reinterpret_cast<uint16_t*>(smem + DhkSmem::ones + (k >> 6) * 2048 + (((k & 63) >> 3) << 4) + (k & 7) * 2) = 0x3F80;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColZ = 0, kColB = 256;

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
      int it = 0, ws = 0;
      uint32_t wph = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int buf = it & 1;
        mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&a_full[buf], 65536);
        const int row0 = p.tile_info[t].row0;
        uint8_t* dst = smem + DhkSmem::A + buf * 65536;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) tma_load_2d(dst + cb * 16384, &tm_dkg, &a_full[buf], cb * 64, row0, pol_stream);
        // W_k is read N-major: a ring slot holds the key features e in [64 kb, 64 kb + 64) (the K slice matching
        // A's block kb) for all 256 columns d, as four 64-column N blocks of [64 k rows][128 B]
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&w_empty[ws], wph ^ 1);
          mbar_expect_tx(&w_full[ws], 32768);
#pragma unroll
          for (int nb = 0; nb < 4; ++nb)
            tma_load_2d(smem + DhkSmem::W + ws * 32768 + nb * 8192, &tm_w, &w_full[ws], nb * 64, kb * 64, pol_keep);
          if (++ws == 2) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_k = umma_idesc(128, 256, 0, 0, 0, 1);     // fp16: A K-major (dkg), B N-major (W_k)
      constexpr uint32_t id_b = umma_idesc(128, 16, 1, 1, 1, 0);      // bf16: A M-major (dz^T), B K-major (ones)
      const uint32_t aW = smem_u32(smem + DhkSmem::W);
      const uint32_t aOne = smem_u32(smem + DhkSmem::ones);
      int it = 0, ws = 0;
      uint32_t wph = 0;
      int pending_buf = -1;            // buffer whose TMA store has been issued but not yet waited for
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const uint32_t ph = it & 1;
        const int buf = it & 1;
        const uint32_t aA = smem_u32(smem + DhkSmem::A + buf * 65536);
        mbar_wait(acc_empty, ph ^ 1);
        mbar_wait(&a_full[buf], (it >> 1) & 1);
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&w_full[ws], wph);
          tc_fence_after();
          const uint32_t wb = aW + ws * 32768;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + kColZ, umma_desc_sw128(aA + kb * 16384 + k * 32, 16, 1024),
                      umma_desc_sw128(wb + k * 2048, 8192, 1024), id_k, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&w_empty[ws]);
          if (++ws == 2) { ws = 0; wph ^= 1; }
        }
        umma_commit(acc_full);
        // the previous tile's store has had a whole MMA phase to read its buffer: release it for the tile after this one
        if (pending_buf >= 0) { tma_store_wait_read(); mbar_arrive(&a_empty[pending_buf]); pending_buf = -1; }
        mbar_wait(w_ready, ph);
        tc_fence_after();
        const TileInfo ti = p.tile_info[t];
        const bool full_tile = ti.nvalid == kTileM;
        if (full_tile) {
#pragma unroll
          for (int cb = 0; cb < 4; ++cb)
            tma_store_2d(&tm_dz, smem + DhkSmem::A + buf * 65536 + cb * 16384, cb * 64, ti.row0);
          tma_store_commit();
        }
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16(tmem_base + kColB + mh * 16, umma_desc_sw128(aA + mh * 2 * 16384 + kk * 2048, 16384, 1024),
                      umma_desc_sw128(aOne + (kk >> 2) * 2048 + (kk & 3) * 32, 16, 1024), id_b, kk != 0 ? 1u : 0u);
        umma_commit(b_bar);
        umma_commit(&a_empty[buf]);
        if (full_tile) pending_buf = buf; else mbar_arrive(&a_empty[buf]);
      }
      if (pending_buf >= 0) { tma_store_wait_read(); mbar_arrive(&a_empty[pending_buf]); }
    }
  } else {
    const int qd = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int r = qd * 32 + lane;
    float inv_gs;
    (void)pow2_scale_gate(p.dg_max, &inv_gs);
    const float ks = p.keep_scale * inv_gs;
    int it = 0, prev_t = -1;
    auto read_db = [&](int tt, uint32_t parity) {
      mbar_wait(b_bar, parity);
      tc_fence_after();
      uint32_t bv[16];
      tmem_ld_32x32b_x16(tmem_base + kColB + ch * 16 + (static_cast<uint32_t>(qd * 32) << 16), bv);
      tmem_ld_wait();
      p.part_db[static_cast<size_t>(tt) * kD + ch * 128 + qd * 32 + lane] = __uint_as_float(bv[0]);
      tc_fence_before();
    };
    // This thread's 256 B of the incoming dz tile (value path of the gradient; a unit masked by ReLU / dropout arrives as
    // -0.0, see bag_bwd_dz_kernel<kDzNacDh>) live in registers one tile AHEAD: each 64 B group is re-loaded for the next tile
    // as soon as the current tile has consumed it, so the loads are in flight during a whole tile's MMAs and epilogue.
    auto row_ptr = [&](const TileInfo& tn) {
      // rows past the end of the packed bag do not exist in the global buffers: clamp the address, mask the value
      return reinterpret_cast<const uint4*>(p.dz + static_cast<size_t>(min(tn.row0 + r, p.total_rows - 1)) * kD + ch * 128);
    };
    TileInfo ti_next = t_begin < t_end ? p.tile_info[t_begin] : TileInfo{};
    uint4 zin[16];
    if (t_begin < t_end) {
      const uint4* zp0 = row_ptr(ti_next);
#pragma unroll
      for (int q = 0; q < 16; q += 2) ld_global_256(zp0 + q, zin[q], zin[q + 1]);
    }
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const TileInfo ti = ti_next;
      const bool has_next = t + 1 < t_end;
      if (has_next) ti_next = p.tile_info[t + 1];
      const uint32_t ph = it & 1;
      const bool valid = r < ti.nvalid;
      const bool direct = ti.nvalid != kTileM;
      const size_t grow = static_cast<size_t>(min(ti.row0 + r, p.total_rows - 1));
      if (prev_t >= 0) read_db(prev_t, ph ^ 1);
      mbar_wait(acc_full, ph);
      tc_fence_after();
      const uint4* zn = row_ptr(ti_next);
      uint8_t* tile = smem + DhkSmem::A + (it & 1) * 65536;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int col0 = ch * 128 + c4 * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + kColZ + (static_cast<uint32_t>(qd * 32) << 16) + col0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = q * 8;
          const uint4 zq = zin[c4 * 4 + q];
          const uint32_t zi[4] = {zq.x, zq.y, zq.z, zq.w};
          uint32_t ow[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool live0 = (zi[e] & 0xFFFFu) != 0x8000u, live1 = (zi[e] >> 16) != 0x8000u;
            const float z0 = live0 ? bf16lo_to_f32(zi[e]) + __uint_as_float(v[j + 2 * e]) * ks : 0.f;
            const float z1 = live1 ? bf16hi_to_f32(zi[e]) + __uint_as_float(v[j + 2 * e + 1]) * ks : 0.f;
            ow[e] = valid ? pack_bf16x2(z0, z1) : 0u;
          }
          const uint4 o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          const int j16 = (col0 + j) >> 3;
          const int cb = j16 >> 3, jj = j16 & 7;
          *reinterpret_cast<uint4*>(tile + cb * 16384 + r * 128 + ((jj ^ (r & 7)) << 4)) = o;
          if (direct && valid) *reinterpret_cast<uint4*>(p.dz + grow * kD + col0 + j) = o;
        }
        if (has_next) {
          // (one full 32-byte sector per load: the thread owns the row)
#pragma unroll
          for (int q = 0; q < 4; q += 2) ld_global_256(zn + c4 * 4 + q, zin[c4 * 4 + q], zin[c4 * 4 + q + 1]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(acc_empty); mbar_arrive(w_ready); }
      prev_t = t;
    }
    if (prev_t >= 0) read_db(prev_t, (it - 1) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

cudaError_t launch_bag_dhk(const CUtensorMap& tm_dkg, const CUtensorMap& tm_w, const CUtensorMap& tm_dz,
                           const BagDhkParams& prm, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bag_dhk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDhkSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (prm.num_tiles <= 0) return cudaSuccess;
  const int grid = prm.num_tiles < num_sms ? prm.num_tiles : num_sms;
  bag_dhk_kernel<<<grid, kGateThreads, kDhkSmemBytes, stream>>>(tm_dkg, tm_w, tm_dz, prm);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bag_gate(const CUtensorMap& tm_h, const CUtensorMap& tm_hlo, const CUtensorMap& tm_w,
                            const BagGateParams& prm, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bag_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGateSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (prm.num_tiles <= 0) return cudaSuccess;
  const int grid = prm.num_tiles < num_sms ? prm.num_tiles : num_sms;
  static int l2pf = -1;
  if (l2pf < 0) { const char* e = getenv("MPO_GATE_L2PF"); l2pf = (e != nullptr && atoi(e) != 0) ? 1 : 0; }
  BagGateParams prm_l = prm;
  prm_l.l2_prefetch = l2pf;
  bag_gate_kernel<<<grid, kGateFwdThreads, kGateSmemBytes, stream>>>(tm_h, tm_hlo, tm_w, prm_l);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mpo
