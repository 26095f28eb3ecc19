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
#include "tail_dev.cuh"

namespace mpo {


// ------------------------------------------------------------------------------------------------
// Row-expand stage on tcgen05, three modes sharing one pipeline.  Per 128-patch tile (one persistent CTA per SM,
// contiguous tile ranges per CTA):
//   TMA      : a saved fp16 tile [128 x 256] (H, or tanh(k) in mode 2) -> shared memory (three buffers, two tiles ahead)
//   MMA-G    : G[128 x 16]    = H dP^T            (dP of the slide as fp16 hi/lo rows, scaled per query to ~1)
//   threads  : per-patch scalars (softmax weights and their gradients), written as tiny fp16 operands
//   MMA-dZ   : Z[128 x 256]   = [a | ds] [dP ; qk]   (fp16, ONE K = 16 step; both operands M / N-major so that they take
//              4 + 8 KB instead of the 16 + 32 KB of 128-byte-swizzled K-major rows -- the room for the third tile buffer.
//              Power-of-two scales keep fp16's range: every row of [dP ; qk] is brought to a maximum in [1, 2) per slide,
//              [a | ds] carries the reciprocals and one more scale per tile that the output warps take off the
//              accumulators; two fp16 roundings per product are a quarter of the bf16 rounding of the stored dz)
//   MMA-dqk  : Q^T[256 x 16]  = tile^T ds            (tile read M-major, power-of-two scale on ds)
//   threads  : out = f(Z, tile) -> 16-bit, written IN PLACE over the tile; TMA store to HBM
//   MMA-db   : b[256] = out^T 1                      (out tile read M-major against a ones operand)
// kDzMcat    (MCAT, autograd of mcat.py:87,97):    out = dz = Z 1[h > 0] keep_scale (bf16); Q = dqk; b = db_H
// kDzNacDh   (NaCAGaT value/fold path, blocks.py:184-192): same with the gated softmax, attention dropout and the
//            gate-side scalars dg_i = ds'_i s_i / 2 written out per patch; no b (the dz tile is not final yet)
// kDzNacDkg  (NaCAGaT gate path, blocks.py:185-186): tile = tanh(k); Z = dg tq; out = dkg = (1 - t^2) Z gs (fp16,
//            gs a batch-wide power of two); Q = dtq = sum_n dg_n tanh(k_n); b = gate part of db_k
// ------------------------------------------------------------------------------------------------
// kDzMcatLite (MCAT, default): the row-scalar half of kDzMcat only -- per patch the 12 coefficients [a_i | ds_i] (fp32)
//            and the 256 ReLU/dropout mask bits of h are written out (80 B per patch instead of a 512 B dz row), Q = dqk;
//            the dz tile itself is regenerated inside the weight-gradient kernel (bag_bwd_dwz_kernel below), so dz never
//            makes the round trip through HBM
constexpr int kDzMcat = 0, kDzNacDh = 1, kDzNacDkg = 2, kDzMcatLite = 3;
constexpr int kDzThreads = 64 + 128 + 256;      // producer + issuer warps, 4 row-scalar warps, 8 output warps
constexpr int kDzBufs = 3;          // tile buffers: two loads in flight while a third tile is being worked on
struct DzSmem {
  static constexpr int tile = 0;                       // 3 x 64 KB  fp16 tile [4][128][64] SW128 (later: the 16-bit output tile)
  static constexpr int C = kDzBufs * 65536;            // fp16 [2 M blocks][16 k][64] M-major (12 k used)     4 KB
  static constexpr int Dm = C + 4096;                  // fp16 [4 N blocks][16 k][64] N-major (12 k used)     8 KB
  static constexpr int DP = Dm + 8192;                 // fp16 [4][16][64] K-major            8 KB
  static constexpr int DS = DP + 8192;                 // fp16 [2][16][64] K-major (K = rows) 4 KB
  static constexpr int ones = DS + 4096;               // 16-bit [2][16][64] row 0 = 1        4 KB
  static constexpr int scal = ones + 4096;             // fp32 scratch (see below)            1 KB
  static constexpr int bars = scal + 1024;
  static constexpr int tmem_slot = bars + 128;
  static constexpr int total = tmem_slot + 16;
};
constexpr int kDzSmemBytes = DzSmem::total + 1024;
constexpr uint32_t kColG = 0, kColQ = 32, kColB = 64, kColZ = 256;

// power-of-two scale that brings m into [1, 2) (1 when m is zero / denormal); *inv receives its reciprocal
__device__ __forceinline__ float pow2_scale(float m, float* inv) {
  const uint32_t e = (__float_as_uint(m) >> 23) & 0xFFu;
  if (e == 0u || e >= 254u) { *inv = 1.f; return 1.f; }
  *inv = __uint_as_float(e << 23);
  return __uint_as_float((254u - e) << 23);
}
// batch-wide scale of the gate-path gradients: |dkg| <= sum_i |dg_i| |tq_i| <= 6 max|dg|
__device__ __forceinline__ float gate_scale(const uint32_t* dg_max, float* inv) {
  return pow2_scale(8.f * __uint_as_float(*dg_max), inv);
}

// hi/lo bf16 pairs of 12 values -> the C / D operand row layout: k 0..11 hi, 16..27 `mid`, 32..43 `last`
__device__ __forceinline__ void split_bf16x12(const float (&v)[12], uint32_t (&hi)[6], uint32_t (&lo)[6]) {
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
    hi[j] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
    lo[j] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
  }
}
__device__ __forceinline__ void store_row48(uint8_t* row, int sw, const uint32_t (&a)[6], const uint32_t (&b)[6],
                                            const uint32_t (&c)[6]) {
  *reinterpret_cast<uint4*>(row + ((0 ^ sw) << 4)) = make_uint4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<uint4*>(row + ((1 ^ sw) << 4)) = make_uint4(a[4], a[5], 0u, 0u);
  *reinterpret_cast<uint4*>(row + ((2 ^ sw) << 4)) = make_uint4(b[0], b[1], b[2], b[3]);
  *reinterpret_cast<uint4*>(row + ((3 ^ sw) << 4)) = make_uint4(b[4], b[5], 0u, 0u);
  *reinterpret_cast<uint4*>(row + ((4 ^ sw) << 4)) = make_uint4(c[0], c[1], c[2], c[3]);
  *reinterpret_cast<uint4*>(row + ((5 ^ sw) << 4)) = make_uint4(c[4], c[5], 0u, 0u);
}
// element (k, m) of an M- or N-major fp16 operand of one K = 16 step: 64-element atoms of [16 k][64] (2048 B, 128-byte
// swizzle: the 16-byte group (m & 63) / 8 of k-row k sits at group ^ (k & 7))
__device__ __forceinline__ void store_mn_f16(uint8_t* base, int k, int m, float v) {
  *reinterpret_cast<__half*>(base + (m >> 6) * 2048 + k * 128 + ((((m & 63) >> 3) ^ (k & 7)) << 4) + (m & 7) * 2) =
      __float2half_rn(v);
}
// 6 values -> fp16 hi/lo entries of a [16 x 128] K-major operand column (rows 0..5 hi, 6..11 lo)
__device__ __forceinline__ void store_col_f16(uint8_t* base, int k, const float (&v)[kQ]) {
  uint8_t* pcol = base + (k >> 6) * 2048 + (k & 7) * 2;
  const int kc = (k & 63) >> 3;
#pragma unroll
  for (int i = 0; i < kQ; ++i) {
    const __half vh = __float2half_rn(v[i]);
    const __half vl = __float2half_rn(v[i] - __half2float(vh));
    *reinterpret_cast<__half*>(pcol + i * 128 + ((kc ^ i) << 4)) = vh;
    *reinterpret_cast<__half*>(pcol + (i + 6) * 128 + ((kc ^ ((i + 6) & 7)) << 4)) = vl;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kDzThreads, 1)
bag_bwd_dz_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
                  const BagBwdDzParams p) {
  pdl_enter();
  constexpr bool kLite = MODE == kDzMcatLite;    // no dz tile: coefficients + mask bits out, dz regenerated by the dW kernel
  constexpr bool kHasG = MODE != kDzNacDkg;      // MMA-G (dots of the tile rows with dP)
  constexpr bool kHasB = MODE != kDzNacDh && !kLite;   // MMA-db (column sums of the output tile)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DzSmem::bars);
  uint64_t* full_bar = bars;          // [3] TMA -> MMA issuers
  uint64_t* empty_bar = bars + 10;    // [3] tile buffer free again (store has read it + last MMA reading it retired)
  uint64_t* g_bar = bars + 5;         // MMA-G retired
  uint64_t* c_bar = bars + 6;         // C / DS (and, after a slide change, D / DP) operands written (4 warp arrivals)
  uint64_t* z_bar = bars + 7;         // MMA-dZ + MMA-dqk retired
  uint64_t* w_bar = bars + 8;         // output tile written in place, Q read back (8 warp arrivals)
  uint64_t* b_bar = bars + 9;         // MMA-db retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DzSmem::tmem_slot);
  float* scal = reinterpret_cast<float*>(smem + DzSmem::scal);
  // scal: [0..7] delta_i, [8..15] lse_i, [16..23] 1/dp_scale_i, [24..31] 1/qk_scale_i, [32..79] warp partials [4][12],
  //       [80..103] warp maxima of |qk_i| [4][6], [128..159] row-warp max [4][8] (entry 6: max of the row warp's scaled
  //       C entries), [160..167] dsuma_i, [168..199] row-warp sums of ds [4][8],
  //       [208..223] ds un-scale of the Q read-back, by tile parity [2][8], [224..225] un-scale of Z, by tile parity

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous tile range of this CTA (consecutive tiles mostly share a slide -> per-slide operands are rebuilt rarely)
  const int per = (p.num_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int t_begin = min(p.num_tiles, static_cast<int>(blockIdx.x) * per);
  const int t_end = min(p.num_tiles, t_begin + per);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_out);
    for (int s = 0; s < kDzBufs; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 2); }
    mbar_init(g_bar, 1); mbar_init(c_bar, 4); mbar_init(z_bar, 1); mbar_init(w_bar, 8);
    mbar_init(b_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  // zero the small operand buffers once (padding rows / columns stay zero), set the ones operand
  for (int o = threadIdx.x * 16; o < DzSmem::scal - DzSmem::C; o += kDzThreads * 16)
    *reinterpret_cast<uint4*>(smem + DzSmem::C + o) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (threadIdx.x < 128) {
    // ones operand: B[n = 0][k = threadIdx.x] = 1.0 in the format of the output tile, K-major, 2 blocks of 64 k
    const int k = threadIdx.x;
    *reinterpret_cast<uint16_t*>(smem + DzSmem::ones + (k >> 6) * 2048 + (((k & 63) >> 3) << 4) + (k & 7) * 2) =
        MODE == kDzNacDkg ? static_cast<uint16_t>(0x3C00) : static_cast<uint16_t>(0x3F80);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The tile loop is software-pipelined over three groups of warps and three tile buffers (round 2, third session; before
  // it the eight compute warps did the row scalars and the output tile of one tile back to back, with the two MMA round
  // trips in between -- scalars + MMA + output + hand-offs in series, 4.8 us per tile -- over two buffers.  Measured one
  // at a time neither change moves the kernel: the pipeline alone is then bound by how long a buffer stays occupied, 5.1 us
  // per tile; the third buffer alone by the serial chain, 4.8 us; profiles/r2d_probe_dz_timing_switches.txt):
  //   row-scalar warps (2..5)  tile t:   wait z(t-1) -> [slide change: D / DP] -> MMA-G(t) -> [a | ds] -> C / DS -> c(t)
  //   issuer (warp 1)          tile t:   wait w(t-1), store(t-1) -> wait c(t) -> MMA-dZ(t), MMA-dqk(t) -> z(t) -> MMA-db(t-1)
  //   output warps (6..13)     tile t:   wait z(t) -> out(t) in place, Q(t), b(t-1) -> w(t)
  // so the row scalars of tile t + 1 run while the output tile of tile t is being written.
  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: input tiles
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int buf = it % kDzBufs;
        const uint32_t ph = (it / kDzBufs) & 1;
        mbar_wait(&empty_bar[buf], ph ^ 1);
        uint8_t* dst = smem + DzSmem::tile + buf * 65536;
        mbar_expect_tx(&full_bar[buf], 65536);
        const int row0 = p.tile_info[t].row0;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) tma_load_2d(dst + cb * 16384, &tm_in, &full_bar[buf], cb * 64, row0, pol);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA-dZ / dqk / db and TMA-store issuer
    if (lane == 0) {
      constexpr uint32_t id_z = umma_idesc(128, 256, 0, 0, 1, 1);      // fp16 x fp16, A M-major, B N-major, K = 16
      constexpr uint32_t id_q = umma_idesc(128, 16, 0, 0, 1, 0);       // fp16, A M-major (tile^T)
      constexpr uint32_t id_b = MODE == kDzNacDkg ? umma_idesc(128, 16, 0, 0, 1, 0)    // fp16 output tile
                                                  : umma_idesc(128, 16, 1, 1, 1, 0);   // bf16 output tile
      const uint32_t aC = smem_u32(smem + DzSmem::C), aD = smem_u32(smem + DzSmem::Dm);
      const uint32_t aDS = smem_u32(smem + DzSmem::DS);
      const uint32_t aOne = smem_u32(smem + DzSmem::ones);
      // output tile `tt` (pipeline index `ip`) has been written: store it ...
      auto store_tile = [&](int tt, int ip) -> bool {
        mbar_wait(w_bar, ip & 1);
        tc_fence_after();
        const TileInfo ti = p.tile_info[tt];
        const bool full_tile = !kLite && ti.nvalid == kTileM;
        if (full_tile && !(p.debug & 16)) {
          uint8_t* src = smem + DzSmem::tile + (ip % kDzBufs) * 65536;
#pragma unroll
          for (int cb = 0; cb < 4; ++cb) tma_store_2d(&tm_out, src + cb * 16384, cb * 64, ti.row0);
          tma_store_commit();
        }
        return full_tile;
      };
      // ... and take its column sums; after that (and the store's read) the buffer goes back to the producer
      auto release_tile = [&](int ip, bool stored) {
        const int buf = ip % kDzBufs;
        const uint32_t aT = smem_u32(smem + DzSmem::tile + buf * 65536);
        if (kHasB) {
          if (!(p.debug & 2)) {
#pragma unroll
            for (int mh = 0; mh < 2; ++mh)
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma_bf16(tmem_base + kColB + mh * 16, umma_desc_sw128(aT + mh * 2 * 16384 + kk * 2048, 16384, 1024),
                          umma_desc_sw128(aOne + (kk >> 2) * 2048 + (kk & 3) * 32, 16, 1024), id_b, kk != 0 ? 1u : 0u);
          }
          umma_commit(b_bar);
        }
        umma_commit(&empty_bar[buf]);            // arrival 1 of 2: no MMA reads the tile any more
        if (stored) tma_store_wait_read();
        mbar_arrive(&empty_bar[buf]);            // arrival 2 of 2: the store has read it
      };
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int buf = it % kDzBufs;
        const uint32_t ph = (it / kDzBufs) & 1, tph = it & 1;
        const uint32_t aT = smem_u32(smem + DzSmem::tile + buf * 65536);
        // the Z / Q accumulators are free once the previous output tile is written
        const bool stored = it > 0 ? store_tile(t - 1, it - 1) : false;
        mbar_wait(c_bar, tph);
        mbar_wait(&full_bar[buf], ph);
        tc_fence_after();
        if (!kLite && !(p.debug & 4)) {
          // one K = 16 step: 64-element M / N atoms 2048 B apart, the two 8-k-row swizzle atoms of an atom 1024 B apart
          umma_bf16(tmem_base + kColZ, umma_desc_sw128(aC, 2048, 1024), umma_desc_sw128(aD, 2048, 1024), id_z, 0u);
        }
        if (!(p.debug & 1)) {
#pragma unroll
          for (int mh = 0; mh < 2; ++mh)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16(tmem_base + kColQ + mh * 16, umma_desc_sw128(aT + mh * 2 * 16384 + kk * 2048, 16384, 1024),
                        umma_desc_sw128(aDS + (kk >> 2) * 2048 + (kk & 3) * 32, 16, 1024), id_q, kk != 0 ? 1u : 0u);
        }
        umma_commit(z_bar);
        if (it > 0) release_tile(it - 1, stored);
      }
      if (it > 0) {
        const bool stored = store_tile(t_end - 1, it - 1);
        release_tile(it - 1, stored);
      }
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- 4 row-scalar warps (one patch row per thread)
    const int st = threadIdx.x - 64;           // 0..127 : features st, st + 128 when building per-slide operands
    const int qd = warp & 3;                   // TMEM lane quadrant
    const int r = qd * 32 + lane;              // patch row of the tile
    const bool leader = warp == 2 && lane == 0;     // issues MMA-G
    uint8_t* Cs = smem + DzSmem::C;
    uint8_t* Ds = smem + DzSmem::Dm;
    uint8_t* DPs = smem + DzSmem::DP;
    uint8_t* DSs = smem + DzSmem::DS;
    float gs = 1.f, inv_gs = 1.f;
    if (MODE == kDzNacDkg) gs = gate_scale(p.dg_max, &inv_gs);
    const uint32_t seed = p.seed_dev != nullptr ? (p.seed ^ __ldg(p.seed_dev)) : p.seed;
    constexpr uint32_t id_g = umma_idesc(128, 16, 0, 0, 0, 0);       // fp16 x fp16, both K-major
    const uint32_t aDP = smem_u32(smem + DzSmem::DP);
    auto issue_g = [&](int ip) {               // MMA-G of the tile in pipeline slot ip (leader only)
      mbar_wait(&full_bar[ip % kDzBufs], (ip / kDzBufs) & 1);
      tc_fence_after();
      if (!(p.debug & 32)) {
        const uint32_t a_tile = smem_u32(smem + DzSmem::tile + (ip % kDzBufs) * 65536);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + kColG, umma_desc_sw128(a_tile + kb * 16384 + k * 32, 16, 1024),
                      umma_desc_sw128(aDP + kb * 2048 + k * 32, 16, 1024), id_g, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(g_bar);
    };
    int cur_slide = -1;
    int it = 0;
    bool g_early = false;
    // The per-row scalars of a tile (scores / gate / map gradient, or dg) are plain global loads feeding the head of the
    // tile's dependency chain: they are issued one tile ahead, and the tile table is read one tile ahead as well, so no
    // DRAM latency sits between the tiles.
    float nsc[kQ], npg[kQ], ndm[kQ];
    auto prefetch_rows = [&](const TileInfo& tn) {
      const bool v = r < tn.nvalid;
      const size_t g0 = static_cast<size_t>(tn.row0 + r);
#pragma unroll
      for (int i = 0; i < kQ; ++i) {
        const size_t o = static_cast<size_t>(i) * p.total_rows + g0;
        if (MODE == kDzNacDkg) {
          nsc[i] = v ? __ldg(p.dg + o) : 0.f;
        } else {
          nsc[i] = v ? __ldg(p.scores + o) : 0.f;
          npg[i] = (MODE == kDzNacDh && v) ? __ldg(p.pgate + o) : 1.f;
          ndm[i] = (p.d_amap != nullptr && v) ? __ldg(p.d_amap + o) : 0.f;
        }
      }
    };
    TileInfo ti_next = t_begin < t_end ? p.tile_info[t_begin] : TileInfo{};
    // mode 2 reads dg, written by the mode-1 launch before it: nothing of this launch produces it, so one tile ahead is safe
    if (t_begin < t_end) prefetch_rows(ti_next);
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const TileInfo ti = ti_next;
      if (t + 1 < t_end) ti_next = p.tile_info[t + 1];
      const uint32_t tph = it & 1;
      // C / DS (and D across a slide change) are free again once MMA-dZ / MMA-dqk of the previous tile have retired
      if (it > 0) mbar_wait(z_bar, tph ^ 1);
      // ---- per-slide operands
      if (ti.slide != cur_slide) {
        cur_slide = ti.slide;
        const size_t sb = static_cast<size_t>(ti.slide) * kQ * kD;
        if (MODE == kDzNacDkg) {
          // D[k = i][n = et] = tanh(q_i)[et] (|.| <= 1: no scale); the qk rows 6..11 stay zero
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int et = st + h * 128;
#pragma unroll
            for (int i = 0; i < kQ; ++i) store_mn_f16(Ds, i, et, tanhf(p.qp[sb + i * kD + et]));
          }
        } else {
          float dp[2][kQ], qv[2][kQ], red[12], qmax[kQ];
#pragma unroll
          for (int j = 0; j < 12; ++j) red[j] = 0.f;
#pragma unroll
          for (int i = 0; i < kQ; ++i) qmax[i] = 0.f;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int et = st + h * 128;
#pragma unroll
            for (int i = 0; i < kQ; ++i) {
              dp[h][i] = p.dpooled[sb + i * kD + et];
              qv[h][i] = p.qk[sb + i * kD + et];
              // delta_i partial; consistent with g = dP . fp16(h): the hi-only pooled vector when a remainder part exists
              red[i] += dp[h][i] * (p.pooled[sb + i * kD + et] - (p.pooled_lo != nullptr ? p.pooled_lo[sb + i * kD + et] : 0.f));
              red[6 + i] = fmaxf(red[6 + i], fabsf(dp[h][i]));       // max |dP_i|
              qmax[i] = fmaxf(qmax[i], fabsf(qv[h][i]));             // max |qk_i|
            }
          }
#pragma unroll
          for (int i = 0; i < kQ; ++i) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              red[i] += __shfl_xor_sync(0xffffffffu, red[i], o);
              red[6 + i] = fmaxf(red[6 + i], __shfl_xor_sync(0xffffffffu, red[6 + i], o));
              qmax[i] = fmaxf(qmax[i], __shfl_xor_sync(0xffffffffu, qmax[i], o));
            }
          }
          named_bar_sync(1, 128);                            // previous tile no longer reads scal
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 12; ++j) scal[32 + (warp - 2) * 12 + j] = red[j];
#pragma unroll
            for (int i = 0; i < kQ; ++i) scal[80 + (warp - 2) * 6 + i] = qmax[i];
          }
          named_bar_sync(1, 128);
          float dscale[kQ], qscale[kQ];
#pragma unroll
          for (int i = 0; i < kQ; ++i) {
            float s = 0.f, m = 0.f, mq = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              s += scal[32 + w * 12 + i]; m = fmaxf(m, scal[32 + w * 12 + 6 + i]); mq = fmaxf(mq, scal[80 + w * 6 + i]);
            }
            float inv, invq;
            dscale[i] = pow2_scale(m, &inv);
            qscale[i] = pow2_scale(mq, &invq);
            if (st == 0) scal[24 + i] = invq;
            if (st == 0) {
              float dsum = 0.f;
              if (MODE == kDzNacDh && p.dsuma != nullptr) {   // delta_i = dP_i . pooled_i + dsuma_i suma_i
                dsum = p.dsuma[ti.slide * kQ + i];
                s += dsum * p.suma[ti.slide * kQ + i];
              }
              if (p.d_amap != nullptr) s += p.amap_dot[ti.slide * kQ + i];   // softmax Jacobian of the map gradient
              scal[i] = s; scal[8 + i] = p.lse[ti.slide * kQ + i]; scal[16 + i] = inv; scal[160 + i] = dsum;
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int et = st + h * 128;
            // D[k][n = et]: k 0..5 dP_i * scale_i, k 6..11 qk_i * qk_scale_i (power-of-two scales bring every row's
            // maximum into [1, 2); the C rows carry the reciprocals)
#pragma unroll
            for (int i = 0; i < kQ; ++i) {
              store_mn_f16(Ds, i, et, dp[h][i] * dscale[i]);
              store_mn_f16(Ds, 6 + i, et, qv[h][i] * qscale[i]);
            }
            // DP[n][k = et]: rows 0..5 hi, 6..11 lo of dP_i * scale_i (fp16), K-major in 4 blocks of 64 k
            float v6[kQ];
#pragma unroll
            for (int i = 0; i < kQ; ++i) v6[i] = dp[h][i] * dscale[i];
            store_col_f16(DPs, et, v6);
          }
        }
        fence_proxy_async_smem();
        g_early = false;                         // (never set across a slide change)
      }
      named_bar_sync(1, 128);                  // scal[0..23] visible to everyone, D / DP complete
      if (kHasG && leader && !g_early) issue_g(it);

      // ---- per-row scalars
      const bool valid = r < ti.nvalid;
      const size_t grow = static_cast<size_t>(ti.row0 + r);
      {
        float c12[12];                         // C row: [a'_i | ds_i]  (mode 2: [dg_i | 0])
        float ds[kQ];                          // DS column (mode 2: dg_i gs)
        float dgmax = 0.f;                     // mode 1: max_i |dg_i| of this row
        if (MODE == kDzNacDkg) {
#pragma unroll
          for (int i = 0; i < kQ; ++i) {
            const float dg = nsc[i];
            c12[i] = dg; c12[6 + i] = 0.f;
            ds[i] = dg * gs;
          }
        } else {
          float sc[kQ], pg[kQ], dmap[kQ];
#pragma unroll
          for (int i = 0; i < kQ; ++i) { sc[i] = nsc[i]; pg[i] = npg[i]; dmap[i] = ndm[i]; }
          mbar_wait(g_bar, tph);
          tc_fence_after();
          uint32_t gv[16];
          tmem_ld_32x32b_x16(tmem_base + kColG + (static_cast<uint32_t>(qd * 32) << 16), gv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < kQ; ++i) {
            const float g = (__uint_as_float(gv[i]) + __uint_as_float(gv[i + 6])) * scal[16 + i] + dmap[i];
            if (MODE == kDzMcat || kLite) {
              const float a = valid ? __expf(sc[i] - scal[8 + i]) : 0.f;
              ds[i] = a * (g - scal[i]);
              c12[i] = a;
            } else {
              // gated softmax with attention dropout: a' = a msc, da = msc (g + dsuma), ds' = a (da - delta)
              const float a = valid ? __expf(sc[i] * pg[i] - scal[8 + i]) : 0.f;
              float msc = 1.f;
              if (p.attn_thr != 0) {
                const uint32_t rb = rng_u32(seed, 1u, static_cast<uint32_t>(i) * static_cast<uint32_t>(p.total_rows) +
                                                          static_cast<uint32_t>(grow)) & 0xFFu;
                msc = rb < p.attn_thr ? 0.f : p.attn_scale;
              }
              const float dsp = a * (msc * (g + scal[160 + i]) - scal[i]);
              ds[i] = dsp * pg[i];
              c12[i] = a * msc;
              const float dgi = 0.5f * dsp * sc[i];
              if (valid) p.dg[static_cast<size_t>(i) * p.total_rows + grow] = dgi;
              dgmax = fmaxf(dgmax, fabsf(dgi));
            }
            c12[6 + i] = ds[i];
          }
        }
        float amax[kQ];
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          float m = fabsf(ds[i]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          amax[i] = m;
        }
        if (MODE == kDzNacDh) {
          // per-tile sums of ds_i (gradient of the key-bias score term kc_i) and the batch-wide max |dg|
#pragma unroll
          for (int i = 0; i < kQ; ++i) {
            float sm = ds[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
            if (lane == 0) scal[168 + qd * 8 + i] = sm;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) dgmax = fmaxf(dgmax, __shfl_xor_sync(0xffffffffu, dgmax, o));
          if (lane == 0 && dgmax > 0.f) atomicMax(p.dg_max, __float_as_uint(dgmax));   // non-negative floats order like uints
        }
        // C entries before the per-tile scale: [a'_i / dp_scale_i | ds_i / qk_scale_i]  (mode 2: [dg_i | -])
        float cu[12];
        float cmax = 0.f;
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          if (MODE == kDzNacDkg) { cu[i] = c12[i]; cu[6 + i] = 0.f; }
          else { cu[i] = c12[i] * scal[16 + i]; cu[6 + i] = c12[6 + i] * scal[24 + i]; }
          cmax = fmaxf(cmax, fmaxf(fabsf(cu[i]), fabsf(cu[6 + i])));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < kQ; ++i) scal[128 + qd * 8 + i] = amax[i];
          scal[128 + qd * 8 + 6] = cmax;
        }
        if (kLite) {
          // the coefficients of dz_n = mask_n * keep_scale * sum_j c12[j] D[j] leave as fp32 (48 B per patch)
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(p.c12_out + grow * 12);
            dst[0] = make_float4(c12[0], c12[1], c12[2], c12[3]);
            dst[1] = make_float4(c12[4], c12[5], c12[6], c12[7]);
            dst[2] = make_float4(c12[8], c12[9], c12[10], c12[11]);
          }
        }
        tc_fence_before();                     // every warp has read G before the leader may issue the next MMA-G
        named_bar_sync(2, 128);                // tile maxima of |ds_i| and of the C entries from the four row warps
        if (!kLite) {
          // C[k][m = r] as fp16 with one power-of-two scale per tile (un-done on the accumulators by the output warps)
          float inv_c;
          const float sc_t = pow2_scale(fmaxf(fmaxf(scal[128 + 6], scal[136 + 6]), fmaxf(scal[144 + 6], scal[152 + 6])), &inv_c);
          if (r == 0) scal[224 + tph] = inv_c;
#pragma unroll
          for (int j = 0; j < (MODE == kDzNacDkg ? 6 : 12); ++j) store_mn_f16(Cs, j, r, cu[j] * sc_t);
        }
        float v6[kQ];
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          if (MODE == kDzNacDkg) {
            v6[i] = ds[i];                     // already scaled by the batch-wide gs
            if (r == 0) scal[208 + tph * 8 + i] = inv_gs;
          } else {
            const float m = fmaxf(fmaxf(scal[128 + i], scal[136 + i]), fmaxf(scal[144 + i], scal[152 + i]));
            float inv;
            const float s = pow2_scale(m, &inv);
            v6[i] = ds[i] * s;
            if (r == 0) scal[208 + tph * 8 + i] = inv;    // un-scale factor for the Q read-back
          }
        }
        store_col_f16(DSs, r, v6);
        if (MODE == kDzNacDh && r < kQ)
          p.part_dkc[static_cast<size_t>(t) * 8 + r] = scal[168 + r] + scal[176 + r] + scal[184 + r] + scal[192 + r];
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(c_bar);
      if (t + 1 < t_end) prefetch_rows(ti_next);
      // MMA-G of the next tile already now (same slide: same dP operand), so that it has retired when that tile's scalars
      // start; every row warp has consumed this tile's G (named barrier 2 above)
      g_early = false;
      if (kHasG && t + 1 < t_end && ti_next.slide == ti.slide) {
        if (leader) { tc_fence_after(); issue_g(it + 1); }
        g_early = true;
      }
    }
  } else {
    // ---------------------------------------------------------------- 8 output warps (row r, column half ch)
    const int qd = warp & 3;                   // TMEM lane quadrant
    const int ch = (warp - 6) >> 2;            // column half (output tile) / feature half (Q, b read-back)
    const int r = qd * 32 + lane;              // patch row of the tile
    float gs = 1.f, inv_gs = 1.f;
    if (MODE == kDzNacDkg) gs = gate_scale(p.dg_max, &inv_gs);
    int it = 0;
    int prev_t = -1;
    // b partial of tile `tt` (its MMA-db is issued behind the NEXT tile's MMA-dZ): read at the end of the next tile, so
    // nobody waits for that MMA -- by then it has long retired
    auto read_db = [&](int tt, uint32_t parity) {
      mbar_wait(b_bar, parity);
      tc_fence_after();
      uint32_t bv[16];
      tmem_ld_32x32b_x16(tmem_base + kColB + ch * 16 + (static_cast<uint32_t>(qd * 32) << 16), bv);
      tmem_ld_wait();
      p.part_db[static_cast<size_t>(tt) * kD + ch * 128 + qd * 32 + lane] = __uint_as_float(bv[0]) * inv_gs;
      tc_fence_before();
    };
    TileInfo ti_next = t_begin < t_end ? p.tile_info[t_begin] : TileInfo{};
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const TileInfo ti = ti_next;
      if (t + 1 < t_end) ti_next = p.tile_info[t + 1];
      const int buf = it % kDzBufs;
      const uint32_t tph = it & 1;
      uint8_t* tile = smem + DzSmem::tile + buf * 65536;
      const bool valid = r < ti.nvalid;
      const size_t grow = static_cast<size_t>(ti.row0 + r);
      mbar_wait(z_bar, tph);
      tc_fence_after();
      uint16_t* grow_out = static_cast<uint16_t*>(p.out) + grow * kD;
      const bool direct = ti.nvalid != kTileM;     // ragged tile: rows are stored by the threads, not by TMA
      const float zs = scal[224 + tph] * (MODE == kDzNacDkg ? gs : p.keep_scale);   // un-scale of Z x the mode's output factor
      if (kLite) {
        // mask bits of this row's 128 features of column half ch: bit = h > 0 (ReLU and dropout zero the same way)
        uint32_t w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j16 = 0; j16 < 16; ++j16) {
          const int g16 = ch * 16 + j16;             // 16-byte chunk (8 features) of the 512 B row
          const int cb = g16 >> 3, jj = g16 & 7;
          const uint4 hv = *reinterpret_cast<const uint4*>(tile + cb * 16384 + r * 128 + ((jj ^ (r & 7)) << 4));
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
          uint32_t bits = 0u;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // fp16 pair: positive <=> sign bit clear and magnitude non-zero
            bits |= ((hw[e] & 0x8000u) == 0u && (hw[e] & 0x7FFFu) != 0u) ? (1u << (2 * e)) : 0u;
            bits |= ((hw[e] & 0x80000000u) == 0u && (hw[e] & 0x7FFF0000u) != 0u) ? (1u << (2 * e + 1)) : 0u;
          }
          w4[j16 >> 2] |= bits << ((j16 & 3) * 8);
        }
        if (valid) *reinterpret_cast<uint4*>(p.mask_out + grow * 8 + ch * 4) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
#pragma unroll 1
      for (int c4 = 0; c4 < ((kLite || (p.debug & 8)) ? 0 : 4); ++c4) {
        const int col0 = ch * 128 + c4 * 32;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + kColZ + (static_cast<uint32_t>(qd * 32) << 16) + col0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const int j16 = (col0 + j) >> 3;
          const int cb = j16 >> 3, jj = j16 & 7;
          uint4* slot = reinterpret_cast<uint4*>(tile + cb * 16384 + r * 128 + ((jj ^ (r & 7)) << 4));
          const uint4 hv = *slot;
          const float2 h0 = unpack_f16x2(hv.x), h1 = unpack_f16x2(hv.y), h2 = unpack_f16x2(hv.z), h3 = unpack_f16x2(hv.w);
          const float hh[8] = {h0.x, h0.y, h1.x, h1.y, h2.x, h2.y, h3.x, h3.y};
          float o8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float z = __uint_as_float(v[j + e]);
            if (MODE == kDzNacDkg) o8[e] = (1.f - hh[e] * hh[e]) * z * zs;
            // NaCAGaT: -0.0 marks a unit masked by ReLU / dropout (bag_dhk_kernel reads the mask from this tile instead
            // of re-reading the saved activations, 0.27 GB); the +0 addend turns a live unit's exact -0 product into +0
            else if (MODE == kDzNacDh) o8[e] = hh[e] > 0.f ? fmaf(z, zs, 0.f) : -0.f;
            else o8[e] = hh[e] > 0.f ? z * zs : 0.f;
          }
          uint4 o;
          if (MODE == kDzNacDkg) {
            o.x = pack_f16x2(o8[0], o8[1]); o.y = pack_f16x2(o8[2], o8[3]);
            o.z = pack_f16x2(o8[4], o8[5]); o.w = pack_f16x2(o8[6], o8[7]);
          } else {
            o.x = pack_bf16x2(o8[0], o8[1]); o.y = pack_bf16x2(o8[2], o8[3]);
            o.z = pack_bf16x2(o8[4], o8[5]); o.w = pack_bf16x2(o8[6], o8[7]);
          }
          *slot = o;
          if (direct && valid) *reinterpret_cast<uint4*>(grow_out + col0 + j) = o;
        }
      }
      // Q partial of this tile: feature f = ch * 128 + qd * 32 + lane
      {
        uint32_t qv[16];
        tmem_ld_32x32b_x16(tmem_base + kColQ + ch * 16 + (static_cast<uint32_t>(qd * 32) << 16), qv);
        tmem_ld_wait();
        float* dst = p.part_dqk + static_cast<size_t>(t) * (kQ * kD) + ch * 128 + qd * 32 + lane;
#pragma unroll
        for (int i = 0; i < kQ; ++i)
          dst[i * kD] = (__uint_as_float(qv[i]) + __uint_as_float(qv[i + 6])) * scal[208 + tph * 8 + i];
      }
      // column sums of the previous output tile, before this tile's w_bar arrival lets the next MMA-db overwrite them
      if (kHasB && prev_t >= 0) read_db(prev_t, tph ^ 1);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(w_bar);
      prev_t = t;
    }
    if (kHasB && prev_t >= 0) read_db(prev_t, (it - 1) & 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// dqk[b][i][d] = sum over the slide's tiles;  db_H[d] += sum over all tiles (gradient accumulation)
// 16 tile groups x 16 float4 columns per block (blockIdx.z = quarter of the columns): all loads of a thread in flight
__global__ void __launch_bounds__(256)
bag_bwd_reduce_kernel(const int* __restrict__ tile_prefix, const float* __restrict__ part_dqk,
                      const float* __restrict__ part_db, float* __restrict__ dqk, float* __restrict__ grad_bias,
                      int B, int num_tiles) {
  pdl_enter();
  __shared__ float4 acc_s[16][16];
  const int tid = threadIdx.x, tg = tid >> 4, dl = tid & 15, dq = dl + 16 * static_cast<int>(blockIdx.z);
  const bool is_bias = static_cast<int>(blockIdx.x) >= B;
  const int b = blockIdx.x, i = blockIdx.y;
  // bias blocks: (gridDim.x - B) * 6 chunks share the tiles round-robin and finish with atomics
  const int nchunk = (static_cast<int>(gridDim.x) - B) * kQ;
  const int chunk = (static_cast<int>(blockIdx.x) - B) * kQ + i;
  const int t0 = is_bias ? chunk * 16 : tile_prefix[b];
  const int t1 = is_bias ? num_tiles : tile_prefix[b + 1];
  const int tstep = is_bias ? nchunk * 16 : 16;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int t = t0 + tg; t < t1; t += tstep) {
    const float* src = is_bias ? part_db + static_cast<size_t>(t) * kD : part_dqk + (static_cast<size_t>(t) * kQ + i) * kD;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + dq);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  acc_s[tg][dl] = acc;
  __syncthreads();
  if (tid < 16) {
    float4 r = acc_s[0][tid];
#pragma unroll
    for (int g = 1; g < 16; ++g) { r.x += acc_s[g][tid].x; r.y += acc_s[g][tid].y; r.z += acc_s[g][tid].z; r.w += acc_s[g][tid].w; }
    if (is_bias) {
      float* g4 = grad_bias + dq * 4;
      atomicAdd(g4 + 0, r.x); atomicAdd(g4 + 1, r.y); atomicAdd(g4 + 2, r.z); atomicAdd(g4 + 3, r.w);
    } else {
      reinterpret_cast<float4*>(dqk + (static_cast<size_t>(b) * kQ + i) * kD)[dq] = r;
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
                  float* __restrict__ grad_w,   // [256][ld] fp32, accumulated
                  int total_rows, int num_splits, int ncb, int ld, uint32_t idesc,
                  const uint32_t* __restrict__ dg_max) {   // non-null: the A operand carries the batch-wide gate scale
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms, by pointer arithmetic so the compiler keeps the shared state space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * kDwStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kDwStages;
  uint64_t* done_bar = bars + 2 * kDwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDwStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb = blockIdx.x % ncb;        // which 256-column block of the B operand / of the gradient
  const int sp = blockIdx.x / ncb;        // split-K index over patch rows
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
      int stage = 0; uint32_t phase = 0;       // idesc: M 128, N 256, A and B both M/N-major, bf16 or fp16 operands
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
      float inv = 1.f;
      if (dg_max != nullptr) (void)gate_scale(dg_max, &inv);
      mbar_wait(done_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int mh = 0; mh < 2; ++mh) {
        const int f = mh * 128 + qd * 32 + lane;
        float* dst = grad_w + static_cast<size_t>(f) * ld + cb * 256;
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + mh * 256 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j),
                         "f"(__uint_as_float(v[j]) * inv), "f"(__uint_as_float(v[j + 1]) * inv),
                         "f"(__uint_as_float(v[j + 2]) * inv), "f"(__uint_as_float(v[j + 3]) * inv)
                         : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// dW_H += dz^T X with dz REGENERATED on the fly (MCAT): no dz buffer in HBM.
//
// A cluster of 4 CTAs owns a contiguous range of 128-patch tiles; CTA cb of the cluster owns the gradient block
// dW_H[:, 256 cb : 256 cb + 256] (all 512 TMEM columns, as in bag_bwd_dw_kernel) and streams the matching X columns.
// The dz operand of a 64-patch chunk is [64 patches x 256 features] bf16, needed in full by all four CTAs: CTA cb
// regenerates ITS feature block 64 cb .. 64 cb + 63 from the per-patch coefficients c12 = [a | ds] (bag_bwd_dz_kernel
// <kDzMcatLite>), the per-slide operand D = [dP ; qk] and the saved mask bits,
//     dz[n][f] = mask[n][f] keep_scale sum_j c12[n][j] D[j][f]            (fp32 FFMA2, then one rounding to bf16)
// and hands the 8 KB box to all four CTAs through the L2: the threads write it (plain row-major, coalesced) into a
// per-CTA scratch ring in global memory, and ONE multicast TMA load issued by the same CTA lands it, swizzled, at the same
// dz-slot offset of every CTA of the cluster and completes on every CTA's dz_full barrier.  Each element is computed once
// per cluster; the scratch (24 KB per CTA) never leaves the L2.
// (Measured dead ends: per-thread st.shared::cluster stores of the box into the four shared memories -- 5x slower than
// the whole old backward, DSMEM takes the stores packet by packet; bulk DSMEM copies shared::cta -> shared::cluster --
// 387 us for the kernel whatever the slot count: 24 KB out + 24 KB in per 64-patch chunk at the ~17 B/clk the SM-to-SM
// network sustains is 2.9 k cycles against 1.44 k cycles of MMA.)  Per CTA: warp 0 TMA (X, 3 stages of 32 KB), warp 1 MMA issuer, warps 2..9 regenerate (and, at the end, warps
// 2..5 flush the accumulators with red.global.add).  db_H[f] = sum_n dz[n][f] falls out of the regenerating threads.
// ------------------------------------------------------------------------------------------------
constexpr int kZStages = 3;
constexpr int kZSlots = 3;                            // dz slots: regenerate + ship (DSMEM) + MMA are three overlapping phases
constexpr int kZXBytes = 4 * kDwBox;                    // X chunk [64 patches x 256 columns]     32 KB
constexpr int kZDzBytes = 4 * kDwBox;                   // dz chunk [64 patches x 256 features]   32 KB
constexpr int kZThreads = 64 + 256;
constexpr int kZC12Bytes = 64 * 48;                     // coefficients of the chunk's 64 patches (fp32 [64][12])
constexpr int kZMaskBytes = 64 * 32;                    // mask bits of the chunk's 64 patches (uint32 [64][8])
constexpr int kZStageBytes = kZXBytes;                  // X chunk
constexpr int kZCStages = 6;                            // coefficient / mask ring: its own, deeper, so that regenerating
constexpr int kZCBytes = kZC12Bytes + kZMaskBytes;      // runs up to kZSlots chunks ahead of the MMAs (5 KB per chunk)
struct DwzSmem {
  static constexpr int x = 0;                           // 3 x 32 KB
  static constexpr int dz = kZStages * kZStageBytes;    // 3 x 32 KB
  static constexpr int cring = dz + kZSlots * kZDzBytes;      // 6 x 5 KB
  static constexpr int red = cring + kZCStages * kZCBytes;    // fp32 [8][64] column-sum partials
  static constexpr int bars = red + 8 * 64 * 4;
  static constexpr int tmem_slot = bars + 256;
  static constexpr int total = tmem_slot + 16;
};
constexpr int kDwzSmemBytes = DwzSmem::total + 1024;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  return ra;
}
// bulk copy local shared memory -> shared memory of another CTA of the cluster; its bytes complete on that CTA's mbarrier
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// 1-D bulk copy global -> shared memory of this CTA, completing on an mbarrier (size and addresses multiples of 16 B)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(kZThreads, 1)
bag_bwd_dwz_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_scr,
                   const BagBwdDwzParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DwzSmem::bars);
  uint64_t* x_full = bars;             // [3] TMA -> MMA
  uint64_t* x_empty = bars + 3;        // [3] MMA -> TMA
  uint64_t* dz_full = bars + 6;        // [3] 1 arrival (expect 4 x 8 KB) + the bytes of the four multicast box loads
  uint64_t* dz_empty = bars + 9;       // [3] 4 arrivals: the MMA issuers of the 4 CTAs (multicast commit)
  uint64_t* done_bar = bars + 12;
  uint64_t* c_full = bars + 13;        // [6] coefficient rows landed
  uint64_t* c_empty = bars + 19;       // [6] 4 arrivals: the regenerating warps of the chunk's group have read them
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + DwzSmem::tmem_slot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cb = cluster_ctarank();                       // column block of dW_H / feature block regenerated here
  const int cluster_id = static_cast<int>(blockIdx.x) >> 2;
  const int nclusters = static_cast<int>(gridDim.x) >> 2;
  const int per = (p.num_tiles + nclusters - 1) / nclusters;
  const int t_begin = min(p.num_tiles, cluster_id * per);
  const int t_end = min(p.num_tiles, t_begin + per);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_scr);
    for (int s = 0; s < kZStages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    for (int s = 0; s < kZCStages; ++s) { mbar_init(&c_full[s], 1); mbar_init(&c_empty[s], 4); }
    for (int s = 0; s < kZSlots; ++s) { mbar_init(&dz_full[s], 1); mbar_init(&dz_empty[s], 4); }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // every CTA's barriers exist before a peer arrives on them or writes its dz slots
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();
      int c = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const TileInfo ti = p.tile_info[t];
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && ti.nvalid <= 64) break;
          const int stage = c % kZStages;
          const uint32_t ph = (c / kZStages) & 1;
          mbar_wait(&x_empty[stage], ph ^ 1);
          uint8_t* sx = smem + DwzSmem::x + stage * kZStageBytes;
          mbar_expect_tx(&x_full[stage], kZXBytes);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_2d(sx + j * kDwBox, &tm_x, &x_full[stage], static_cast<int>(cb) * 256 + j * 64, ti.row0 + half * 64, pol_stream);
          ++c;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 1, 1);      // A (dz^T) and B (X) both MN-major
      int c = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int nvalid = p.tile_info[t].nvalid;
        for (int half = 0; half < 2; ++half) {
          if (half == 1 && nvalid <= 64) break;
          const int stage = c % kZStages, slot = c % kZSlots;
          mbar_wait(&x_full[stage], (c / kZStages) & 1);
          mbar_wait_cluster(&dz_full[slot], (c / kZSlots) & 1);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(smem + DwzSmem::x + stage * kZStageBytes);
          const uint32_t a_addr = smem_u32(smem + DwzSmem::dz + slot * kZDzBytes);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t db = umma_desc_sw128(b_addr + kk * 2048, kDwBox, 1024);
#pragma unroll
            for (int mh = 0; mh < 2; ++mh)
              umma_bf16(tmem_base + mh * 256, umma_desc_sw128(a_addr + mh * 2 * kDwBox + kk * 2048, kDwBox, 1024), db, idesc,
                        (c | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&x_empty[stage]);
          umma_commit_mcast(&dz_empty[slot], 0xF);        // this CTA no longer reads the slot: tell all four producers
          ++c;
        }
      }
      umma_commit(done_bar);
    }
  } else {
    // ---------------------------------------------------------------- regenerating warps 2..9
    const int rw = warp - 2;                 // 0..7: rows 8 rw .. 8 rw + 7 of the chunk
    const int f0 = static_cast<int>(cb) * 64 + 2 * lane;       // the two features this thread regenerates
    const int wsel = lane >> 4;              // which of the CTA's two mask words holds them
    const uint32_t bit0 = 1u << ((2 * lane) & 31);
    // byte offset of the pair inside a k-row of the CTA's box: 16-byte chunk lane >> 2 (swizzled per row), 4 (lane & 3) within
    // this CTA's scratch ring in global memory: kZSlots boxes of [64 patches][64 features] bf16, row-major
    const int scr_row0 = static_cast<int>(blockIdx.x) * (kZSlots * 64);
    uint8_t* scr = reinterpret_cast<uint8_t*>(p.scratch) + static_cast<size_t>(scr_row0) * 128;
    float2 D[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) D[j] = make_float2(0.f, 0.f);
    float bsum0 = 0.f, bsum1 = 0.f;
    int cur_slide = -1;
    int c = 0;
    // coefficient / mask prefetch (thread 0 of the regenerating warps): a second walk over the chunk sequence that stays
    // kZCStages - 1 chunks ahead; the rows of a chunk are contiguous, two 1-D bulk copies per chunk
    int pf_c = 0, pf_t = t_begin, pf_half = 0;
    auto prefetch_one = [&]() {
      if (pf_t >= t_end) return;
      const TileInfo pt = p.tile_info[pf_t];
      const int st = pf_c % kZCStages;
      mbar_wait(&c_empty[st], ((pf_c / kZCStages) & 1) ^ 1);
      const size_t r0 = static_cast<size_t>(pt.row0 + pf_half * 64);
      const uint32_t nrows = static_cast<uint32_t>(min(64, p.total_rows - static_cast<int>(r0)));   // rows that exist
      uint8_t* dst = smem + DwzSmem::cring + st * kZCBytes;
      mbar_expect_tx(&c_full[st], nrows * 80u);
      bulk_load_1d(dst, p.c12 + r0 * 12, nrows * 48u, &c_full[st]);
      bulk_load_1d(dst + kZC12Bytes, p.mask + r0 * 8, nrows * 32u, &c_full[st]);
      ++pf_c;
      if (pf_half == 0 && pt.nvalid > 64) pf_half = 1; else { pf_half = 0; ++pf_t; }
    };
    // Two groups of four warps alternate over the chunks (group g takes the chunks with c % 2 == g, 16 rows per warp): the
    // publication of a chunk -- global stores, proxy fence (waits for the stores' acknowledgement from the L2, ~1.5 k
    // cycles), multicast load -- of one group overlaps the arithmetic of the other.  One group alone was regenerate-bound
    // (2.8 k cycles per chunk against 1.44 k cycles of MMA).
    const int grp = rw >> 2, gw = rw & 3;
    const bool is_pf = rw == 0 && lane == 0;
    if (is_pf) for (int i = 0; i < kZCStages - 2; ++i) prefetch_one();
    for (int t = t_begin; t < t_end; ++t) {
      const TileInfo ti = p.tile_info[t];
      if (ti.slide != cur_slide) {
        cur_slide = ti.slide;
        const size_t sb = static_cast<size_t>(ti.slide) * kQ * kD + f0;
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          D[i] = *reinterpret_cast<const float2*>(p.dpooled + sb + i * kD);
          D[6 + i] = *reinterpret_cast<const float2*>(p.qk + sb + i * kD);
          D[i].x *= p.keep_scale; D[i].y *= p.keep_scale; D[6 + i].x *= p.keep_scale; D[6 + i].y *= p.keep_scale;
        }
      }
      for (int half = 0; half < 2; ++half) {
        if (half == 1 && ti.nvalid <= 64) break;
        if ((c & 1) != grp) { ++c; continue; }
        const int slot = c % kZSlots, stage = c % kZCStages;
        const int nv = min(64, ti.nvalid - half * 64);
        if (is_pf) { prefetch_one(); prefetch_one(); }
        // coefficients and mask words of this warp's 16 rows from the prefetch ring (warp-uniform shared-memory
        // addresses: broadcast reads).  Rows past the tile's end are another slide's: masked to zero.
        mbar_wait(&c_full[stage], (c / kZCStages) & 1);
        const uint8_t* sc = smem + DwzSmem::cring + stage * kZCBytes;
        uint32_t zv[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int kr = gw * 16 + q;
          const float4* src = reinterpret_cast<const float4*>(sc + kr * 48);
          const float4 c0 = src[0], c1 = src[1], c2 = src[2];
          const uint32_t mk = kr < nv ? *reinterpret_cast<const uint32_t*>(sc + kZC12Bytes + kr * 32 + (cb * 2 + wsel) * 4) : 0u;
          const float cv[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
          float z0 = 0.f, z1 = 0.f;
#pragma unroll
          for (int j = 0; j < 12; ++j) { z0 = fmaf(cv[j], D[j].x, z0); z1 = fmaf(cv[j], D[j].y, z1); }
          z0 = (mk & bit0) ? z0 : 0.f;
          z1 = (mk & (bit0 << 1)) ? z1 : 0.f;
          bsum0 += z0; bsum1 += z1;
          zv[q] = pack_bf16x2(z0, z1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&c_empty[stage]);           // this warp is done with the stage's coefficient rows
        mbar_wait(&dz_empty[slot], ((c / kZSlots) & 1) ^ 1);   // all four CTAs have consumed this slot's previous chunk
        // a warp writes one 128-byte row of the box (64 features) per instruction
#pragma unroll
        for (int q = 0; q < 16; ++q)
          *reinterpret_cast<uint32_t*>(scr + (slot * 64 + gw * 16 + q) * 128 + lane * 4) = zv[q];
        // generic-proxy global stores -> visible to the async proxy (the multicast TMA load below reads them back)
        asm volatile("fence.proxy.async.global;" ::: "memory");
        named_bar_sync(1 + grp, 128);
        if (gw == 0 && lane == 0) {
          // one arrival that announces the four boxes (this CTA's and the three peers') landing in this slot
          mbar_expect_tx(&dz_full[slot], 4 * kDwBox);
          tma_load_2d_mcast(smem + DwzSmem::dz + slot * kZDzBytes + cb * kDwBox, &tm_scr, &dz_full[slot], 0,
                            scr_row0 + slot * 64, static_cast<uint16_t>(0xF), policy_evict_last());
        }
        ++c;
      }
    }
    // db_H: column sums of the features regenerated here (each dz element exists exactly once per cluster)
    float* red = reinterpret_cast<float*>(smem + DwzSmem::red);
    red[rw * 64 + 2 * lane] = bsum0;
    red[rw * 64 + 2 * lane + 1] = bsum1;
    named_bar_sync(3, 256);
    if (rw == 0) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) { s0 += red[w * 64 + 2 * lane]; s1 += red[w * 64 + 2 * lane + 1]; }
      if (t_begin < t_end) { atomicAdd(p.grad_b + f0, s0); atomicAdd(p.grad_b + f0 + 1, s1); }
    }
    // accumulators -> gradient (warps 2..5 cover the four TMEM lane quadrants)
    if (rw < 4 && t_begin < t_end) {
      const int qd = warp & 3;
      mbar_wait(done_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int mh = 0; mh < 2; ++mh) {
        const int f = mh * 128 + qd * 32 + lane;
        float* dst = p.grad_w + static_cast<size_t>(f) * kDIn + cb * 256;
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + mh * 256 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + j), "f"(__uint_as_float(v[j])),
                         "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                         : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // nobody leaves while a peer may still store into its dz slots / arrive on its barriers
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int bag_bwd_dwz_max_clusters(int num_sms);

cudaError_t launch_bag_bwd_dwz(const CUtensorMap& tm_x, const CUtensorMap& tm_scr, const BagBwdDwzParams& prm, int num_sms,
                               cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bag_bwd_dwz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwzSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (prm.num_tiles <= 0) return cudaSuccess;
  // clusters of 4 only fit 33 at a time on a B200 (GPC sizes are not multiples of 4: 132 of the 148 SMs); a 34th cluster
  // would run as a second wave and double the kernel (measured), so the grid is what is resident at once
  int clusters = bag_bwd_dwz_max_clusters(num_sms);
  if (clusters > prm.num_tiles) clusters = prm.num_tiles;
  if (clusters < 1) clusters = 1;
  bag_bwd_dwz_kernel<<<clusters * 4, kZThreads, kDwzSmemBytes, stream>>>(tm_x, tm_scr, prm);
  count_launch();
  return cudaGetLastError();
}

// bytes of global scratch the kernel needs for `clusters` clusters (one ring of kZSlots boxes per CTA)
size_t bag_bwd_dwz_scratch_bytes(int clusters) { return static_cast<size_t>(clusters) * 4 * kZSlots * kDwBox; }

int bag_bwd_dwz_max_clusters(int num_sms) {
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cudaFuncSetAttribute(bag_bwd_dwz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwzSmemBytes);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms / 4 * 4);
    cfg.blockDim = dim3(kZThreads);
    cfg.dynamicSmemBytes = kDwzSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, bag_bwd_dwz_kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = num_sms / 4 - 4; }
    max_clusters = n < 1 ? 1 : n;
  }
  return max_clusters;
}

// ------------------------------------------------------------------------------------------------
// dkc[b][i] = sum over the slide's tiles of the per-tile sums of ds_i (one warp per slide and query)
__global__ void bag_bwd_dkc_kernel(const int* __restrict__ tile_prefix, const float* __restrict__ part_dkc,
                                   float* __restrict__ dkc) {
  const int b = blockIdx.x, i = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float v = 0.f;
  for (int t = tile_prefix[b] + lane; t < tile_prefix[b + 1]; t += 32) v += part_dkc[static_cast<size_t>(t) * 8 + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) dkc[b * kQ + i] = v;
}
cudaError_t launch_bag_bwd_dkc(const int* tile_prefix, const float* part_dkc, float* dkc, int B, cudaStream_t stream) {
  if (B <= 0) return cudaSuccess;
  bag_bwd_dkc_kernel<<<B, kQ * 32, 0, stream>>>(tile_prefix, part_dkc, dkc);
  count_launch();
  return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_dz_mode(const CUtensorMap& tm_in, const CUtensorMap& tm_out, const BagBwdDzParams& prm,
                                  int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bag_bwd_dz_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDzSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (prm.num_tiles <= 0) return cudaSuccess;
  const int grid = prm.num_tiles < num_sms ? prm.num_tiles : num_sms;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("MPO_DZ_DEBUG"); dbg = e ? atoi(e) : 0; }
  BagBwdDzParams prm_d = prm;
  prm_d.debug = dbg;
  cudaError_t e = launch_step(bag_bwd_dz_kernel<MODE>, dim3(grid), dim3(kDzThreads), kDzSmemBytes, stream, tm_in, tm_out, prm_d);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}
cudaError_t launch_bag_bwd_dz(int mode, const CUtensorMap& tm_in, const CUtensorMap& tm_out, const BagBwdDzParams& prm,
                              int num_sms, cudaStream_t stream) {
  switch (mode) {
    case kDzMcat: return launch_dz_mode<kDzMcat>(tm_in, tm_out, prm, num_sms, stream);
    case kDzNacDh: return launch_dz_mode<kDzNacDh>(tm_in, tm_out, prm, num_sms, stream);
    case kDzNacDkg: return launch_dz_mode<kDzNacDkg>(tm_in, tm_out, prm, num_sms, stream);
    case kDzMcatLite: return launch_dz_mode<kDzMcatLite>(tm_in, tm_out, prm, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_bag_bwd_reduce(const int* tile_prefix, const float* part_dqk, const float* part_db, float* dqk,
                                  float* grad_bias, int B, int num_tiles, cudaStream_t stream) {
  bag_bwd_reduce_kernel<<<dim3(B + (part_db != nullptr ? 16 : 0), kQ, 4), 256, 0, stream>>>(tile_prefix, part_dqk, part_db,
                                                                                        dqk, grad_bias, B, num_tiles);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_bag_bwd_dw(const CUtensorMap& tm_dz, const CUtensorMap& tm_x, float* grad_w, int total_rows,
                              int ncols, int ld, bool f16, const uint32_t* dg_max, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bag_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (total_rows <= 0) return cudaSuccess;
  const int chunks = (total_rows + kDwBK - 1) / kDwBK;
  const int ncb = ncols / 256;
  int splits = num_sms / ncb;
  if (splits > chunks) splits = chunks;
  if (splits < 1) splits = 1;
  const uint32_t idesc = f16 ? umma_idesc(128, 256, 0, 0, 1, 1) : umma_idesc_bf16(128, 256, 1, 1);
  cudaError_t e = launch_step(bag_bwd_dw_kernel, dim3(ncb * splits), dim3(kDwThreads), kDwSmemBytes, stream, tm_dz, tm_x, grad_w,
                              total_rows, splits, ncb, ld, idesc, dg_max);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mpo
