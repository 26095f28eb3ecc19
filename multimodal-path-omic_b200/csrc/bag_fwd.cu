// Bag-pass forward for MCAT (reference: models/mcat/mcat.py:87 `H_bag = self.H(wsi)` and :97 co-attention).
//
// One persistent CTA per SM streams a range of CONSECUTIVE 128-patch tiles of the packed bf16 bag (so the folded queries
// of a slide are loaded once per CTA and slide, not once per tile); everything GEMM-shaped runs on tcgen05:
//   TMA (warp 0)  : X tile [128 x 1024] and W_H [256 x 1024] in 64-wide K blocks, 128B-swizzled, 3-stage ring
//   MMA (warp 1)  : H = X W_H^T, tcgen05.mma M=128 N=256 K=16, fp32 accumulators in TMEM, two accumulator stages
//   epilogue (16 w): (1) TMEM -> registers; +bias, ReLU, (dropout); the H tile is written ONCE to shared memory as
//                       fp16 in the canonical 128B-swizzle layout (and TMA-stored to HBM when the backward needs it);
//                   (2) the six folded-query dots per patch are taken in the same register pass, in fp32 (an fp16
//                       tensor-core product would cost the attention map its 1e-3 parity on sharp softmaxes);
//                   (3) tile-local softmax statistics (warp shuffles), weights p written as a tiny fp16 B operand;
//                   (4) pooled[6 x 256] = p^T H: a second tcgen05.mma on the staged bytes, read M-major
//                       (A = H^T, N = 16, K = 128 patches), read back one feature per thread -> per-tile partial.
// The K/V projections of the reference's nn.MultiheadAttention are folded away exactly:
//   score_in = h_n . (W_k^T q_i)/sqrt(d)   (the b_k term is constant over n and cancels in the softmax)
//   out_i    = W_v (sum_n a_in h_n) + b_v  (done in the tail on the pooled vector)
// A second kernel merges the per-tile partials by log-sum-exp per slide.
#include "mpo_ptx.cuh"
#include "mpo_common.cuh"
#include "launchers.h"
#include "tail_dev.cuh"
#include <cstdlib>

namespace mpo {

constexpr int kStages = 3;
constexpr int kABytes = kTileM * kBK * 2;            // 16 KB
constexpr int kBBytes = kD * kBK * 2;                // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;       // 48 KB
constexpr int kStagingBytes = kTileM * kD * 2;       // 64 KB  fp16 H tile, K-major SW128 (4 blocks of 128x64)
constexpr int kEpiThreads = 512;                     // 16 epilogue warps: 4 per TMEM lane quadrant, 64 columns each
constexpr int kCG = kEpiThreads / 128;               // column groups
constexpr int kFwdThreads = 64 + kEpiThreads;

struct FwdSmem {
  // offsets from the 1024-aligned base
  static constexpr int stages = 0;
  static constexpr int staging = kStages * kStageBytes;                 // 147456: fp16 H tile [4][128][64]
  static constexpr int Pb = staging + kStagingBytes;                    // fp16 [2 k-blocks][16 rows][64]  (4 KB)
  static constexpr int qk = Pb + 2 * 16 * 128;                          // fp32 [6][256] folded queries of the slide
  static constexpr int bias = qk + kQ * kD * 4;                         // fp32 [256]
  static constexpr int spart = bias + kD * 4;                           // fp32 [128][8] score partials of column half 1
  static constexpr int wred = spart + kTileM * 8 * 4;                   // fp32 [2][4][8]
  static constexpr int bars = wred + 2 * 4 * 8 * 4;                     // mbarriers
  static constexpr int tmem_slot = bars + 16 * 8;
  static constexpr int total = tmem_slot + 16;
};
constexpr int kFwdSmemBytes = FwdSmem::total + 1024;  // + slack for manual 1024 B alignment


// kC = CTAs per cluster.  With kC > 1 every CTA still owns its own tiles, but the cluster walks the K blocks in
// lock step and each CTA fetches only 1/kC of every W_H block, multicasting it to all kC shared memories: the L2 ->
// SM traffic for the weights drops by kC (W_H re-reads, not the bag, are what saturates the L2 slices at kC = 1).
// kPair (kC = 2): the two CTAs share every MMA (tcgen05 cta_group::2, M = 256 = two adjacent tiles): each CTA stages its own
// X block and HALF of the W_H block (four 32 KB stages instead of three of 48 KB), rank 0 issues all MMAs; the pooled
// product is one N = 32 pair MMA over both CTAs' weight operands (scripts/probes/pair_mma.cu).
template <int kC, bool kPair = false>
__global__ void __launch_bounds__(kFwdThreads, 1)
bag_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_h, const BagFwdParams p) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms, by pointer arithmetic so the compiler keeps the shared state space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::bars);
  constexpr int kNS = kPair ? 4 : kStages;                  // ring depth
  constexpr int kSB = kPair ? 2 * kABytes : kStageBytes;    // bytes per stage in this CTA
  uint64_t* full_bar = bars;                 // [kNS]   (pair: only rank 0's are used)
  uint64_t* empty_bar = bars + kNS;          // [kNS]
  uint64_t* tfull_bar = bars + 2 * kNS;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]     (pair: only rank 0's are used, both CTAs' warps arrive)
  uint64_t* d_bar = tempty_bar + 2;          // pooled MMA done
  uint64_t* pready_bar = d_bar + 1;          // pair, rank 0: both CTAs' pooled operands are in place
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FwdSmem::tmem_slot);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int s = 0; s < kNS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kPair ? 1 : kC);     // one tcgen05.commit arrival from every MMA-issuing CTA
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], (kPair ? 2 : 1) * (kEpiThreads / 32));
    }
    mbar_init(d_bar, 1);
    mbar_init(pready_bar, 2);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (kPair) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (kC > 1) cluster_sync_all();     // remote CTAs' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t cta_rank = kC > 1 ? cluster_ctarank() : 0;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << kC) - 1);
  // every CTA of a cluster runs the same number of pipeline iterations; surplus ones are dummies without MMAs
  // work units: tiles, or tile pairs (2u, 2u + 1 -> rank 0, 1) in pair mode
  const int nunits = kPair ? p.num_tiles / 2 : p.num_tiles;
  const int nworkers = kPair ? static_cast<int>(gridDim.x) / 2 : static_cast<int>(gridDim.x);
  const int iters = (nunits + nworkers - 1) / nworkers;
  // a CTA owns `iters` CONSECUTIVE tiles: it stays inside one slide for up to `iters` tiles, so the folded queries are
  // reloaded once or twice per CTA instead of once per tile (a stride of gridDim tiles lands in a new slide every time)
  const int unit0 = (kPair ? static_cast<int>(blockIdx.x) / 2 : static_cast<int>(blockIdx.x)) * iters;
  const int nreal = max(0, min(iters, nunits - unit0));
  auto tile_at = [&](int it) { return kPair ? 2 * (unit0 + it) + static_cast<int>(cta_rank) : unit0 + it; };

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first();   // the bag is read once per pass
      const uint64_t pol_keep = policy_evict_last();      // W_H is re-read by every tile: keep it in L2
      int stage = 0;
      uint32_t phase = 0;
      constexpr int kWRows = kD / kC;                     // W_H rows this CTA fetches per K block
      int row_pref = nreal > 0 ? p.tile_info[tile_at(0)].row0 : -1;        // loaded one tile ahead of its use
      for (int it = 0; it < iters; ++it) {
        int row0 = row_pref >= 0 ? row_pref : 0;
        if (p.debug & 2) row0 = (blockIdx.x & 7) * kTileM;
        row_pref = it + 1 < nreal ? p.tile_info[tile_at(it + 1)].row0 : -1;
        const int row_next = row_pref;
        for (int kb = 0; kb < kKBlocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + FwdSmem::stages + stage * kSB;
          if (kPair) {
            if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * kSB);      // the bytes of both CTAs
            tma2_load_2d(sa, &tm_x, &full_bar[stage], kb * kBK, row0, pol_stream);
            tma2_load_2d(sa + kABytes, &tm_w, &full_bar[stage], kb * kBK, static_cast<int>(cta_rank) * (kD / 2), pol_keep);
            if (++stage == kNS) { stage = 0; phase ^= 1; }
            continue;
          }
          const bool skip_w = (p.debug & 1) && it > 0;
          if ((p.debug & 8) && it > 0) {            // timing experiment: no TMA traffic at all
            mbar_arrive(&full_bar[stage]);
            if (++stage == kNS) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full_bar[stage], skip_w ? kABytes : kStageBytes);
          tma_load_2d(sa, &tm_x, &full_bar[stage], kb * kBK, row0, pol_stream);
          if ((p.debug & 4) && row_next >= 0) tma_prefetch_l2_2d(&tm_x, kb * kBK, row_next);
          if (skip_w) {
          } else if (kC == 1) {
            tma_load_2d(sa + kABytes, &tm_w, &full_bar[stage], kb * kBK, 0, pol_keep);
          } else {
            tma_load_2d_mcast(sa + kABytes + cta_rank * (kWRows * 128), &tm_w, &full_bar[stage], kb * kBK,
                              cta_rank * kWRows, kMask, pol_keep);
          }
          if (++stage == kNS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0 && (!kPair || cta_rank == 0)) {
      constexpr uint32_t idesc = umma_idesc_bf16(kPair ? 2 * kTileM : kTileM, kD, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        const bool real = it < nreal;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        if (real) {
          mbar_wait(&tempty_bar[as], aphase ^ 1);   // epilogue has drained this accumulator stage
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + as * kD;
        for (int kb = 0; kb < kKBlocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (real && !(p.debug & 16)) {            // (debug bit 4: timing experiment without the main MMAs)
            const uint32_t a_addr = smem_u32(smem + FwdSmem::stages + stage * kSB);
            const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint64_t da = umma_desc_sw128(a_addr + k * 32, 16, 1024);
              const uint64_t db = umma_desc_sw128(b_addr + k * 32, 16, 1024);
              if (kPair) umma2_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          // frees this smem stage -- in every CTA of the cluster -- once these MMAs retire
          if (kPair) umma2_commit_mcast(&empty_bar[stage], kMask);
          else if (kC == 1) umma_commit(&empty_bar[stage]);
          else umma_commit_mcast(&empty_bar[stage], kMask);
          if (++stage == kNS) { stage = 0; phase ^= 1; }
        }
        if (real) {                                 // accumulator complete -> epilogue (of both CTAs in pair mode)
          if (kPair) umma2_commit_mcast(&tfull_bar[as], kMask);
          else umma_commit(&tfull_bar[as]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps (2..17)
    const int et = threadIdx.x - 64;           // 0..511
    const int qd = warp & 3;                   // TMEM lane quadrant this warp may access
    const int ch = (warp - 2) >> 2;            // column group (64 columns) in step (1); feature half in step (4) when < 2
    const int r = qd * 32 + lane;              // tile row (patch) owned in steps (1)-(3)
    float* bias_s = reinterpret_cast<float*>(smem + FwdSmem::bias);
    float* qk_s = reinterpret_cast<float*>(smem + FwdSmem::qk);
    float* spart_s = reinterpret_cast<float*>(smem + FwdSmem::spart);
    float* wmax_s = reinterpret_cast<float*>(smem + FwdSmem::wred);
    float* wsum_s = wmax_s + 4 * 8;
    uint8_t* staging = smem + FwdSmem::staging;
    uint8_t* Pb = smem + FwdSmem::Pb;
    constexpr uint32_t idesc_d = umma_idesc(kTileM, 16, 0, 0, 1, 0);   // D2 = H^T p^T  : fp16, A M-major, B K-major

    if (et < kD) {
      bias_s[et] = p.bias[et];
      // zero the small B operand once.  Rows 12..15 (the N padding) may later hold score partials (step 2 borrows the
      // buffer): they only feed accumulator columns 12..15, which nobody reads
      *reinterpret_cast<uint4*>(Pb + et * 16) = make_uint4(0, 0, 0, 0);
    }
    const uint32_t seed = p.seed_dev != nullptr ? (p.seed ^ __ldg(p.seed_dev)) : p.seed;
    int cur_slide = -1;
    int it = 0;
    int prev_t = -1;
    // pooled partial of tile `tt` (accumulator stage `pas`): read one iteration late, so that nobody waits for the
    // pooled MMA -- it retires while the next tile's accumulators are being awaited -- then hand the stage back
    auto read_pooled = [&](int tt, int pas, uint32_t parity) {
      mbar_wait(d_bar, parity);
      tc_fence_after();
      if (ch < 2) {
        uint32_t dv[16];
        tmem_ld_32x32b_x16(tmem_base + pas * kD + (static_cast<uint32_t>(qd * 32) << 16) +
                               (kPair ? ch * 32 + 16 * static_cast<int>(cta_rank) : ch * 16), dv);
        tmem_ld_wait();
        float* dst = p.part_pool + static_cast<size_t>(tt) * (kQ * kD) + ch * 128 + qd * 32 + lane;
#pragma unroll
        for (int i = 0; i < kQ; ++i) dst[i * kD] = __uint_as_float(dv[i]) + __uint_as_float(dv[i + 6]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && !(p.debug & 256)) { if (kPair) mbar_arrive_pair0(&tempty_bar[pas]); else mbar_arrive(&tempty_bar[pas]); }
    };
    TileInfo ti_next = nreal > 0 ? p.tile_info[tile_at(0)] : TileInfo{};
    for (; it < nreal; ++it) {
      const int t = tile_at(it);
      const TileInfo ti = ti_next;
      if (it + 1 < nreal) ti_next = p.tile_info[tile_at(it + 1)];     // in flight during this tile's epilogue
      if (prev_t >= 0) read_pooled(prev_t, (it - 1) & 1, (it - 1) & 1);   // also: the staged tile / P operand are free again
      prev_t = -1;
      if (et == 32) tma_store_wait_read();      // the previous tile's H store no longer reads the staged tile
      if (ti.slide != cur_slide) {
        cur_slide = ti.slide;
        const float* src = p.qk + static_cast<size_t>(ti.slide) * kQ * kD;
#pragma unroll
        for (int j = 0; j < kQ * kD / kEpiThreads; ++j) qk_s[et + j * kEpiThreads] = src[et + j * kEpiThreads];
      }
      named_bar_sync(1, kEpiThreads);           // qk visible; staging / P operand of the previous tile are free

      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const uint32_t tphase = it & 1;
      const uint32_t acc_col = tmem_base + as * kD;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      if (p.debug & 32) {                       // timing experiment: loads + MMAs alone, no epilogue work
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (kPair) mbar_arrive_pair0(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
        continue;
      }

      // ---- (1)+(2) accumulator -> h = relu(acc + bias) (dropout); fp32 score partials; fp16 tile in shared memory
      float s[kQ];
#pragma unroll
      for (int i = 0; i < kQ; ++i) s[i] = 0.f;
      const uint32_t grow = static_cast<uint32_t>(ti.row0 + r);
#pragma unroll 1
      for (int c4 = 0; c4 < 8 / kCG; ++c4) {
        const int col0 = ch * (kD / kCG) + c4 * 32;
        uint32_t v[32];
        if (p.debug & 512) {          // timing experiment (wrong results): no accumulator reads
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0x3f000000u + static_cast<uint32_t>(j + lane);
        } else {
          tmem_ld_32x32b_x32(acc_col + (static_cast<uint32_t>(qd * 32) << 16) + col0, v);
          tmem_ld_wait();
        }
        uint4 lo_even = make_uint4(0u, 0u, 0u, 0u);     // remainder chunk of the even 8-column group, stored with the odd one
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float h[8];
          const float4 b0 = *reinterpret_cast<const float4*>(bias_s + col0 + j);
          const float4 b1 = *reinterpret_cast<const float4*>(bias_s + col0 + j + 4);
          h[0] = fmaxf(__uint_as_float(v[j + 0]) + b0.x, 0.f);
          h[1] = fmaxf(__uint_as_float(v[j + 1]) + b0.y, 0.f);
          h[2] = fmaxf(__uint_as_float(v[j + 2]) + b0.z, 0.f);
          h[3] = fmaxf(__uint_as_float(v[j + 3]) + b0.w, 0.f);
          h[4] = fmaxf(__uint_as_float(v[j + 4]) + b1.x, 0.f);
          h[5] = fmaxf(__uint_as_float(v[j + 5]) + b1.y, 0.f);
          h[6] = fmaxf(__uint_as_float(v[j + 6]) + b1.z, 0.f);
          h[7] = fmaxf(__uint_as_float(v[j + 7]) + b1.w, 0.f);
          if (p.drop_thr != 0) {
            // one 32-bit draw covers four consecutive features of this patch row
            const uint32_t base = (grow * kD + col0 + j) >> 2;
            const uint32_t r0 = rng_u32(seed, 0u, base);
            const uint32_t r1 = rng_u32(seed, 0u, base + 1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              h[e] = (((r0 >> (8 * e)) & 0xFFu) < p.drop_thr) ? 0.f : h[e] * p.drop_scale;
              h[4 + e] = (((r1 >> (8 * e)) & 0xFFu) < p.drop_thr) ? 0.f : h[4 + e] * p.drop_scale;
            }
          }
#pragma unroll
          for (int i = 0; i < kQ; ++i) {
            const float4 q0 = *reinterpret_cast<const float4*>(qk_s + i * kD + col0 + j);
            const float4 q1 = *reinterpret_cast<const float4*>(qk_s + i * kD + col0 + j + 4);
            s[i] = fmaf(h[0], q0.x, s[i]);
            s[i] = fmaf(h[1], q0.y, s[i]);
            s[i] = fmaf(h[2], q0.z, s[i]);
            s[i] = fmaf(h[3], q0.w, s[i]);
            s[i] = fmaf(h[4], q1.x, s[i]);
            s[i] = fmaf(h[5], q1.y, s[i]);
            s[i] = fmaf(h[6], q1.z, s[i]);
            s[i] = fmaf(h[7], q1.w, s[i]);
          }
          uint4 pk;
          pk.x = pack_f16x2(h[0], h[1]);
          pk.y = pack_f16x2(h[2], h[3]);
          pk.z = pack_f16x2(h[4], h[5]);
          pk.w = pack_f16x2(h[6], h[7]);
          if (p.h_lo_out != nullptr && grow < static_cast<uint32_t>(p.total_rows)) {
            // NaCAGaT: the fp16 remainder h - fp16(h), so that the key projection K = H W_k^T (tensor cores, in
            // bag_gate_kernel) sees ~22-bit activations: its tanh gate dots feed a softmax argument directly
            const float2 a0 = unpack_f16x2(pk.x), a1 = unpack_f16x2(pk.y), a2 = unpack_f16x2(pk.z), a3 = unpack_f16x2(pk.w);
            uint4 lo;
            lo.x = pack_f16x2(h[0] - a0.x, h[1] - a0.y);
            lo.y = pack_f16x2(h[2] - a1.x, h[3] - a1.y);
            lo.z = pack_f16x2(h[4] - a2.x, h[5] - a2.y);
            lo.w = pack_f16x2(h[6] - a3.x, h[7] - a3.y);
            // the thread owns the row: two chunks = one full 32-byte sector per store instead of two half-filled ones
            if ((j & 8) == 0) lo_even = lo;
            else st_global_256(p.h_lo_out + static_cast<size_t>(grow) * kD + col0 + j - 8, lo_even, lo);
          }
          const int j16 = (col0 + j) >> 3;           // 16-byte chunk index within the 512 B row
          const int cb = j16 >> 3, jj = j16 & 7;     // 64-feature block, chunk within the 128 B swizzle row
          *reinterpret_cast<uint4*>(staging + cb * (kTileM * 128) + r * 128 + ((jj ^ (r & 7)) << 4)) = pk;
        }
      }
      // score partials of the four column groups -> group 0, in two levels (two 4 KB scratch sets: spart and, until
      // step 3 rewrites it, the P operand buffer): groups 1 and 3 publish, 0 and 2 add; group 2 publishes, 0 adds
      float* scr = (ch == 1 || ch == 2) ? spart_s : reinterpret_cast<float*>(Pb);
      if (ch == 1 || ch == 3) {
        *reinterpret_cast<float4*>(scr + r * 8) = make_float4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<float2*>(scr + r * 8 + 4) = make_float2(s[4], s[5]);
      }
      fence_proxy_async_smem();      // generic-proxy writes of the tile -> visible to UMMA / TMA (async proxy)
      tc_fence_before();
      if (p.debug & 256) {           // timing experiment (WRONG pooled vectors): the accumulator stage goes back to the MMA
        __syncwarp();                // warp as soon as it has been drained, not after the pooled product has been read
        if (lane == 0) { if (kPair) mbar_arrive_pair0(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
      }
      named_bar_sync(1, kEpiThreads);
      if (ch == 0 || ch == 2) {
        const float* src = ch == 0 ? spart_s : reinterpret_cast<const float*>(Pb);
        const float4 o0 = *reinterpret_cast<const float4*>(src + r * 8);
        const float2 o1 = *reinterpret_cast<const float2*>(src + r * 8 + 4);
        s[0] += o0.x; s[1] += o0.y; s[2] += o0.z; s[3] += o0.w; s[4] += o1.x; s[5] += o1.y;
      }
      named_bar_sync(3, kEpiThreads);
      if (ch == 2) {
        *reinterpret_cast<float4*>(spart_s + r * 8) = make_float4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<float2*>(spart_s + r * 8 + 4) = make_float2(s[4], s[5]);
      }
      named_bar_sync(1, kEpiThreads);
      if ((p.debug & 128) && p.h_out != nullptr) {
        // timing experiment: the saved activations leave through the LSU instead of the TMA unit.  The 12 warps that
        // do not run the tile soft-max copy the staged tile, one 512 B patch row per instruction (lane = 16 B chunk)
        if (ch != 0) {
          const int cbl = lane >> 3, jl = lane & 7;
          uint4 hv[11];
#pragma unroll
          for (int k = 0; k < 11; ++k) {
            const int row = (warp - 6) + 12 * k;
            if (row < kTileM)
              hv[k] = *reinterpret_cast<const uint4*>(staging + cbl * (kTileM * 128) + row * 128 + ((jl ^ (row & 7)) << 4));
          }
#pragma unroll
          for (int k = 0; k < 11; ++k) {
            const int row = (warp - 6) + 12 * k;
            if (row < ti.nvalid)
              __stcs(reinterpret_cast<uint4*>(p.h_out + static_cast<size_t>(ti.row0 + row) * kD + lane * 8), hv[k]);
          }
        }
      } else if (et == 32 && p.h_out != nullptr) {
        // keep the activations for the backward pass: one TMA store per 64-feature block, straight from the tile
#pragma unroll
        // evict-first: the tile is next read by the backward pass, a bag-pass (> L2) later; measured 4 % on the kernel
        const uint64_t pol_store = policy_evict_first();
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
          if (p.debug & 64) tma_store_2d(&tm_h, staging + cb * (kTileM * 128), cb * 64, ti.row0);
          else tma_store_2d_hint(&tm_h, staging + cb * (kTileM * 128), cb * 64, ti.row0, pol_store);
        }
        tma_store_commit();
      }

      // ---- (3) tile-local softmax statistics; one patch row per thread of column half 0
      const bool valid = r < ti.nvalid;
      if (ch == 0) {
        const float4 o0 = *reinterpret_cast<const float4*>(spart_s + r * 8);
        const float2 o1 = *reinterpret_cast<const float2*>(spart_s + r * 8 + 4);
        s[0] += o0.x; s[1] += o0.y; s[2] += o0.z; s[3] += o0.w; s[4] += o1.x; s[5] += o1.y;
#pragma unroll
        for (int i = 0; i < kQ; ++i)
          if (valid) p.scores[static_cast<size_t>(i) * p.total_rows + ti.row0 + r] = s[i];
      }
      if (p.skip_pool) {
        // NaCAGaT: the softmax runs on the gated scores in bag_gate_kernel; this pass only projects and scores
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (kPair) mbar_arrive_pair0(&tempty_bar[as]); else mbar_arrive(&tempty_bar[as]); }
        continue;
      }
      if (ch == 0) {
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          float m = valid ? s[i] : -INFINITY;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          if (lane == 0) wmax_s[qd * 8 + i] = m;
        }
        named_bar_sync(2, 128);
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          const float m = fmaxf(fmaxf(wmax_s[i], wmax_s[8 + i]), fmaxf(wmax_s[16 + i], wmax_s[24 + i]));
          const float pr = valid ? __expf(s[i] - m) : 0.f;
          // p^T as a B operand [16 rows x 128 patches], K-major (patches), 128B swizzle, 2 blocks of 64 patches.
          // fp16 has 11 bits: rows 0..5 carry the leading part, rows 6..11 the remainder (the N = 16 tile has room),
          // so the pooled sum sees ~22-bit weights and tiny bags keep their 1e-3 parity.
          const __half ph = __float2half_rn(pr);
          const __half pl = __float2half_rn(pr - __half2float(ph));
          uint8_t* pcol = Pb + (r >> 6) * 2048 + (r & 7) * 2;
          *reinterpret_cast<__half*>(pcol + i * 128 + ((((r & 63) >> 3) ^ i) << 4)) = ph;
          *reinterpret_cast<__half*>(pcol + (i + 6) * 128 + ((((r & 63) >> 3) ^ ((i + 6) & 7)) << 4)) = pl;
          if (i < 4) {   // rows 12..15 held score partials: clear this patch's column (keeps the padding finite)
            *reinterpret_cast<__half*>(pcol + (i + 12) * 128 + ((((r & 63) >> 3) ^ ((i + 12) & 7)) << 4)) = __float2half_rn(0.f);
          }
          float l = __half2float(ph) + __half2float(pl);   // the sum uses the weights the MMA will see
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
          if (lane == 0) wsum_s[qd * 8 + i] = l;
          if (et == 0) p.part_ml[static_cast<size_t>(t) * 12 + i] = m;
        }
        fence_proxy_async_smem();
        tc_fence_before();
      }
      named_bar_sync(1, kEpiThreads);
      if (et < kQ) {
        p.part_ml[static_cast<size_t>(t) * 12 + 6 + et] =
            wsum_s[et] + wsum_s[8 + et] + wsum_s[16 + et] + wsum_s[24 + et];
      }

      // ---- (4) pooled = p^T H on the tensor core: A = the same staged bytes read M-major (features), K = 128 patches
      if (kPair) {
        // one pair MMA, N = 32: rank r's weight rows are B rows 16r..16r+15, so its product sits in columns 16r..16r+15
        if (et == 0) {
          mbar_arrive_pair0(pready_bar);
          if (cta_rank == 0) {
            mbar_wait(pready_bar, tphase);
            tc_fence_after();
            constexpr uint32_t idesc_p = umma_idesc(2 * kTileM, 32, 0, 0, 1, 0);
            const uint32_t a0 = smem_u32(staging), b0 = smem_u32(Pb);
#pragma unroll
            for (int mh = 0; mh < 2; ++mh)
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma2_f16(acc_col + mh * 32,
                          umma_desc_sw128(a0 + mh * 2 * (kTileM * 128) + kk * 2048, kTileM * 128, 1024),
                          umma_desc_sw128(b0 + (kk >> 2) * 2048 + (kk & 3) * 32, 16, 1024), idesc_p, kk != 0 ? 1u : 0u);
            umma2_commit_mcast(d_bar, kMask);
          }
        }
      } else if (et == 0) {
        tc_fence_after();
        const uint32_t a0 = smem_u32(staging), b0 = smem_u32(Pb);
#pragma unroll
        for (int mh = 0; mh < 2; ++mh)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16(acc_col + mh * 16,
                      umma_desc_sw128(a0 + mh * 2 * (kTileM * 128) + kk * 2048, kTileM * 128, 1024),
                      umma_desc_sw128(b0 + (kk >> 2) * 2048 + (kk & 3) * 32, 16, 1024), idesc_d, kk != 0 ? 1u : 0u);
        umma_commit(d_bar);
      }
      prev_t = t;
    }
    if (prev_t >= 0) read_pooled(prev_t, (it - 1) & 1, (it - 1) & 1);
    if (et == 32) tma_store_wait_read();
  }

  tc_fence_before();
  __syncthreads();
  if (kC > 1) cluster_sync_all();     // nobody leaves while a peer may still multicast into / arrive on its smem
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// LSE merge of per-tile partials -> pooled[B][6][256] (normalised) and lse[B][6]
// (the same combine is reused across GPUs for a patch-range-sharded bag, see lse_combine in bag_aux.cu)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bag_merge_kernel(const int* __restrict__ tile_prefix,   // [B+1] first tile of each slide
                 const float* __restrict__ part_ml, const int mls,   // per-tile stats, mls floats per tile (12 or 18)
                 const float* __restrict__ part_pool, float* __restrict__ pooled, float* __restrict__ lse,
                 float* __restrict__ suma,     // suma (NaCAGaT, mls = 18): sum_n of the dropped-and-rescaled weights
                 const float* __restrict__ part_pool2, float* __restrict__ pooled2) {   // optional second partial set
  pdl_enter();
  // one block per (slide, query, quarter of the feature columns); 16 tile groups x 16 float4 columns: a 128-tile slide
  // is 8 independent loads per thread, all in flight at once
  __shared__ float red[8];
  __shared__ float4 acc_s[16][16];
  const int b = blockIdx.x, i = blockIdx.y, tid = threadIdx.x;
  const int t0 = tile_prefix[b], t1 = tile_prefix[b + 1];
  float m = -INFINITY;
  for (int t = t0 + tid; t < t1; t += 256) m = fmaxf(m, __ldg(part_ml + static_cast<size_t>(t) * mls + i));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  float M = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) M = fmaxf(M, red[w]);
  __syncthreads();
  float l = 0.f;
  for (int t = t0 + tid; t < t1; t += 256)
    l = fmaf(__ldg(part_ml + static_cast<size_t>(t) * mls + 6 + i), __expf(__ldg(part_ml + static_cast<size_t>(t) * mls + i) - M), l);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if ((tid & 31) == 0) red[tid >> 5] = l;
  __syncthreads();
  float L = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) L += red[w];
  const int tg = tid >> 4, dl = tid & 15, dq = dl + 16 * static_cast<int>(blockIdx.z);
  for (int set = 0; set < (part_pool2 != nullptr ? 2 : 1); ++set) {      // block-uniform
    const float* pp = set == 0 ? part_pool : part_pool2;
    float* out = set == 0 ? pooled : pooled2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int t = t0 + tg; t < t1; t += 16) {
      const float w = __expf(__ldg(part_ml + static_cast<size_t>(t) * mls + i) - M);
      const float4 v = __ldg(reinterpret_cast<const float4*>(pp + (static_cast<size_t>(t) * kQ + i) * kD) + dq);
      acc.x = fmaf(v.x, w, acc.x); acc.y = fmaf(v.y, w, acc.y); acc.z = fmaf(v.z, w, acc.z); acc.w = fmaf(v.w, w, acc.w);
    }
    if (set) __syncthreads();
    acc_s[tg][dl] = acc;
    __syncthreads();
    if (tid < 16) {
      const float inv = 1.f / L;
      float4 r = acc_s[0][tid];
#pragma unroll
      for (int g = 1; g < 16; ++g) { r.x += acc_s[g][tid].x; r.y += acc_s[g][tid].y; r.z += acc_s[g][tid].z; r.w += acc_s[g][tid].w; }
      r.x *= inv; r.y *= inv; r.z *= inv; r.w *= inv;
      reinterpret_cast<float4*>(out + (static_cast<size_t>(b) * kQ + i) * kD)[dq] = r;
    }
  }
  if (tid == 0 && blockIdx.z == 0) lse[b * kQ + i] = M + __logf(L);
  if (suma != nullptr && blockIdx.z == 0) {       // block-uniform
    __syncthreads();
    float ld = 0.f;
    for (int t = t0 + tid; t < t1; t += 256)
      ld = fmaf(__ldg(part_ml + static_cast<size_t>(t) * mls + 12 + i), __expf(__ldg(part_ml + static_cast<size_t>(t) * mls + i) - M), ld);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
    if ((tid & 31) == 0) red[tid >> 5] = ld;
    __syncthreads();
    if (tid == 0) {
      float LD = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) LD += red[w];
      suma[b * kQ + i] = LD / L;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers (C++ side; the extern "C" surface is in api.cu)
// ------------------------------------------------------------------------------------------------
template <int kC, bool kPair = false>
static cudaError_t launch_fwd_cluster(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const CUtensorMap& tm_h,
                                      const BagFwdParams& prm,
                                      int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  static int max_clusters = 0;
  auto kern = bag_fwd_kernel<kC, kPair>;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kFwdThreads);
  cfg.dynamicSmemBytes = kFwdSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes);
    if (e != cudaSuccess) return e;
    cfg.gridDim = dim3(num_sms / kC * kC);
    e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess) return e;
    if (max_clusters < 1) return cudaErrorInvalidConfiguration;
    attr_set = true;
  }
  int clusters = (prm.num_tiles + kC - 1) / kC;
  if (clusters > max_clusters) clusters = max_clusters;
  cfg.gridDim = dim3(clusters * kC);
  cfg.numAttrs = step_pdl_attr(attr, 1, stream);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm_x, tm_w, tm_h, prm);
  count_launch();
  return e;
}

int fwd_cluster_size() {
  static int c = -1;
  if (c < 0) {
    const char* env = getenv("MPO_FWD_CLUSTER");
    c = env ? atoi(env) : 2;
    if (c != 1 && c != 2 && c != 4) c = 2;
  }
  return c;
}

cudaError_t launch_bag_fwd(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const CUtensorMap& tm_h,
                           const BagFwdParams& prm, int num_sms, cudaStream_t stream) {
  if (prm.num_tiles <= 0) return cudaSuccess;
  static int pair = -1;
  if (pair < 0) { const char* e = getenv("MPO_FWD_PAIR"); pair = (e && atoi(e) != 0) ? 1 : 0; }
  if (pair && fwd_cluster_size() == 2 && prm.num_tiles % 2 == 0 && (prm.debug & ~(32 | 64 | 128 | 256 | 512)) == 0)
    return launch_fwd_cluster<2, true>(tm_x, tm_w, tm_h, prm, num_sms, stream);
  switch (fwd_cluster_size()) {
    case 1: return launch_fwd_cluster<1>(tm_x, tm_w, tm_h, prm, num_sms, stream);
    case 4: return launch_fwd_cluster<4>(tm_x, tm_w, tm_h, prm, num_sms, stream);
    default: return launch_fwd_cluster<2>(tm_x, tm_w, tm_h, prm, num_sms, stream);
  }
}

cudaError_t launch_bag_merge(const int* tile_prefix, const float* part_ml, int ml_stride, const float* part_pool,
                             float* pooled, float* lse, float* suma, int B, cudaStream_t stream,
                             const float* part_pool2, float* pooled2) {
  if (B <= 0) return cudaSuccess;
  cudaError_t e = launch_step(bag_merge_kernel, dim3(B, kQ, 4), dim3(256), 0, stream, tile_prefix, part_ml, ml_stride, part_pool,
                              pooled, lse, suma, part_pool2, pooled2);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mpo
