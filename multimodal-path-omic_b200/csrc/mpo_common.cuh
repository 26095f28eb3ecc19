// Shared constants and parameter blocks of the slide hot path (medium model: 1024 -> 256, 6 omic queries).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace mpo {

constexpr int kDIn = 1024;    // patch feature width          (reference: models/mcat/mcat.py:25)
constexpr int kD = 256;       // model width, medium           (reference: models/mcat/mcat.py:18-19)
constexpr int kQ = 6;         // omic signature queries
constexpr int kTileM = 128;   // patches per bag tile (one UMMA M)
constexpr int kBK = 64;       // K elements per pipeline stage (128 B of bf16 = one swizzle row)
constexpr int kKBlocks = kDIn / kBK;

// one entry per 128-patch tile of the packed bag
struct TileInfo {
  int slide;     // slide index in the batch
  int row0;      // first packed-bag row of the tile
  int nvalid;    // rows of this tile that belong to `slide` (1..128)
  int tile_in_slide;
};

struct BagFwdParams {
  const TileInfo* tile_info;   // [num_tiles]
  int num_tiles;
  int total_rows;
  const float* bias;           // [256]                       H.0.bias
  const float* qk;             // [B][6][256]                 folded queries (W_k^T q_i / 16)
  float* scores;               // [6][total_rows]             raw scores (pre-softmax)
  float* part_ml;              // [num_tiles][12]             tile max (6) and tile sum of exp (6)
  float* part_pool;            // [num_tiles][6][256]         sum_n exp(s - m_tile) h_n
  __half* h_out;               // [total_rows][256] or null   saved activations (fp16) for the backward pass
  __half* h_lo_out;            // [total_rows][256] or null   fp16 remainder h - fp16(h) (NaCAGaT key projection)
  uint32_t seed;               // dropout stream (train mode)
  const uint32_t* seed_dev;    // when non-null the stream id is read from device memory (CUDA-graph replays)
  uint32_t drop_thr;           // drop an element when its 8 random bits < drop_thr (0 = eval)
  float drop_scale;            // 1 / keep probability
  int skip_pool;               // 1: write activations and raw scores only (NaCAGaT: softmax runs on gated scores)
  int debug;                   // timing experiments only (env MPO_FWD_DEBUG): bit0 skip W loads, bit1 X from L2,
                               // bit2 L2-prefetch the next tile's X, bit3 no TMA loads, bit4 no main MMAs,
                               // bit5 no epilogue work, bit6 h_saved store without the evict-first hint,
                               // bit7 h_saved stored through the LSU (st.global.cs) instead of the TMA unit,
                               // bit8 accumulator stage released before the pooled product (wrong results),
                               // bit9 epilogue without its accumulator reads (wrong results)
};

// NaCAGaT gate pass (bag_gate.cu)
struct BagGateParams {
  const TileInfo* tile_info;
  int num_tiles;
  int total_rows;
  const float* qp;             // [B][6][256]   projected queries q_i (tanh taken in the kernel)
  const float* kc;             // [B][6]        key-bias score term b_k . q_i / 16
  const float* bias_k;         // [256]         co_attention.in_proj_bias[256:512]
  float* scores;               // [6][total_rows] in: h.qk_i ; rewritten as s = h.qk_i + kc_i when pgate != null
  float* scores_g;             // [6][total_rows] out: gated scores s' = s P
  float* pgate;                // [6][total_rows] out: P (kept for the backward pass) or null
  __half* t_out;               // [total_rows][256] out: tanh(k) fp16 (kept for the backward pass) or null
  float* part_ml;              // [num_tiles][18]  tile max, sum of exp, sum of dropped-and-rescaled exp
  float* part_pool;            // [num_tiles][6][256]  sum_n p_n (fp16(h_n) + remainder)
  float* part_pool_lo;         // [num_tiles][6][256]  the remainder part alone (for the backward pass) or null
  uint32_t seed;
  const uint32_t* seed_dev;
  uint32_t drop_thr;           // attention dropout (blocks.py:189-190): drop when 8 random bits < drop_thr
  float drop_scale;
  int l2_prefetch;             // prefetch the next tile into L2 while this one occupies the tile buffer (MPO_GATE_L2PF)
};

// NaCAGaT backward, key-projection path (bag_gate.cu: bag_dhk_kernel)
struct BagDhkParams {
  const TileInfo* tile_info;
  int num_tiles;
  int total_rows;
  const __half* h;             // [total_rows][256]  saved activations (ReLU / dropout mask)
  __nv_bfloat16* dz;           // [total_rows][256]  in: value/fold part of dz ; out: the complete dz
  float* part_db;              // [num_tiles][256]   per-tile column sums of dz
  const uint32_t* dg_max;      // bits of the batch-wide max |dg| (scale of dkg)
  float keep_scale;
};

// weight gradient with regenerated dz (bag_bwd.cu: bag_bwd_dwz_kernel)
struct BagBwdDwzParams {
  const TileInfo* tile_info;
  int num_tiles;
  int total_rows;
  const float* c12;            // [total_rows][12]  from bag_bwd_dz_kernel<kDzMcatLite>
  const uint32_t* mask;        // [total_rows][8]
  const float* dpooled;        // [B][6][256]
  const float* qk;             // [B][6][256]
  float* grad_w;               // [256][1024]  accumulated
  float* grad_b;               // [256]        accumulated
  void* scratch;               // bf16 [CTAs x 3 x 64][64]: the regenerated boxes on their way through the L2
  float keep_scale;
};

struct BagBwdDzParams {
  const TileInfo* tile_info;
  int num_tiles;
  int total_rows;
  const float* scores;         // [6][total_rows]  raw scores s (NaCAGaT: including the key-bias term)
  const float* lse;            // [B][6]           log-sum-exp of the softmax argument (s, or s P for NaCAGaT)
  const float* pooled;         // [B][6][256]
  const float* pooled_lo;      // [B][6][256] or null: part of `pooled` that came from the fp16 remainders of h (NaCAGaT)
  const float* dpooled;        // [B][6][256]
  const float* qk;             // [B][6][256]
  void* out;                   // [total_rows][256] 16-bit output tile rows (dz bf16 / dkg fp16), for ragged tiles
  float* part_dqk;             // [num_tiles][6][256]  per-tile Q partials (dqk, or dtq in mode 2)
  float* part_db;              // [num_tiles][256]     per-tile column sums of the output (modes 0 and 2)
  float keep_scale;            // 1/(1-p) in train mode, 1 in eval
  // NaCAGaT (modes 1 and 2)
  const float* pgate;          // [6][total_rows]  P
  const float* suma;           // [B][6]           sum_n a'_in            (with dsuma; null = no attention dropout)
  const float* dsuma;          // [B][6]           gradient of suma
  const float* qp;             // [B][6][256]      projected queries (mode 2: tanh taken in the kernel)
  float* dg;                   // [6][total_rows]  mode 1 writes, mode 2 reads the gate-dot gradients
  uint32_t* dg_max;            // bits of max |dg| over the batch (mode 1: atomicMax; mode 2: scale source)
  float* part_dkc;             // [num_tiles][8]   per-tile sums of ds_i (mode 1)
  // lite mode (MCAT): coefficients and mask bits instead of a dz tile
  float* c12_out;              // [total_rows][12] fp32  a_0..a_5, ds_0..ds_5
  uint32_t* mask_out;          // [total_rows][8]        bit f of a row: h[row][f] > 0
  // optional gradient arriving on the returned attention map (e.g. the CESAR regulariser, models/loss.py:88-101)
  const float* d_amap;         // [6][total_rows]  dL/dA (A = the map as returned: post-dropout for NaCAGaT) or null
  const float* amap_dot;       // [B][6]           sum_n A_in dL/dA_in   (with d_amap)
  uint32_t seed;
  const uint32_t* seed_dev;
  uint32_t attn_thr;           // attention dropout threshold (0 = off) and rescale factor
  float attn_scale;
  int debug;                   // timing switches (MPO_DZ_DEBUG; results are wrong when set): 1 no MMA-dqk, 2 no MMA-db,
                               // 4 no MMA-dZ, 8 no output-tile epilogue, 16 no TMA store, 32 no MMA-G
};

}  // namespace mpo
