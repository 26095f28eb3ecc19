/* mpo_b200.h -- C ABI of the B200-native slide hot path (libmpo_b200.so).
 *
 * The reference (mattiagualtieri/multimodal-path-omic) is pure PyTorch and has no FFI layer of its own; its
 * boundary for this path is the Python class surface.  Each entry point below therefore
 * cites the reference *module code* whose device work it replaces; INTEGRATION.md shows the ctypes binding the
 * reference-side modules use to call it.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`; the caller owns all buffers
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing synchronises
 *   - return value: 0 on success, negative MPO_E_* otherwise; mpo_last_error() gives the message
 *   - no entry point has a CPU fallback: without a CUDA device they fail with MPO_E_CUDA
 *   - gradients named grad_* are ACCUMULATED (+=) like torch .grad; everything else is overwritten
 *   - medium model only: patch features 1024, model width 256, 6 omic queries (mcat.py:16-21)
 */
#ifndef MPO_B200_H
#define MPO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPO_OK 0
#define MPO_E_ARG (-1)
#define MPO_E_CUDA (-2)
#define MPO_E_UNSUPPORTED (-3)

#define MPO_D_IN 1024
#define MPO_D 256
#define MPO_Q 6
#define MPO_TILE 128

const char* mpo_last_error(void);
int mpo_version(void);

/* Packed bag: the bf16 patch features of all slides of a batch, slide after slide, no padding rows.
 * tile_info[t] = {slide, first packed row, valid rows (1..128), tile index within the slide};
 * tile_prefix[b] = index of the first tile of slide b (tile_prefix[num_slides] = num_tiles). */
typedef struct mpo_bag {
  const void* x;              /* bf16 [total_rows][1024]                                        */
  int64_t total_rows;
  const int32_t* tile_info;   /* int32 [num_tiles][4]                                           */
  const int32_t* tile_prefix; /* int32 [num_slides + 1]                                         */
  int32_t num_tiles;
  int32_t num_slides;
} mpo_bag;

/* fp32 -> bf16 (round to nearest even).  Replaces the implicit fp32 bag of dataset/dataset.py:126 and the
 * fp32 H.0.weight read of mcat.py:87 by their bf16 streaming copies. */
int mpo_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream);

/* Bag-pass forward.  Replaces mcat.py:87 (`H_bag = self.H(wsi)`: Linear 1024->256 + ReLU + Dropout) fused with the
 * bag-facing part of mcat.py:97 (`self.co_attention(query=G_bag, key=H_bag, value=H_bag)`): scores, softmax over
 * the patches and the attention-weighted sum of H.  qk[b][i] = W_k^T q_i / 16 is the folded query (tail, below).
 *   scores    fp32 [6][total_rows]       raw (pre-softmax) scores, column n = packed row n
 *   part_ml   fp32 [num_tiles][12]       workspace
 *   part_pool fp32 [num_tiles][6][256]   workspace
 *   pooled    fp32 [num_slides][6][256]  sum_n a_in h_n
 *   lse       fp32 [num_slides][6]       log-sum-exp of the scores of each query
 *   h_saved   bf16 [total_rows][256] or NULL (inference): activations kept for mpo_bag_bwd
 *   drop_p    dropout probability on H in train mode (0 = eval); seed selects the mask stream */
int mpo_bag_fwd(const mpo_bag* bag, const void* w_h_bf16, const float* bias_h, const float* qk, float* scores,
                float* part_ml, float* part_pool, float* pooled, float* lse, void* h_saved, uint32_t seed,
                float drop_p, void* stream);

/* Normalised co-attention map A[i][n] = exp(scores[i][n] - lse[slide(n)][i])  (attention_scores['coattn'],
 * mcat.py:97,140).  amap fp32 [6][total_rows]. */
int mpo_attn_map(const mpo_bag* bag, const float* scores, const float* lse, float* amap, void* stream);

/* Bag-pass backward (autograd of mcat.py:87,97 w.r.t. H.0.weight, H.0.bias and the folded queries).
 *   dz_ws     bf16 [total_rows][256]     workspace (gradient at the pre-activation, feeds the dW_H GEMM)
 *   part_dqk  fp32 [num_tiles][6][256], part_db fp32 [num_tiles][256]  workspaces
 *   dqk       fp32 [num_slides][6][256]  gradient of the folded queries (overwritten)
 *   grad_w_h  fp32 [256][1024], grad_b_h fp32 [256]   accumulated */
int mpo_bag_bwd(const mpo_bag* bag, const void* h_saved, const float* scores, const float* lse, const float* pooled,
                const float* dpooled, const float* qk, void* dz_ws, float* part_dqk, float* part_db, float* dqk,
                float* grad_w_h, float* grad_b_h, float drop_p, void* stream);

/* Cross-shard log-sum-exp combine for one bag split by patch range over `nshards` ranks:
 * lse_in fp32 [nshards][6], pooled_in fp32 [nshards][6][256] (each shard's normalised result, e.g. after an
 * all-gather) -> lse_out [6], pooled_out [6][256]. */
int mpo_lse_combine(const float* lse_in, const float* pooled_in, int32_t nshards, float* lse_out, float* pooled_out,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPO_B200_H */
