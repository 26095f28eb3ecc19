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
/* number of CUDA kernels this library has launched since the last reset (bench.py's gpu_launches) */
int64_t mpo_launch_count(int32_t reset);

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
 *   h_saved   fp16 [total_rows][256] or NULL (inference): activations kept for mpo_bag_bwd (fp16, not bf16:
 *             11 mantissa bits keep the pooled vectors of small bags inside the 1e-3 parity gate)
 *   pooled == NULL selects the projection-only pass used by NaCAGaT (h_saved and scores are written, part_ml,
 *             part_pool, lse are ignored): the softmax then runs on the gated scores in mpo_bag_gate_fwd
 *   h_lo      fp16 [total_rows][256] or NULL: the remainder h - fp16(h); NaCAGaT's key projection (mpo_bag_gate_fwd)
 *             runs on the pair (h_saved, h_lo) so that its tanh gate sees ~22-bit activations
 *   drop_p    dropout probability on H in train mode (0 = eval); seed selects the mask stream; when seed_dev is
 *             non-NULL the stream id is seed ^ *seed_dev, read on the device (so a captured CUDA graph draws a new
 *             mask on every replay; advance it with mpo_advance_seed) */
int mpo_bag_fwd(const mpo_bag* bag, const void* w_h_bf16, const float* bias_h, const float* qk, float* scores,
                float* part_ml, float* part_pool, float* pooled, float* lse, void* h_saved, void* h_lo, uint32_t seed,
                const uint32_t* seed_dev, float drop_p, void* stream);
/* fp32 -> fp16 (round to nearest even): the fp16 streaming copy of NaCAGaT's key projection W_k */
int mpo_cast_f16(const float* src, void* dst_f16, int64_t n, void* stream);

/* NaCAGaT gate pass.  Replaces the attention core of models/blocks.py:156-192 (PreGatingContextualAttention /
 * multi_head_attention_forward): key projection k = W_k h + b_k, pre-gate P = (tanh(q).tanh(k) + 1)/2, gated scores
 * s' = s P, softmax over the patches, attention dropout (train) and the weighted sum of H.  Runs after mpo_bag_fwd
 * called with pooled == NULL (projection + raw folded scores h.qk only; h_saved is then mandatory).
 *   w_k_f16   fp16 [256][256]           co_attention.in_proj_weight[256:512] (mpo_cast_f16 of the fp32 master)
 *   bias_k    fp32 [256]                co_attention.in_proj_bias[256:512]
 *   qp, kc    from mpo_tail_pre_fwd     q_i = W_q g_i + b_q  and  kc_i = b_k . q_i / 16
 *   scores    fp32 [6][total_rows]      in: h.qk_i ; when pgate != NULL rewritten as s = h.qk_i + kc_i (for the backward)
 *   scores_g  fp32 [6][total_rows]      out: s'
 *   pgate     fp32 [6][total_rows] and t_saved fp16 [total_rows][256] (tanh(k)): kept for mpo_bag_bwd_nacagat, or both NULL
 *   part_ml   fp32 [num_tiles][18], part_pool fp32 [num_tiles][6][256]   workspaces
 *   pooled    fp32 [B][6][256] sum_n a'_in h_n ; lse fp32 [B][6] of s' ; suma fp32 [B][6] sum_n a'_in (1 without dropout)
 *   part_pool_lo fp32 [num_tiles][6][256] workspace and pooled_lo fp32 [B][6][256], or both NULL (inference): the part of
 *             `pooled` contributed by the fp16 remainders h_lo -- the gated softmax is often sharp, so pooled is formed
 *             from the ~22-bit pair (h_saved, h_lo); mpo_bag_bwd_nacagat works on h_saved alone and needs the split
 *   attn_drop_p  dropout on the attention weights in train mode (blocks.py:52,189-190: 0.25), 0 in eval; the mask
 *             stream is (seed ^ *seed_dev) as in mpo_bag_fwd, site 1 */
int mpo_bag_gate_fwd(const mpo_bag* bag, const void* h_saved, const void* h_lo, const void* w_k_f16, const float* bias_k, const float* qp,
                     const float* kc, float* scores, float* scores_g, float* pgate, void* t_saved, float* part_ml,
                     float* part_pool, float* pooled, float* lse, float* suma, float* part_pool_lo, float* pooled_lo,
                     uint32_t seed, const uint32_t* seed_dev, float attn_drop_p, void* stream);
/* NaCAGaT attention map: dropout(softmax(s')) -- the reference returns the post-dropout weights (blocks.py:189-199) */
int mpo_attn_map_dropout(const mpo_bag* bag, const float* scores_g, const float* lse, float* amap, uint32_t seed,
                         const uint32_t* seed_dev, float attn_drop_p, void* stream);

/* *seed_dev = hash(*seed_dev + golden ratio): one tiny kernel, stream-ordered (graph-capturable) */
int mpo_advance_seed(uint32_t* seed_dev, void* stream);

/* Normalised co-attention map A[i][n] = exp(scores[i][n] - lse[slide(n)][i])  (attention_scores['coattn'],
 * mcat.py:97,140).  amap fp32 [6][total_rows]. */
int mpo_attn_map(const mpo_bag* bag, const float* scores, const float* lse, float* amap, void* stream);

/* Bag-pass backward (autograd of mcat.py:87,97 w.r.t. H.0.weight, H.0.bias and the folded queries).
 *   dz_ws     bf16 [total_rows][256]     workspace (gradient at the pre-activation, feeds the dW_H GEMM)
 *   part_dqk  fp32 [num_tiles][6][256], part_db fp32 [num_tiles][256]  workspaces
 *   dqk       fp32 [num_slides][6][256]  gradient of the folded queries (overwritten)
 *   grad_w_h  fp32 [256][1024], grad_b_h fp32 [256]   accumulated
 *   d_amap    fp32 [6][total_rows] or NULL: a gradient arriving on the returned co-attention map (attention_scores['coattn']
 *             fed to a loss, e.g. models/loss.py:88-101) and amap_dot fp32 [num_slides][6] = sum_n A_in d_amap_in
 *             (mpo_attn_map_dot); both or neither */
int mpo_bag_bwd(const mpo_bag* bag, const void* h_saved, const float* scores, const float* lse, const float* pooled,
                const float* dpooled, const float* qk, void* dz_ws, float* part_dqk, float* part_db, float* dqk,
                float* grad_w_h, float* grad_b_h, const float* d_amap, const float* amap_dot, float drop_p, void* stream);

/* dot[b][i] = sum over the patches n of slide b of amap[i][n] * other[i][n]   (both fp32 [6][total_rows]; dot is
 * overwritten).  other = d_amap gives the softmax-Jacobian term of a map gradient; other = amap gives the squared
 * Frobenius norm per query that the attention-norm regulariser needs. */
int mpo_attn_map_dot(const mpo_bag* bag, const float* amap, const float* other, float* dot, void* stream);

/* Attention-norm regulariser of CrossEntropySurvivalAttnRegLoss (models/loss.py:88-101): for every slide b
 *   reg[b] = lambda_reg * ||A_b||_2 (Frobenius norm of the slide's [6, N_b] map) and
 *   d_amap[i][n] = grad_scale * lambda_reg * A_in / ||A_b||_2        (overwritten; feeds mpo_bag_bwd*)
 * sumsq fp32 [num_slides][6] = mpo_attn_map_dot(amap, amap). */
int mpo_cesar_reg(const mpo_bag* bag, const float* amap, const float* sumsq, float lambda_reg, float grad_scale,
                  float* reg, float* d_amap, void* stream);

/* SurvivalClassificationTobitLoss (models/loss.py:62-85) on the soft-maxed class probabilities Y [B][n_classes]:
 * uncensored: -log(Y[label] + eps); censored: -log(sum_{j >= label} Y[j] + eps).  dY is scaled by grad_scale. */
int mpo_sct_loss(const float* Y, const int64_t* label, const float* censor, float eps, float grad_scale, float* loss,
                 float* dY, int32_t B, int32_t n_classes, void* stream);

/* l1_reg (models/utils.py:33-40) over a flat fp32 buffer: *sum_out += sum |p| (the caller zeroes it); and its autograd:
 * grad[i] += scale * sign(p[i]). */
int mpo_l1_sum(const float* p, int64_t n, float* sum_out, void* stream);
int mpo_l1_grad(const float* p, float* grad, int64_t n, float scale, void* stream);

/* NaCAGaT bag-pass backward (autograd of nacagat.py:83,93 / blocks.py:156-192 w.r.t. H.0.*, the key projection and the
 * query-side operands).  All pointers are device pointers; workspaces are caller-owned. */
typedef struct mpo_nacagat_bwd {
  /* kept by mpo_bag_fwd (projection-only) and mpo_bag_gate_fwd */
  const void* h_saved;        /* fp16 [total_rows][256]                                     */
  const void* t_saved;        /* fp16 [total_rows][256]  tanh(k)                            */
  const float* scores;        /* fp32 [6][total_rows]    s (with the key-bias term)         */
  const float* pgate;         /* fp32 [6][total_rows]    P                                  */
  const float* lse;           /* fp32 [B][6]             of s' = s P                        */
  const float* pooled;        /* fp32 [B][6][256]                                           */
  const float* suma;          /* fp32 [B][6] or NULL (no attention dropout)                 */
  const float* pooled_lo;     /* fp32 [B][6][256] or NULL: remainder part of pooled (mpo_bag_gate_fwd) */
  /* upstream gradients (mpo_tail_post_bwd) */
  const float* dpooled;       /* fp32 [B][6][256]                                           */
  const float* dsuma;         /* fp32 [B][6] or NULL, together with suma                    */
  const float* d_amap;        /* fp32 [6][total_rows] or NULL: gradient on the returned (post-dropout) map */
  const float* amap_dot;      /* fp32 [B][6] = mpo_attn_map_dot(map, d_amap), together with d_amap */
  /* query-side operands (mpo_tail_pre_fwd) and the fp16 key projection */
  const float* qk;            /* fp32 [B][6][256]                                           */
  const float* qp;            /* fp32 [B][6][256]                                           */
  const void* w_k_f16;        /* fp16 [256][256]                                            */
  /* workspaces */
  void* dz_ws;                /* bf16 [total_rows][256]                                     */
  void* dkg_ws;               /* fp16 [total_rows][256]                                     */
  float* dg_ws;               /* fp32 [6][total_rows]                                       */
  float* part_dqk;            /* fp32 [num_tiles][6][256]                                   */
  float* part_dtq;            /* fp32 [num_tiles][6][256]                                   */
  float* part_db;             /* fp32 [num_tiles][256]                                      */
  float* part_dbk;            /* fp32 [num_tiles][256]                                      */
  float* part_dkc;            /* fp32 [num_tiles][8]                                        */
  uint32_t* dg_max;           /* one word                                                   */
  /* results: dqk, dkc, dtq are overwritten (inputs of mpo_tail_pre_bwd); grad_* are accumulated */
  float* dqk;                 /* fp32 [B][6][256]                                           */
  float* dkc;                 /* fp32 [B][6]                                                */
  float* dtq;                 /* fp32 [B][6][256]  gradient w.r.t. tanh(q_i)                */
  float* grad_w_h;            /* fp32 [256][1024]                                           */
  float* grad_b_h;            /* fp32 [256]                                                 */
  float* grad_w_k;            /* fp32 [256][256]   gate part of co_attention.in_proj_weight[256:512]' gradient */
  float* grad_b_k;            /* fp32 [256]        gate part of co_attention.in_proj_bias[256:512]'s gradient  */
  float drop_p;               /* bag dropout of the forward pass                            */
  float attn_drop_p;          /* attention dropout of the forward pass                      */
  uint32_t seed;              /* mask stream of the forward pass                            */
  const uint32_t* seed_dev;
} mpo_nacagat_bwd;
int mpo_bag_bwd_nacagat(const mpo_bag* bag, const mpo_nacagat_bwd* args, void* stream);

/* Optimizer step of the reference drivers (torch.optim.Adam with L2 weight decay, models/mcat/main.py:298-299;
 * SURVEY 8a row a14) over ONE flat fp32 buffer holding every parameter (and the matching flat gradient buffer the
 * backward entry points accumulate into), fused with zeroing the gradients of the next accumulation window.
 * *step_dev is the number of steps taken so far (bias correction); the call increments it on the device, so the
 * step can sit inside a captured CUDA graph.  n must be a multiple of 4 and the buffers 16-byte aligned. */
#define MPO_ADAM_NO_BUMP 2   /* zero_grad flag bit: leave *step_dev alone (an earlier call of the same step reads it too) */
int mpo_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t* step_dev, int32_t zero_grad, void* stream);
/* The same update for a bucket whose gradients the post stage completes (co_attention.out_proj and everything behind
 * it, models/mcat/mcat.py:97-138): queued BEHIND the side-stream weight-gradient kernel of mpo_tail_post_step, so it
 * runs next to the bag backward pass; mpo_tail_pre_bwd joins it back into the step's stream.  *step_dev is read, not
 * incremented (the step's last mpo_adam_step does that).  When no side-stream work is pending (per-op tail, or
 * MPO_POST_STEP_INLINE_WGRAD) the update simply runs on `stream`. */
int mpo_tail_side_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int32_t* step_dev, int32_t zero_grad, void* stream);

/* Cross-shard log-sum-exp combine for one bag split by patch range over `nshards` ranks:
 * lse_in fp32 [nshards][6], pooled_in fp32 [nshards][6][256] (each shard's normalised result, e.g. after an
 * all-gather) -> lse_out [6], pooled_out [6][256]. */
int mpo_lse_combine(const float* lse_in, const float* pooled_in, int32_t nshards, float* lse_out, float* pooled_out,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Peer-memory collectives over NVLink / NVSwitch (one process per GPU).  The reference has no distributed
 * code at all (its nn.DataParallel wrapper is a no-op at batch 1, models/mcat/main.py:267-268); these entry points carry
 * the two exchanges of the B200 design -- the gradient reduction of data-parallel training over slides and the soft-max
 * state combine of one bag sharded by patch range -- as plain kernels over CUDA-IPC mapped peer memory, so that a whole
 * step replays as one CUDA graph (csrc/peer.cu).
 *
 * Set-up (once): every rank allocates an exchange buffer of mpo_peer_exchange_bytes() with mpo_peer_alloc, exports it
 * (mpo_peer_export -> 64 opaque bytes, sent to the other ranks by any means, e.g. torch.distributed.all_gather_object),
 * opens the others' (mpo_peer_open) and fills struct mpo_peer_group: data[p] = rank p's exchange buffer as mapped here,
 * flags[p] = data[p] + mpo_peer_flags_offset(); for mpo_peer_adam_step also the flat gradient / parameter buffers of
 * every rank (allocated with mpo_peer_alloc, exported and opened the same way).  epochs: local device memory,
 * 8 zero-initialised uint32 counters.  All ranks must issue the same sequence of calls per slot.
 * ------------------------------------------------------------------------------------------------ */
#define MPO_PEER_MAX 8
#define MPO_PEER_HANDLE_BYTES 64
typedef struct mpo_peer_group {
  int32_t world, rank;
  void* data[MPO_PEER_MAX];             /* exchange buffers (peer-mapped; data[rank] is the local one)   */
  uint32_t* flags[MPO_PEER_MAX];        /* flags[p][slot * 8 + r]: written by rank r, polled by rank p   */
  void* grad[MPO_PEER_MAX];             /* flat fp32 gradient buffer of every rank  (mpo_peer_adam_step) */
  void* param[MPO_PEER_MAX];            /* flat fp32 parameter buffer of every rank (mpo_peer_adam_step) */
  uint32_t* epochs;                     /* local: one call counter per slot                              */
} mpo_peer_group;
int64_t mpo_peer_exchange_bytes(void);
int64_t mpo_peer_flags_offset(void);
int mpo_peer_alloc(int64_t bytes, void** ptr);           /* cudaMalloc'd (IPC-exportable), zero-filled */
int mpo_peer_free(void* ptr);
int mpo_peer_export(const void* ptr, void* handle_out);  /* handle_out: MPO_PEER_HANDLE_BYTES */
int mpo_peer_open(const void* handle, void** ptr_out);
int mpo_peer_close(void* ptr);
int mpo_peer_warmup(void);                               /* load the kernels ahead of a CUDA-graph capture */
/* all ranks' earlier work on `stream` is complete and visible to every rank's later work (signal + wait kernel) */
int mpo_peer_barrier(const mpo_peer_group* g, int32_t slot, void* stream);
/* Patch-range sharded bag: publish this rank's (lse [6], pooled [6][256]) to every peer, wait for theirs, merge:
 * lse_out [6], pooled_out [6][256] are identical on every rank.  An empty shard passes lse = -inf, pooled = 0.
 * Replaces all-gather + mpo_lse_combine by one launch. */
int mpo_peer_lse_combine(const mpo_peer_group* g, int32_t slot, const float* lse_local, const float* pooled_local,
                         float* lse_out, float* pooled_out, void* stream);
/* Data-parallel optimizer step over elements [lo, hi) of the flat buffers (multiples of 4): barrier(slot) ->
 * rank r sums slice r (mpo_peer_slice) of all ranks' gradients, scaled by grad_scale, applies Adam with L2 weight decay
 * (models/mcat/main.py:298-299; arithmetic of mpo_adam_step) using exp_avg / exp_avg_sq [i - state_offset], and stores the
 * new parameters into every rank's parameter buffer -> barrier(slot + 1) -> this rank's gradients [lo, hi) are zeroed.
 * bump_step: increment *step_dev afterwards (pass 1 on the last bucket of an optimizer step). */
int mpo_peer_adam_step(const mpo_peer_group* g, int32_t slot, int64_t lo, int64_t hi, float* exp_avg, float* exp_avg_sq,
                       int64_t state_offset, float lr, float beta1, float beta2, float eps, float weight_decay,
                       float grad_scale, int32_t* step_dev, int32_t bump_step, void* stream);
void mpo_peer_slice(int64_t lo, int64_t hi, int32_t world, int32_t rank, int64_t* s0, int64_t* s1);

/* ------------------------------------------------------------------------------------------------
 * Slide tail: everything that runs on the 6 omic tokens per slide, batched over the B slides of a step.
 * Parameters are plain fp32 device arrays in torch's own layouts (the nn.Parameter storage of the drop-in
 * modules), so reference checkpoints load unchanged.  g* are the gradient buffers (may be NULL
 * for inference); gradients are accumulated.
 * ------------------------------------------------------------------------------------------------ */
typedef struct mpo_lin { const float* w; const float* b; float* gw; float* gb; } mpo_lin;      /* nn.Linear / nn.Bilinear */
typedef struct mpo_norm { const float* g; const float* b; float* gg; float* gb; } mpo_norm;    /* nn.LayerNorm(256)       */

typedef struct mpo_encoder_layer {      /* nn.TransformerEncoderLayer(256, nhead 8, ff 512, relu, post-norm): mcat.py:51-53 */
  mpo_lin in_proj;                      /* [768,256]  self_attn.in_proj_{weight,bias} */
  mpo_lin out_proj;                     /* [256,256]  self_attn.out_proj              */
  mpo_lin linear1;                      /* [512,256] */
  mpo_lin linear2;                      /* [256,512] */
  mpo_norm norm1, norm2;
} mpo_encoder_layer;

typedef struct mpo_pool_head {          /* AttentionNetGated (blocks.py:13-48) + rho (mcat.py:57) */
  mpo_lin att_a, att_b;                 /* [256,256] each (Tanh / Sigmoid branches)   */
  mpo_lin att_c;                        /* [1,256]   */
  mpo_lin rho;                          /* [256,256] + ReLU */
} mpo_pool_head;

typedef struct mpo_cag {                /* ContextualAttentionGate(dim 256, hidden 256): blocks.py:232-253 */
  mpo_lin fc1, fc2, fc3, fc_c;          /* [256,256] each, ELU */
  mpo_norm G, E;
} mpo_cag;

typedef struct mpo_bilinear {           /* BilinearFusion(256,256, hidden 32, mm 64, out 256): fusion.py:44-113 */
  mpo_lin h1, z1, o1, h2, z2, o2;       /* h [32,256]; z nn.Bilinear w [32,256,256] b [32]; o [32,32] */
  mpo_lin fc1;                          /* [64,1089] */
  mpo_lin fc2;                          /* [256,130] */
} mpo_bilinear;

#define MPO_VARIANT_MCAT 0
#define MPO_VARIANT_NACAGAT 1
#define MPO_FUSION_CONCAT 0
#define MPO_FUSION_BILINEAR 1
#define MPO_FUSION_GATED_CONCAT 2
#define MPO_LOSS_NLL 0
#define MPO_LOSS_CES 1

typedef struct mpo_model {
  int32_t variant;                      /* MPO_VARIANT_*  */
  int32_t fusion;                       /* MPO_FUSION_*   */
  int32_t n_classes;                    /* survival bins (4) */
  int32_t omic_dims[MPO_Q];             /* input width of each SNN encoder */
  mpo_lin H;                            /* [256,1024] H.0 (fp32 master; kernels stream its bf16 copy) */
  mpo_lin snn[MPO_Q][2];                /* G.i.0.0 [256,d_i] and G.i.1.0 [256,256], ELU (mcat.py:32-45) */
  mpo_lin coattn_in;                    /* co_attention.in_proj [768,256] (q | k | v blocks) */
  mpo_lin coattn_out;                   /* co_attention.out_proj [256,256] */
  mpo_cag cag;                          /* NaCAGaT only */
  mpo_encoder_layer path_tr[2], omic_tr[2];
  mpo_pool_head path_pool, omic_pool;
  mpo_lin fusion0, fusion2;             /* ConcatFusion / GatedConcatFusion MLP [256,512], [256,256] (fusion.py:7-19, 28-33) */
  mpo_lin gate[2];                      /* fusion == gated_concat: Linear(256,1) + Sigmoid per input (fusion.py:25-27,36-39).
                                         * The reference keeps these in a plain Python list: not in the state_dict, never
                                         * trained; gw / gb may be NULL (their gradients are then not formed) */
  mpo_bilinear bil;                     /* fusion == bilinear */
  mpo_lin classifier;                   /* [n_classes,256] */
} mpo_model;

/* One step's token-side inputs/outputs.  omics[i] is fp32 [B][omic_dims[i]]. */
typedef struct mpo_tail_io {
  int32_t num_slides;
  const float* omics[MPO_Q];
  float* ws;                            /* fp32 workspace of mpo_tail_ws_floats() elements */
  /* bag-stage interface */
  float* qp;                            /* out of pre_fwd : [B][6][256] projected queries q = W_q g + b_q */
  float* qk;                            /* out of pre_fwd : [B][6][256] folded queries W_k^T q / 16   */
  float* kc;                            /* out of pre_fwd : [B][6] key-bias score term b_k.q/16 (NaCAGaT; else NULL) */
  const float* pooled;                  /* in  to post_fwd: [B][6][256] from mpo_bag_fwd             */
  const float* suma;                    /* in  to post_fwd: [B][6] sum_n a'_in (NaCAGaT with attention dropout) or NULL (= 1) */
  float* dsuma;                         /* out of post_bwd: [B][6] gradient of suma (written when suma != NULL)       */
  float* dpooled;                       /* out of post_bwd: [B][6][256]                               */
  const float* dqk;                     /* in  to pre_bwd : [B][6][256] from mpo_bag_bwd             */
  const float* dkc;                     /* in  to pre_bwd : [B][6]      (NaCAGaT; else NULL)          */
  const float* dtq;                     /* in  to pre_bwd : [B][6][256] gradient w.r.t. tanh(q) (NaCAGaT; else NULL) */
  /* model outputs (mcat.py:126-142) */
  float* hazards; float* S; float* Y;   /* [B][n_classes] each                                        */
  float* att_path; float* att_omic;     /* [B][6] raw pooling logits (attention_scores['path'/'omic']) */
  /* train mode: the dropout layers of the tail (SNN AlphaDropout mcat.py:38,42; encoder layers mcat.py:51-53; pooling
   * heads blocks.py:34-36 (p fixed at 0.25); rho mcat.py:57; bilinear fusion fusion.py:58-76 (p 0.25)).  drop_p is the
   * model's `dropout` (0 = eval: every dropout of the tail is the identity); masks come from the stateless RNG keyed by
   * (seed ^ *seed_dev, site, element), so the backward entry points regenerate them from the same three fields. */
  float drop_p;
  uint32_t seed;
  const uint32_t* seed_dev;
  /* 1 = train mode.  Separate from drop_p because the pooling heads (blocks.py:34-36) and the bilinear fusion
   * (fusion.py:58-76) hard-code p = 0.25: a model built with dropout = 0 still drops there while training. */
  int32_t train;
} mpo_tail_io;

/* ABI self-check for bindings: sizeof(mpo_bag) (which = 0), sizeof(mpo_model) (1), sizeof(mpo_tail_io) (2),
 * sizeof(mpo_nacagat_bwd) (3), sizeof(mpo_ge_model) (4), sizeof(mpo_peer_group) (5) */
int64_t mpo_sizeof(int32_t which);
/* workspace size in floats for a batch of B slides */
int64_t mpo_tail_ws_floats(const mpo_model* m, int32_t B);
/* debugging/test aid: offset (in floats) and length of a named intermediate inside the workspace; -1 if unknown */
int64_t mpo_tail_ws_lookup(const mpo_model* m, int32_t B, const char* name, int64_t* length);

/* mcat.py:90-92 (SNN encoders -> G_bag), the query in-projection of mcat.py:97 and the key fold qk = W_k^T q / 16 */
int mpo_tail_pre_fwd(const mpo_model* m, const mpo_tail_io* io, void* stream);
/* mcat.py:97 (value/out projection of the pooled vectors), :101-138 (encoders, pooling, fusion, survival head) */
int mpo_tail_post_fwd(const mpo_model* m, const mpo_tail_io* io, void* stream);
/* models/loss.py:31-43 (NLL, kind 0) and :5-28 (CES, kind 1): loss [B], dhaz/dS [B][n_classes] scaled by grad_scale */
int mpo_surv_loss(int32_t kind, const float* hazards, const float* S, const int64_t* label, const float* censor,
                  float alpha, float eps, float grad_scale, float* loss, float* dhaz, float* dS, int32_t B,
                  int32_t n_classes, void* stream);
/* autograd of post_fwd given d(hazards), d(S), d(Y) (each [B][n_classes], any may be NULL): parameter gradients,
 * io->dpooled, and the token-side gradients kept in the workspace for pre_bwd */
int mpo_tail_post_bwd(const mpo_model* m, const mpo_tail_io* io, const float* dhaz, const float* dS, const float* dY,
                      void* stream);
/* post_fwd + mpo_surv_loss + post_bwd as one call (the training step of models/mcat/main.py:39-70 between the bag
 * forward and the bag backward): for MCAT with concat fusion this is ONE cluster kernel (csrc/tail_fused.cu) plus the
 * grouped weight-gradient kernel; other configurations run the three stages back to back.  Arguments as in
 * mpo_surv_loss; io->dpooled is written, parameter gradients are accumulated.  On the fused path the post stage's
 * weight gradients run on an internal side stream next to the bag backward pass and are joined back into `stream`
 * by mpo_tail_pre_bwd: the step's gradients are complete (stream-ordered) after mpo_tail_pre_bwd, which a training
 * step always calls (mpo_bag_bwd needs nothing but io->dpooled from this call).  With MPO_POST_STEP_INLINE_WGRAD in
 * `flags` they stay on `stream` instead: a data-parallel trainer can then all-reduce the post stage's gradient bucket
 * while the bag backward pass runs. */
#define MPO_POST_STEP_INLINE_WGRAD 1   /* flags: keep the post stage's weight gradients on `stream` (complete on return order) */
int mpo_tail_post_step(const mpo_model* m, const mpo_tail_io* io, int32_t kind, const int64_t* label, const float* censor,
                       float alpha, float eps, float grad_scale, float* loss, float* dhaz, float* dS, int32_t flags,
                       void* stream);
/* autograd of pre_fwd: consumes io->dqk and the workspace gradients, finishes co_attention.in_proj and SNN grads */
int mpo_tail_pre_bwd(const mpo_model* m, const mpo_tail_io* io, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GE-NaCAGaT (models/ge_nacagat/ge_nacagat.py:9-72): WSI-only gene-expression classifier.  One slide per call.
 * H projection = mpo_bag_fwd with pooled == NULL (h_saved + h_lo); then 1-head N x N self-attention over the patches
 * (ge_nacagat.py:27,49 -- the map is returned, as in the reference), the 2-layer 8-head encoder over N tokens (:30-32,53),
 * gated attention pooling over N (:56-60), classifier and soft-max (:66-68).  First functional version: the attention
 * matrices are materialised in fp32 (N <= 46340), fp32 CUDA-core GEMMs; tensor cores in the projection and dW_H only.
 * ------------------------------------------------------------------------------------------------ */
typedef struct mpo_ge_model {
  int32_t n_classes;                    /* 3 */
  mpo_lin H;                            /* H.0 [256,1024]: only gw / gb are used here (dW_H, db_H)          */
  mpo_lin sa_in;                        /* self_attention.in_proj_{weight,bias} [768,256]                     */
  mpo_lin sa_out;                       /* self_attention.out_proj [256,256]                                  */
  mpo_encoder_layer tr[2];              /* path_transformer.layers.{0,1}                                      */
  mpo_pool_head pool;                   /* path_attention_head + path_rho                                     */
  mpo_lin classifier;                   /* [n_classes,256]                                                    */
} mpo_ge_model;
/* workspace size in floats for a slide of N patches (dominated by 18 N x N attention matrices) */
int64_t mpo_ge_ws_floats(int64_t N);
/* attn fp32 [N][N] (attention_scores['attn']), path fp32 [N] raw pooling logits (attention_scores['path']), Y [n_classes] */
/* train != 0: the dropout layers of the path draw masks from the stateless RNG keyed by (seed, site, element): the four
 * sites of each encoder layer at rate drop_p (attention probabilities, dropout1, feed-forward, dropout2:
 * ge_nacagat.py:30-32), the pooling head at its hard-coded 0.25 (blocks.py:34-36) and rho at drop_p (ge_nacagat.py:36).
 * mpo_ge_bwd regenerates them from the same (drop_p, seed, train). */
int mpo_ge_fwd(const mpo_ge_model* m, int64_t N, const void* h_hi, const void* h_lo, float* ws, float* attn, float* path,
               float* Y, float drop_p, uint32_t seed, int32_t train, void* stream);
/* the reference driver's loss: nn.CrossEntropyLoss on the soft-maxed Y (models/ge_nacagat/main.py:29,33); dY scaled by grad_scale */
int mpo_ge_ce_loss(const float* Y, const int64_t* label, int32_t n_classes, float grad_scale, float* loss, float* dY,
                   void* stream);
/* autograd of mpo_ge_fwd (+ the projection) given dY; gradients accumulated into m->*.gw/gb; dz_ws bf16 [N][256] */
int mpo_ge_bwd(const mpo_ge_model* m, const mpo_bag* bag, const void* h_hi, float* ws, const float* attn, const float* path,
               const float* Y, const float* dY, void* dz_ws, float keep_scale, float drop_p, uint32_t seed, int32_t train,
               void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone operators.  The reference's building blocks (models/blocks.py:13-48 AttentionNetGated, :232-253
 * ContextualAttentionGate; models/fusion.py:7-113 ConcatFusion / GatedConcatFusion / BilinearFusion) are nn.Modules that
 * can be called on their own -- its unit tests do (models/blocks.py:304-325, models/fusion.py:116-170).  On this path they
 * run fused inside a slide pass; these entry points expose the same device kernels one operator at a time for the
 * stand-alone `forward` of those classes (inference: no gradients).  All pointers are fp32 device memory.
 *   mpo_op_linear    y[r][o] = dropout(act(sum_i x[r][i] w[o][i] + b[o]))   act: 0 none, 1 ReLU, 2 ELU, 3 tanh, 4 sigmoid
 *   mpo_op_layernorm y = LayerNorm(x) over rows of `cols` features (biased variance)
 *   mpo_op_ewise     y = a + b | a * b | act(a)                              op: MPO_OP_ADD, MPO_OP_MUL, MPO_OP_ACT + act
 *   mpo_op_rowscale  y[r][c] = x[r][c] g[r]
 *   mpo_op_dropout   y = dropout_p(x) from the stateless (seed, site, element) mask stream
 *   mpo_op_bil_gate  nn.Bilinear(256, 256, 32) gate of BilinearFusion (fusion.py:86-89): z = x1^T W x2 + b through
 *                    U = x2 W^T [rows][32 * 256] (by mpo_op_linear), g = sigmoid(z), gh = g * h
 *   mpo_op_bil_kron  kp [rows][33 * 33] = dropout([o1, 1] x [o2, 1]) (fusion.py:100-107); cat [rows][130]: columns 64..129
 *                    receive [o1, 1, o2, 1] (the skip connection, fusion.py:110-111) */
#define MPO_OP_ADD 0
#define MPO_OP_MUL 1
#define MPO_OP_ACT 16
int mpo_op_linear(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy, int32_t rows,
                  int32_t in, int32_t out, int32_t act, float drop_p, uint32_t seed, uint32_t site, void* stream);
int mpo_op_layernorm(const float* x, const float* gamma, const float* beta, float* y, int32_t rows, int32_t cols, float eps,
                     void* stream);
int mpo_op_ewise(int32_t op, const float* a, const float* b, float* y, int64_t n, void* stream);
int mpo_op_rowscale(const float* x, const float* g, float* y, int32_t rows, int32_t cols, void* stream);
int mpo_op_dropout(const float* x, float* y, int64_t n, float p, uint32_t seed, uint32_t site, void* stream);
int mpo_op_bil_gate(const float* x1, const float* U, const float* bias, const float* h, float* g, float* gh, int32_t rows,
                    void* stream);
int mpo_op_bil_kron(const float* o1, const float* o2, float* kp, float* cat, int32_t rows, float drop_p, uint32_t seed,
                    uint32_t site, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPO_B200_H */
