// extern "C" surface for the bag stage (declared in include/mpo_b200.h) + small auxiliary kernels.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include "../../include/mpo_b200.h"
#include "mpo_ptx.cuh"
#include "mpo_common.cuh"
#include "launchers.h"
#include "tail_dev.cuh"

namespace mpo {

thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

int fail(int code, const char* fmt, const char* detail) {
  snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}
int check_cuda(cudaError_t e, const char* where) {
  if (e == cudaSuccess) return MPO_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return MPO_E_CUDA;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 0;
  }
  return n;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D row-major bf16 tensor [rows][cols]; box = box_cols x box_rows, 128-byte swizzle (box_cols * 2 B == 128 B)
int make_tmap_16b_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols,
                     uint32_t box_rows, bool is_f16) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(MPO_E_CUDA, "%s", "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, is_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return MPO_E_CUDA;
  }
  return MPO_OK;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols,
                      uint32_t box_rows) {
  return make_tmap_16b_2d(out, base, rows, cols, box_cols, box_rows, false);
}

// 2-D row-major fp32 tensor [rows][cols] with a row pitch of `pitch` elements (pitch * 4 B a multiple of 16 B);
// box = 32 x box_rows (128 B wide), 128-byte swizzle: the store target of the tensor-core GEMM's epilogue
int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows) {
  return make_tmap_f32_2d_sw(out, base, rows, cols, pitch, box_rows, true);
}
// same with the swizzle selectable (false: the box lands dense, rows of 128 B): the weight ring of the fused tail
int make_tmap_f32_2d_sw(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_rows,
                        bool swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(MPO_E_CUDA, "%s", "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", static_cast<int>(r));
    return MPO_E_CUDA;
  }
  return MPO_OK;
}

// fp32 matrix [rows][ld] (ld a multiple of 32) seen as [ld / 32][rows][32]: one box {32, 32 rows, kblocks} delivers a
// [32 rows][32 * kblocks] block of the matrix as `kblocks` consecutive 128-byte-swizzled [32][32] tiles with ONE copy
// (the forward chunks of the fused tail's weight ring)
int make_tmap_f32_rows32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t ld, uint32_t kblocks) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(MPO_E_CUDA, "%s", "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t dims[3] = {32, rows, ld / 32};
  cuuint64_t strides[2] = {ld * 4, 128};
  cuuint32_t box[3] = {32, 32, kblocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err), "cuTensorMapEncodeTiled (fp32, 3-D row view) failed with CUresult %d", static_cast<int>(r));
    return MPO_E_CUDA;
  }
  return MPO_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  pdl_enter();
  const int64_t i4 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i4 + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(src + i4);
    const float4 b = *reinterpret_cast<const float4*>(src + i4 + 4);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + i4) = o;
  } else {
    for (int64_t i = i4; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

__global__ void cast_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2half_rn(src[i]);
}

// NaCAGaT: amap = dropout(softmax(s')) -- the reference returns the post-dropout weights in train mode (blocks.py:189-199)
__global__ void attn_map_drop_kernel(const TileInfo* __restrict__ tile_info, const float* __restrict__ scores,
                                     const float* __restrict__ lse, float* __restrict__ amap, int total_rows,
                                     uint32_t seed, const uint32_t* __restrict__ seed_dev, uint32_t thr, float scale) {
  const TileInfo ti = tile_info[blockIdx.x];
  const int r = threadIdx.x;
  if (r >= ti.nvalid) return;
  const uint32_t sd = seed_dev != nullptr ? (seed ^ __ldg(seed_dev)) : seed;
#pragma unroll
  for (int i = 0; i < kQ; ++i) {
    const size_t o = static_cast<size_t>(i) * total_rows + ti.row0 + r;
    float a = __expf(scores[o] - lse[ti.slide * kQ + i]);
    if (thr != 0) {
      const uint32_t rb = rng_u32(sd, 1u, static_cast<uint32_t>(i) * static_cast<uint32_t>(total_rows) +
                                               static_cast<uint32_t>(ti.row0 + r)) & 0xFFu;
      a = rb < thr ? 0.f : a * scale;
    }
    amap[o] = a;
  }
}

// one block per tile: amap[i][row] = exp(scores[i][row] - lse[slide][i])
__global__ void attn_map_kernel(const TileInfo* __restrict__ tile_info, const float* __restrict__ scores,
                                const float* __restrict__ lse, float* __restrict__ amap, int total_rows) {
  const TileInfo ti = tile_info[blockIdx.x];
  const int r = threadIdx.x;
  if (r >= ti.nvalid) return;
#pragma unroll
  for (int i = 0; i < kQ; ++i) {
    const size_t o = static_cast<size_t>(i) * total_rows + ti.row0 + r;
    amap[o] = __expf(scores[o] - lse[ti.slide * kQ + i]);
  }
}

// lse_in [S][6], pooled_in [S][6][256] -> combined
__global__ void lse_combine_kernel(const float* __restrict__ lse_in, const float* __restrict__ pooled_in, int S,
                                   float* __restrict__ lse_out, float* __restrict__ pooled_out) {
  const int i = blockIdx.x, d = threadIdx.x;
  float M = -INFINITY;
  for (int s = 0; s < S; ++s) M = fmaxf(M, lse_in[s * kQ + i]);
  float L = 0.f, acc = 0.f;
  for (int s = 0; s < S; ++s) {
    const float w = __expf(lse_in[s * kQ + i] - M);
    L += w;
    acc = fmaf(pooled_in[(static_cast<size_t>(s) * kQ + i) * kD + d], w, acc);
  }
  pooled_out[i * kD + d] = acc / L;
  if (d == 0) lse_out[i] = M + __logf(L);
}

}  // namespace mpo

using namespace mpo;

extern "C" {

const char* mpo_last_error(void) { return g_err; }
int mpo_version(void) { return 100; }
int64_t mpo_launch_count(int32_t reset) {
  const int64_t n = static_cast<int64_t>(g_launches);
  if (reset) g_launches = 0;
  return n;
}

int mpo_cast_bf16(const float* src, void* dst, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!src || !dst))) return fail(MPO_E_ARG, "%s", "mpo_cast_bf16: null pointer");
  if (n == 0) return MPO_OK;
  const int64_t groups = (n + 7) / 8;
  const int threads = 256;
  const int64_t blocks = (groups + threads - 1) / threads;
  mpo::launch_step(cast_bf16_kernel, dim3(static_cast<unsigned>(blocks)), dim3(threads), 0, static_cast<cudaStream_t>(stream),
      src, static_cast<__nv_bfloat16*>(dst), n);
  count_launch();
  return check_cuda(cudaGetLastError(), "mpo_cast_bf16");
}

static int check_bag(const mpo_bag* bag, const char* who) {
  if (!bag) return fail(MPO_E_ARG, "%s: bag is NULL", who);
  if (bag->num_slides < 0 || bag->num_tiles < 0 || bag->total_rows < 0) return fail(MPO_E_ARG, "%s: negative size", who);
  if (bag->total_rows > 0 && (!bag->x || !bag->tile_info || !bag->tile_prefix))
    return fail(MPO_E_ARG, "%s: bag pointers are NULL", who);
  if ((reinterpret_cast<uintptr_t>(bag->x) & 15) != 0) return fail(MPO_E_ARG, "%s: bag must be 16-byte aligned", who);
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s: no CUDA device (this library has no CPU fallback)", who);
  return MPO_OK;
}

int mpo_bag_fwd(const mpo_bag* bag, const void* w_h_bf16, const float* bias_h, const float* qk, float* scores,
                float* part_ml, float* part_pool, float* pooled, float* lse, void* h_saved, void* h_lo, uint32_t seed,
                const uint32_t* seed_dev, float drop_p, void* stream) {
  int rc = check_bag(bag, "mpo_bag_fwd");
  if (rc) return rc;
  if (bag->total_rows == 0 || bag->num_tiles == 0) return MPO_OK;
  const bool project_only = pooled == nullptr;      // NaCAGaT: activations + raw scores, softmax in mpo_bag_gate_fwd
  if (!w_h_bf16 || !bias_h || !qk || !scores) return fail(MPO_E_ARG, "%s", "mpo_bag_fwd: null pointer");
  if (!project_only && (!part_ml || !part_pool || !lse)) return fail(MPO_E_ARG, "%s", "mpo_bag_fwd: null pointer");
  if (project_only && (!h_saved || !h_lo))
    return fail(MPO_E_ARG, "%s", "mpo_bag_fwd: a projection-only pass needs h_saved and h_lo");
  if (drop_p < 0.f || drop_p >= 1.f) return fail(MPO_E_ARG, "%s", "mpo_bag_fwd: drop_p must be in [0,1)");
  CUtensorMap tm_x, tm_w;
  rc = make_tmap_bf16_2d(&tm_x, bag->x, static_cast<uint64_t>(bag->total_rows), kDIn, kBK, kTileM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_w, w_h_bf16, kD, kDIn, kBK, kD / fwd_cluster_size());   // one multicast slice per CTA
  if (rc) return rc;
  BagFwdParams p;
  p.tile_info = reinterpret_cast<const TileInfo*>(bag->tile_info);
  p.num_tiles = bag->num_tiles;
  p.total_rows = static_cast<int>(bag->total_rows);
  p.bias = bias_h;
  p.qk = qk;
  p.scores = scores;
  p.part_ml = part_ml;
  p.part_pool = part_pool;
  p.h_out = static_cast<__half*>(h_saved);
  p.h_lo_out = static_cast<__half*>(h_lo);
  p.seed = seed;
  p.seed_dev = seed_dev;
  p.skip_pool = project_only ? 1 : 0;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("MPO_FWD_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
  p.drop_thr = static_cast<uint32_t>(drop_p * 256.f + 0.5f);
  p.drop_scale = p.drop_thr ? 256.f / static_cast<float>(256 - p.drop_thr) : 1.f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tm_h = tm_x;   // placeholder when nothing is saved (never dereferenced by the kernel)
  if (h_saved != nullptr) {
    rc = make_tmap_16b_2d(&tm_h, h_saved, static_cast<uint64_t>(bag->total_rows), kD, 64, kTileM, true);
    if (rc) return rc;
  }
  rc = check_cuda(launch_bag_fwd(tm_x, tm_w, tm_h, p, num_sms(), st), "bag_fwd_kernel");
  if (rc || project_only) return rc;
  return check_cuda(launch_bag_merge(bag->tile_prefix, part_ml, 12, part_pool, pooled, lse, nullptr, bag->num_slides, st),
                    "bag_merge_kernel");
}

__global__ void advance_seed_kernel(uint32_t* s) { mpo::pdl_enter(); *s = mpo::hash_u32(*s + 0x9E3779B9u); }

int mpo_advance_seed(uint32_t* seed_dev, void* stream) {
  if (!seed_dev) return fail(MPO_E_ARG, "%s", "mpo_advance_seed: NULL pointer");
  mpo::launch_step(advance_seed_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), seed_dev);
  count_launch();
  return check_cuda(cudaGetLastError(), "advance_seed_kernel");
}

int mpo_attn_map(const mpo_bag* bag, const float* scores, const float* lse, float* amap, void* stream) {
  int rc = check_bag(bag, "mpo_attn_map");
  if (rc) return rc;
  if (bag->num_tiles == 0) return MPO_OK;
  if (!scores || !lse || !amap) return fail(MPO_E_ARG, "%s", "mpo_attn_map: null pointer");
  attn_map_kernel<<<bag->num_tiles, kTileM, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const TileInfo*>(bag->tile_info), scores, lse, amap, static_cast<int>(bag->total_rows));
  count_launch();
  return check_cuda(cudaGetLastError(), "attn_map_kernel");
}

int mpo_cast_f16(const float* src, void* dst, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!src || !dst))) return fail(MPO_E_ARG, "%s", "mpo_cast_f16: null pointer");
  if (n == 0) return MPO_OK;
  cast_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__half*>(dst), n);
  count_launch();
  return check_cuda(cudaGetLastError(), "mpo_cast_f16");
}

static void drop_params(float drop_p, uint32_t* thr, float* scale) {
  *thr = static_cast<uint32_t>(drop_p * 256.f + 0.5f);
  *scale = *thr ? 256.f / static_cast<float>(256 - *thr) : 1.f;
}

int mpo_bag_gate_fwd(const mpo_bag* bag, const void* h_saved, const void* h_lo, const void* w_k_f16, const float* bias_k, const float* qp,
                     const float* kc, float* scores, float* scores_g, float* pgate, void* t_saved, float* part_ml,
                     float* part_pool, float* pooled, float* lse, float* suma, float* part_pool_lo, float* pooled_lo,
                     uint32_t seed, const uint32_t* seed_dev, float attn_drop_p, void* stream) {
  int rc = check_bag(bag, "mpo_bag_gate_fwd");
  if (rc) return rc;
  if (bag->total_rows == 0 || bag->num_tiles == 0) return MPO_OK;
  if (!h_saved || !h_lo || !w_k_f16 || !bias_k || !qp || !kc || !scores || !scores_g || !part_ml || !part_pool || !pooled ||
      !lse || !suma)
    return fail(MPO_E_ARG, "%s", "mpo_bag_gate_fwd: null pointer");
  if ((pgate == nullptr) != (t_saved == nullptr))
    return fail(MPO_E_ARG, "%s", "mpo_bag_gate_fwd: pgate and t_saved are kept (or dropped) together");
  if ((part_pool_lo == nullptr) != (pooled_lo == nullptr))
    return fail(MPO_E_ARG, "%s", "mpo_bag_gate_fwd: part_pool_lo and pooled_lo are kept (or dropped) together");
  if (attn_drop_p < 0.f || attn_drop_p >= 1.f) return fail(MPO_E_ARG, "%s", "mpo_bag_gate_fwd: drop_p must be in [0,1)");
  CUtensorMap tm_h, tm_hlo, tm_w;
  rc = make_tmap_16b_2d(&tm_h, h_saved, static_cast<uint64_t>(bag->total_rows), kD, 64, kTileM, true);
  if (rc) return rc;
  rc = make_tmap_16b_2d(&tm_hlo, h_lo, static_cast<uint64_t>(bag->total_rows), kD, 64, kTileM, true);
  if (rc) return rc;
  rc = make_tmap_16b_2d(&tm_w, w_k_f16, kD, kD, 64, kD, true);
  if (rc) return rc;
  BagGateParams p;
  p.tile_info = reinterpret_cast<const TileInfo*>(bag->tile_info);
  p.num_tiles = bag->num_tiles;
  p.total_rows = static_cast<int>(bag->total_rows);
  p.qp = qp; p.kc = kc; p.bias_k = bias_k;
  p.scores = scores; p.scores_g = scores_g; p.pgate = pgate;
  p.t_out = static_cast<__half*>(t_saved);
  p.part_ml = part_ml; p.part_pool = part_pool; p.part_pool_lo = part_pool_lo;
  p.seed = seed; p.seed_dev = seed_dev;
  drop_params(attn_drop_p, &p.drop_thr, &p.drop_scale);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = check_cuda(launch_bag_gate(tm_h, tm_hlo, tm_w, p, num_sms(), st), "bag_gate_kernel");
  if (rc) return rc;
  return check_cuda(launch_bag_merge(bag->tile_prefix, part_ml, 18, part_pool, pooled, lse, suma, bag->num_slides, st,
                                     part_pool_lo, pooled_lo),
                    "bag_merge_kernel");
}

int mpo_attn_map_dropout(const mpo_bag* bag, const float* scores_g, const float* lse, float* amap, uint32_t seed,
                         const uint32_t* seed_dev, float attn_drop_p, void* stream) {
  int rc = check_bag(bag, "mpo_attn_map_dropout");
  if (rc) return rc;
  if (bag->num_tiles == 0) return MPO_OK;
  if (!scores_g || !lse || !amap) return fail(MPO_E_ARG, "%s", "mpo_attn_map_dropout: null pointer");
  uint32_t thr; float scale;
  drop_params(attn_drop_p, &thr, &scale);
  attn_map_drop_kernel<<<bag->num_tiles, kTileM, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const TileInfo*>(bag->tile_info), scores_g, lse, amap, static_cast<int>(bag->total_rows), seed,
      seed_dev, thr, scale);
  count_launch();
  return check_cuda(cudaGetLastError(), "attn_map_drop_kernel");
}

}  // extern "C"
namespace {
// The per-tile partial sums of the backward kernels (dqk, dkc, dtq, bias gradients) are folded by small reduction
// kernels whose results nothing inside the bag backward pass reads.  They run on an internal side stream, forked behind
// the kernel that wrote the partials and joined at the end of the call (event edges: capturable), so that they overlap
// the next bag kernel instead of sitting between two of them.  MPO_BAG_SIDE=0 keeps everything on the caller's stream.
struct BagSide {
  cudaStream_t s = nullptr;
  cudaEvent_t fork[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t join = nullptr;
  int state = 0;     // 0 not created, 1 ready, -1 unavailable
};
BagSide& bag_side() {
  static BagSide pool[32];
  int dev = 0;
  cudaGetDevice(&dev);
  BagSide& b = pool[dev & 31];
  if (b.state == 0) {
    const char* env = getenv("MPO_BAG_SIDE");
    bool ok = !(env && atoi(env) == 0);
    ok = ok && cudaStreamCreateWithFlags(&b.s, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreateWithFlags(&b.fork[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&b.join, cudaEventDisableTiming) == cudaSuccess;
    b.state = ok ? 1 : -1;
  }
  return b;
}
// the stream for a reduction that depends on everything queued on `st` so far (fork number `i` of this call)
cudaStream_t side_fork(BagSide& b, int i, cudaStream_t st, cudaError_t* err) {
  if (b.state != 1) return st;
  cudaError_t e = cudaEventRecord(b.fork[i & 3], st);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(b.s, b.fork[i & 3], 0);
  if (e != cudaSuccess) { *err = e; return st; }
  return b.s;
}
// `st` waits for everything queued on the side stream
cudaError_t side_join(BagSide& b, cudaStream_t st) {
  if (b.state != 1) return cudaSuccess;
  cudaError_t e = cudaEventRecord(b.join, b.s);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(st, b.join, 0);
  mpo::pdl_bar_next(st);        // the next kernel of the step waits for the side stream: launched fully serialized
  return e;
}
}  // namespace
extern "C" {

int mpo_bag_bwd(const mpo_bag* bag, const void* h_saved, const float* scores, const float* lse, const float* pooled,
                const float* dpooled, const float* qk, void* dz_ws, float* part_dqk, float* part_db, float* dqk,
                float* grad_w_h, float* grad_b_h, const float* d_amap, const float* amap_dot, float drop_p, void* stream) {
  int rc = check_bag(bag, "mpo_bag_bwd");
  if (rc) return rc;
  if (bag->total_rows == 0 || bag->num_tiles == 0) return MPO_OK;
  if (!h_saved || !scores || !lse || !pooled || !dpooled || !qk || !dz_ws || !part_dqk || !part_db || !dqk ||
      !grad_w_h || !grad_b_h)
    return fail(MPO_E_ARG, "%s", "mpo_bag_bwd: null pointer");
  if ((d_amap == nullptr) != (amap_dot == nullptr))
    return fail(MPO_E_ARG, "%s", "mpo_bag_bwd: d_amap and amap_dot go together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BagBwdDzParams p = {};
  p.d_amap = d_amap;
  p.amap_dot = amap_dot;
  p.tile_info = reinterpret_cast<const TileInfo*>(bag->tile_info);
  p.num_tiles = bag->num_tiles;
  p.total_rows = static_cast<int>(bag->total_rows);
  p.scores = scores;
  p.lse = lse;
  p.pooled = pooled;
  p.dpooled = dpooled;
  p.qk = qk;
  p.out = dz_ws;
  p.part_dqk = part_dqk;
  p.part_db = part_db;
  const uint32_t thr = static_cast<uint32_t>(drop_p * 256.f + 0.5f);
  p.keep_scale = thr ? 256.f / static_cast<float>(256 - thr) : 1.f;
  CUtensorMap tm_h, tm_dzs;
  rc = make_tmap_16b_2d(&tm_h, h_saved, static_cast<uint64_t>(bag->total_rows), kD, 64, kTileM, true);
  if (rc) return rc;
  // MPO_BWD_REGEN=1: dz never exists in HBM.  The row-scalar kernel leaves 80 B per patch (12 fp32 coefficients + 256
  // mask bits, inside the dz workspace the caller provides) and the weight-gradient kernel regenerates the dz operand,
  // once per 4-CTA cluster (bag_bwd_dwz_kernel).  Parity-green and 0.54 GB less HBM traffic per 32 x 16k step, but
  // measured SLOWER than the two-kernel path with a materialised bf16 dz (0.414 vs 0.375 ms): the exchange of
  // the regenerated boxes inside the cluster is latency-bound.  Off by default.
  static int regen = -1;
  if (regen < 0) { const char* e = getenv("MPO_BWD_REGEN"); regen = e ? atoi(e) : 0; }
  // the regenerating kernel needs a 24 KB scratch ring per CTA behind the 80 B per patch: both live in the caller's dz
  // workspace (512 B per patch), which is large enough from ~8 k patches on; smaller batches take the two-kernel path
  int clusters = bag_bwd_dwz_max_clusters(num_sms());
  if (clusters > bag->num_tiles) clusters = bag->num_tiles;
  const size_t scr_off = (static_cast<size_t>(bag->total_rows) * 80 + 1023) / 1024 * 1024;
  bool use_regen = regen != 0;
  if (scr_off + bag_bwd_dwz_scratch_bytes(clusters) > static_cast<size_t>(bag->total_rows) * kD * 2) use_regen = false;
  if ((reinterpret_cast<uintptr_t>(dz_ws) & 127) != 0) use_regen = false;        // TMA global addresses: 16 B, keep 128
  if (use_regen) {
    float* c12 = static_cast<float*>(dz_ws);
    uint32_t* mask = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(dz_ws) + static_cast<size_t>(bag->total_rows) * 48);
    p.c12_out = c12;
    p.mask_out = mask;
    p.out = nullptr;
    p.part_db = nullptr;
    rc = check_cuda(launch_bag_bwd_dz(3, tm_h, tm_h, p, num_sms(), st), "bag_bwd_dz_kernel<lite>");
    if (rc) return rc;
    rc = check_cuda(launch_bag_bwd_reduce(bag->tile_prefix, part_dqk, nullptr, dqk, nullptr, bag->num_slides,
                                          bag->num_tiles, st), "bag_bwd_reduce_kernel");
    if (rc) return rc;
    CUtensorMap tm_x64;
    rc = make_tmap_bf16_2d(&tm_x64, bag->x, static_cast<uint64_t>(bag->total_rows), kDIn, 64, 64);
    if (rc) return rc;
    BagBwdDwzParams z = {};
    z.tile_info = p.tile_info; z.num_tiles = p.num_tiles; z.total_rows = p.total_rows;
    z.c12 = c12; z.mask = mask; z.dpooled = dpooled; z.qk = qk;
    z.grad_w = grad_w_h; z.grad_b = grad_b_h; z.keep_scale = p.keep_scale;
    z.scratch = static_cast<uint8_t*>(dz_ws) + scr_off;
    CUtensorMap tm_scr;
    rc = make_tmap_bf16_2d(&tm_scr, z.scratch, static_cast<uint64_t>(clusters) * 4 * 3 * 64, 64, 64, 64);
    if (rc) return rc;
    return check_cuda(launch_bag_bwd_dwz(tm_x64, tm_scr, z, num_sms(), st), "bag_bwd_dwz_kernel");
  }
  rc = make_tmap_bf16_2d(&tm_dzs, dz_ws, static_cast<uint64_t>(bag->total_rows), kD, 64, kTileM);
  if (rc) return rc;
  rc = check_cuda(launch_bag_bwd_dz(0, tm_h, tm_dzs, p, num_sms(), st), "bag_bwd_dz_kernel");
  if (rc) return rc;
  BagSide& side = bag_side();
  cudaError_t ferr = cudaSuccess;
  cudaStream_t rs = side_fork(side, 0, st, &ferr);        // the reduction overlaps the weight-gradient kernel
  if (ferr != cudaSuccess) return check_cuda(ferr, "mpo_bag_bwd: side-stream fork");
  rc = check_cuda(launch_bag_bwd_reduce(bag->tile_prefix, part_dqk, part_db, dqk, grad_b_h, bag->num_slides,
                                        bag->num_tiles, rs),
                  "bag_bwd_reduce_kernel");
  if (rc) return rc;
  CUtensorMap tm_dz, tm_x;
  rc = make_tmap_bf16_2d(&tm_dz, dz_ws, static_cast<uint64_t>(bag->total_rows), kD, 64, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tm_x, bag->x, static_cast<uint64_t>(bag->total_rows), kDIn, 64, 64);
  if (rc) return rc;
  rc = check_cuda(launch_bag_bwd_dw(tm_dz, tm_x, grad_w_h, static_cast<int>(bag->total_rows), kDIn, kDIn, false, nullptr,
                                    num_sms(), st),
                  "bag_bwd_dw_kernel");
  if (rc) return rc;
  return rs != st ? check_cuda(side_join(side, st), "mpo_bag_bwd: side-stream join") : MPO_OK;
}

int mpo_bag_bwd_nacagat(const mpo_bag* bag, const mpo_nacagat_bwd* a, void* stream) {
  int rc = check_bag(bag, "mpo_bag_bwd_nacagat");
  if (rc) return rc;
  if (!a) return fail(MPO_E_ARG, "%s", "mpo_bag_bwd_nacagat: argument block is NULL");
  if (bag->total_rows == 0 || bag->num_tiles == 0) return MPO_OK;
  if (!a->h_saved || !a->t_saved || !a->scores || !a->pgate || !a->lse || !a->pooled || !a->dpooled || !a->qk || !a->qp ||
      !a->w_k_f16 || !a->dz_ws || !a->dkg_ws || !a->dg_ws || !a->part_dqk || !a->part_dtq || !a->part_db || !a->part_dbk ||
      !a->part_dkc || !a->dg_max || !a->dqk || !a->dkc || !a->dtq || !a->grad_w_h || !a->grad_b_h || !a->grad_w_k ||
      !a->grad_b_k)
    return fail(MPO_E_ARG, "%s", "mpo_bag_bwd_nacagat: null pointer");
  if ((a->suma == nullptr) != (a->dsuma == nullptr))
    return fail(MPO_E_ARG, "%s", "mpo_bag_bwd_nacagat: suma and dsuma go together");
  if ((a->d_amap == nullptr) != (a->amap_dot == nullptr))
    return fail(MPO_E_ARG, "%s", "mpo_bag_bwd_nacagat: d_amap and amap_dot go together");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint64_t R = static_cast<uint64_t>(bag->total_rows);
  const int ns = num_sms();
  CUtensorMap tm_h, tm_t, tm_dz, tm_dkg, tm_wk;
  if ((rc = make_tmap_16b_2d(&tm_h, a->h_saved, R, kD, 64, kTileM, true))) return rc;
  if ((rc = make_tmap_16b_2d(&tm_t, a->t_saved, R, kD, 64, kTileM, true))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm_dz, a->dz_ws, R, kD, 64, kTileM))) return rc;
  if ((rc = make_tmap_16b_2d(&tm_dkg, a->dkg_ws, R, kD, 64, kTileM, true))) return rc;
  if ((rc = make_tmap_16b_2d(&tm_wk, a->w_k_f16, kD, kD, 64, 64, true))) return rc;
  rc = check_cuda(cudaMemsetAsync(a->dg_max, 0, sizeof(uint32_t), st), "memset dg_max");
  if (rc) return rc;

  BagBwdDzParams p = {};
  p.tile_info = reinterpret_cast<const TileInfo*>(bag->tile_info);
  p.num_tiles = bag->num_tiles;
  p.total_rows = static_cast<int>(bag->total_rows);
  p.scores = a->scores; p.lse = a->lse; p.pooled = a->pooled; p.pooled_lo = a->pooled_lo; p.dpooled = a->dpooled; p.qk = a->qk;
  p.pgate = a->pgate; p.suma = a->suma; p.dsuma = a->dsuma; p.qp = a->qp;
  p.dg = a->dg_ws; p.dg_max = a->dg_max; p.part_dkc = a->part_dkc;
  p.d_amap = a->d_amap; p.amap_dot = a->amap_dot;
  p.seed = a->seed; p.seed_dev = a->seed_dev;
  drop_params(a->attn_drop_p, &p.attn_thr, &p.attn_scale);
  { uint32_t thr; drop_params(a->drop_p, &thr, &p.keep_scale); }
  // value / fold path: dz_part, dqk, dkc, dg
  p.out = a->dz_ws; p.part_dqk = a->part_dqk; p.part_db = nullptr;
  if ((rc = check_cuda(launch_bag_bwd_dz(1, tm_h, tm_dz, p, ns, st), "bag_bwd_dz_kernel<nacagat dh>"))) return rc;
  BagSide& side = bag_side();
  cudaError_t ferr = cudaSuccess;
  cudaStream_t rs = side_fork(side, 0, st, &ferr);        // the reductions overlap the next bag kernel
  if (ferr != cudaSuccess) return check_cuda(ferr, "mpo_bag_bwd_nacagat: side-stream fork");
  if ((rc = check_cuda(launch_bag_bwd_reduce(bag->tile_prefix, a->part_dqk, nullptr, a->dqk, nullptr, bag->num_slides,
                                             bag->num_tiles, rs), "bag_bwd_reduce_kernel"))) return rc;
  if ((rc = check_cuda(launch_bag_bwd_dkc(bag->tile_prefix, a->part_dkc, a->dkc, bag->num_slides, rs),
                       "bag_bwd_dkc_kernel"))) return rc;
  // gate path: dkg = (1 - tanh(k)^2) (dg tq) gs, dtq, gate part of db_k
  p.out = a->dkg_ws; p.part_dqk = a->part_dtq; p.part_db = a->part_dbk;
  if ((rc = check_cuda(launch_bag_bwd_dz(2, tm_t, tm_dkg, p, ns, st), "bag_bwd_dz_kernel<nacagat dkg>"))) return rc;
  rs = side_fork(side, 1, st, &ferr);
  if (ferr != cudaSuccess) return check_cuda(ferr, "mpo_bag_bwd_nacagat: side-stream fork");
  if ((rc = check_cuda(launch_bag_bwd_reduce(bag->tile_prefix, a->part_dtq, a->part_dbk, a->dtq, a->grad_b_k,
                                             bag->num_slides, bag->num_tiles, rs), "bag_bwd_reduce_kernel"))) return rc;
  // key-projection path into dz, db_H
  BagDhkParams d = {};
  d.tile_info = p.tile_info; d.num_tiles = p.num_tiles; d.total_rows = p.total_rows;
  d.h = static_cast<const __half*>(a->h_saved);
  d.dz = static_cast<__nv_bfloat16*>(a->dz_ws);
  d.part_db = a->part_db; d.dg_max = a->dg_max; d.keep_scale = p.keep_scale;
  CUtensorMap tm_dkg_a = tm_dkg;
  if ((rc = check_cuda(launch_bag_dhk(tm_dkg_a, tm_wk, tm_dz, d, ns, st), "bag_dhk_kernel"))) return rc;
  rs = side_fork(side, 2, st, &ferr);
  if (ferr != cudaSuccess) return check_cuda(ferr, "mpo_bag_bwd_nacagat: side-stream fork");
  if ((rc = check_cuda(launch_bag_bwd_reduce(bag->tile_prefix, a->part_dqk, a->part_db, a->dqk, a->grad_b_h, 0,
                                             bag->num_tiles, rs), "bag_bwd_reduce_kernel (bias)"))) return rc;
  // weight gradients: dW_H += dz^T X (bf16), dW_k += dkg^T H / gs (fp16)
  CUtensorMap tm_dz64, tm_x64, tm_dkg64, tm_h64;
  if ((rc = make_tmap_bf16_2d(&tm_dz64, a->dz_ws, R, kD, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm_x64, bag->x, R, kDIn, 64, 64))) return rc;
  if ((rc = make_tmap_16b_2d(&tm_dkg64, a->dkg_ws, R, kD, 64, 64, true))) return rc;
  if ((rc = make_tmap_16b_2d(&tm_h64, a->h_saved, R, kD, 64, 64, true))) return rc;
  if ((rc = check_cuda(launch_bag_bwd_dw(tm_dz64, tm_x64, a->grad_w_h, p.total_rows, kDIn, kDIn, false, nullptr, ns, st),
                       "bag_bwd_dw_kernel"))) return rc;
  if ((rc = check_cuda(launch_bag_bwd_dw(tm_dkg64, tm_h64, a->grad_w_k, p.total_rows, kD, kD, true, a->dg_max, ns, st),
                       "bag_bwd_dw_kernel (W_k)"))) return rc;
  return rs != st ? check_cuda(side_join(side, st), "mpo_bag_bwd_nacagat: side-stream join") : MPO_OK;
}

// Adam with L2 weight decay over one flat fp32 parameter buffer (torch.optim.Adam semantics, the optimizer of the
// reference drivers: models/mcat/main.py:298-299), fused with zeroing the gradient buffer for the next window.
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 float lr, float b1, float b2, float eps, float wd, const int32_t* __restrict__ step_dev, int zero_grad) {
  mpo::pdl_enter();
  const float t = static_cast<float>(*step_dev + 1);
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x * 4) {
    float4 pv = *reinterpret_cast<float4*>(p + i), gv = *reinterpret_cast<float4*>(g + i);
    float4 mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ge = gp[e] + wd * pp[e];
      mp[e] = mp[e] + (ge - mp[e]) * (1.f - b1);
      vp[e] = b2 * vp[e] + (1.f - b2) * ge * ge;
      pp[e] -= step_size * mp[e] / (sqrtf(vp[e]) / bc2_sqrt + eps);
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__global__ void bump_step_kernel(int32_t* s) { mpo::pdl_enter(); *s += 1; }

}  // extern "C"
namespace mpo {
// Adam kernel over [0, n) on `st`; bump: increment *step_dev afterwards (one more 1-thread launch)
int launch_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                float eps, float weight_decay, int32_t* step_dev, bool zero_grad, bool bump, cudaStream_t st) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step_dev || n < 0 || (n & 3))
    return fail(MPO_E_ARG, "%s", "mpo_adam_step: null pointer or n not a multiple of 4");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_adam_step: no CUDA device (this library has no CPU fallback)");
  if (n > 0) {
    const int64_t groups = n / 4;
    int blocks = static_cast<int>((groups + 255) / 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    mpo::launch_step(adam_step_kernel, dim3(blocks), dim3(256), 0, st, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                             step_dev, zero_grad ? 1 : 0);
    count_launch();
  }
  if (bump) { mpo::launch_step(bump_step_kernel, dim3(1), dim3(1), 0, st, step_dev); count_launch(); }
  return check_cuda(cudaGetLastError(), "adam_step_kernel");
}
}  // namespace mpo
extern "C" {

int mpo_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int32_t* step_dev, int32_t zero_grad, void* stream) {
  if (n == 0 && !(zero_grad & MPO_ADAM_NO_BUMP)) return MPO_OK;
  return mpo::launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_dev,
                          (zero_grad & 1) != 0, (zero_grad & MPO_ADAM_NO_BUMP) == 0, static_cast<cudaStream_t>(stream));
}

int mpo_lse_combine(const float* lse_in, const float* pooled_in, int32_t nshards, float* lse_out, float* pooled_out,
                    void* stream) {
  if (nshards <= 0 || !lse_in || !pooled_in || !lse_out || !pooled_out)
    return fail(MPO_E_ARG, "%s", "mpo_lse_combine: bad arguments");
  lse_combine_kernel<<<kQ, kD, 0, static_cast<cudaStream_t>(stream)>>>(lse_in, pooled_in, nshards, lse_out, pooled_out);
  count_launch();
  return check_cuda(cudaGetLastError(), "lse_combine_kernel");
}

}  // extern "C"
