// fp32-accurate GEMM on tcgen05 for the bag-scale products of GE-NaCAGaT (reference: models/ge_nacagat/ge_nacagat.py:49,53:
// N x N self-attention over the patches and the 8-head N-token encoder layers).
//
//   C[M x N] (+)= alpha * A[M x K] * B[N x K]^T          fp32 in, fp32 out, fp32 accumulation in TMEM
//
// The operands are fp32 activations whose products feed soft-max arguments and returned attention maps (1e-3 parity
// gate), so each is split into a bf16 pair (hi = bf16(x), lo = bf16(x - hi)) by split_bf16_kernel and the product runs as
// three bf16 MMAs per K step: hi*hi + hi*lo + lo*hi (the dropped lo*lo term is 2^-16 of the result).  Either operand may
// be K-major (rows of K, as stored for Q / K / dctx / P / dS) or MN-major (rows of the GEMM's K dimension, as stored for
// V or for P^T / dS^T read straight out of P / dS), so no transposed copy of an N x N matrix is ever made.
// Persistent: each CTA walks 128 x BN output tiles (BN = 128 or 64) with stride gridDim.x.  Warp 0 = TMA producer
// (3-stage ring of 64-deep K blocks, 128B swizzle, running ahead across tile boundaries), warp 1 = MMA issuer (two
// accumulator stages of BN TMEM columns, so the epilogue of tile i overlaps the loads and MMAs of tile i + 1), warps 2..5 =
// epilogue: TMEM -> registers -> alpha -> a 128B-swizzled [128 x 32] fp32 slab in shared memory -> one TMA store per slab
// (coalesced 128 B rows; a thread-per-row direct store touches 32 different rows per instruction and was 9x off the
// write roofline for the N x N outputs).  Outputs whose row pitch is not a multiple of 16 B take the direct-store path.
#include "../../include/mpo_b200.h"
#include "mpo_ptx.cuh"
#include "mpo_common.cuh"
#include "launchers.h"
#include "tail_dev.cuh"

namespace mpo {

constexpr int kTcStages = 3;
constexpr int kTcBK = 64;
constexpr int kTcBM = 128;
constexpr int kTcABytes = kTcBM * kTcBK * 2;        // 16 KB per hi / lo tile of A
constexpr int kTcBBytesMax = 128 * kTcBK * 2;       // 16 KB per hi / lo tile of B at BN = 128
constexpr int kTcStageBytes = 2 * kTcABytes + 2 * kTcBBytesMax;   // 64 KB
constexpr int kTcSlabBytes = kTcBM * 32 * 4;        // one [128 x 32] fp32 output slab: 16 KB
constexpr int kTcThreads = 64 + 128;
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 2 * kTcSlabBytes + 256 + 1024;
static_assert(kTcSmemBytes <= 227 * 1024, "tensor-core GEMM shared memory");

struct TcGemmParams {
  float* C;
  long long ldc;
  int M, N, K;
  float alpha;
  int accumulate;
  int bn;            // 128 or 64
  int a_mn, b_mn;    // 1: the operand is MN-major (its array has the GEMM's K dimension as rows)
  int tiles_m, num_tiles;
  int tma_store;     // 1: the epilogue stores through tm_c
  int ksplit;        // > 1: split-K work items (tile, split); partial products are added with red.global.add
};
// work item -> output tile and K-block range (every split owns at least one K block: ksplit <= nkb)
struct TcItem { int m0, n0, kb0, kb1; };
__device__ __forceinline__ TcItem tc_item(const TcGemmParams& p, int item, int nkb) {
  const int tile = item % p.num_tiles, split = item / p.num_tiles;
  TcItem w;
  w.m0 = (tile % p.tiles_m) * kTcBM;
  w.n0 = (tile / p.tiles_m) * p.bn;
  w.kb0 = static_cast<int>(static_cast<long long>(split) * nkb / p.ksplit);
  w.kb1 = static_cast<int>(static_cast<long long>(split + 1) * nkb / p.ksplit);
  return w;
}

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const __grid_constant__ CUtensorMap tm_c, const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* slab = smem + kTcStages * kTcStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slab + 2 * kTcSlabBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kTcStages;
  uint64_t* tfull = bars + 2 * kTcStages;          // [2] accumulator stage complete
  uint64_t* tempty = bars + 2 * kTcStages + 2;     // [2] accumulator stage drained (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + kTcBK - 1) / kTcBK;
  const int b_bytes = p.bn * kTcBK * 2;
  const int num_items = p.num_tiles * p.ksplit;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
    if (p.tma_store) tma_prefetch_desc(&tm_c);
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol = policy_evict_last();          // operand panels are re-read by the other tiles of the row / column
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const TcItem w = tc_item(p, item, nkb);
        const int m0 = w.m0, n0 = w.n0;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kTcStageBytes;
          uint8_t* sb = sa + 2 * kTcABytes;
          mbar_expect_tx(&full_bar[stage], 2 * kTcABytes + 2 * b_bytes);
          const int k0 = kb * kTcBK;
          if (!p.a_mn) {        // K-major: one box {64 k, 128 rows}
            tma_load_2d(sa, &tm_a_hi, &full_bar[stage], k0, m0, pol);
            tma_load_2d(sa + kTcABytes, &tm_a_lo, &full_bar[stage], k0, m0, pol);
          } else {              // M-major: two boxes {64 m, 64 k rows}
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              tma_load_2d(sa + j * 8192, &tm_a_hi, &full_bar[stage], m0 + j * 64, k0, pol);
              tma_load_2d(sa + kTcABytes + j * 8192, &tm_a_lo, &full_bar[stage], m0 + j * 64, k0, pol);
            }
          }
          if (!p.b_mn) {        // K-major: one box {64 k, bn rows}
            tma_load_2d(sb, &tm_b_hi, &full_bar[stage], k0, n0, pol);
            tma_load_2d(sb + kTcBBytesMax, &tm_b_lo, &full_bar[stage], k0, n0, pol);
          } else {              // N-major: bn / 64 boxes {64 n, 64 k rows}
            for (int j = 0; j < p.bn / 64; ++j) {
              tma_load_2d(sb + j * 8192, &tm_b_hi, &full_bar[stage], n0 + j * 64, k0, pol);
              tma_load_2d(sb + kTcBBytesMax + j * 8192, &tm_b_lo, &full_bar[stage], n0 + j * 64, k0, pol);
            }
          }
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTcBM, static_cast<uint32_t>(p.bn), static_cast<uint32_t>(p.a_mn),
                                             static_cast<uint32_t>(p.b_mn));
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const TcItem w = tc_item(p, item, nkb);
        const int acc = it & 1;
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(acc * 128);
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * kTcStageBytes), a_lo = a_hi + kTcABytes;
          const uint32_t b_hi = a_hi + 2 * kTcABytes, b_lo = b_hi + kTcBBytesMax;
#pragma unroll
          for (int k = 0; k < kTcBK / 16; ++k) {
            // K-major: 16 k = 32 B along the 128 B row; MN-major: 16 k rows = two 8-row swizzle atoms = 2048 B down the box,
            // LBO = 8 KB between the 64-element MN atoms
            const uint32_t ao = p.a_mn ? k * 2048 : k * 32, bo = p.b_mn ? k * 2048 : k * 32;
            const uint32_t albo = p.a_mn ? 8192 : 16, blbo = p.b_mn ? 8192 : 16;
            const uint64_t dah = umma_desc_sw128(a_hi + ao, albo, 1024), dal = umma_desc_sw128(a_lo + ao, albo, 1024);
            const uint64_t dbh = umma_desc_sw128(b_hi + bo, blbo, 1024), dbl = umma_desc_sw128(b_lo + bo, blbo, 1024);
            umma_bf16(tacc, dah, dbh, idesc, (kb != w.kb0 || k != 0) ? 1u : 0u);
            umma_bf16(tacc, dah, dbl, idesc, 1u);
            umma_bf16(tacc, dal, dbh, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // epilogue: thread = one output row of the tile (TMEM lane), 32 columns per tcgen05.ld
    const int qd = warp & 3;
    const int r = qd * 32 + lane;                  // row inside the tile
    const bool elected = threadIdx.x == 64;        // first epilogue thread issues the slab stores
    int it = 0;
    uint32_t nslab = 0;                            // slabs written so far by this CTA (staging buffer = nslab & 1)
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const TcItem w = tc_item(p, item, nkb);
      const int acc = it & 1;
      const int m0 = w.m0, n0 = w.n0;
      const int row = m0 + r;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + static_cast<uint32_t>(acc * 128) + (static_cast<uint32_t>(qd * 32) << 16);
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tacc + c0, v);
        tmem_ld_wait();
        if (c0 + 32 >= p.bn) {                     // last read of this accumulator stage: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
        if (p.tma_store) {
          if (n0 + c0 < p.N) {                     // uniform over the CTA: a slab wholly outside the matrix is skipped
            uint8_t* sl = slab + (nslab & 1u) * kTcSlabBytes;
            if (elected && nslab >= 2) tma_store_wait_read1();      // the store that last read this buffer has finished
            named_bar_sync(1, 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float4 o;
              o.x = __uint_as_float(v[4 * c]) * p.alpha; o.y = __uint_as_float(v[4 * c + 1]) * p.alpha;
              o.z = __uint_as_float(v[4 * c + 2]) * p.alpha; o.w = __uint_as_float(v[4 * c + 3]) * p.alpha;
              *reinterpret_cast<float4*>(sl + r * 128 + ((c ^ (r & 7)) << 4)) = o;
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (elected) { tma_store_2d(&tm_c, sl, n0 + c0, m0); tma_store_commit(); }
            ++nslab;
          }
        } else if (row < p.M) {
          float* crow = p.C + static_cast<long long>(row) * p.ldc + n0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = n0 + c0 + j;
            if (col < p.N) {
              const float rr = __uint_as_float(v[j]) * p.alpha;
              if (p.ksplit > 1) atomicAdd(crow + c0 + j, rr);
              else crow[c0 + j] = p.accumulate ? crow[c0 + j] + rr : rr;
            }
          }
        }
      }
    }
    if (elected) tma_store_wait_read();            // shared memory stays valid until the last store has read it
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// hi = bf16(x), lo = bf16(x - hi) of a strided fp32 matrix [rows x cols] into two compact bf16 arrays of pitch `pitch`
// (a multiple of 64, columns cols..pitch-1 zero-filled: the GEMM's TMA boxes are 64 elements wide).
// DROP: x = dropout(src) with the mask of element index base + row * cols + col regenerated on the fly (the train-mode
// attention-probability dropout of the N-token encoder layers: no dropped fp32 copy of an N x N matrix is written).
template <bool DROP>
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, long long ld, int rows, int cols, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo, int pitch, DropSpec drop, uint32_t base) {
  const int groups = pitch >> 3;
  const long long total = static_cast<long long>(rows) * groups;
  uint32_t seedv = 0;
  if (DROP) seedv = drop_seed(drop);
  const bool vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / groups), c0 = static_cast<int>(i - static_cast<long long>(r) * groups) * 8;
    const float* s = src + static_cast<long long>(r) * ld + c0;
    float x[8];
    if (vec && c0 + 8 <= cols) {
      const float4 a = *reinterpret_cast<const float4*>(s), b = *reinterpret_cast<const float4*>(s + 4);
      x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = c0 + e < cols ? s[e] : 0.f;
    }
    if (DROP) {
      const uint32_t eb = base + static_cast<uint32_t>(r) * static_cast<uint32_t>(cols) + static_cast<uint32_t>(c0);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c0 + e < cols) x[e] = drop_fwd(x[e], drop, seedv, eb + e);
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = x[2 * e], x1 = x[2 * e + 1];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
      h[e] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      l[e] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
    const long long o = static_cast<long long>(r) * pitch + c0;
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

cudaError_t launch_split_bf16(const float* src, long long ld, int rows, int cols, void* hi, void* lo, int pitch,
                              cudaStream_t stream, const DropSpec* drop, uint32_t base) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(rows) * (pitch >> 3);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (drop != nullptr && drop->thr != 0)
    split_bf16_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        src, ld, rows, cols, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), pitch, *drop, base);
  else
    split_bf16_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        src, ld, rows, cols, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), pitch, DropSpec{}, 0u);
  count_launch();
  return cudaGetLastError();
}

// A operand: arrays a_hi / a_lo of a_rows x a_pitch bf16.  K-major: a_rows = M, columns = K.  M-major: a_rows = K, columns = M.
// B operand likewise (K-major: N x K; N-major: K x N).  All pitches are multiples of 64.
int launch_tc_gemm(const void* a_hi, const void* a_lo, int a_rows, int a_pitch, bool a_mn, const void* b_hi, const void* b_lo,
                   int b_rows, int b_pitch, bool b_mn, float* C, long long ldc, int M, int N, int K, float alpha, bool accumulate,
                   cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "gemm_tc_kernel attribute");
    attr_set = true;
  }
  if (M <= 0 || N <= 0 || K <= 0) return MPO_OK;
  const int bn = N > 64 ? 128 : 64;
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo, tc;
  int rc;
  const uint32_t a_box_rows = a_mn ? 64 : 128, b_box_rows = b_mn ? 64 : static_cast<uint32_t>(bn);
  if ((rc = make_tmap_bf16_2d(&ta_hi, a_hi, static_cast<uint64_t>(a_rows), static_cast<uint64_t>(a_pitch), 64, a_box_rows))) return rc;
  if ((rc = make_tmap_bf16_2d(&ta_lo, a_lo, static_cast<uint64_t>(a_rows), static_cast<uint64_t>(a_pitch), 64, a_box_rows))) return rc;
  if ((rc = make_tmap_bf16_2d(&tb_hi, b_hi, static_cast<uint64_t>(b_rows), static_cast<uint64_t>(b_pitch), 64, b_box_rows))) return rc;
  if ((rc = make_tmap_bf16_2d(&tb_lo, b_lo, static_cast<uint64_t>(b_rows), static_cast<uint64_t>(b_pitch), 64, b_box_rows))) return rc;
  // TMA store of the output needs 16-byte aligned rows; anything else (odd bag sizes) takes the direct-store epilogue
  const bool tma_store = !accumulate && (reinterpret_cast<uintptr_t>(C) & 15u) == 0 && (ldc & 3) == 0;
  if (tma_store) {
    if ((rc = make_tmap_f32_2d(&tc, C, static_cast<uint64_t>(M), static_cast<uint64_t>(N), static_cast<uint64_t>(ldc), kTcBM))) return rc;
  } else {
    tc = ta_hi;     // unused
  }
  TcGemmParams p;
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.accumulate = accumulate ? 1 : 0; p.bn = bn;
  p.a_mn = a_mn ? 1 : 0; p.b_mn = b_mn ? 1 : 0;
  p.tiles_m = (M + kTcBM - 1) / kTcBM;
  p.num_tiles = p.tiles_m * ((N + bn - 1) / bn);
  p.tma_store = tma_store ? 1 : 0;
  const int sms = num_sms();
  // accumulating products with few output tiles and a bag-long reduction (the weight gradients of the N-token layers):
  // split K so that every SM has a work item; the partial products are added atomically
  p.ksplit = 1;
  const int nkb = (K + kTcBK - 1) / kTcBK;
  if (accumulate && p.num_tiles < sms && nkb >= 8) {
    int ks = sms / p.num_tiles;
    if (ks > nkb / 4) ks = nkb / 4;
    if (ks > 1) p.ksplit = ks;
  }
  const int items = p.num_tiles * p.ksplit;
  const int grid = items < sms ? items : sms;
  gemm_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, tc, p);
  count_launch();
  return check_cuda(cudaGetLastError(), "gemm_tc_kernel");
}

}  // namespace mpo
