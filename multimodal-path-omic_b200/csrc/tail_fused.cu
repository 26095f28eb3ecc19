// Fused cluster tail: the 6-token part of MCAT / NaCAGaT with concat fusion (reference models/mcat/mcat.py:90-138,
// models/nacagat/nacagat.py:80-138, models/blocks.py:51-111,232-253 and their autograd) as
//   snn_fwd_kernel  the six SNN encoders, batched over the slides (cluster of 8 CTAs per omic group and 32 slides)
//   pre_kernel      query projection, key fold (and NaCAGaT's key-bias term)              (before the bag forward)
//   path_kernel x2  launch 1: omic and path encoders + pooling as two concurrent cluster roles; launch 2: fusion,
//                   survival head, loss, and the whole data-gradient chain back to d(pooled) / dG, again by role
//                                                                              (between bag forward and backward)
//   pre_bwd_kernel  fold / query projection data gradients; snn_bwd_kernel: SNN data gradients (after the bag backward)
//   wgrad_kernel    every weight / bias / LayerNorm gradient of the tail as one grouped launch (twice per step)
// One thread-block cluster of 4 CTAs owns S slides (M = 6 S token rows; S = 1 up to 35 slides per step).  Activations
// are replicated in the shared memory of the 4 CTAs; every linear layer is split by output columns (two 32-column
// blocks = two attention heads per CTA), its weight slice is streamed L2 -> shared memory through a 3-slot ring of 16 KB
// chunks that runs ahead across layer boundaries (path kernels: one TMA tensor copy per chunk behind a per-slot
// mbarrier; pre / SNN kernels and odd shapes: per-thread cp.async), results are broadcast to the peers through
// distributed shared memory, and the per-row work (LayerNorm, the 6x6 attention of the CTA's heads, pooling soft-max,
// survival head, loss) happens in place.  The GEMM blocks are paced by the shared-memory pipe, not by latency or
// arithmetic (profiles/r2d_tail_phase_timing.txt): hence the half-warp step (gemm_step2) and the TMA ring.  The one-slide kernels fit twice per SM (113 KB, 128 registers), so the two roles hide each
// other's latencies.  Slides never mix inside these kernels, so no grid-wide synchronisation is needed; only the
// weight gradients sum over slides, and they are deferred to wgrad_kernel.  fp32 CUDA-core arithmetic throughout
// (packed FFMA2), the 1e-3 parity gate of SURVEY.md 8c.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/mpo_b200.h"
#include "launchers.h"
#include "tail_dev.cuh"
#include "tail_fused.h"
#include "tail_ws.h"

namespace mpo {
namespace fused {
using tailws::Ws;

constexpr int E = 256;
constexpr int FF = 512;
constexpr int CL = 4;              // CTAs per cluster: 33 clusters of 4 CTAs (215 KB each) fit a B200 at once, only 15 of 8
constexpr int NV = 8;              // 32-column blocks of a 256-wide layer (= heads of the encoder layers)
constexpr int NB = NV / CL;        // column blocks (heads) owned by one CTA: "virtual ranks" rank * NB + nb
constexpr int NT = 256;            // threads per CTA
constexpr int NW = NT / 32;
constexpr int KC = 128;            // reduction extent of one weight chunk (128: the one-slide kernels fit twice per SM)
constexpr int KW = KC / (NT / 32); // reduction elements of a chunk per warp
constexpr int WLD = KC + 4;        // padded row pitch of a forward chunk [32][WLD]
constexpr int CHUNK = 32 * WLD;    // floats per ring slot (a data-gradient chunk [256][32] fits as well)
constexpr int NSTAGE = 3;
// Depth of the weight ring by slides per cluster S of the kernel: the one-slide kernels (M = 6) sit twice on an SM and have room
// for 3 slots; the two-slide kernels (S = 2, M = 12) own the SM and take 6, i.e. 5 chunks (80 KB) in flight per CTA
template <int S> struct RingDepth { static constexpr int n = S == 2 ? 6 : NSTAGE; };
constexpr int MAXK = 8;            // survival bins supported by the fused kernels
constexpr int OMIC_LD = 608;       // shared-memory pitch of one omic input row (d_i <= 608)

enum : int { T_FWD = 0, T_DGRAD = 1 };

// The weight stream of a kernel is a flat list of chunks, built on the host in the order the device code consumes
// them.  T_FWD (y = x W^T): 32 rows n0..n0+31 of W (output features) x kc reduction elements, rows contiguous;
// T_DGRAD (dx = dz W): kc rows of W (reduction over output features) x 32 columns n0..n0+31.
// n0 = rank * rank_mul + n_off + blk * blk_stride, folded into `base` + rank * rank_stride.
struct Chunk {
  const float* base;     // first element of the chunk for rank 0
  int rank_stride;       // floats to add per cluster rank
  short ld;              // row pitch of W (in_features)
  short kc_type;         // kc (reduction elements, multiple of 4, <= KC) | type << 15
};
constexpr int MAX_CHUNKS = 160;    // per launch and role (path role of NaCAGaT: 100 forward, 124 backward)
// TMA form of a chunk (same 16 bytes; Program::tma != 0): the chunk is ONE box of tensor map `map` -- data gradient:
// 32 columns x KC rows of the 2-D matrix, dense; forward: 32 rows x KC reduction elements of the matrix seen as
// [ld / 32][rows][32], i.e. KC / 32 tiles [32 rows][32 floats] with the 128-byte swizzle
struct ChunkT {
  int c0, c1;            // inner (column) and outer (row) coordinate of the chunk for rank 0
  int map_rmul;          // tensor-map index | elements to add per cluster rank << 8 (to c1 forward, to c0 data gradient)
  short ld;
  short kc_type;
};
static_assert(sizeof(ChunkT) == sizeof(Chunk), "one table format for both forms");
constexpr int NMAP = 24;           // distinct (weight matrix, chunk type) pairs per launch and role
struct Program { Chunk c[MAX_CHUNKS]; int n; int tma; };

// tensor maps of the weight matrices, cached by (base, pitch, chunk type): parameters keep their storage between steps
struct MapKey { const float* w; int ld, type; };
struct MapCache {
  static constexpr int N = 256;
  MapKey key[N]; CUtensorMap map[N]; int n = 0;
  const CUtensorMap* get(const float* w, int ld, int type) {
    for (int i = 0; i < n; ++i) if (key[i].w == w && key[i].ld == ld && key[i].type == type) return &map[i];
    if (n >= N) n = 0;                   // (a process that cycles through > 256 weight matrices simply re-encodes)
    // rows: an upper bound (full chunks never leave the matrix; the extent only has to cover every row that is read)
    const int rc = type == T_FWD ? make_tmap_f32_rows32(&map[n], w, 1u << 20, static_cast<uint64_t>(ld), KC / 32)
                                 : make_tmap_f32_2d_sw(&map[n], w, 1u << 20, static_cast<uint64_t>(ld),
                                                       static_cast<uint64_t>(ld), KC, false);
    if (rc != MPO_OK) return nullptr;
    key[n] = MapKey{w, ld, type};
    return &map[n++];
  }
};
inline MapCache& map_cache() { static MapCache c; return c; }

struct ProgBuilder {
  Program& p;
  bool ok = true;
  // TMA form: the tensor maps of this program (kernel parameter space) and the matrices they stand for
  CUtensorMap* maps = nullptr;
  int nmaps = 0;
  MapKey mkey[NMAP];
  int map_index(const float* w, int ld, int type) {
    for (int i = 0; i < nmaps; ++i) if (mkey[i].w == w && mkey[i].ld == ld && mkey[i].type == type) return i;
    if (nmaps >= NMAP) return -1;
    const CUtensorMap* m = map_cache().get(w, ld, type);
    if (m == nullptr) return -1;
    maps[nmaps] = *m; mkey[nmaps] = MapKey{w, ld, type};
    return nmaps++;
  }
  void add(int type, const float* w, int ld, int K, int nblk, int rank_mul, int blk_stride, int n_off) {
    int mi = -1;
    if (p.tma) {
      // full chunks of 16-byte aligned, 16-byte pitched matrices only; anything else turns the whole program back
      // into the cp.async form (the caller rebuilds it)
      if (K % KC != 0 || (ld & 31) != 0 || (reinterpret_cast<uintptr_t>(w) & 15u) != 0 || rank_mul > 255 ||
          (mi = map_index(w, ld, type)) < 0) { ok = false; return; }
    }
    for (int blk = 0; blk < nblk; ++blk)
      for (int k0 = 0; k0 < K; k0 += KC) {
        if (p.n >= MAX_CHUNKS) { ok = false; return; }
        Chunk& c = p.c[p.n++];
        const int n0 = n_off + blk * blk_stride, kc = K - k0 < KC ? K - k0 : KC;
        if (p.tma) {
          ChunkT& t = reinterpret_cast<ChunkT&>(c);
          if (type == T_FWD) { t.c0 = k0; t.c1 = n0; } else { t.c0 = n0; t.c1 = k0; }
          t.map_rmul = mi | (rank_mul << 8);
        } else if (type == T_FWD) { c.base = w + static_cast<size_t>(n0) * ld + k0; c.rank_stride = rank_mul * ld; }
        else { c.base = w + static_cast<size_t>(k0) * ld + n0; c.rank_stride = rank_mul; }
        c.ld = static_cast<short>(ld);
        c.kc_type = static_cast<short>(kc | (type << 15));
      }
  }
  void fwd(const float* w, int ld, int K, int nblk = 1, int rank_mul = 32, int blk_stride = 32, int n_off = 0) {
    add(T_FWD, w, ld, K, nblk, rank_mul, blk_stride, n_off);
  }
  void dgrad(const float* w, int ld, int K, int nblk = 1, int rank_mul = 32, int blk_stride = 32, int n_off = 0) {
    add(T_DGRAD, w, ld, K, nblk, rank_mul, blk_stride, n_off);
  }
};

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// Phase timing of the path kernels (-DMPO_TAIL_PROF builds only; scripts/gpu_tail_prof.py): thread 0 of CTA 0 adds clock64
// intervals to g_prof[8 * pass + phase]; phase 0 whole kernel, 1 gemm_block total, 2 ring wait + CTA barrier, 3 chunk
// issue, 4 FFMA2 block, 5 partial store + CTA barrier, 6 cluster barriers, 7 gemm_block calls
#ifdef MPO_TAIL_PROF
__device__ unsigned long long g_prof[32];
__device__ int g_prof_pass;
#define PROF_ON (blockIdx.x == 0 && threadIdx.x == 0)
#define PROF_T() (PROF_ON ? clock64() : 0ll)
#define PROF_ADD(i, v) do { if (PROF_ON) g_prof[8 * g_prof_pass + (i)] += static_cast<unsigned long long>(v); } while (0)
#else
#define PROF_T() 0ll
#define PROF_ADD(i, v) do { } while (0)
#endif
__device__ __forceinline__ int cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return static_cast<int>(r);
}
__device__ __forceinline__ void cluster_sync() {
  [[maybe_unused]] const long long pt0 = PROF_T();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  PROF_ADD(6, PROF_T() - pt0);
}
// store v at the same shared-memory offset in every CTA of the cluster
template <int N>
__device__ __forceinline__ void bcast_n(float* local, float v) {
  const uint32_t a = smem_addr(local);
#pragma unroll
  for (int rk = 0; rk < N; ++rk) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rk));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
  }
}
__device__ __forceinline__ void bcast(float* local, float v) {
  const uint32_t a = smem_addr(local);
#pragma unroll
  for (int rk = 0; rk < CL; ++rk) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rk));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
  }
}
// explicit shared-memory loads: pointers that cross a noinline call boundary lose their address space and would be
// read with generic loads
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
// 16 bytes of shared memory as two packed f32x2 operands
__device__ __forceinline__ void lds2x64(uint32_t a, unsigned long long& lo, unsigned long long& hi) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(a));
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
// packed fp32 FMA (FFMA2 on sm_100): {a.lo * b.lo + c.lo, a.hi * b.hi + c.hi}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// cp.async ring over the chunk sequence of a Program (copied to shared memory at kernel start); chunk c lives in
// slot c % NSTAGE and chunks cons and cons+1 are always in flight or landed
struct Pipe {
  uint32_t tbl;        // shared-memory address of the chunk table (bases already offset by the cluster rank)
  uint32_t ring;       // shared-memory address of the ring
  int n, rank;
  int cons;            // next chunk to consume; chunks cons .. cons + NST - 2 are in flight or landed
  int slot;            // ring slot of chunk `cons` (kept incrementally: no modulo in the loop)
  // TMA form (maps != nullptr; the path kernels): a chunk arrives as one or four 2-D tensor copies issued by thread 0
  // behind the slot's mbarrier -- whole 128 B lines into shared memory (4 wavefronts per 512 B against ~10 for the
  // sector-wise cp.async fill) and no per-thread address work.  (32 row-wise 1-D bulk copies per chunk were tried first:
  // right for the shared-memory pipe, but their issue alone took ~2 200 cycles per chunk.)
  const CUtensorMap* maps;
  uint32_t bars;       // shared-memory address of one mbarrier per slot
  uint32_t phase;      // bit s: parity the next wait on slot s's mbarrier expects
  uint32_t stride;     // bytes per ring slot (dense 16 KB slots in the TMA form: 1024-byte aligned swizzle atoms)
};
constexpr uint32_t TMA_SLOT = 32 * KC * 4;
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (ok == 0u);
}
// partial chunks (kc < KC: the omic input layers and ragged reduction tails) -- rare, kept out of line
__device__ __noinline__ void pipe_issue_partial(const float* src, int ld, int kc, int type, uint32_t dst) {
  const int t = threadIdx.x;
  if (type == T_FWD) {
    const int per_row = kc >> 2;
    for (int p = t; p < 32 * per_row; p += NT) {
      const int row = p / per_row, c4 = p - row * per_row;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (row * WLD + c4 * 4) * 4),
                   "l"(src + static_cast<size_t>(row) * ld + c4 * 4) : "memory");
    }
  } else {
    // rows (t >> 3) + 32 j of the chunk, 16 B at column 4 (t & 7)
    const float* s0 = src + static_cast<size_t>(t >> 3) * ld + (t & 7) * 4;
    const uint32_t d0 = dst + ((t >> 3) * 32 + (t & 7) * 4) * 4;
    const size_t sstep = static_cast<size_t>(32) * ld;
    for (int j = 0; j * 32 + (t >> 3) < kc; ++j)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + j * (32 * 32 * 4)), "l"(s0 + j * sstep) : "memory");
  }
}
// Puts chunk c into ring slot `slot`: every thread copies 4 x 16 B.  Inlined into the GEMM blocks with one code path for
// both chunk types (selects, no branches): as an out-of-line call that decoded the table entry, took a modulo and
// converted the global-memory descriptor for every copy it cost ~680 cycles per chunk, 22 % of the path kernels
// (phase timing of the -DMPO_TAIL_PROF build, profiles/r2d_tail_phase_timing.txt).
// (1-D bulk copies by one warp were tried instead: 256 copies of 128 B per data-gradient chunk are far slower)
__device__ __forceinline__ void pipe_issue(uint32_t tbl, uint32_t ring, int n, int c, int slot, uint32_t bars,
                                           const CUtensorMap* maps) {
  if (maps != nullptr) {
    if (c < n && threadIdx.x == 0) {
      uint32_t c0, c1, mi, pk;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c0), "=r"(c1), "=r"(mi), "=r"(pk) : "r"(tbl + c * 16));
      const uint32_t dst = ring + slot * TMA_SLOT, bar = bars + slot * 8;
      const unsigned long long mp = reinterpret_cast<unsigned long long>(maps + (mi & 0xffu));
      // (the slot was last READ, with generic loads that every warp finished before the CTA barrier in front of this
      // call: a write-after-read across proxies needs no proxy fence)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TMA_SLOT) : "memory");
      if ((pk >> 31) != 0u) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(mp), "r"(bar), "r"(c0), "r"(c1) : "memory");
      } else {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(mp), "r"(bar), "r"(0), "r"(c1), "r"(c0 >> 5) : "memory");
      }
    }
    return;
  }
  if (c < n) {
    const int t = threadIdx.x;
    uint32_t b_lo, b_hi, pk;
    [[maybe_unused]] uint32_t rs;      // (the per-rank stride: folded into the base by pipe_init)
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(b_lo), "=r"(b_hi), "=r"(rs), "=r"(pk) : "r"(tbl + c * 16));
    const float* src = reinterpret_cast<const float*>(static_cast<unsigned long long>(b_lo) |
                                                      (static_cast<unsigned long long>(b_hi) << 32));
    const int ld = pk & 0xffff, kc = (pk >> 16) & 0x7fff;
    const bool dg = (pk >> 31) != 0u;
    const uint32_t dst = ring + slot * (CHUNK * 4);
    if (kc == KC) {
      // forward chunk [32 rows][KC]: rows t / 32 + 8 j, 16 B at column 4 (t % 32), row pitch WLD in shared memory;
      // data-gradient chunk [KC rows][32]: rows t / 8 + 32 j, 16 B at column 4 (t % 8), dense
      const int row = dg ? (t >> 3) : (t >> 5);
      const int seg = dg ? (t & 7) : (t & 31);
      const uint32_t d0 = dst + (row * (dg ? 32 : WLD) + seg * 4) * 4;
      const uint32_t dstep = dg ? 32 * 32 * 4 : 8 * WLD * 4;
      const float* s0 = src + (row * ld + seg * 4);
      const size_t sstep = static_cast<size_t>((dg ? 32 : 8) * ld);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + j * dstep), "l"(s0 + j * sstep) : "memory");
    } else {
      pipe_issue_partial(src, ld, kc, dg ? T_DGRAD : T_FWD, dst);
    }
  }
  cp_async_commit();
}
// copies the chunk table to shared memory (folding the cluster rank into the chunk bases) and puts the first NST - 1
// chunks in flight
template <int NST = NSTAGE>
__device__ __forceinline__ void pipe_init(Pipe& pp, const Chunk* chunks, int n, float* tbl_smem, float* ring_smem, int rank,
                                          const CUtensorMap* maps = nullptr, float* bars_smem = nullptr) {
  for (int i = threadIdx.x; i < n; i += NT) {
    Chunk c = chunks[i];
    if (maps != nullptr) {
      ChunkT& t = reinterpret_cast<ChunkT&>(c);
      const int add = rank * (t.map_rmul >> 8);
      if (((t.kc_type >> 15) & 1) == T_FWD) t.c1 += add; else t.c0 += add;
    } else {
      c.base += static_cast<size_t>(rank) * c.rank_stride;
    }
    reinterpret_cast<Chunk*>(tbl_smem)[i] = c;
  }
  pp.maps = maps;
  pp.bars = maps != nullptr ? smem_addr(bars_smem) : 0u;
  pp.phase = 0u;
  pp.stride = maps != nullptr ? TMA_SLOT : static_cast<uint32_t>(CHUNK * 4);
  if (maps != nullptr && threadIdx.x == 0) {
    for (int sl = 0; sl < NST; ++sl) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pp.bars + sl * 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pp.tbl = smem_addr(tbl_smem); pp.ring = smem_addr(ring_smem); pp.n = n; pp.rank = rank; pp.cons = 0; pp.slot = 0;
  // swizzle atoms want 1024-byte aligned slots: the dense TMA slots leave NST * 512 B of the ring region to align with
  if (maps != nullptr) pp.ring = (pp.ring + 1023u) & ~1023u;
#pragma unroll
  for (int c = 0; c < NST - 1; ++c) pipe_issue(pp.tbl, pp.ring, pp.n, c, c, pp.bars, pp.maps);
}
template <int NST = NSTAGE>
__device__ __forceinline__ void pipe_init(Pipe& pp, const Program& prog, float* tbl_smem, float* ring_smem, int rank,
                                          const CUtensorMap* maps = nullptr, float* bars_smem = nullptr) {
  pipe_init<NST>(pp, prog.c, prog.n, tbl_smem, ring_smem, rank, maps, bars_smem);
}

struct Dev {
  int rank, t, lane, warp;
  int s0, B, grow0, Rtot;     // first slide of the cluster, slides in the batch, first token row, token rows in the batch
  uint32_t seedv;
  float* ws;
  float* red;                 // [NW][M][32] cross-warp reduction scratch
  Pipe* pipe;                 // the only state that crosses the noinline GEMM calls (d itself stays in registers)
};
struct RowCtx { int rank, lane, warp, grow0, Rtot; uint32_t seedv; };    // by-value context of the noinline row helpers
__device__ __forceinline__ RowCtx row_ctx(const Dev& d) { return RowCtx{d.rank, d.lane, d.warp, d.grow0, d.Rtot, d.seedv}; }

// red[warp][r][lane] = sum over this warp's k-slices of x[r][k] Wslice(k, lane), for the next ceil(Ktot / KC) chunks.
// acc[r] = {sum over even k, sum over odd k} as one packed register pair: two FFMA2 per row and 4 k.  TYPE and the
// pitch LDX of x are compile-time so that every shared-memory load of a k-step is base register + immediate.
template <int M, int TYPE, int LDX>
__device__ __forceinline__ void gemm_step(uint32_t wk, uint32_t xk, unsigned long long (&acc)[M]) {
  unsigned long long w0, w1;
  if (TYPE == T_FWD) {
    lds2x64(wk, w0, w1);
  } else {
    w0 = pack2(lds32(wk), lds32(wk + 128));
    w1 = pack2(lds32(wk + 256), lds32(wk + 384));
  }
#pragma unroll
  for (int r = 0; r < M; ++r) {
    unsigned long long x0, x1;
    lds2x64(xk + r * LDX * 4, x0, x1);
    acc[r] = fma2(w1, x1, fma2(w0, x0, acc[r]));
  }
}
// Half-warp form (M <= 12): lanes 0..15 take reduction quad 2 j, lanes 16..31 quad 2 j + 1, each for 16 columns at a time
// (group A = the lane's own half of the 32 columns, group B = the other half, so that the two halves of the warp never
// meet in a bank).  An activation load then fetches TWO distinct quads per instruction for the same two wavefronts a
// full broadcast costs: 20 shared-memory wavefronts per 8 reduction elements instead of 32 (the path kernels are paced by
// the shared-memory pipe, profiles/r2d_tail_phase_timing.txt), and 8 loads instead of 14 per 24 FFMA2.
template <int M, int TYPE, int LDX>
__device__ __forceinline__ void gemm_step2(uint32_t wa, uint32_t wb, uint32_t xk, unsigned long long (&acc)[M][2]) {
  unsigned long long a0, a1, b0, b1;
  if (TYPE == T_FWD) {
    lds2x64(wa, a0, a1);
    lds2x64(wb, b0, b1);
  } else {
    a0 = pack2(lds32(wa), lds32(wa + 128));
    a1 = pack2(lds32(wa + 256), lds32(wa + 384));
    b0 = pack2(lds32(wb), lds32(wb + 128));
    b1 = pack2(lds32(wb + 256), lds32(wb + 384));
  }
#pragma unroll
  for (int r = 0; r < M; ++r) {
    unsigned long long x0, x1;
    lds2x64(xk + r * LDX * 4, x0, x1);
    acc[r][0] = fma2(a1, x1, fma2(a0, x0, acc[r][0]));
    acc[r][1] = fma2(b1, x1, fma2(b0, x0, acc[r][1]));
  }
}
template <int M, int TYPE, int LDX, int NST = NSTAGE>
__device__ __noinline__ void gemm_block(Pipe& pp, uint32_t red0, const float* __restrict__ xs, int Ktot) {
  constexpr bool kHalf = M <= 12;        // half-warp form (gemm_step2); the 32-row SNN blocks keep one column per lane
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hc = lane & 15, hh = lane >> 4;
  const uint32_t tbl = pp.tbl, ring = pp.ring;
  const int nchunks = pp.n;
  const uint32_t bars = pp.bars, stride = pp.stride;
  const CUtensorMap* const maps = pp.maps;
  const bool tma = maps != nullptr;
  uint32_t phase = pp.phase;
  int cons = pp.cons, slot = pp.slot;
  [[maybe_unused]] const long long pt_in = PROF_T();
  [[maybe_unused]] long long pt_d = pt_in;
  unsigned long long acc[kHalf ? 1 : M];
  unsigned long long acc2[kHalf ? M : 1][2];
#pragma unroll
  for (int r = 0; r < (kHalf ? 1 : M); ++r) acc[r] = 0ull;
#pragma unroll
  for (int r = 0; r < (kHalf ? M : 1); ++r) acc2[r][0] = acc2[r][1] = 0ull;
  for (int kb = 0; kb < Ktot; kb += KC) {
    [[maybe_unused]] const long long pt_a = PROF_T();
    const int kc = min(KC, Ktot - kb);
    if (tma) {
      mbar_wait_u32(bars + slot * 8, (phase >> slot) & 1u);
      phase ^= 1u << slot;
    } else {
      cp_async_wait<NST - 2>();
    }
    __syncthreads();
    [[maybe_unused]] const long long pt_b = PROF_T();
    pipe_issue(tbl, ring, nchunks, cons + NST - 1, slot == 0 ? NST - 1 : slot - 1, bars, maps);     // the slot freed by chunk cons - 1
    [[maybe_unused]] const long long pt_c = PROF_T();
    PROF_ADD(2, pt_b - pt_a); PROF_ADD(3, pt_c - pt_b);
    const uint32_t wsm = ring + slot * stride;
    ++cons;
    slot = slot + 1 == NST ? 0 : slot + 1;
    const int kbeg = warp * KW;
    if constexpr (kHalf) {
      // this warp's KW reduction elements as KW / 8 pairs of quads; quad 2 j + hh is this lane's
      const int k0 = kbeg + 4 * hh;
      const uint32_t xk = smem_addr(xs) + (kb + k0) * 4;
      const int ca = 16 * hh + hc, cb = 16 * (1 - hh) + hc;
      const uint32_t wa = TYPE == T_FWD ? wsm + (ca * WLD + k0) * 4 : wsm + (k0 * 32 + ca) * 4;
      const uint32_t wb = TYPE == T_FWD ? wsm + (cb * WLD + k0) * 4 : wsm + (k0 * 32 + cb) * 4;
      constexpr int WSTEP = TYPE == T_FWD ? 32 : 8 * 128;
      if (TYPE == T_FWD && tma) {
        // four boxes [32 rows][32 floats] with the 128-byte swizzle: the 16-byte quad q of row c sits at q ^ (c & 7);
        // this warp's 16 reduction elements are quads 4 (warp & 1) .. + 3 of box warp / 2 (full chunks only)
        const uint32_t ba = wsm + (warp >> 1) * 4096 + ca * 128, bb = wsm + (warp >> 1) * 4096 + cb * 128;
        const int q0 = (warp & 1) * 4 + hh, sw = hc & 7;
#pragma unroll
        for (int j = 0; j < KW / 8; ++j)
          gemm_step2<M, TYPE, LDX>(ba + (((q0 + 2 * j) ^ sw) << 4), bb + (((q0 + 2 * j) ^ sw) << 4), xk + j * 32, acc2);
      } else if (kc == KC) {
#pragma unroll
        for (int j = 0; j < KW / 8; ++j) gemm_step2<M, TYPE, LDX>(wa + j * WSTEP, wb + j * WSTEP, xk + j * 32, acc2);
      } else {
        const int nq = (min(kbeg + KW, kc) - kbeg) >> 2;       // valid quads of this warp's slice, may be <= 0
        for (int j = 0; j < KW / 8; ++j)
          if (2 * j + hh < nq) gemm_step2<M, TYPE, LDX>(wa + j * WSTEP, wb + j * WSTEP, xk + j * 32, acc2);
      }
    } else {
      // this warp's KW reduction elements: k-step stride is 16 B in x and in a forward chunk row, 4 rows in a dgrad chunk
      const uint32_t xk = smem_addr(xs) + (kb + kbeg) * 4;
      const uint32_t wk = TYPE == T_FWD ? wsm + (lane * WLD + kbeg) * 4 : wsm + (kbeg * 32 + lane) * 4;
      constexpr int WSTEP = TYPE == T_FWD ? 16 : 4 * 128;
      if (kc == KC) {
#pragma unroll
        for (int j = 0; j < KW / 4; ++j) gemm_step<M, TYPE, LDX>(wk + j * WSTEP, xk + j * 16, acc);
      } else {
        const int nst = (min(kbeg + KW, kc) - kbeg) >> 2;       // may be <= 0
        for (int j = 0; j < nst; ++j) gemm_step<M, TYPE, LDX>(wk + j * WSTEP, xk + j * 16, acc);
      }
    }
#ifdef MPO_TAIL_PROF
    {   // (the accumulators have to be complete for the interval to mean anything)
      unsigned long long sink = 0;
#pragma unroll
      for (int r = 0; r < (kHalf ? 1 : M); ++r) sink ^= acc[r];
#pragma unroll
      for (int r = 0; r < (kHalf ? M : 1); ++r) sink ^= acc2[r][0] ^ acc2[r][1];
      if (sink == 0x123456789abcdefull) g_prof[31] = 1;
    }
#endif
    pt_d = PROF_T();
    PROF_ADD(4, pt_d - pt_c);
  }
  pp.cons = cons; pp.slot = slot; pp.phase = phase;
  // this warp's k-slice partials; reduce_epi() sums the 8 slices
  const uint32_t red = red0 + ((warp * M) * 32 + lane) * 4;
#pragma unroll
  for (int r = 0; r < M; ++r) {
    float v;
    if constexpr (kHalf) {
      // column `lane` = group A of this lane + group B of the lane in the other half
      float alo, ahi, blo, bhi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(alo), "=f"(ahi) : "l"(acc2[r][0]));
      asm("mov.b64 {%0, %1}, %2;" : "=f"(blo), "=f"(bhi) : "l"(acc2[r][1]));
      v = (alo + ahi) + __shfl_xor_sync(0xffffffffu, blo + bhi, 16);
    } else {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[r]));
      v = lo + hi;
    }
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(red + r * 128), "f"(v) : "memory");
  }
  __syncthreads();
  [[maybe_unused]] const long long pt_e = PROF_T();
  PROF_ADD(5, pt_e - pt_d); PROF_ADD(1, pt_e - pt_in); PROF_ADD(7, 1);
}
// sums the 8 k-slices; thread (warp, lane) finishes outputs (row warp + 8 i, column lane): epi(row, i, value)
template <int M, class Epi>
__device__ __forceinline__ void reduce_epi(Dev& d, Epi epi) {
#pragma unroll
  for (int i = 0; i < (M + NW - 1) / NW; ++i) {
    const int r = d.warp + NW * i;
    if (r < M) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) v += d.red[(w * M + r) * 32 + d.lane];
      epi(r, i, v);
    }
  }
}

// rows grow0 .. grow0+M-1 of a [rows][256] global array -> shared [M][256] (zeros past the end of the batch)
template <int M>
__device__ __forceinline__ void load_rows(const Dev& d, float* dst, const float* src, int grow0, int rtot) {
  for (int i = d.t; i < M * 64; i += NT) {
    const int r = i >> 6, c4 = i & 63;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow0 + r < rtot) v = __ldcg(reinterpret_cast<const float4*>(src + static_cast<size_t>(grow0 + r) * E) + c4);
    reinterpret_cast<float4*>(dst)[r * 64 + c4] = v;
  }
}
__device__ __forceinline__ DropSpec site_of(const DropSpec& base, uint32_t site) {
  DropSpec s = base;
  s.site = site;
  return s;
}

// LayerNorm over 256 features (eps 1e-5, biased variance), one warp per row, replicated in every CTA of the cluster;
// row r is written to global memory by the CTA of rank r % 8
template <int M>
__device__ __noinline__ void ln_fwd_rows(const RowCtx d, const float* src, float* dst, const float* gamma, const float* beta,
                                            float* y_g, float* xh_g, float* rs_g) {
  float gm[8], bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { gm[j] = __ldg(gamma + d.lane + 32 * j); bt[j] = __ldg(beta + d.lane + 32 * j); }
  for (int r = d.warp; r < M; r += NW) {
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = src[r * E + d.lane + 32 * j]; s += v[j]; }
    s = warp_sum(s);
    const float mu = s * (1.f / 256.f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float dd = v[j] - mu; q = fmaf(dd, dd, q); }
    q = warp_sum(q);
    const float rs = rsqrtf(q * (1.f / 256.f) + 1e-5f);
    const int grow = d.grow0 + r;
    const bool own = grow < d.Rtot && (r % CL) == d.rank;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = d.lane + 32 * j;
      const float xh = (v[j] - mu) * rs;
      const float y = fmaf(xh, gm[j], bt[j]);
      dst[r * E + c] = y;
      if (own) { xh_g[static_cast<size_t>(grow) * E + c] = xh; y_g[static_cast<size_t>(grow) * E + c] = y; }
    }
    if (own && d.lane == 0) rs_g[grow] = rs;
  }
}
// dr = LayerNorm backward of dy (shared), dd = dr through the dropout layer in front of the residual branch
template <int M>
__device__ __noinline__ void ln_bwd_rows(const RowCtx d, const float* dy, float* dr, float* dd, const float* gamma,
                                            const float* xh_g, const float* rs_g, float* dy_g, float* dd_g,
                                            const DropSpec dsp) {
  float gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gm[j] = __ldg(gamma + d.lane + 32 * j);
  for (int r = d.warp; r < M; r += NW) {
    const int grow = d.grow0 + r;
    const bool valid = grow < d.Rtot;
    const bool own = valid && (r % CL) == d.rank;
    float xh[8], dxh[8], g[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = d.lane + 32 * j;
      xh[j] = valid ? __ldcg(xh_g + static_cast<size_t>(grow) * E + c) : 0.f;
      g[j] = dy[r * E + c];
      dxh[j] = g[j] * gm[j];
      s1 += dxh[j];
      s2 = fmaf(dxh[j], xh[j], s2);
    }
    s1 = warp_sum(s1) * (1.f / 256.f);
    s2 = warp_sum(s2) * (1.f / 256.f);
    const float rs = valid ? __ldcg(rs_g + grow) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = d.lane + 32 * j;
      const float v = rs * (dxh[j] - s1 - xh[j] * s2);
      float vd = v;
      if (dsp.thr != 0) vd *= drop_grad(dsp, d.seedv, static_cast<uint32_t>(grow) * E + c);
      dr[r * E + c] = v;
      dd[r * E + c] = vd;
      if (own) { dy_g[static_cast<size_t>(grow) * E + c] = g[j]; dd_g[static_cast<size_t>(grow) * E + c] = vd; }
    }
  }
}

// ------------------------------------------------------------------------------------------------ parameter blocks
struct EncP { const float *w_in, *b_in, *w_out, *b_out, *w1, *b1, *w2, *b2, *g1, *be1, *g2, *be2; };
struct EncW { int x, qkv, probs, ctx, y1, xh1, rs1, f, y2, xh2, rs2, dy2, df2, df, dy1, dsa, dqkv; };
struct PoolP { const float *wa, *ba, *wb, *bb, *wc, *bc, *wr, *br; float *gwc, *gbc; };
struct PoolW { int a, b, w, hp, dzr, da, db; };

// ContextualAttentionGate of NaCAGaT (models/blocks.py:232-253)
struct CagP { const float *w1, *b1, *w2, *b2, *w3, *b3, *wc, *bc, *gG, *bG, *gE, *bE; };
struct CagW { int f1, f2, f3, u, w, Gg, Gxh, Grs, Ee, Exh, Ers, m, C, t0, dGg, dEe, df1, df2, df3; };
__device__ __forceinline__ float elu_f(float v) { return v > 0.f ? v : expm1f(v); }
__device__ __forceinline__ float elu_d(float y) { return y > 0.f ? 1.f : y + 1.f; }      // from the ELU output

struct PathParams {
  CUtensorMap maps[2][NMAP];   // TMA form of the weight ring: the tensor maps the chunks of prog[role] index
  Program prog[2];             // chunk streams of the path-role ([0]) and omic-role ([1]) clusters
  int nroles;                  // 2: adjacent clusters 2g / 2g+1 run the path / omic branch of slide group g concurrently
  int off_dG2;                 // NaCAGaT: the CAG's dQ (added to dG by pre_bwd_kernel; the omic role owns dG)
  EncP enc[4];                 // path.0, path.1, omic.0, omic.1
  EncW encw[4];
  PoolP pool[2];               // path, omic
  PoolW poolw[2];
  const float *wv, *bv, *wo, *bo, *wf0, *bf0, *wf2, *bf2, *wcl, *bcl;
  float* ws;
  const float* pooled;
  float* dpooled;
  float *hazards, *S, *Y, *att_path, *att_omic;
  int off_G, off_v, off_hc, off_cat, off_z1, off_z2, off_logits, off_dlogits, off_dz1, off_dz2, off_dhc, off_dv, off_dG;
  // loss (F_LOSS) or upstream gradients (F_BWD without F_LOSS)
  int loss_kind;
  float loss_alpha, loss_eps, grad_scale;
  const int64_t* label;
  const float* censor;
  float *loss, *dhaz_out, *dS_out;
  const float *dhaz_in, *dS_in, *dY_in;
  DropSpec d_model, d_quarter;
  int B, K, flags;
  // NaCAGaT: CAG(G, q) added to the co-attention output (blocks.py:110-111); attention dropout leaves sum_n a' != 1
  int nac;
  CagP cag;
  CagW cagw;
  const float* qp;             // [B][6][256] projected queries (Q-hat of the CAG)
  const float* suma;           // [B][6] or null
  float* dsuma;                // [B][6] or null
  int off_dqp;
};
static_assert(sizeof(PathParams) <= 20000, "kernel parameter space (large kernel parameters, CUDA 12.1+)");
static_assert(sizeof(Program) <= 2608, "two chunk programs have to fit the kernel parameter space");

// ------------------------------------------------------------------------------------------------ encoder layer
// nn.TransformerEncoderLayer(256, nhead 8, ff 512, relu, post-norm) as built at models/mcat/mcat.py:51-53.
// in: x in XA (replicated); out: y2 in XA.  CTA `rank` owns head `rank` of the self-attention.
template <int S>
__device__ __forceinline__ void enc_fwd(Dev& d, const EncP& p, const EncW& w, int eidx, const DropSpec& dm, float* XA,
                                        float* XB, float* XC, float* BIG, float* QKVL) {
  constexpr int M = 6 * S;
  float* ws = d.ws;
  const uint32_t s0 = SITE_ENC + 4 * eidx;      // attention probabilities, dropout1, feed-forward dropout, dropout2
  // every bias of the layer up front: one exposed global-load latency per layer instead of one in front of each GEMM
  float bias_in[NB][3], bias_out[NB], bias_1[2 * NB], bias_2[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int c = (d.rank * NB + nb) * 32 + d.lane;
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) bias_in[nb][blk] = __ldg(p.b_in + blk * E + c);
    bias_out[nb] = __ldg(p.b_out + c);
    bias_2[nb] = __ldg(p.b2 + c);
  }
#pragma unroll
  for (int blk = 0; blk < 2 * NB; ++blk) bias_1[blk] = __ldg(p.b1 + d.rank * (64 * NB) + blk * 32 + d.lane);
  // packed in-projection: this CTA computes q, k, v of its heads
#pragma unroll
  for (int nb = 0; nb < NB; ++nb)
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {
      const int col = blk * E + (d.rank * NB + nb) * 32 + d.lane;
      const float bias = bias_in[nb][blk];
      gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XA, E);
      reduce_epi<M>(d, [&](int r, int, float v) {
        v += bias;
        QKVL[(nb * M + r) * 96 + blk * 32 + d.lane] = v;
        const int grow = d.grow0 + r;
        if (grow < d.Rtot) ws[w.qkv + static_cast<size_t>(grow) * 768 + col] = v;
      });
    }
  __syncthreads();
  // 6 x 6 attention of head `rank`, one warp per (slide, query token); lane = head dimension
  {
    const DropSpec da = site_of(dm, s0);
    const float scale = 0.17677669529663687f;   // 1/sqrt(32)
    for (int it = d.warp; it < NB * M; it += NW) {
      const int nb = it / M, r1 = it - nb * M, head = d.rank * NB + nb;
      const float* Q = QKVL + nb * M * 96;
      const int sl = r1 / 6, l1 = r1 - sl * 6;
      const float q = Q[r1 * 96 + d.lane];
      float sc[6];
      float mx = -INFINITY;
#pragma unroll
      for (int l2 = 0; l2 < 6; ++l2) {
        sc[l2] = warp_sum(q * Q[(sl * 6 + l2) * 96 + 32 + d.lane]) * scale;
        mx = fmaxf(mx, sc[l2]);
      }
      float sum = 0.f;
#pragma unroll
      for (int l2 = 0; l2 < 6; ++l2) { sc[l2] = expf(sc[l2] - mx); sum += sc[l2]; }
      const float inv = 1.f / sum;
      const int slide = d.s0 + sl;
      const uint32_t pi0 = ((static_cast<uint32_t>(slide) * 8 + head) * 6 + l1) * 6;
      float c = 0.f;
#pragma unroll
      for (int l2 = 0; l2 < 6; ++l2) {
        const float pr = sc[l2] * inv;
        if (d.lane == 0 && slide < d.B) ws[w.probs + pi0 + l2] = pr;       // kept before dropout
        const float pd = da.thr != 0 ? drop_fwd(pr, da, d.seedv, pi0 + l2) : pr;
        c = fmaf(pd, Q[(sl * 6 + l2) * 96 + 64 + d.lane], c);
      }
      const int col = head * 32 + d.lane;
      bcast(XB + r1 * E + col, c);
      if (slide < d.B) ws[w.ctx + static_cast<size_t>(d.grow0 + r1) * E + col] = c;
    }
  }
  cluster_sync();
  // out-projection + dropout1 + residual
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    const float bias = bias_out[nb];
    const DropSpec d1 = site_of(dm, s0 + 1);
    gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XB, E);
    reduce_epi<M>(d, [&](int r, int, float v) {
      v += bias;
      if (d1.thr != 0) v = drop_fwd(v, d1, d.seedv, static_cast<uint32_t>(d.grow0 + r) * E + col);
      bcast(XC + r * E + col, v + XA[r * E + col]);
    });
  }
  cluster_sync();
  ln_fwd_rows<M>(row_ctx(d), XC, XA, p.g1, p.be1, ws + w.y1, ws + w.xh1, ws + w.rs1);
  __syncthreads();
  // feed-forward
#pragma unroll
  for (int blk = 0; blk < 2 * NB; ++blk) {
    const int col = d.rank * (64 * NB) + blk * 32 + d.lane;
    const float bias = bias_1[blk];
    const DropSpec d2 = site_of(dm, s0 + 2);
    gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XA, E);
    reduce_epi<M>(d, [&](int r, int, float v) {
      v = fmaxf(v + bias, 0.f);
      const int grow = d.grow0 + r;
      if (d2.thr != 0) v = drop_fwd(v, d2, d.seedv, static_cast<uint32_t>(grow) * FF + col);
      bcast(BIG + r * FF + col, v);
      if (grow < d.Rtot) ws[w.f + static_cast<size_t>(grow) * FF + col] = v;
    });
  }
  cluster_sync();
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    const float bias = bias_2[nb];
    const DropSpec d3 = site_of(dm, s0 + 3);
    gemm_block<M, T_FWD, FF, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), BIG, FF);
    reduce_epi<M>(d, [&](int r, int, float v) {
      v += bias;
      if (d3.thr != 0) v = drop_fwd(v, d3, d.seedv, static_cast<uint32_t>(d.grow0 + r) * E + col);
      bcast(XC + r * E + col, v + XA[r * E + col]);
    });
  }
  cluster_sync();
  ln_fwd_rows<M>(row_ctx(d), XC, XA, p.g2, p.be2, ws + w.y2, ws + w.xh2, ws + w.rs2);
  __syncthreads();
}

// in: dy2 in XA (replicated); out: dx in XA (and in global memory at dx_g when non-null)
template <int S>
__device__ __forceinline__ void enc_bwd(Dev& d, const EncP& p, const EncW& w, int eidx, const DropSpec& dm, float* XA,
                                        float* XB, float* XC, float* BIG, float* DCTX, float* QKVL, float* SCR,
                                        float* dx_g) {
  constexpr int M = 6 * S;
  float* ws = d.ws;
  const uint32_t s0 = SITE_ENC + 4 * eidx;
  // norm2: dr2 -> XB, gradient of linear2's output (through dropout2) -> XC
  ln_bwd_rows<M>(row_ctx(d), XA, XB, XC, p.g2, ws + w.xh2, ws + w.rs2, ws + w.dy2, ws + w.df2, site_of(dm, s0 + 3));
  __syncthreads();
  // linear2 data gradient with the ReLU / feed-forward-dropout derivative: gradient at linear1's pre-activation
  for (int blk = 0; blk < 2 * NB; ++blk) {
    const int col = d.rank * (64 * NB) + blk * 32 + d.lane;
    const DropSpec d2 = site_of(dm, s0 + 2);
    float fv[(M + NW - 1) / NW];
#pragma unroll
    for (int i = 0; i < (M + NW - 1) / NW; ++i) {
      const int grow = d.grow0 + d.warp + NW * i;
      fv[i] = (d.warp + NW * i < M && grow < d.Rtot) ? __ldcg(ws + w.f + static_cast<size_t>(grow) * FF + col) : 0.f;
    }
    gemm_block<M, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XC, E);
    reduce_epi<M>(d, [&](int r, int i, float v) {
      const int grow = d.grow0 + r;
      if (d2.thr != 0) v *= drop_grad(d2, d.seedv, static_cast<uint32_t>(grow) * FF + col);
      v = fv[i] > 0.f ? v : 0.f;          // a kept element is positive exactly when its ReLU output was
      bcast(BIG + r * FF + col, v);
      if (grow < d.Rtot) ws[w.df + static_cast<size_t>(grow) * FF + col] = v;
    });
  }
  cluster_sync();
  // linear1 data gradient + the residual branch: dy1
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    gemm_block<M, T_DGRAD, FF, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), BIG, FF);
    reduce_epi<M>(d, [&](int r, int, float v) { bcast(XA + r * E + col, v + XB[r * E + col]); });
  }
  cluster_sync();
  // norm1: dr1 -> XB, gradient of the attention block's output (through dropout1) -> XC
  ln_bwd_rows<M>(row_ctx(d), XA, XB, XC, p.g1, ws + w.xh1, ws + w.rs1, ws + w.dy1, ws + w.dsa, site_of(dm, s0 + 1));
  __syncthreads();
  // out-projection data gradient: this CTA's column blocks are its own heads
  for (int nb = 0; nb < NB; ++nb) {
    gemm_block<M, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XC, E);
    reduce_epi<M>(d, [&](int r, int, float v) { DCTX[(nb * M + r) * 32 + d.lane] = v; });
  }
  __syncthreads();
  // attention backward of head `rank`, spread over all warps: (A) dA = dctx v^T as 36 S warp dot products,
  // (B) soft-max backward per (slide, query) row, (C) dq = ds k, dk = ds^T q, dv = (a mask)^T dctx
  {
    const DropSpec da = site_of(dm, s0);
    const float scale = 0.17677669529663687f;
    constexpr int NP = NB * S * 36;   // (head, slide, query, key) entries of this CTA
    float* PRm = SCR;                 // saved probabilities   [NB][S][6][6]
    float* DAm = SCR + NP;            // dctx . v
    float* AMG = SCR + 2 * NP;        // a * dropout derivative
    float* DSm = SCR + 3 * NP;        // gradient of the scaled scores
    for (int i = d.t; i < NB * M * 96; i += NT) {
      const int nb = i / (M * 96), j = i - nb * (M * 96), r = j / 96, c = j - r * 96, grow = d.grow0 + r;
      QKVL[i] = grow < d.Rtot ? __ldcg(ws + w.qkv + static_cast<size_t>(grow) * 768 + (c >> 5) * 256 +
                                       (d.rank * NB + nb) * 32 + (c & 31)) : 0.f;
    }
    for (int p = d.t; p < NP; p += NT) {
      const int nb = p / (S * 36), j = p - nb * (S * 36), sl = j / 36, slide = d.s0 + sl;
      PRm[p] = slide < d.B ? __ldcg(ws + w.probs + (static_cast<size_t>(slide) * 8 + d.rank * NB + nb) * 36 + (j - sl * 36)) : 0.f;
    }
    __syncthreads();
    for (int p = d.warp; p < NP; p += NW) {
      const int nb = p / (S * 36), j = p - nb * (S * 36), sl = j / 36, q6 = j - sl * 36, l1 = q6 / 6, l2 = q6 - l1 * 6;
      const float v = warp_sum(DCTX[(nb * M + sl * 6 + l1) * 32 + d.lane] * QKVL[(nb * M + sl * 6 + l2) * 96 + 64 + d.lane]);
      if (d.lane == 0) DAm[p] = v;
    }
    __syncthreads();
    if (d.t < NB * M) {
      const int nb = d.t / M, r = d.t - nb * M, sl = r / 6, l1 = r - sl * 6, p0 = nb * S * 36 + sl * 36 + l1 * 6;
      const uint32_t pw = ((static_cast<uint32_t>(d.s0 + sl) * 8 + d.rank * NB + nb) * 6 + l1) * 6;
      float a[6], dA[6];
      float dot = 0.f;
#pragma unroll
      for (int l2 = 0; l2 < 6; ++l2) {
        a[l2] = PRm[p0 + l2];
        const float mg = da.thr != 0 ? drop_grad(da, d.seedv, pw + l2) : 1.f;
        dA[l2] = DAm[p0 + l2] * mg;
        dot = fmaf(dA[l2], a[l2], dot);
        AMG[p0 + l2] = a[l2] * mg;
      }
#pragma unroll
      for (int l2 = 0; l2 < 6; ++l2) DSm[p0 + l2] = a[l2] * (dA[l2] - dot) * scale;
    }
    __syncthreads();
    for (int o = d.warp; o < NB * 3 * M; o += NW) {
      const int nb = o / (3 * M), j = o - nb * (3 * M), which = j / M, r = j - which * M, sl = r / 6, l = r - sl * 6;
      const float* Q = QKVL + nb * M * 96;
      const float* DC = DCTX + nb * M * 32;
      const int pb = nb * S * 36 + sl * 36;
      float acc = 0.f;
      if (which == 0) {
#pragma unroll
        for (int l2 = 0; l2 < 6; ++l2) acc = fmaf(DSm[pb + l * 6 + l2], Q[(sl * 6 + l2) * 96 + 32 + d.lane], acc);
      } else if (which == 1) {
#pragma unroll
        for (int l1 = 0; l1 < 6; ++l1) acc = fmaf(DSm[pb + l1 * 6 + l], Q[(sl * 6 + l1) * 96 + d.lane], acc);
      } else {
#pragma unroll
        for (int l1 = 0; l1 < 6; ++l1) acc = fmaf(AMG[pb + l1 * 6 + l], DC[(sl * 6 + l1) * 32 + d.lane], acc);
      }
      const int col = (d.rank * NB + nb) * 32 + d.lane;
      bcast(BIG + r * 768 + which * 256 + col, acc);
      if (d.grow0 + r < d.Rtot) ws[w.dqkv + static_cast<size_t>(d.grow0 + r) * 768 + which * 256 + col] = acc;
    }
  }
  cluster_sync();
  // in-projection data gradient + the residual branch: dx
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    gemm_block<M, T_DGRAD, 768, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), BIG, 768);
    reduce_epi<M>(d, [&](int r, int, float v) {
      v += XB[r * E + col];
      bcast(XA + r * E + col, v);
      const int grow = d.grow0 + r;
      if (dx_g != nullptr && grow < d.Rtot) dx_g[static_cast<size_t>(grow) * E + col] = v;
    });
  }
  cluster_sync();
}

// ------------------------------------------------------------------------------------------------ pooling + rho
// AttentionNetGated (models/blocks.py:42-48) + soft-max pooling + rho (models/mcat/mcat.py:105-109).
// in: tokens in XA; out: rho output broadcast into CAT[s][p * 256 + .]
template <int S>
__device__ __forceinline__ void pool_fwd(Dev& d, const PoolP& p, const PoolW& w, int pidx, const DropSpec& dm,
                                         const DropSpec& dq, float* att_out, int off_cat, float* XA, float* AL, float* BL,
                                         float* PA, float* AW, float* HP, float* CAT) {
  constexpr int M = 6 * S;
  float* ws = d.ws;
  for (int nb = 0; nb < NB; ++nb)
    for (int br = 0; br < 2; ++br) {
      const int col = (d.rank * NB + nb) * 32 + d.lane;
      const float bias = __ldg((br == 0 ? p.ba : p.bb) + col);
      const DropSpec ds = site_of(dq, SITE_POOL + 2 * pidx + br);
      float* dst = (br == 0 ? AL : BL) + nb * M * 32;
      const int off = br == 0 ? w.a : w.b;
      gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XA, E);
      reduce_epi<M>(d, [&](int r, int, float v) {
        v += bias;
        v = br == 0 ? tanhf(v) : 1.f / (1.f + expf(-v));
        const int grow = d.grow0 + r;
        if (ds.thr != 0) v = drop_fwd(v, ds, d.seedv, static_cast<uint32_t>(grow) * E + col);
        dst[r * 32 + d.lane] = v;
        if (grow < d.Rtot) ws[off + static_cast<size_t>(grow) * E + col] = v;
      });
    }
  __syncthreads();
  for (int it = d.warp; it < NB * M; it += NW) {
    const int nb = it / M, r = it - nb * M, vr = d.rank * NB + nb;
    const float v = warp_sum(AL[(nb * M + r) * 32 + d.lane] * BL[(nb * M + r) * 32 + d.lane] * __ldg(p.wc + vr * 32 + d.lane));
    if (d.lane == 0) bcast(PA + vr * M + r, v);
  }
  cluster_sync();
  if (d.t < M) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) a += PA[k * M + d.t];
    a += __ldg(p.bc);
    PA[NV * M + d.t] = a;                       // raw logits A
    const int grow = d.grow0 + d.t;
    if (d.rank == 0 && grow < d.Rtot) att_out[grow] = a;
  }
  __syncthreads();
  if (d.t < S) {
    const float* A = PA + NV * M + d.t * 6;
    float m = A[0];
#pragma unroll
    for (int l = 1; l < 6; ++l) m = fmaxf(m, A[l]);
    float e[6], sum = 0.f;
#pragma unroll
    for (int l = 0; l < 6; ++l) { e[l] = expf(A[l] - m); sum += e[l]; }
#pragma unroll
    for (int l = 0; l < 6; ++l) {
      const float wl = e[l] / sum;
      AW[d.t * 6 + l] = wl;
      if (d.rank == 0 && d.s0 + d.t < d.B) ws[w.w + (d.s0 + d.t) * 6 + l] = wl;
    }
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < S; ++s) {
    float h = 0.f;
#pragma unroll
    for (int l = 0; l < 6; ++l) h = fmaf(AW[s * 6 + l], XA[(s * 6 + l) * E + d.t], h);
    HP[s * E + d.t] = h;
    if (d.rank == 0 && d.s0 + s < d.B) ws[w.hp + static_cast<size_t>(d.s0 + s) * E + d.t] = h;
  }
  __syncthreads();
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    const float bias = __ldg(p.br + col);
    const DropSpec dr = site_of(dm, SITE_RHO + pidx);
    gemm_block<S, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), HP, E);
    reduce_epi<S>(d, [&](int s, int, float v) {
      v = fmaxf(v + bias, 0.f);
      const int slide = d.s0 + s;
      if (dr.thr != 0) v = drop_fwd(v, dr, d.seedv, static_cast<uint32_t>(slide) * E + col);
      bcast(CAT + s * 2 * E + pidx * E + col, v);
      if (slide < d.B) ws[off_cat + static_cast<size_t>(slide) * 2 * E + pidx * E + col] = v;
    });
  }
  cluster_sync();
}

// in: DZR[s][p*256 + .] = gradient at rho's pre-activation; out: gradient of the tokens in XA
template <int S>
__device__ __forceinline__ void pool_bwd(Dev& d, const PoolP& p, const PoolW& w, int pidx, const DropSpec& dq,
                                         const float* tok_g, float* XA, float* BIG, float* AL, float* BL, float* PA,
                                         float* AW, float* DHP, const float* DZR) {
  constexpr int M = 6 * S;
  float* ws = d.ws;
  // rho data gradient
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    gemm_block<S, T_DGRAD, 2 * E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), DZR + pidx * E, E);
    reduce_epi<S>(d, [&](int s, int, float v) { bcast(DHP + s * E + col, v); });
  }
  // this pooling head's tokens, own columns of the two gate branches, pooling weights
  load_rows<M>(d, XA, tok_g, d.grow0, d.Rtot);
  for (int i = d.t; i < NB * M * 32; i += NT) {
    const int nb = i / (M * 32), j = i - nb * (M * 32), r = j >> 5, n = j & 31;
    const int grow = d.grow0 + r;
    const bool valid = grow < d.Rtot;
    AL[i] = valid ? __ldcg(ws + w.a + static_cast<size_t>(grow) * E + (d.rank * NB + nb) * 32 + n) : 0.f;
    BL[i] = valid ? __ldcg(ws + w.b + static_cast<size_t>(grow) * E + (d.rank * NB + nb) * 32 + n) : 0.f;
  }
  if (d.t < M) AW[d.t] = (d.grow0 + d.t < d.Rtot) ? __ldcg(ws + w.w + d.grow0 + d.t) : 0.f;
  cluster_sync();
  // dw[r] = dhp . x[r]
  for (int r = d.warp; r < M; r += NW) {
    const int s = r / 6;
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) v = fmaf(DHP[s * E + d.lane + 32 * j], XA[r * E + d.lane + 32 * j], v);
    v = warp_sum(v);
    if (d.lane == 0) PA[r] = v;
  }
  __syncthreads();
  if (d.t < S) {
    float dot = 0.f;
#pragma unroll
    for (int l = 0; l < 6; ++l) dot = fmaf(PA[d.t * 6 + l], AW[d.t * 6 + l], dot);
#pragma unroll
    for (int l = 0; l < 6; ++l) PA[M + d.t * 6 + l] = AW[d.t * 6 + l] * (PA[d.t * 6 + l] - dot);      // dA
  }
  __syncthreads();
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    const float wc = __ldg(p.wc + col);
    const DropSpec dsa = site_of(dq, SITE_POOL + 2 * pidx), dsb = site_of(dq, SITE_POOL + 2 * pidx + 1);
    float gw = 0.f;
#pragma unroll
    for (int i = 0; i < (M + NW - 1) / NW; ++i) {
      const int r = d.warp + NW * i;
      if (r < M) {
        const int grow = d.grow0 + r;
        const uint32_t o = static_cast<uint32_t>(grow) * E + col;
        const float dA = PA[M + r];
        const float av = AL[(nb * M + r) * 32 + d.lane], bv = BL[(nb * M + r) * 32 + d.lane];   // after their dropout layers
        const float dab = dA * wc;
        float ga = 1.f, gb = 1.f, a0 = av, b0 = bv;
        if (dsa.thr != 0) { ga = drop_grad(dsa, d.seedv, o); a0 = drop_invert(av, dsa); }
        if (dsb.thr != 0) { gb = drop_grad(dsb, d.seedv, o); b0 = drop_invert(bv, dsb); }
        const float da_pre = dab * bv * ga * (1.f - a0 * a0);
        const float db_pre = dab * av * gb * b0 * (1.f - b0);
        bcast(BIG + r * FF + col, da_pre);
        bcast(BIG + r * FF + E + col, db_pre);
        if (grow < d.Rtot) {
          ws[w.da + static_cast<size_t>(grow) * E + col] = da_pre;
          ws[w.db + static_cast<size_t>(grow) * E + col] = db_pre;
          gw = fmaf(dA, av * bv, gw);
        }
      }
    }
    // attention_c parameter gradients: one atomic per (cluster, column)
    __syncthreads();
    d.red[d.warp * 32 + d.lane] = gw;
    __syncthreads();
    if (d.warp == 0 && p.gwc != nullptr) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < NW; ++k) v += d.red[k * 32 + d.lane];
      atomicAdd(p.gwc + col, v);
    }
  }
  if (d.rank == 0 && d.t == 0 && p.gbc != nullptr) {
    float v = 0.f;
    for (int r = 0; r < M; ++r) if (d.grow0 + r < d.Rtot) v += PA[M + r];
    atomicAdd(p.gbc, v);
  }
  cluster_sync();
  // gradient of the tokens: both gate branches' data gradients + the value path of the pooling
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    gemm_block<M, T_DGRAD, FF, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), BIG, FF);          // chunks: attention_a then attention_b
    reduce_epi<M>(d, [&](int r, int, float v) {
      v = fmaf(AW[r], DHP[(r / 6) * E + col], v);
      bcast(XA + r * E + col, v);
    });
  }
  cluster_sync();
}

// ------------------------------------------------------------------------------------------------ path kernel
template <int S>
struct PathSmem {
  static constexpr int M = 6 * S;
  static constexpr int ring = 0;
  static constexpr int XA = ring + RingDepth<S>::n * CHUNK;
  static constexpr int XB = XA + M * E;
  static constexpr int XC = XB + M * E;
  static constexpr int BIG = XC + M * E;
  static constexpr int RED = BIG + M * 768;
  static constexpr int QKVL = RED + NW * M * 32;
  static constexpr int AL = QKVL + NB * M * 96;
  static constexpr int BL = AL + NB * M * 32;
  static constexpr int PA = BL + NB * M * 32;         // [NV + 1][M]
  static constexpr int AW = PA + (NV + 1) * M + 4;
  static constexpr int HP = ((AW + M + 3) / 4) * 4;
  static constexpr int CAT = HP + S * E;
  static constexpr int Z1 = CAT + S * 2 * E;
  static constexpr int Z2 = Z1 + S * E;
  static constexpr int DZS = Z2 + S * E;
  static constexpr int DZ1 = DZS + S * E;
  static constexpr int DZR = DZ1 + S * E;
  static constexpr int DHP = DZR + S * 2 * E;
  static constexpr int SM = DHP + S * E;              // small per-slide vectors: 8 x [S][MAXK]
  static constexpr int TBL = ((SM + 8 * S * MAXK + 3) / 4) * 4;     // chunk table
  static constexpr int BARS = TBL + MAX_CHUNKS * 4;                 // one mbarrier per ring slot (bulk forward chunks)
  static constexpr int total = BARS + 16;
};

static_assert(PathSmem<2>::total * sizeof(float) <= 227 * 1024, "path kernel shared memory");
static_assert(PathSmem<1>::total * sizeof(float) <= 113 * 1024, "two one-slide path CTAs have to fit one SM");
template <int S>
__global__ void __launch_bounds__(NT, S == 1 ? 2 : 1) path_kernel(const __grid_constant__ PathParams P) {
  pdl_enter();
  constexpr int M = 6 * S;
  using L = PathSmem<S>;
  extern __shared__ __align__(16) float sm[];
#ifdef MPO_TAIL_PROF
  if (PROF_ON) g_prof_pass = (P.flags & F_FWD) ? 0 : 1;
  const long long pt_k0 = PROF_T();
#endif
  Dev d;
  d.rank = cluster_rank();
  d.t = threadIdx.x; d.lane = d.t & 31; d.warp = d.t >> 5;
  const int cid = blockIdx.x / CL;
  const int role = P.nroles == 2 ? (cid & 1) : 0;          // 0: path branch (and everything that is saved), 1: omic branch
  const int br_hi = P.nroles == 2 ? role : 1, br_lo = P.nroles == 2 ? role : 0;
  const bool saver = role == 0;
  d.s0 = (cid / P.nroles) * S;
  d.B = P.B; d.grow0 = d.s0 * 6; d.Rtot = 6 * P.B;
  d.seedv = drop_seed(P.d_model);
  d.ws = P.ws;
  d.red = sm + L::RED;
  Pipe pipe;
  d.pipe = &pipe;
  pipe_init<RingDepth<S>::n>(pipe, P.prog[role], sm + L::TBL, sm + L::ring, d.rank,
                             P.prog[role].tma != 0 ? &P.maps[role][0] : nullptr, sm + L::BARS);
  float* ws = P.ws;
  float *XA = sm + L::XA, *XB = sm + L::XB, *XC = sm + L::XC, *BIG = sm + L::BIG, *QKVL = sm + L::QKVL;
  float *AL = sm + L::AL, *BL = sm + L::BL, *PA = sm + L::PA, *AW = sm + L::AW, *HP = sm + L::HP, *CAT = sm + L::CAT;
  float *Z1 = sm + L::Z1, *Z2 = sm + L::Z2, *DZS = sm + L::DZS, *DZ1 = sm + L::DZ1, *DZR = sm + L::DZR, *DHP = sm + L::DHP;
  float* LG = sm + L::SM;                 // logits
  float* HZ = LG + S * MAXK;              // hazards
  float* SV = HZ + S * MAXK;              // S
  float* YV = SV + S * MAXK;              // Y
  float* DH = YV + S * MAXK;              // d hazards
  float* DS = DH + S * MAXK;              // d S
  float* DYv = DS + S * MAXK;             // d Y
  float* DLG = DYv + S * MAXK;            // d logits
  const int K = P.K;
  const DropSpec& dm = P.d_model;
  const DropSpec& dq = P.d_quarter;
  cluster_sync();                         // every CTA of the cluster is running before any remote store

  if (P.flags & F_FWD) {
    // omic branch first (encoders + pooling over the SNN tokens, mcat.py:102,111-115), then the path branch (value /
    // output projections of the pooled vectors -- the folded form of mcat.py:97 -- encoders, pooling); one copy of the
    // code, run twice
#pragma unroll 1
    for (int br = br_hi; br >= br_lo; --br) {
      if (br == 1) {
        load_rows<M>(d, XA, ws + P.off_G, d.grow0, d.Rtot);
      } else {
        load_rows<M>(d, XA, P.pooled, d.grow0, d.Rtot);
        for (int nb = 0; nb < NB; ++nb) {
          const int col = (d.rank * NB + nb) * 32 + d.lane;
          const float bias = __ldg(P.bv + col);
          float sa[(M + NW - 1) / NW];        // v_i = W_v pooled_i + (sum_n a'_in) b_v   (blocks.py:189-192)
#pragma unroll
          for (int i = 0; i < (M + NW - 1) / NW; ++i) {
            const int grow = d.grow0 + d.warp + NW * i;
            sa[i] = (P.suma != nullptr && d.warp + NW * i < M && grow < d.Rtot) ? __ldcg(P.suma + grow) : 1.f;
          }
          gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XA, E);
          reduce_epi<M>(d, [&](int r, int i, float v) {
            v = fmaf(bias, sa[i], v);
            bcast(XB + r * E + col, v);
            if (d.grow0 + r < d.Rtot) ws[P.off_v + static_cast<size_t>(d.grow0 + r) * E + col] = v;
          });
        }
        cluster_sync();
        for (int nb = 0; nb < NB; ++nb) {
          const int col = (d.rank * NB + nb) * 32 + d.lane;
          const float bias = __ldg(P.bo + col);
          gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XB, E);
          reduce_epi<M>(d, [&](int r, int, float v) {
            v += bias;
            bcast(XA + r * E + col, v);
            if (d.grow0 + r < d.Rtot) ws[P.off_hc + static_cast<size_t>(d.grow0 + r) * E + col] = v;
          });
        }
        cluster_sync();
        if (P.nac) {
          // C = fc_c( LN_G(ELU(fc1 G + fc2 q)) * LN_E(ELU(fc3 q)) ), every fc = Linear + ELU; hc += C
          const CagP& cp = P.cag;
          const CagW& cw = P.cagw;
          load_rows<M>(d, XB, ws + P.off_G, d.grow0, d.Rtot);
          load_rows<M>(d, XC, P.qp, d.grow0, d.Rtot);
          for (int nb = 0; nb < NB; ++nb) {
            const int col = (d.rank * NB + nb) * 32 + d.lane;
            const float b1 = __ldg(cp.b1 + col), b2 = __ldg(cp.b2 + col), b3 = __ldg(cp.b3 + col);
            float f1v[(M + NW - 1) / NW];
            gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XB, E);
            reduce_epi<M>(d, [&](int r, int i, float v) {
              v = elu_f(v + b1);
              f1v[i] = v;
              if (d.grow0 + r < d.Rtot) ws[cw.f1 + static_cast<size_t>(d.grow0 + r) * E + col] = v;
            });
            gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XC, E);
            reduce_epi<M>(d, [&](int r, int i, float v) {
              v = elu_f(v + b2);
              const float u = elu_f(f1v[i] + v);
              bcast(BIG + r * E + col, u);
              if (d.grow0 + r < d.Rtot) {
                ws[cw.f2 + static_cast<size_t>(d.grow0 + r) * E + col] = v;
                ws[cw.u + static_cast<size_t>(d.grow0 + r) * E + col] = u;
              }
            });
            gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XC, E);
            reduce_epi<M>(d, [&](int r, int, float v) {
              v = elu_f(v + b3);
              const float wv = elu_f(v);
              bcast(BIG + M * E + r * E + col, wv);
              if (d.grow0 + r < d.Rtot) {
                ws[cw.f3 + static_cast<size_t>(d.grow0 + r) * E + col] = v;
                ws[cw.w + static_cast<size_t>(d.grow0 + r) * E + col] = wv;
              }
            });
          }
          cluster_sync();
          ln_fwd_rows<M>(row_ctx(d), BIG, XB, cp.gG, cp.bG, ws + cw.Gg, ws + cw.Gxh, ws + cw.Grs);
          ln_fwd_rows<M>(row_ctx(d), BIG + M * E, XC, cp.gE, cp.bE, ws + cw.Ee, ws + cw.Exh, ws + cw.Ers);
          __syncthreads();
          for (int i = d.t; i < M * E; i += NT) {
            const int r = i / E, grow = d.grow0 + r;
            const float mv = XB[i] * XC[i];
            XB[i] = mv;
            if (grow < d.Rtot && (r % CL) == d.rank) ws[cw.m + static_cast<size_t>(grow) * E + (i - r * E)] = mv;
          }
          for (int nb = 0; nb < NB; ++nb) {
            const int col = (d.rank * NB + nb) * 32 + d.lane;
            const float bc = __ldg(cp.bc + col);
            gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XB, E);
            reduce_epi<M>(d, [&](int r, int, float v) {
              v = elu_f(v + bc);
              const float hv = XA[r * E + col] + v;
              bcast(BIG + 2 * M * E + r * E + col, hv);
              if (d.grow0 + r < d.Rtot) {
                ws[cw.C + static_cast<size_t>(d.grow0 + r) * E + col] = v;
                ws[P.off_hc + static_cast<size_t>(d.grow0 + r) * E + col] = hv;
              }
            });
          }
          cluster_sync();
          for (int i = d.t; i < M * E; i += NT) XA[i] = BIG[2 * M * E + i];
          __syncthreads();
        }
      }
#pragma unroll 1
      for (int l = 0; l < 2; ++l) enc_fwd<S>(d, P.enc[2 * br + l], P.encw[2 * br + l], 2 * br + l, dm, XA, XB, XC, BIG, QKVL);
      pool_fwd<S>(d, P.pool[br], P.poolw[br], br, dm, dq, br == 1 ? P.att_omic : P.att_path, P.off_cat, XA, AL, BL, PA, AW,
                  HP, CAT);
    }
  }
  if (P.flags & F_HEAD) {
    if (!(P.flags & F_FWD) || P.nroles == 2) {
      // [h_path | h_omic] was written by the two roles of the previous launch
      for (int i = d.t; i < S * 2 * E; i += NT) {
        const int s = i / (2 * E);
        CAT[i] = (d.s0 + s < d.B) ? __ldcg(ws + P.off_cat + static_cast<size_t>(d.s0 + s) * 2 * E + (i - s * 2 * E)) : 0.f;
      }
      __syncthreads();
    }
    // ---- concat fusion (fusion.py:17-19)
    for (int nb = 0; nb < NB; ++nb) {
      const int col = (d.rank * NB + nb) * 32 + d.lane;
      const float bias = __ldg(P.bf0 + col);
      gemm_block<S, T_FWD, 2 * E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), CAT, 2 * E);
      reduce_epi<S>(d, [&](int s, int, float v) {
        v = fmaxf(v + bias, 0.f);
        bcast(Z1 + s * E + col, v);
        if (saver && d.s0 + s < d.B) ws[P.off_z1 + static_cast<size_t>(d.s0 + s) * E + col] = v;
      });
    }
    cluster_sync();
    for (int nb = 0; nb < NB; ++nb) {
      const int col = (d.rank * NB + nb) * 32 + d.lane;
      const float bias = __ldg(P.bf2 + col);
      gemm_block<S, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), Z1, E);
      reduce_epi<S>(d, [&](int s, int, float v) {
        v = fmaxf(v + bias, 0.f);
        bcast(Z2 + s * E + col, v);
        if (saver && d.s0 + s < d.B) ws[P.off_z2 + static_cast<size_t>(d.s0 + s) * E + col] = v;
      });
    }
    cluster_sync();
    // ---- classifier + survival head (mcat.py:126-138), replicated in every CTA
    for (int pr = d.warp; pr < S * K; pr += NW) {
      const int s = pr / K, k = pr - s * K;
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) v = fmaf(Z2[s * E + d.lane + 32 * j], __ldg(P.wcl + k * E + d.lane + 32 * j), v);
      v = warp_sum(v);
      if (d.lane == 0) LG[s * MAXK + k] = v + __ldg(P.bcl + k);
    }
    __syncthreads();
    if (d.t < S) {
      const int s = d.t, slide = d.s0 + s;
      float m = -INFINITY;
      for (int j = 0; j < K; ++j) m = fmaxf(m, LG[s * MAXK + j]);
      float sum = 0.f, sv = 1.f;
      for (int j = 0; j < K; ++j) {
        const float z = LG[s * MAXK + j];
        const float hz = 1.f / (1.f + expf(-z));
        HZ[s * MAXK + j] = hz;
        sv *= (1.f - hz);
        SV[s * MAXK + j] = sv;
        sum += expf(z - m);
      }
      for (int j = 0; j < K; ++j) YV[s * MAXK + j] = expf(LG[s * MAXK + j] - m) / sum;
      if (saver && d.rank == 0 && slide < d.B) {
        for (int j = 0; j < K; ++j) {
          P.hazards[slide * K + j] = HZ[s * MAXK + j];
          P.S[slide * K + j] = SV[s * MAXK + j];
          P.Y[slide * K + j] = YV[s * MAXK + j];
          ws[P.off_logits + slide * K + j] = LG[s * MAXK + j];
        }
      }
    }
    __syncthreads();
  }

  if (P.flags & F_BWD) {
    if (!(P.flags & F_HEAD)) {
      // activations of the forward pass come back from the workspace
      for (int i = d.t; i < S * E; i += NT) {
        const int s = i / E, c = i - s * E;
        const bool valid = d.s0 + s < d.B;
        Z1[i] = valid ? __ldcg(ws + P.off_z1 + static_cast<size_t>(d.s0 + s) * E + c) : 0.f;
        Z2[i] = valid ? __ldcg(ws + P.off_z2 + static_cast<size_t>(d.s0 + s) * E + c) : 0.f;
        CAT[s * 2 * E + c] = valid ? __ldcg(ws + P.off_cat + static_cast<size_t>(d.s0 + s) * 2 * E + c) : 0.f;
        CAT[s * 2 * E + E + c] = valid ? __ldcg(ws + P.off_cat + static_cast<size_t>(d.s0 + s) * 2 * E + E + c) : 0.f;
      }
      if (d.t < S * K) {
        const int s = d.t / K, j = d.t - s * K, slide = d.s0 + s;
        const bool valid = slide < d.B;
        HZ[s * MAXK + j] = valid ? P.hazards[slide * K + j] : 0.5f;
        SV[s * MAXK + j] = valid ? P.S[slide * K + j] : 0.5f;
        YV[s * MAXK + j] = valid ? P.Y[slide * K + j] : 0.f;
      }
      __syncthreads();
    }
    // ---- loss gradients (models/loss.py:5-43) or the caller's upstream gradients
    if (d.t < S) {
      const int s = d.t, slide = d.s0 + s;
      const bool valid = slide < d.B;
      for (int j = 0; j < K; ++j) { DH[s * MAXK + j] = 0.f; DS[s * MAXK + j] = 0.f; DYv[s * MAXK + j] = 0.f; }
      if (P.flags & F_LOSS) {
        if (valid) {
          const int y = static_cast<int>(P.label[slide]);
          const float c = P.censor[slide];
          const float alpha = P.loss_alpha, eps = P.loss_eps, gs = P.grad_scale;
          float l = NAN;
          if (y >= 0 && y < K) {
            const float s_prev = y == 0 ? 1.f : SV[s * MAXK + y - 1];
            const float h_y = HZ[s * MAXK + y];
            const float unc = -(1.f - c) * (logf(fmaxf(s_prev, eps)) + logf(fmaxf(h_y, eps)));
            float w_unc;
            if (P.loss_kind == MPO_LOSS_NLL) {
              const float s_y = SV[s * MAXK + y];
              const float cen = -c * logf(fmaxf(s_y, eps));
              l = (1.f - alpha) * (cen + unc) + alpha * unc;
              w_unc = 1.f;
              if (s_y > eps) DS[s * MAXK + y] += gs * (1.f - alpha) * (-c / s_y);
            } else {
              const float s_y = fmaxf(SV[s * MAXK + y], eps);
              const float ce = -(c * logf(s_y) + (1.f - c) * logf(1.f - s_y));
              l = (1.f - alpha) * ce + alpha * unc;
              w_unc = alpha;
              if (SV[s * MAXK + y] > eps) DS[s * MAXK + y] += gs * (1.f - alpha) * (-(c / s_y) + (1.f - c) / (1.f - s_y));
            }
            if (y >= 1 && s_prev > eps) DS[s * MAXK + y - 1] += gs * w_unc * (-(1.f - c) / s_prev);
            if (h_y > eps) DH[s * MAXK + y] += gs * w_unc * (-(1.f - c) / h_y);
          }
          if (saver && d.rank == 0) {
            P.loss[slide] = l;
            for (int j = 0; j < K; ++j) {
              if (P.dhaz_out) P.dhaz_out[slide * K + j] = DH[s * MAXK + j];
              if (P.dS_out) P.dS_out[slide * K + j] = DS[s * MAXK + j];
            }
          }
        }
      } else if (valid) {
        for (int j = 0; j < K; ++j) {
          if (P.dhaz_in) DH[s * MAXK + j] = P.dhaz_in[slide * K + j];
          if (P.dS_in) DS[s * MAXK + j] = P.dS_in[slide * K + j];
          if (P.dY_in) DYv[s * MAXK + j] = P.dY_in[slide * K + j];
        }
      }
      // survival head backward
      float dotY = 0.f;
      for (int j = 0; j < K; ++j) dotY = fmaf(DYv[s * MAXK + j], YV[s * MAXK + j], dotY);
      float tail = 0.f;
      for (int tt = K - 1; tt >= 0; --tt) {
        const float hz = HZ[s * MAXK + tt];
        tail = fmaf(DS[s * MAXK + tt], SV[s * MAXK + tt], tail);
        float dh = DH[s * MAXK + tt] - tail / (1.f - hz);
        float dl = dh * hz * (1.f - hz) + YV[s * MAXK + tt] * (DYv[s * MAXK + tt] - dotY);
        DLG[s * MAXK + tt] = dl;
        if (saver && d.rank == 0 && valid) ws[P.off_dlogits + slide * K + tt] = dl;
      }
    }
    __syncthreads();
    // ---- classifier data gradient with fusion layer 2's ReLU derivative (replicated; thread = feature)
#pragma unroll
    for (int s = 0; s < S; ++s) {
      float g = 0.f;
      for (int k = 0; k < K; ++k) g = fmaf(DLG[s * MAXK + k], __ldg(P.wcl + k * E + d.t), g);
      g = Z2[s * E + d.t] > 0.f ? g : 0.f;
      DZS[s * E + d.t] = g;
      if (saver && d.rank == 0 && d.s0 + s < d.B) ws[P.off_dz2 + static_cast<size_t>(d.s0 + s) * E + d.t] = g;
    }
    __syncthreads();
    for (int nb = 0; nb < NB; ++nb) {
      const int col = (d.rank * NB + nb) * 32 + d.lane;
      gemm_block<S, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), DZS, E);
      reduce_epi<S>(d, [&](int s, int, float v) {
        v = Z1[s * E + col] > 0.f ? v : 0.f;
        bcast(DZ1 + s * E + col, v);
        if (saver && d.s0 + s < d.B) ws[P.off_dz1 + static_cast<size_t>(d.s0 + s) * E + col] = v;
      });
    }
    cluster_sync();
    // fusion layer 0 data gradient: the gradient of [h_path | h_omic], taken through rho's dropout and ReLU
    for (int blk = 0; blk < 2 * NB; ++blk) {
      const int c2 = d.rank * (64 * NB) + blk * 32 + d.lane;     // column of the [., 512] concat
      const int pidx = c2 >> 8, c = c2 & 255;
      const DropSpec dr = site_of(dm, SITE_RHO + pidx);
      gemm_block<S, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), DZ1, E);
      reduce_epi<S>(d, [&](int s, int, float v) {
        const int slide = d.s0 + s;
        float hv = CAT[s * 2 * E + c2];
        if (dr.thr != 0) { v *= drop_grad(dr, d.seedv, static_cast<uint32_t>(slide) * E + c); hv = drop_invert(hv, dr); }
        v = hv > 0.f ? v : 0.f;
        bcast(DZR + s * 2 * E + c2, v);
        if (saver && slide < d.B) ws[P.poolw[pidx].dzr + static_cast<size_t>(slide) * E + c] = v;
      });
    }
    cluster_sync();
    // omic branch backward -> dG (completed by pre_bwd_kernel), then the path branch backward -> d(pooled)
#pragma unroll 1
    for (int br = br_hi; br >= br_lo; --br) {
      pool_bwd<S>(d, P.pool[br], P.poolw[br], br, dq, ws + P.encw[2 * br + 1].y2, XA, BIG, AL, BL, PA, AW, DHP, DZR);
#pragma unroll 1
      for (int l = 1; l >= 0; --l)
        enc_bwd<S>(d, P.enc[2 * br + l], P.encw[2 * br + l], 2 * br + l, dm, XA, XB, XC, BIG, AL, QKVL, BL,
                   l == 1 ? nullptr : ws + (br == 1 ? P.off_dG : P.off_dhc));
    }
    if (br_lo == 0) {            // the path branch goes on to d(pooled) (and through the CAG for NaCAGaT)
    for (int nb = 0; nb < NB; ++nb) {
      const int col = (d.rank * NB + nb) * 32 + d.lane;
      gemm_block<M, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XA, E);
      reduce_epi<M>(d, [&](int r, int, float v) {
        bcast(XB + r * E + col, v);
        if (d.grow0 + r < d.Rtot) ws[P.off_dv + static_cast<size_t>(d.grow0 + r) * E + col] = v;
      });
    }
    cluster_sync();
    if (P.dsuma != nullptr) {       // d(sum_n a'_in) = dv_i . b_v
      for (int r = d.warp; r < M; r += NW) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) v = fmaf(XB[r * E + d.lane + 32 * j], __ldg(P.bv + d.lane + 32 * j), v);
        v = warp_sum(v);
        if (d.lane == 0 && d.rank == 0 && d.grow0 + r < d.Rtot) P.dsuma[d.grow0 + r] = v;
      }
    }
    for (int nb = 0; nb < NB; ++nb) {
      const int col = (d.rank * NB + nb) * 32 + d.lane;
      gemm_block<M, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XB, E);
      reduce_epi<M>(d, [&](int r, int, float v) {
        if (d.grow0 + r < d.Rtot) P.dpooled[static_cast<size_t>(d.grow0 + r) * E + col] = v;
      });
    }
    if (P.nac) {
      // CAG backward (autograd of blocks.py:247-253): dC = d(hc) -> dQ accumulated into dG, dQ-hat written to dqp
      const CagP& cp = P.cag;
      const CagW& cw = P.cagw;
      cluster_sync();
      load_rows<M>(d, XA, ws + P.off_dhc, d.grow0, d.Rtot);
      __syncthreads();
      for (int i = d.t; i < M * E; i += NT) {          // t0 = dC * ELU'(C): gradient at fc_c's pre-activation
        const int r = i / E, c = i - r * E, grow = d.grow0 + r;
        const bool valid = grow < d.Rtot;
        const float Cv = valid ? __ldcg(ws + cw.C + static_cast<size_t>(grow) * E + c) : 0.f;
        const float t0 = XA[i] * elu_d(Cv);
        XB[i] = t0;
        if (valid && (r % CL) == d.rank) ws[cw.t0 + static_cast<size_t>(grow) * E + c] = t0;
      }
      for (int nb = 0; nb < NB; ++nb) {                // dm = t0 W_c ; dGg = dm * Ee, dEe = dm * Gg
        const int col = (d.rank * NB + nb) * 32 + d.lane;
        float ee[(M + NW - 1) / NW], gg[(M + NW - 1) / NW];
#pragma unroll
        for (int i = 0; i < (M + NW - 1) / NW; ++i) {
          const int r = d.warp + NW * i, grow = d.grow0 + r;
          const bool valid = r < M && grow < d.Rtot;
          ee[i] = valid ? __ldcg(ws + cw.Ee + static_cast<size_t>(grow) * E + col) : 0.f;
          gg[i] = valid ? __ldcg(ws + cw.Gg + static_cast<size_t>(grow) * E + col) : 0.f;
        }
        gemm_block<M, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XB, E);
        reduce_epi<M>(d, [&](int r, int i, float v) {
          bcast(BIG + r * E + col, v * ee[i]);
          bcast(BIG + M * E + r * E + col, v * gg[i]);
        });
      }
      cluster_sync();
      // du -> XB, dw -> XC (XA is scratch for the unused dropout output; df1 is rewritten below)
      ln_bwd_rows<M>(row_ctx(d), BIG, XB, XA, cp.gG, ws + cw.Gxh, ws + cw.Grs, ws + cw.dGg, ws + cw.df1, DropSpec{});
      ln_bwd_rows<M>(row_ctx(d), BIG + M * E, XC, XA, cp.gE, ws + cw.Exh, ws + cw.Ers, ws + cw.dEe, ws + cw.df1, DropSpec{});
      __syncthreads();
      for (int i = d.t; i < M * E; i += NT) {          // gradients at the pre-activations of fc1, fc2, fc3 (all columns)
        const int r = i / E, c = i - r * E, grow = d.grow0 + r;
        const bool valid = grow < d.Rtot;
        const size_t o = static_cast<size_t>(grow) * E + c;
        const float u = valid ? __ldcg(ws + cw.u + o) : 0.f, f1 = valid ? __ldcg(ws + cw.f1 + o) : 0.f;
        const float f2 = valid ? __ldcg(ws + cw.f2 + o) : 0.f, wv = valid ? __ldcg(ws + cw.w + o) : 0.f;
        const float f3 = valid ? __ldcg(ws + cw.f3 + o) : 0.f;
        const float dsum = XB[i] * elu_d(u);
        const float d1 = dsum * elu_d(f1), d2 = dsum * elu_d(f2), d3 = XC[i] * elu_d(wv) * elu_d(f3);
        BIG[r * FF + c] = d3;
        BIG[r * FF + E + c] = d2;
        XA[i] = d1;
        if (valid && (r % CL) == d.rank) { ws[cw.df1 + o] = d1; ws[cw.df2 + o] = d2; ws[cw.df3 + o] = d3; }
      }
      for (int nb = 0; nb < NB; ++nb) {                // dQ-hat = df3 W_3 + df2 W_2
        const int col = (d.rank * NB + nb) * 32 + d.lane;
        gemm_block<M, T_DGRAD, FF, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), BIG, FF);
        reduce_epi<M>(d, [&](int r, int, float v) {
          if (d.grow0 + r < d.Rtot) ws[P.off_dqp + static_cast<size_t>(d.grow0 + r) * E + col] = v;
        });
      }
      for (int nb = 0; nb < NB; ++nb) {                // dQ = df1 W_1 (pre_bwd_kernel adds it to the omic branch's dG)
        const int col = (d.rank * NB + nb) * 32 + d.lane;
        gemm_block<M, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XA, E);
        reduce_epi<M>(d, [&](int r, int, float v) {
          if (d.grow0 + r < d.Rtot) ws[P.off_dG2 + static_cast<size_t>(d.grow0 + r) * E + col] = v;
        });
      }
    }
    }
  }
  cp_async_wait<0>();
  cluster_sync();                         // no CTA exits while a peer may still store into its shared memory
#ifdef MPO_TAIL_PROF
  PROF_ADD(0, PROF_T() - pt_k0);
  if (PROF_ON) g_prof_pass = 2;           // the pre / SNN kernels in between land in a scratch bucket
#endif
}

// ------------------------------------------------------------------------------------------------ pre kernels
struct PreParams {
  Program prog;
  const float* omics[MPO_Q];
  int omic_dims[MPO_Q];
  const float* b1[MPO_Q];
  const float* b2[MPO_Q];
  const float* bq;
  float* ws;
  float *qp, *qk;
  const float* dqk;
  int off_snn_h[MPO_Q], off_G, off_dqp, off_dG, off_snn_dz1[MPO_Q], off_snn_dz2[MPO_Q];
  DropSpec d_alpha;
  int B;
  // NaCAGaT: key-bias score term kc = q . b_k / 16 (forward) and the extra gradient paths into q (backward)
  int nac;
  const float* bk;             // co_attention.in_proj_bias[256:512]
  float* kc;                   // [B][6] out of pre_kernel
  const float* dkc;            // [B][6]      from the bag backward
  const float* dtq;            // [B][6][256] gradient w.r.t. tanh(q) from the bag backward
  int off_dG2;                 // the CAG's dQ, left by the path kernel
  int skip_snn;                // the SNN encoders run batched over the slides in snn_fwd_kernel / snn_bwd_kernel
};
static_assert(sizeof(PreParams) <= 12000, "kernel parameter space (large kernel parameters, CUDA 12.1+)");

template <int S>
struct PreSmem {
  static constexpr int M = 6 * S;
  static constexpr int ring = 0;
  static constexpr int XO = ring + RingDepth<S>::n * CHUNK;       // [6][S][OMIC_LD] omic inputs
  static constexpr int H1 = XO + MPO_Q * S * OMIC_LD;    // [6][S][256]
  static constexpr int G = H1 + MPO_Q * S * E;           // [M][256]
  static constexpr int QP = G + M * E;                   // [M][256]
  static constexpr int RED = QP + M * E;
  static constexpr int TBL = RED + NW * M * 32;
  static constexpr int total = TBL + MAX_CHUNKS * 4;
};

// SNN encoders (mcat.py:32-45,90-92), query in-projection (rows 0..255 of co_attention.in_proj) and the key fold
// qk = W_k^T q / 16
template <int S>
__global__ void __launch_bounds__(NT, 1) pre_kernel(const __grid_constant__ PreParams P) {
  pdl_enter();
  constexpr int M = 6 * S;
  using L = PreSmem<S>;
  extern __shared__ __align__(16) float sm[];
  Dev d;
  d.rank = cluster_rank();
  d.t = threadIdx.x; d.lane = d.t & 31; d.warp = d.t >> 5;
  d.s0 = (blockIdx.x / CL) * S;
  d.B = P.B; d.grow0 = d.s0 * 6; d.Rtot = 6 * P.B;
  d.seedv = drop_seed(P.d_alpha);
  d.ws = P.ws;
  d.red = sm + L::RED;
  Pipe pipe;
  d.pipe = &pipe;
  pipe_init<RingDepth<S>::n>(pipe, P.prog, sm + L::TBL, sm + L::ring, d.rank);
  float* ws = P.ws;
  float *XO = sm + L::XO, *H1 = sm + L::H1, *G = sm + L::G, *QP = sm + L::QP;
  if (P.skip_snn) load_rows<M>(d, G, ws + P.off_G, d.grow0, d.Rtot);      // G_bag comes from snn_fwd_kernel
  for (int i = 0; i < (P.skip_snn ? 0 : MPO_Q); ++i) {
    const int dim = P.omic_dims[i];
    for (int j = d.t; j < S * dim; j += NT) {
      const int s = j / dim, c = j - s * dim;
      XO[(i * S + s) * OMIC_LD + c] = (d.s0 + s < d.B) ? __ldg(P.omics[i] + static_cast<size_t>(d.s0 + s) * dim + c) : 0.f;
    }
  }
  cluster_sync();
  for (int i = 0; i < (P.skip_snn ? 0 : MPO_Q); ++i)
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    const float bias = __ldg(P.b1[i] + col);
    const DropSpec ds = site_of(P.d_alpha, SITE_SNN + 2 * i);
    gemm_block<S, T_FWD, OMIC_LD, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XO + i * S * OMIC_LD, P.omic_dims[i]);
    reduce_epi<S>(d, [&](int s, int, float v) {
      v += bias;
      v = v > 0.f ? v : expm1f(v);
      const int slide = d.s0 + s;
      if (ds.thr != 0) v = drop_fwd(v, ds, d.seedv, static_cast<uint32_t>(slide) * E + col);
      bcast(H1 + (i * S + s) * E + col, v);
      if (slide < d.B) ws[P.off_snn_h[i] + static_cast<size_t>(slide) * E + col] = v;
    });
  }
  cluster_sync();
  for (int i = 0; i < (P.skip_snn ? 0 : MPO_Q); ++i)
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    const float bias = __ldg(P.b2[i] + col);
    const DropSpec ds = site_of(P.d_alpha, SITE_SNN + 2 * i + 1);
    gemm_block<S, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), H1 + i * S * E, E);
    reduce_epi<S>(d, [&](int s, int, float v) {
      v += bias;
      v = v > 0.f ? v : expm1f(v);
      const int slide = d.s0 + s;
      if (ds.thr != 0) v = drop_fwd(v, ds, d.seedv, static_cast<uint32_t>(slide) * E + col);
      bcast(G + (s * 6 + i) * E + col, v);
      if (slide < d.B) ws[P.off_G + static_cast<size_t>(slide * 6 + i) * E + col] = v;
    });
  }
  cluster_sync();
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    const float bias = __ldg(P.bq + col);
    gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), G, E);
    reduce_epi<M>(d, [&](int r, int, float v) {
      v += bias;
      bcast(QP + r * E + col, v);
      if (d.grow0 + r < d.Rtot) P.qp[static_cast<size_t>(d.grow0 + r) * E + col] = v;
    });
  }
  cluster_sync();
  if (P.nac && d.rank == 0) {
    for (int r = d.warp; r < M; r += NW) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) v = fmaf(QP[r * E + d.lane + 32 * j], __ldg(P.bk + d.lane + 32 * j), v);
      v = warp_sum(v);
      if (d.lane == 0 && d.grow0 + r < d.Rtot) P.kc[d.grow0 + r] = v * (1.f / 16.f);
    }
  }
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    gemm_block<M, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), QP, E);
    reduce_epi<M>(d, [&](int r, int, float v) {
      if (d.grow0 + r < d.Rtot) P.qk[static_cast<size_t>(d.grow0 + r) * E + col] = v * (1.f / 16.f);
    });
  }
  cp_async_wait<0>();
  cluster_sync();
}

// autograd of pre_kernel: fold, query projection (+ the omic branch's dG), SNN data gradients
template <int S>
__global__ void __launch_bounds__(NT, 1) pre_bwd_kernel(const __grid_constant__ PreParams P) {
  pdl_enter();
  constexpr int M = 6 * S;
  using L = PreSmem<S>;
  extern __shared__ __align__(16) float sm[];
  Dev d;
  d.rank = cluster_rank();
  d.t = threadIdx.x; d.lane = d.t & 31; d.warp = d.t >> 5;
  d.s0 = (blockIdx.x / CL) * S;
  d.B = P.B; d.grow0 = d.s0 * 6; d.Rtot = 6 * P.B;
  d.seedv = drop_seed(P.d_alpha);
  d.ws = P.ws;
  d.red = sm + L::RED;
  Pipe pipe;
  d.pipe = &pipe;
  pipe_init<RingDepth<S>::n>(pipe, P.prog, sm + L::TBL, sm + L::ring, d.rank);
  float* ws = P.ws;
  float *XA = sm + L::G, *XB = sm + L::QP, *DZ2 = sm + L::H1;     // DZ2: [M][256], row s*6+i
  load_rows<M>(d, XA, P.dqk, d.grow0, d.Rtot);
  cluster_sync();
  // dq[r][e] = sum_d dqk[r][d] W_k[e][d] / 16
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    // NaCAGaT: + the CAG's dQ-hat (left in dqp by the path kernel) + dkc b_k / 16 + dtq (1 - tanh(q)^2)
    float extra[(M + NW - 1) / NW];
#pragma unroll
    for (int i = 0; i < (M + NW - 1) / NW; ++i) {
      const int r = d.warp + NW * i, grow = d.grow0 + r;
      extra[i] = 0.f;
      if (P.nac && r < M && grow < d.Rtot) {
        const size_t o = static_cast<size_t>(grow) * E + col;
        const float tq = tanhf(__ldcg(P.qp + o));
        extra[i] = __ldcg(ws + P.off_dqp + o) + __ldcg(P.dkc + grow) * __ldg(P.bk + col) * (1.f / 16.f) +
                   __ldcg(P.dtq + o) * (1.f - tq * tq);
      }
    }
    gemm_block<M, T_FWD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XA, E);
    reduce_epi<M>(d, [&](int r, int i, float v) {
      v = fmaf(v, 1.f / 16.f, extra[i]);
      bcast(XB + r * E + col, v);
      if (d.grow0 + r < d.Rtot) ws[P.off_dqp + static_cast<size_t>(d.grow0 + r) * E + col] = v;
    });
  }
  cluster_sync();
  // dG = dq W_q + (omic branch), then through the second SNN layer's ELU + AlphaDropout
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    float dg0[(M + NW - 1) / NW], gv[(M + NW - 1) / NW];
#pragma unroll
    for (int i = 0; i < (M + NW - 1) / NW; ++i) {
      const int r = d.warp + NW * i, grow = d.grow0 + r;
      const bool valid = r < M && grow < d.Rtot;
      dg0[i] = valid ? __ldcg(ws + P.off_dG + static_cast<size_t>(grow) * E + col) : 0.f;
      if (valid && P.nac) dg0[i] += __ldcg(ws + P.off_dG2 + static_cast<size_t>(grow) * E + col);
      gv[i] = valid ? __ldcg(ws + P.off_G + static_cast<size_t>(grow) * E + col) : 0.f;
    }
    gemm_block<M, T_DGRAD, E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), XB, E);
    reduce_epi<M>(d, [&](int r, int i, float v) {
      v += dg0[i];
      if (P.skip_snn) {                 // total gradient of G_bag, consumed by snn_bwd_kernel
        if (d.grow0 + r < d.Rtot) ws[P.off_dG + static_cast<size_t>(d.grow0 + r) * E + col] = v;
        return;
      }
      const int s = r / 6, om = r - s * 6, slide = d.s0 + s;
      const DropSpec ds = site_of(P.d_alpha, SITE_SNN + 2 * om + 1);
      float y = gv[i];
      if (ds.thr != 0) { v *= drop_grad(ds, d.seedv, static_cast<uint32_t>(slide) * E + col); y = drop_invert(y, ds); }
      v *= (y > 0.f ? 1.f : y + 1.f);
      bcast(DZ2 + r * E + col, v);
      if (slide < d.B) ws[P.off_snn_dz2[om] + static_cast<size_t>(slide) * E + col] = v;
    });
  }
  cluster_sync();
  for (int i = 0; i < (P.skip_snn ? 0 : MPO_Q); ++i)
  for (int nb = 0; nb < NB; ++nb) {
    const int col = (d.rank * NB + nb) * 32 + d.lane;
    float hv[S];
#pragma unroll
    for (int s = 0; s < S; ++s)
      hv[s] = (d.s0 + s < d.B) ? __ldcg(ws + P.off_snn_h[i] + static_cast<size_t>(d.s0 + s) * E + col) : 0.f;
    const DropSpec ds = site_of(P.d_alpha, SITE_SNN + 2 * i);
    gemm_block<S, T_DGRAD, 6 * E, RingDepth<S>::n>(*d.pipe, smem_addr(d.red), DZ2 + i * E, E);
    reduce_epi<S>(d, [&](int s, int, float v) {
      const int slide = d.s0 + s;
      float y = hv[s];
      if (ds.thr != 0) { v *= drop_grad(ds, d.seedv, static_cast<uint32_t>(slide) * E + col); y = drop_invert(y, ds); }
      v *= (y > 0.f ? 1.f : y + 1.f);
      if (slide < d.B) ws[P.off_snn_dz1[i] + static_cast<size_t>(slide) * E + col] = v;
    });
  }
  cp_async_wait<0>();
  cluster_sync();
}

// ------------------------------------------------------------------------------------------------ SNN encoders
// The six SNN encoders (mcat.py:32-45,90-92) see ONE row per slide, so inside a per-slide cluster they are 62 tiny
// one-row GEMM chunks.  Batched over the slides instead: a cluster of 8 CTAs per omic group and 32 slides, CTA r owning
// the 32 output columns 32 r .. 32 r + 31 of both layers; the hidden layer is exchanged through distributed shared
// memory.  Same weight ring, FFMA2 blocks and dropout indices as the per-slide kernels.
constexpr int SNN_CL = 8, SNN_ROWS = 32;
struct SnnParams {
  Chunk c[MPO_Q][8];           // per omic group: layer 1 (ceil(d / KC) chunks) then layer 2 (forward); layer-2 dgrad (backward)
  int n[MPO_Q];
  const float* omics[MPO_Q];
  int omic_dims[MPO_Q];
  const float* b1[MPO_Q];
  const float* b2[MPO_Q];
  float* ws;
  int off_snn_h[MPO_Q], off_G, off_dG, off_snn_dz1[MPO_Q], off_snn_dz2[MPO_Q];
  DropSpec d_alpha;
  int B;
};
struct SnnSmem {
  static constexpr int ring = 0;
  static constexpr int XO = ring + NSTAGE * CHUNK;              // [32][OMIC_LD] inputs (backward: [32][256] dz2)
  static constexpr int H1 = XO + SNN_ROWS * OMIC_LD;            // [32][256] hidden layer
  static constexpr int RED = H1 + SNN_ROWS * E;
  static constexpr int TBL = RED + NW * SNN_ROWS * 32;
  static constexpr int total = TBL + 8 * 4 + 8;
};
static_assert(SnnSmem::total * sizeof(float) <= 227 * 1024, "SNN kernel shared memory");

__global__ void __launch_bounds__(NT, 1) snn_fwd_kernel(const __grid_constant__ SnnParams P) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  Dev d;
  d.rank = cluster_rank();
  d.t = threadIdx.x; d.lane = d.t & 31; d.warp = d.t >> 5;
  const int cid = blockIdx.x / SNN_CL, om = cid % MPO_Q;
  d.s0 = (cid / MPO_Q) * SNN_ROWS;
  d.B = P.B; d.grow0 = 0; d.Rtot = 0;
  d.seedv = drop_seed(P.d_alpha);
  d.ws = P.ws;
  d.red = sm + SnnSmem::RED;
  Pipe pipe;
  d.pipe = &pipe;
  pipe_init(pipe, P.c[om], P.n[om], sm + SnnSmem::TBL, sm + SnnSmem::ring, d.rank);
  float* ws = P.ws;
  float *XO = sm + SnnSmem::XO, *H1 = sm + SnnSmem::H1;
  const int dim = P.omic_dims[om];
  for (int j = d.t; j < SNN_ROWS * dim; j += NT) {
    const int s = j / dim, c = j - s * dim;
    XO[s * OMIC_LD + c] = (d.s0 + s < d.B) ? __ldg(P.omics[om] + static_cast<size_t>(d.s0 + s) * dim + c) : 0.f;
  }
  cluster_sync();
  const int col = d.rank * 32 + d.lane;
  {
    const float bias = __ldg(P.b1[om] + col);
    const DropSpec ds = site_of(P.d_alpha, SITE_SNN + 2 * om);
    gemm_block<SNN_ROWS, T_FWD, OMIC_LD>(*d.pipe, smem_addr(d.red), XO, dim);
    reduce_epi<SNN_ROWS>(d, [&](int s, int, float v) {
      v = elu_f(v + bias);
      const int slide = d.s0 + s;
      if (ds.thr != 0) v = drop_fwd(v, ds, d.seedv, static_cast<uint32_t>(slide) * E + col);
      bcast_n<SNN_CL>(H1 + s * E + col, v);
      if (slide < d.B) ws[P.off_snn_h[om] + static_cast<size_t>(slide) * E + col] = v;
    });
  }
  cluster_sync();
  {
    const float bias = __ldg(P.b2[om] + col);
    const DropSpec ds = site_of(P.d_alpha, SITE_SNN + 2 * om + 1);
    gemm_block<SNN_ROWS, T_FWD, E>(*d.pipe, smem_addr(d.red), H1, E);
    reduce_epi<SNN_ROWS>(d, [&](int s, int, float v) {
      v = elu_f(v + bias);
      const int slide = d.s0 + s;
      if (ds.thr != 0) v = drop_fwd(v, ds, d.seedv, static_cast<uint32_t>(slide) * E + col);
      if (slide < d.B) ws[P.off_G + static_cast<size_t>(slide * 6 + om) * E + col] = v;
    });
  }
  cp_async_wait<0>();
  cluster_sync();
}

// dG (total gradient of G_bag, from pre_bwd_kernel) -> gradients at the pre-activations of both SNN layers
__global__ void __launch_bounds__(NT, 1) snn_bwd_kernel(const __grid_constant__ SnnParams P) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  Dev d;
  d.rank = cluster_rank();
  d.t = threadIdx.x; d.lane = d.t & 31; d.warp = d.t >> 5;
  const int cid = blockIdx.x / SNN_CL, om = cid % MPO_Q;
  d.s0 = (cid / MPO_Q) * SNN_ROWS;
  d.B = P.B; d.grow0 = 0; d.Rtot = 0;
  d.seedv = drop_seed(P.d_alpha);
  d.ws = P.ws;
  d.red = sm + SnnSmem::RED;
  Pipe pipe;
  d.pipe = &pipe;
  pipe_init(pipe, P.c[om], P.n[om], sm + SnnSmem::TBL, sm + SnnSmem::ring, d.rank);
  float* ws = P.ws;
  float* DZ = sm + SnnSmem::XO;                     // [32][256], every CTA computes all columns (elementwise)
  const int col = d.rank * 32 + d.lane;
  {
    const DropSpec ds = site_of(P.d_alpha, SITE_SNN + 2 * om + 1);
    for (int idx = d.t; idx < SNN_ROWS * E; idx += NT) {
      const int s = idx / E, c = idx - s * E, slide = d.s0 + s;
      const bool valid = slide < d.B;
      const size_t o = static_cast<size_t>(slide * 6 + om) * E + c;
      float g = valid ? __ldcg(ws + P.off_dG + o) : 0.f;
      float y = valid ? __ldcg(ws + P.off_G + o) : 0.f;
      if (ds.thr != 0) { g *= drop_grad(ds, d.seedv, static_cast<uint32_t>(slide) * E + c); y = drop_invert(y, ds); }
      g *= elu_d(y);
      DZ[idx] = g;
      if (valid && (c >> 5) == d.rank) ws[P.off_snn_dz2[om] + static_cast<size_t>(slide) * E + c] = g;
    }
  }
  {
    float hv[SNN_ROWS / NW];
#pragma unroll
    for (int i = 0; i < SNN_ROWS / NW; ++i) {
      const int slide = d.s0 + d.warp + NW * i;
      hv[i] = slide < d.B ? __ldcg(ws + P.off_snn_h[om] + static_cast<size_t>(slide) * E + col) : 0.f;
    }
    const DropSpec ds = site_of(P.d_alpha, SITE_SNN + 2 * om);
    gemm_block<SNN_ROWS, T_DGRAD, E>(*d.pipe, smem_addr(d.red), DZ, E);
    reduce_epi<SNN_ROWS>(d, [&](int s, int i, float v) {
      const int slide = d.s0 + s;
      float y = hv[i];
      if (ds.thr != 0) { v *= drop_grad(ds, d.seedv, static_cast<uint32_t>(slide) * E + col); y = drop_invert(y, ds); }
      v *= elu_d(y);
      if (slide < d.B) ws[P.off_snn_dz1[om] + static_cast<size_t>(slide) * E + col] = v;
    });
  }
  cp_async_wait<0>();
  cluster_sync();
}

// ------------------------------------------------------------------------------------------------ SNN encoders, wide form
// Round 2.  The cluster form above runs 48 CTAs through a 3-deep ring of 16 KB weight chunks and three cluster barriers:
// 35 + 28 us per step for 3.6 MB of weights (latency-bound, 5 % of the issue slots).  Here one layer of the six encoders is
// ONE plain launch of 16 x 6 CTAs per 32 slides: a CTA owns 16 output columns of one omic group, fetches its whole weight
// slice and the 32 input rows with one burst of 16-byte cp.async (everything in flight at once, one wait), splits the
// reduction over its 8 warps (lane = slide row, float4 along the reduction, weights as broadcast LDS.128) and finishes
// with a cross-warp sum and the fused epilogue (bias, ELU, AlphaDropout with the same (site, element) indices as the
// cluster form, so both forms produce identical masks).  Forward = two launches, backward = one.
constexpr int SNN2_COLS = 16;
constexpr int SNN2_RP = 20;                 // padded pitch of one row's 16 partial sums (conflict-free float4 stores)
struct Snn2Params {
  const float* x[MPO_Q];                    // forward: input rows [B][ldx].  backward: unused
  const float* w[MPO_Q];                    // forward: layer weight [256][K].  backward: layer-2 weight [256][256]
  const float* bias[MPO_Q];
  int K[MPO_Q], ldx[MPO_Q];
  float* ws;
  int off_out[MPO_Q];                       // forward: first output element of the group; backward: snn_dz1
  int out_ld;                               // floats between two slides of the output (E, or 6 E for G_bag)
  int off_snn_h[MPO_Q], off_snn_dz2[MPO_Q], off_G, off_dG;
  int site_add;                             // 0: layer 1, 1: layer 2
  DropSpec d_alpha;
  int B;
};
__host__ __device__ constexpr int snn2_kp(int K) { return ((K + 31) / 32) * 32 + 4; }       // row pitch = 4 (mod 32) floats
constexpr int SNN2_SMEM_FLOATS = (SNN_ROWS + SNN2_COLS) * snn2_kp(OMIC_LD) + NW * SNN_ROWS * SNN2_RP;
static_assert(SNN2_SMEM_FLOATS * sizeof(float) <= 227 * 1024, "SNN (wide form) shared memory");

// acc[j] += sum over this warp's share of the reduction of X[lane][r] * W[j][r]
__device__ __forceinline__ void snn2_dot(const float* Xs, const float* Wsm, int kp, int K, int warp, int lane, float (&acc)[SNN2_COLS]) {
  const int groups = K >> 2;                               // float4 groups of the reduction
  const float4* xr = reinterpret_cast<const float4*>(Xs + lane * kp);
  for (int g = warp; g < groups; g += NW) {
    const float4 xv = xr[g];
#pragma unroll
    for (int j = 0; j < SNN2_COLS; ++j) {
      const float4 wv = *reinterpret_cast<const float4*>(Wsm + j * kp + 4 * g);
      acc[j] = fmaf(xv.x, wv.x, acc[j]); acc[j] = fmaf(xv.y, wv.y, acc[j]);
      acc[j] = fmaf(xv.z, wv.z, acc[j]); acc[j] = fmaf(xv.w, wv.w, acc[j]);
    }
  }
}
__device__ __forceinline__ void snn2_store_partials(float* red, int warp, int lane, const float (&acc)[SNN2_COLS]) {
  float4* dst = reinterpret_cast<float4*>(red + (warp * SNN_ROWS + lane) * SNN2_RP);
#pragma unroll
  for (int q = 0; q < SNN2_COLS / 4; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
}
__device__ __forceinline__ float snn2_sum_partials(const float* red, int s, int j) {
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) v += red[(w * SNN_ROWS + s) * SNN2_RP + j];
  return v;
}

__global__ void __launch_bounds__(NT, 1) snn2_fwd_kernel(const __grid_constant__ Snn2Params P) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int om = blockIdx.y, col0 = blockIdx.x * SNN2_COLS, s0 = blockIdx.z * SNN_ROWS;
  const int K = P.K[om], kp = snn2_kp(K), ldx = P.ldx[om];
  float* Xs = sm;                              // [32][kp]
  float* Wsm = Xs + SNN_ROWS * kp;             // [16][kp]
  float* red = Wsm + SNN2_COLS * kp;           // [8][32][SNN2_RP]
  const int g4 = K >> 2;
  const float* xg = P.x[om];
  const float* wg = P.w[om] + static_cast<size_t>(col0) * K;
  for (int i = t; i < SNN_ROWS * g4; i += NT) {
    const int s = i / g4, g = i - s * g4;
    if (s0 + s < P.B) cp_async16(Xs + s * kp + 4 * g, xg + static_cast<size_t>(s0 + s) * ldx + 4 * g);
    else *reinterpret_cast<float4*>(Xs + s * kp + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = t; i < SNN2_COLS * g4; i += NT) {
    const int j = i / g4, g = i - j * g4;
    cp_async16(Wsm + j * kp + 4 * g, wg + static_cast<size_t>(j) * K + 4 * g);
  }
  cp_async_commit();
  const uint32_t seedv = drop_seed(P.d_alpha);
  const DropSpec ds = site_of(P.d_alpha, SITE_SNN + 2 * om + P.site_add);
  cp_async_wait<0>();
  __syncthreads();
  float acc[SNN2_COLS];
#pragma unroll
  for (int j = 0; j < SNN2_COLS; ++j) acc[j] = 0.f;
  snn2_dot(Xs, Wsm, kp, K, warp, lane, acc);
  snn2_store_partials(red, warp, lane, acc);
  __syncthreads();
  float* out = P.ws + P.off_out[om];
#pragma unroll
  for (int i = 0; i < SNN_ROWS * SNN2_COLS / NT; ++i) {
    const int idx = t + i * NT, s = idx / SNN2_COLS, j = idx % SNN2_COLS, col = col0 + j, slide = s0 + s;
    if (slide >= P.B) continue;
    float v = elu_f(snn2_sum_partials(red, s, j) + __ldg(P.bias[om] + col));
    if (ds.thr != 0) v = drop_fwd(v, ds, seedv, static_cast<uint32_t>(slide) * E + col);
    out[static_cast<size_t>(slide) * P.out_ld + col] = v;
  }
}

// dG -> dz2 (stored) -> dh1 = dz2 W2 -> dz1 (stored); a CTA owns 16 columns of dz2 (store) and of dz1
__global__ void __launch_bounds__(NT, 1) snn2_bwd_kernel(const __grid_constant__ Snn2Params P) {
  pdl_enter();
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int om = blockIdx.y, col0 = blockIdx.x * SNN2_COLS, s0 = blockIdx.z * SNN_ROWS;
  constexpr int kp = snn2_kp(E);
  float* Xs = sm;                              // [32][kp]  dG, then dz2
  float* Wsm = Xs + SNN_ROWS * kp;             // [16][kp]  Wsm[j][c] = W2[c][col0 + j]
  float* red = Wsm + SNN2_COLS * kp;           // [8][32][SNN2_RP]; first holds the raw G_bag rows [32][kp]
  float* Gs = red;
  static_assert(NW * SNN_ROWS * SNN2_RP <= SNN_ROWS * snn2_kp(E), "the partial sums alias the staged G rows");
  static_assert((2 * SNN_ROWS + SNN2_COLS) * snn2_kp(E) <= SNN2_SMEM_FLOATS, "SNN backward shared memory");
  float* ws = P.ws;
  constexpr int g4 = E / 4;
  for (int i = t; i < SNN_ROWS * g4; i += NT) {
    const int s = i / g4, g = i - s * g4, slide = s0 + s;
    if (slide < P.B) {
      const size_t o = static_cast<size_t>(slide * 6 + om) * E + 4 * g;
      cp_async16(Xs + s * kp + 4 * g, ws + P.off_dG + o);
      cp_async16(Gs + s * kp + 4 * g, ws + P.off_G + o);
    } else {
      *reinterpret_cast<float4*>(Xs + s * kp + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(Gs + s * kp + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  cp_async_commit();
  // transposed slice of W2: 64-byte row pieces, scattered into the [j][c] layout
  const float* w2 = P.w[om];
  float wreg[SNN2_COLS * E / NT];
#pragma unroll
  for (int i = 0; i < SNN2_COLS * E / NT; ++i) {
    const int idx = t + i * NT, c = idx / SNN2_COLS, j = idx % SNN2_COLS;
    wreg[i] = __ldg(w2 + static_cast<size_t>(c) * E + col0 + j);
  }
  // hidden-layer outputs of this CTA's 16 columns (for the ELU / dropout derivative of dz1)
  float hv[SNN_ROWS * SNN2_COLS / NT];
#pragma unroll
  for (int i = 0; i < SNN_ROWS * SNN2_COLS / NT; ++i) {
    const int idx = t + i * NT, s = idx / SNN2_COLS, j = idx % SNN2_COLS, slide = s0 + s;
    hv[i] = slide < P.B ? __ldcg(ws + P.off_snn_h[om] + static_cast<size_t>(slide) * E + col0 + j) : 0.f;
  }
  const uint32_t seedv = drop_seed(P.d_alpha);
  const DropSpec ds2 = site_of(P.d_alpha, SITE_SNN + 2 * om + 1), ds1 = site_of(P.d_alpha, SITE_SNN + 2 * om);
#pragma unroll
  for (int i = 0; i < SNN2_COLS * E / NT; ++i) {
    const int idx = t + i * NT, c = idx / SNN2_COLS, j = idx % SNN2_COLS;
    Wsm[j * kp + c] = wreg[i];
  }
  cp_async_wait<0>();
  __syncthreads();
  for (int idx = t; idx < SNN_ROWS * E; idx += NT) {
    const int s = idx / E, c = idx - s * E, slide = s0 + s;
    float g = Xs[s * kp + c], y = Gs[s * kp + c];
    if (ds2.thr != 0) { g *= drop_grad(ds2, seedv, static_cast<uint32_t>(slide) * E + c); y = drop_invert(y, ds2); }
    g *= elu_d(y);
    if (slide >= P.B) g = 0.f;
    Xs[s * kp + c] = g;
    if (slide < P.B && (c / SNN2_COLS) == static_cast<int>(blockIdx.x))
      ws[P.off_snn_dz2[om] + static_cast<size_t>(slide) * E + c] = g;
  }
  __syncthreads();                             // dz2 complete, G rows no longer needed (red aliases them)
  float acc[SNN2_COLS];
#pragma unroll
  for (int j = 0; j < SNN2_COLS; ++j) acc[j] = 0.f;
  snn2_dot(Xs, Wsm, kp, E, warp, lane, acc);
  __syncthreads();
  snn2_store_partials(red, warp, lane, acc);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SNN_ROWS * SNN2_COLS / NT; ++i) {
    const int idx = t + i * NT, s = idx / SNN2_COLS, j = idx % SNN2_COLS, col = col0 + j, slide = s0 + s;
    if (slide >= P.B) continue;
    float v = snn2_sum_partials(red, s, j);
    float y = hv[i];
    if (ds1.thr != 0) { v *= drop_grad(ds1, seedv, static_cast<uint32_t>(slide) * E + col); y = drop_invert(y, ds1); }
    v *= elu_d(y);
    ws[P.off_out[om] + static_cast<size_t>(slide) * E + col] = v;
  }
}

// ------------------------------------------------------------------------------------------------ weight gradients
// kind 0: gw[o][i] += alpha * sum_r dz[r][o] x[r][i]  and  gb[o] += sum_r dz[r][o]     (64 x 64 tiles)
// kind 1: LayerNorm: gw[c] += sum_r dz[r][c] x[r][c]  and  gb[c] += sum_r dz[r][c]     (64-column tiles)
struct WJob {
  const float* dz;
  const float* x;
  float* gw;
  float* gb;
  const float* rw;       // optional per-row weights of the bias gradient (gb[o] += sum_r rw[r] dz[r][o]); null = 1
  int lddz, ldx, out, in, rows, tile0, kind;
  float alpha;
};
constexpr int MAX_JOBS = 60;
struct WParams { WJob job[MAX_JOBS]; int njobs; int ntiles; };
static_assert(sizeof(WParams) <= 8000, "kernel parameter space (large kernel parameters, CUDA 12.1+)");

__global__ void __launch_bounds__(256) wgrad_kernel(const __grid_constant__ WParams P) {
  pdl_enter();
  __shared__ __align__(16) float As[2][32][64 + 4];
  __shared__ __align__(16) float Bs[2][32][64 + 4];
  const int t = threadIdx.x;
  int j = 0;
  while (j + 1 < P.njobs && static_cast<int>(blockIdx.x) >= P.job[j + 1].tile0) ++j;
  const WJob& J = P.job[j];
  const int tile = blockIdx.x - J.tile0;
  if (J.kind == 1) {
    const int c = tile * 64 + (t & 63), part = t >> 6;       // 4 row groups
    float s1 = 0.f, s2 = 0.f;
    if (c < J.out)
      for (int r = part; r < J.rows; r += 4) {
        const float g = __ldcg(J.dz + static_cast<size_t>(r) * J.lddz + c);
        s1 = fmaf(g, __ldcg(J.x + static_cast<size_t>(r) * J.ldx + c), s1);
        s2 += g;
      }
    As[0][part][t & 63] = s1;
    Bs[0][part][t & 63] = s2;
    __syncthreads();
    if (part == 0 && c < J.out) {
      J.gw[c] += As[0][0][t] + As[0][1][t] + As[0][2][t] + As[0][3][t];
      J.gb[c] += Bs[0][0][t] + Bs[0][1][t] + Bs[0][2][t] + Bs[0][3][t];
    }
    return;
  }
  const int tiles_in = (J.in + 63) / 64;
  const int o0 = (tile / tiles_in) * 64, i0 = (tile % tiles_in) * 64;
  const int tx = t & 15, ty = t >> 4;            // outputs (o0 + ty*4 .. +3, i0 + tx*4 .. +3)
  const int lc = t & 63, lr = t >> 6;            // staging: column lc, rows lr + 4 q
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float bsum = 0.f;                              // bias gradient of column o0 + lc (threads with lr == 0 finish it)
  float ra[8], rb[8];
  auto load = [&](int r0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int r = r0 + lr + 4 * q;
      ra[q] = (r < J.rows && o0 + lc < J.out) ? __ldcg(J.dz + static_cast<size_t>(r) * J.lddz + o0 + lc) : 0.f;
      rb[q] = (r < J.rows && i0 + lc < J.in) ? __ldcg(J.x + static_cast<size_t>(r) * J.ldx + i0 + lc) : 0.f;
    }
  };
  auto store = [&](int buf) {
#pragma unroll
    for (int q = 0; q < 8; ++q) { As[buf][lr + 4 * q][lc] = ra[q]; Bs[buf][lr + 4 * q][lc] = rb[q]; }
  };
  const int nsteps = (J.rows + 31) / 32;
  load(0);
  store(0);
  __syncthreads();
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    if (s + 1 < nsteps) load((s + 1) * 32);
#pragma unroll
    for (int kk = 0; kk < 32; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) acc[aa][bb] = fmaf(av[aa], bv[bb], acc[aa][bb]);
    }
    if (J.gb != nullptr && i0 == 0 && t < 64) {
      if (J.rw == nullptr) {
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) bsum += As[buf][kk][t];
      } else {
        for (int kk = 0; kk < 32 && s * 32 + kk < J.rows; ++kk) bsum = fmaf(__ldcg(J.rw + s * 32 + kk), As[buf][kk][t], bsum);
      }
    }
    if (s + 1 < nsteps) store(buf ^ 1);
    __syncthreads();
  }
  if (J.gb != nullptr && i0 == 0 && t < 64 && o0 + t < J.out) J.gb[o0 + t] += bsum;
#pragma unroll
  for (int aa = 0; aa < 4; ++aa) {
    const int o = o0 + ty * 4 + aa;
    if (o >= J.out) continue;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int i = i0 + tx * 4 + bb;
      if (i < J.in) J.gw[static_cast<size_t>(o) * J.in + i] += J.alpha * acc[aa][bb];
    }
  }
}

struct JobBuilder {
  WParams& p;
  bool ok = true;
  void push(int kind, const float* dz, int lddz, const float* x, int ldx, float* gw, float* gb, int out, int in, int rows,
            float alpha, const float* rw = nullptr) {
    if (gw == nullptr) return;
    if (p.njobs >= MAX_JOBS) { ok = false; return; }
    WJob& j = p.job[p.njobs++];
    j.dz = dz; j.x = x; j.gw = gw; j.gb = gb; j.lddz = lddz; j.ldx = ldx; j.out = out; j.in = in; j.rows = rows;
    j.kind = kind; j.alpha = alpha; j.tile0 = p.ntiles; j.rw = rw;
    p.ntiles += kind == 1 ? (out + 63) / 64 : ((out + 63) / 64) * ((in + 63) / 64);
  }
  void lin(const float* dz, int lddz, const float* x, int ldx, const mpo_lin& L, int out, int in, int rows, float alpha = 1.f) {
    push(0, dz, lddz, x, ldx, L.gw, L.gb, out, in, rows, alpha);
  }
  void norm(const float* dy, const float* xh, const mpo_norm& N, int rows) { push(1, dy, E, xh, E, N.gg, N.gb, E, E, rows, 1.f); }
};

// ------------------------------------------------------------------------------------------------ host side
DropSpec host_drop(const mpo_tail_io* io, float p, bool alpha) {
  DropSpec d = {};
  if (io->train == 0 || p <= 0.f) return d;            // eval, or a site whose rate is 0
  d.thr = static_cast<uint32_t>(p * 256.f + 0.5f);
  if (d.thr == 0) return d;
  const float pe = static_cast<float>(d.thr) / 256.f;
  d.seed = io->seed; d.seed_dev = io->seed_dev; d.alpha = alpha ? 1 : 0;
  if (alpha) {
    d.scale = 1.f / sqrtf((1.f - pe) * (1.f + pe * kAlphaPrime * kAlphaPrime));
    d.shift = -d.scale * kAlphaPrime * pe;
  } else {
    d.scale = 1.f / (1.f - pe);
    d.shift = 0.f;
  }
  return d;
}

template <typename K, typename PT>
cudaError_t launch_cluster(K kern, const PT& prm, int nclusters, size_t smem_bytes, cudaStream_t st, int cluster_size = CL) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(nclusters * cluster_size));
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = step_pdl_attr(at, 1, st);
  e = cudaLaunchKernelEx(&cfg, kern, prm);
  count_launch();
  return e;
}

// Slides per cluster.  One slide (6 token rows) per cluster of 4 CTAs makes the 32 slides of a step 32 clusters =
// 128 CTAs, a single wave (33 such clusters fit); larger batches take two slides per cluster to halve the waves.
int slides_per_cluster(int B) {
  const char* env = getenv("MPO_TAIL_FUSED_S");
  const int forced = env ? atoi(env) : 0;
  if (forced == 1 || forced == 2) return forced;
  return B <= 35 ? 1 : 2;        // 2 x 35 one-slide clusters (path + omic role) are resident at once (71 fit)
}

bool eligible(const mpo_model* m, const mpo_tail_io* io) {
  const char* env = getenv("MPO_TAIL_FUSED");        // read on every call: tests flip it to compare the two tails
  if (env != nullptr && atoi(env) == 0) return false;
  if (m->fusion != MPO_FUSION_CONCAT) return false;                      // bilinear fusion: per-op tail
  if (m->variant != MPO_VARIANT_MCAT && m->variant != MPO_VARIANT_NACAGAT) return false;
  if (m->n_classes > MAXK) return false;
  if (m->variant == MPO_VARIANT_NACAGAT) {
    const char* e2 = getenv("MPO_TAIL_FUSED_NACAGAT");
    if (e2 != nullptr && atoi(e2) == 0) return false;
  }
  for (int i = 0; i < MPO_Q; ++i)
    if (m->omic_dims[i] % 4 != 0 || m->omic_dims[i] > OMIC_LD || m->omic_dims[i] < 4) return false;
  return true;
}

static int fin(cudaError_t e, const char* where) {
  if (e == cudaSuccess) e = cudaGetLastError();
  return check_cuda(e, where);
}

static void fill_pre(const mpo_model* m, const mpo_tail_io* io, const Ws& w, PreParams& P) {
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < MPO_Q; ++i) {
    P.omics[i] = io->omics[i]; P.omic_dims[i] = m->omic_dims[i];
    P.b1[i] = m->snn[i][0].b; P.b2[i] = m->snn[i][1].b;
    P.off_snn_h[i] = static_cast<int>(w.snn_h[i]);
    P.off_snn_dz1[i] = static_cast<int>(w.snn_dz1[i]); P.off_snn_dz2[i] = static_cast<int>(w.snn_dz2[i]);
  }
  P.bq = m->coattn_in.b;
  P.ws = io->ws; P.qp = io->qp; P.qk = io->qk; P.dqk = io->dqk;
  P.off_G = static_cast<int>(w.G); P.off_dqp = static_cast<int>(w.dqp); P.off_dG = static_cast<int>(w.dG);
  P.d_alpha = host_drop(io, io->drop_p, true);
  P.B = io->num_slides;
  P.nac = m->variant == MPO_VARIANT_NACAGAT ? 1 : 0;
  P.bk = m->coattn_in.b + E;
  P.kc = io->kc; P.dkc = io->dkc; P.dtq = io->dtq;
  P.off_dG2 = static_cast<int>(w.fz_dG2);
}

static bool snn_batched() {
  const char* env = getenv("MPO_TAIL_SNN_BATCH");
  return !(env != nullptr && atoi(env) == 0);
}
static void fill_snn(const mpo_model* m, const mpo_tail_io* io, const Ws& w, SnnParams& P, bool backward) {
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < MPO_Q; ++i) {
    P.omics[i] = io->omics[i]; P.omic_dims[i] = m->omic_dims[i];
    P.b1[i] = m->snn[i][0].b; P.b2[i] = m->snn[i][1].b;
    P.off_snn_h[i] = static_cast<int>(w.snn_h[i]);
    P.off_snn_dz1[i] = static_cast<int>(w.snn_dz1[i]); P.off_snn_dz2[i] = static_cast<int>(w.snn_dz2[i]);
    // one 32-column block per CTA of the 8-CTA cluster
    Program tmp; tmp.n = 0; tmp.tma = 0;
    ProgBuilder pb{tmp};
    if (!backward) { pb.fwd(m->snn[i][0].w, m->omic_dims[i], m->omic_dims[i]); pb.fwd(m->snn[i][1].w, E, E); }
    else pb.dgrad(m->snn[i][1].w, E, E);
    P.n[i] = tmp.n < 8 ? tmp.n : 8;
    for (int k = 0; k < P.n[i]; ++k) P.c[i][k] = tmp.c[k];
  }
  P.ws = io->ws;
  P.off_G = static_cast<int>(w.G); P.off_dG = static_cast<int>(w.dG);
  P.d_alpha = host_drop(io, io->drop_p, true);
  P.B = io->num_slides;
}

// wide form of the SNN kernels (default); MPO_TAIL_SNN_FORM=cluster keeps the 8-CTA cluster kernels
static bool snn_wide() {
  const char* env = getenv("MPO_TAIL_SNN_FORM");
  return !(env != nullptr && strcmp(env, "cluster") == 0);
}
static void fill_snn2(const mpo_model* m, const mpo_tail_io* io, const Ws& w, Snn2Params& P, int layer, bool backward) {
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < MPO_Q; ++i) {
    P.off_snn_h[i] = static_cast<int>(w.snn_h[i]);
    P.off_snn_dz2[i] = static_cast<int>(w.snn_dz2[i]);
    if (backward) {
      P.w[i] = m->snn[i][1].w; P.K[i] = E; P.ldx[i] = E;
      P.off_out[i] = static_cast<int>(w.snn_dz1[i]);
    } else if (layer == 0) {
      P.x[i] = io->omics[i]; P.K[i] = m->omic_dims[i]; P.ldx[i] = m->omic_dims[i];
      P.w[i] = m->snn[i][0].w; P.bias[i] = m->snn[i][0].b;
      P.off_out[i] = static_cast<int>(w.snn_h[i]);
    } else {
      P.x[i] = io->ws + w.snn_h[i]; P.K[i] = E; P.ldx[i] = E;
      P.w[i] = m->snn[i][1].w; P.bias[i] = m->snn[i][1].b;
      P.off_out[i] = static_cast<int>(w.G) + i * E;
    }
  }
  P.out_ld = (!backward && layer == 1) ? 6 * E : E;
  P.site_add = layer;
  P.ws = io->ws;
  P.off_G = static_cast<int>(w.G); P.off_dG = static_cast<int>(w.dG);
  P.d_alpha = host_drop(io, io->drop_p, true);
  P.B = io->num_slides;
}
template <typename K>
static cudaError_t launch_snn2(K kern, const Snn2Params& P, int maxK, cudaStream_t st, bool backward = false) {
  // backward: dG rows + W2 slice + G rows (the partial sums alias the G rows)
  const size_t smem = backward ? static_cast<size_t>(2 * SNN_ROWS + SNN2_COLS) * snn2_kp(E) * sizeof(float)
                               : (static_cast<size_t>(SNN_ROWS + SNN2_COLS) * snn2_kp(maxK) + NW * SNN_ROWS * SNN2_RP) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(SNN2_SMEM_FLOATS * sizeof(float)));
  if (e != cudaSuccess) return e;
  const dim3 grid(E / SNN2_COLS, MPO_Q, (P.B + SNN_ROWS - 1) / SNN_ROWS);
  e = launch_step(kern, grid, dim3(NT), smem, st, P);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

int pre_fwd(const mpo_model* m, const mpo_tail_io* io, const Ws& w, cudaStream_t st) {
  static PreParams P;
  fill_pre(m, io, w, P);
  if (P.nac && !io->kc) return fail(MPO_E_ARG, "%s", "mpo_tail_pre_fwd: kc is NULL (NaCAGaT)");
  const bool batched = snn_batched();
  if (batched && snn_wide()) {
    static Snn2Params S2;
    int maxK = 0;
    for (int i = 0; i < MPO_Q; ++i) maxK = m->omic_dims[i] > maxK ? m->omic_dims[i] : maxK;
    fill_snn2(m, io, w, S2, 0, false);
    int rc0 = fin(launch_snn2(snn2_fwd_kernel, S2, maxK, st), "snn2_fwd_kernel layer 1 (fused tail)");
    if (rc0) return rc0;
    fill_snn2(m, io, w, S2, 1, false);
    rc0 = fin(launch_snn2(snn2_fwd_kernel, S2, E, st), "snn2_fwd_kernel layer 2 (fused tail)");
    if (rc0) return rc0;
  } else if (batched) {
    static SnnParams SP;
    fill_snn(m, io, w, SP, false);
    const int ncl_snn = MPO_Q * ((io->num_slides + SNN_ROWS - 1) / SNN_ROWS);
    cudaError_t e0 = launch_cluster(snn_fwd_kernel, SP, ncl_snn, SnnSmem::total * sizeof(float), st, SNN_CL);
    const int rc0 = fin(e0, "snn_fwd_kernel (fused tail)");
    if (rc0) return rc0;
  }
  P.skip_snn = batched ? 1 : 0;
  P.prog.tma = 0;                  // (the pre / SNN kernels keep the cp.async ring)
  ProgBuilder pb{P.prog};
  if (!batched) {
    for (int i = 0; i < MPO_Q; ++i) pb.fwd(m->snn[i][0].w, m->omic_dims[i], m->omic_dims[i], NB, 32 * NB);
    for (int i = 0; i < MPO_Q; ++i) pb.fwd(m->snn[i][1].w, E, E, NB, 32 * NB);
  }
  pb.fwd(m->coattn_in.w, E, E, NB, 32 * NB);
  pb.dgrad(m->coattn_in.w + static_cast<size_t>(E) * E, E, E, NB, 32 * NB);
  if (!pb.ok) return fail(MPO_E_CUDA, "%s", "fused tail: program table overflow");
  const int B = io->num_slides, S = slides_per_cluster(B), ncl = (B + S - 1) / S;
  cudaError_t e = S == 2 ? launch_cluster(pre_kernel<2>, P, ncl, PreSmem<2>::total * sizeof(float), st)
                         : launch_cluster(pre_kernel<1>, P, ncl, PreSmem<1>::total * sizeof(float), st);
  return fin(e, "pre_kernel (fused tail)");
}

// Side stream for the post stage's weight gradients inside a training step: they depend only on the path kernel and
// are first needed by the optimizer, so they run next to the bag backward pass instead of in front of it.  The fork
// and the join (at the end of pre_bwd) are event edges, so the step stays capturable into one CUDA graph.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int state = 0;          // 0 not created, 1 ready, -1 unavailable
  bool pending = false;   // a weight-gradient launch on `s` has not been joined yet
};
static SideStream& side_stream() {
  static SideStream pool[32];
  int dev = 0;
  cudaGetDevice(&dev);
  SideStream& p = pool[dev & 31];
  if (p.state == 0) {
    const char* env = getenv("MPO_TAIL_STREAMS");
    bool ok = !(env && atoi(env) == 0);
    ok = ok && cudaStreamCreateWithFlags(&p.s, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p.ev_fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p.ev_join, cudaEventDisableTiming) == cudaSuccess;
    p.state = ok ? 1 : -1;
  }
  return p;
}

static int launch_wgrad(const WParams& W, cudaStream_t st) {
  if (W.ntiles == 0) return MPO_OK;
  const cudaError_t e = launch_step(wgrad_kernel, dim3(W.ntiles), dim3(256), 0, st, W);
  count_launch();
  return fin(e, "wgrad_kernel (fused tail)");
}

// Adam over a gradient bucket that the pending side-stream weight gradients complete: queued behind them (next to the bag
// backward pass); the join event moves behind the update, so pre_bwd's join covers it
int side_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
              float eps, float weight_decay, int32_t* step_dev, bool zero_grad, cudaStream_t st) {
  SideStream& side = side_stream();
  if (!(side.state == 1 && side.pending))
    return launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_dev, zero_grad, false, st);
  int rc = launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_dev, zero_grad, false, side.s);
  if (rc) return rc;
  return check_cuda(cudaEventRecord(side.ev_join, side.s), "fused tail: weight-gradient stream event (after Adam)");
}

int pre_bwd(const mpo_model* m, const mpo_tail_io* io, const Ws& w, cudaStream_t st) {
  static PreParams P;
  fill_pre(m, io, w, P);
  if (P.nac && (!io->dkc || !io->dtq)) return fail(MPO_E_ARG, "%s", "mpo_tail_pre_bwd: dkc/dtq are NULL (NaCAGaT)");
  P.prog.tma = 0;                  // (the pre / SNN kernels keep the cp.async ring)
  ProgBuilder pb{P.prog};
  pb.fwd(m->coattn_in.w + static_cast<size_t>(E) * E, E, E, NB, 32 * NB);
  pb.dgrad(m->coattn_in.w, E, E, NB, 32 * NB);
  const bool batched = snn_batched();
  P.skip_snn = batched ? 1 : 0;
  if (!batched)
    for (int i = 0; i < MPO_Q; ++i) pb.dgrad(m->snn[i][1].w, E, E, NB, 32 * NB);
  if (!pb.ok) return fail(MPO_E_CUDA, "%s", "fused tail: program table overflow");
  const int B = io->num_slides, R = 6 * B, S = slides_per_cluster(B), ncl = (B + S - 1) / S;
  cudaError_t e = S == 2 ? launch_cluster(pre_bwd_kernel<2>, P, ncl, PreSmem<2>::total * sizeof(float), st)
                         : launch_cluster(pre_bwd_kernel<1>, P, ncl, PreSmem<1>::total * sizeof(float), st);
  int rc = fin(e, "pre_bwd_kernel (fused tail)");
  if (rc) return rc;
  if (batched && snn_wide()) {
    static Snn2Params S2;
    fill_snn2(m, io, w, S2, 0, true);
    rc = fin(launch_snn2(snn2_bwd_kernel, S2, E, st, true), "snn2_bwd_kernel (fused tail)");
    if (rc) return rc;
  } else if (batched) {
    static SnnParams SP;
    fill_snn(m, io, w, SP, true);
    const int ncl_snn = MPO_Q * ((io->num_slides + SNN_ROWS - 1) / SNN_ROWS);
    cudaError_t e1 = launch_cluster(snn_bwd_kernel, SP, ncl_snn, SnnSmem::total * sizeof(float), st, SNN_CL);
    rc = fin(e1, "snn_bwd_kernel (fused tail)");
    if (rc) return rc;
  }
  static WParams W;
  memset(&W, 0, sizeof(W));
  JobBuilder jb{W};
  float* ws = io->ws;
  for (int i = 0; i < MPO_Q; ++i) {
    jb.lin(ws + w.snn_dz1[i], E, io->omics[i], m->omic_dims[i], m->snn[i][0], E, m->omic_dims[i], B);
    jb.lin(ws + w.snn_dz2[i], E, ws + w.snn_h[i], E, m->snn[i][1], E, E, B);
  }
  // query block of co_attention.in_proj, and the key block through the fold: dW_k[e][d] += sum_r q[r][e] dqk[r][d] / 16
  mpo_lin Lq = m->coattn_in;
  jb.lin(ws + w.dqp, E, ws + w.G, E, Lq, E, E, R);
  mpo_lin Lk = {nullptr, nullptr, m->coattn_in.gw ? m->coattn_in.gw + static_cast<size_t>(E) * E : nullptr, nullptr};
  jb.lin(io->qp, E, io->dqk, E, Lk, E, E, R, 1.f / 16.f);
  if (P.nac && m->coattn_in.gb != nullptr)      // kc = q . b_k / 16 :  db_k[e] += sum_r dkc[r] q[r][e] / 16
    jb.push(0, io->dkc, 1, io->qp, E, m->coattn_in.gb + E, nullptr, 1, E, R, 1.f / 16.f);
  if (!jb.ok) return fail(MPO_E_CUDA, "%s", "fused tail: job table overflow");
  rc = launch_wgrad(W, st);
  if (rc) return rc;
  SideStream& side = side_stream();
  if (side.pending) {          // the post stage's weight gradients of this step (launched by post() on the side stream)
    side.pending = false;
    pdl_bar_next(st);          // the next kernel on st waits for the side stream: launched fully serialized
    return check_cuda(cudaStreamWaitEvent(st, side.ev_join, 0), "fused tail: join of the weight-gradient stream");
  }
  return MPO_OK;
}

int post(const mpo_model* m, const mpo_tail_io* io, const Ws& w, int flags, const LossArgs* loss, const float* dhaz,
         const float* dS, const float* dY, cudaStream_t st, bool side_wgrad) {
  static PathParams P;
  memset(&P, 0, sizeof(P));
  const int B = io->num_slides, R = 6 * B, K = m->n_classes;
  const mpo_encoder_layer* enc[4] = {&m->path_tr[0], &m->path_tr[1], &m->omic_tr[0], &m->omic_tr[1]};
  const long long enc_x[4] = {w.hc, w.enc[0].y2, w.G, w.enc[2].y2};
  for (int e = 0; e < 4; ++e) {
    const mpo_encoder_layer& L = *enc[e];
    P.enc[e] = EncP{L.in_proj.w, L.in_proj.b, L.out_proj.w, L.out_proj.b, L.linear1.w, L.linear1.b, L.linear2.w,
                    L.linear2.b, L.norm1.g, L.norm1.b, L.norm2.g, L.norm2.b};
    const tailws::EncBuf& b = w.enc[e];
    const Ws::FzEnc& f = w.fz_enc[e];
    P.encw[e] = EncW{(int)enc_x[e], (int)b.qkv, (int)b.probs, (int)b.ctx, (int)b.y1, (int)b.xh1, (int)b.rs1, (int)b.f,
                     (int)b.y2, (int)b.xh2, (int)b.rs2, (int)f.dy2, (int)f.df2, (int)f.df, (int)f.dy1, (int)f.dsa,
                     (int)f.dqkv};
  }
  const mpo_pool_head* pool[2] = {&m->path_pool, &m->omic_pool};
  for (int p = 0; p < 2; ++p) {
    const mpo_pool_head& H = *pool[p];
    P.pool[p] = PoolP{H.att_a.w, H.att_a.b, H.att_b.w, H.att_b.b, H.att_c.w, H.att_c.b, H.rho.w, H.rho.b,
                      (flags & F_BWD) ? H.att_c.gw : nullptr, (flags & F_BWD) ? H.att_c.gb : nullptr};
    P.poolw[p] = PoolW{(int)w.pool[p].a, (int)w.pool[p].b, (int)w.pool[p].w, (int)w.pool[p].hp, (int)w.fz_pool[p].dzr,
                       (int)w.fz_pool[p].da, (int)w.fz_pool[p].db};
  }
  const float* Wv = m->coattn_in.w + static_cast<size_t>(2 * E) * E;
  P.wv = Wv; P.bv = m->coattn_in.b + 2 * E;
  P.wo = m->coattn_out.w; P.bo = m->coattn_out.b;
  P.wf0 = m->fusion0.w; P.bf0 = m->fusion0.b; P.wf2 = m->fusion2.w; P.bf2 = m->fusion2.b;
  P.wcl = m->classifier.w; P.bcl = m->classifier.b;
  P.ws = io->ws; P.pooled = io->pooled; P.dpooled = io->dpooled;
  P.hazards = io->hazards; P.S = io->S; P.Y = io->Y; P.att_path = io->att_path; P.att_omic = io->att_omic;
  P.off_G = (int)w.G; P.off_v = (int)w.v; P.off_hc = (int)w.hc; P.off_cat = (int)w.cat; P.off_z1 = (int)w.z1;
  P.off_z2 = (int)w.z2; P.off_logits = (int)w.logits; P.off_dlogits = (int)w.dlogits; P.off_dz1 = (int)w.dz1;
  P.off_dz2 = (int)w.dz2; P.off_dhc = (int)w.dhc; P.off_dv = (int)w.dv; P.off_dG = (int)w.dG;
  if (flags & F_LOSS) {
    P.loss_kind = loss->kind; P.loss_alpha = loss->alpha; P.loss_eps = loss->eps; P.grad_scale = loss->grad_scale;
    P.label = loss->label; P.censor = loss->censor; P.loss = loss->loss; P.dhaz_out = loss->dhaz; P.dS_out = loss->dS;
  }
  P.dhaz_in = dhaz; P.dS_in = dS; P.dY_in = dY;
  P.d_model = host_drop(io, io->drop_p, false);
  P.d_quarter = host_drop(io, 0.25f, false);        // AttentionNetGated hard-codes p = 0.25 (blocks.py:34-36)
  P.B = B; P.K = K; P.flags = flags;
  const bool nac = m->variant == MPO_VARIANT_NACAGAT;
  P.nac = nac ? 1 : 0;
  if (nac) {
    if (!io->qp) return fail(MPO_E_ARG, "%s", "fused tail: qp is NULL (NaCAGaT)");
    if ((flags & F_BWD) && io->suma != nullptr && !io->dsuma) return fail(MPO_E_ARG, "%s", "fused tail: dsuma is NULL although suma is given");
    const mpo_cag& C = m->cag;
    P.cag = CagP{C.fc1.w, C.fc1.b, C.fc2.w, C.fc2.b, C.fc3.w, C.fc3.b, C.fc_c.w, C.fc_c.b, C.G.g, C.G.b, C.E.g, C.E.b};
    P.cagw = CagW{(int)w.cag_f1, (int)w.cag_f2, (int)w.cag_f3, (int)w.cag_u, (int)w.cag_w, (int)w.cag_Gg, (int)w.cag_Gxh,
                  (int)w.cag_Grs, (int)w.cag_Ee, (int)w.cag_Exh, (int)w.cag_Ers, (int)w.cag_m, (int)w.cag_C,
                  (int)w.fz_cag[0], (int)w.fz_cag[1], (int)w.fz_cag[2], (int)w.fz_cag[3], (int)w.fz_cag[4], (int)w.fz_cag[5]};
    P.qp = io->qp;
    P.suma = io->suma;
    P.dsuma = io->suma != nullptr ? io->dsuma : nullptr;
  }
  P.off_dqp = (int)w.dqp;

  // chunk order = consumption order of the device code; a CTA owns NB column blocks of every 256-wide layer
  auto build_program = [&](Program& prog, int lflags, int br_hi, int br_lo, CUtensorMap* maps) -> bool {
    prog.n = 0;
    prog.tma = maps != nullptr ? 1 : 0;
    ProgBuilder pb{prog};
    pb.maps = maps;
    auto enc_f = [&](const mpo_encoder_layer& L) {
      for (int nb = 0; nb < NB; ++nb) pb.fwd(L.in_proj.w, E, E, 3, 32 * NB, E, nb * 32);     // q, k, v of head rank * NB + nb
      pb.fwd(L.out_proj.w, E, E, NB, 32 * NB);
      pb.fwd(L.linear1.w, E, E, 2 * NB, 64 * NB, 32);
      pb.fwd(L.linear2.w, FF, FF, NB, 32 * NB);
    };
    auto enc_b = [&](const mpo_encoder_layer& L) {
      pb.dgrad(L.linear2.w, FF, E, 2 * NB, 64 * NB, 32);
      pb.dgrad(L.linear1.w, E, FF, NB, 32 * NB);
      pb.dgrad(L.out_proj.w, E, E, NB, 32 * NB);
      pb.dgrad(L.in_proj.w, E, 3 * E, NB, 32 * NB);
    };
    auto pool_f = [&](const mpo_pool_head& H) {
      for (int nb = 0; nb < NB; ++nb) {
        pb.fwd(H.att_a.w, E, E, 1, 32 * NB, 32, nb * 32);
        pb.fwd(H.att_b.w, E, E, 1, 32 * NB, 32, nb * 32);
      }
      pb.fwd(H.rho.w, E, E, NB, 32 * NB);
    };
    auto pool_b = [&](const mpo_pool_head& H) {
      pb.dgrad(H.rho.w, E, E, NB, 32 * NB);
      for (int nb = 0; nb < NB; ++nb) {
        pb.dgrad(H.att_a.w, E, E, 1, 32 * NB, 32, nb * 32);
        pb.dgrad(H.att_b.w, E, E, 1, 32 * NB, 32, nb * 32);
      }
    };
    if (lflags & F_FWD) {
      for (int br = br_hi; br >= br_lo; --br) {
        if (br == 1) {
          enc_f(m->omic_tr[0]); enc_f(m->omic_tr[1]); pool_f(m->omic_pool);
        } else {
          pb.fwd(Wv, E, E, NB, 32 * NB); pb.fwd(m->coattn_out.w, E, E, NB, 32 * NB);
          if (nac) {
            for (int nb = 0; nb < NB; ++nb) {
              pb.fwd(m->cag.fc1.w, E, E, 1, 32 * NB, 32, nb * 32);
              pb.fwd(m->cag.fc2.w, E, E, 1, 32 * NB, 32, nb * 32);
              pb.fwd(m->cag.fc3.w, E, E, 1, 32 * NB, 32, nb * 32);
            }
            pb.fwd(m->cag.fc_c.w, E, E, NB, 32 * NB);
          }
          enc_f(m->path_tr[0]); enc_f(m->path_tr[1]); pool_f(m->path_pool);
        }
      }
    }
    if (lflags & F_HEAD) { pb.fwd(m->fusion0.w, 2 * E, 2 * E, NB, 32 * NB); pb.fwd(m->fusion2.w, E, E, NB, 32 * NB); }
    if (lflags & F_BWD) {
      pb.dgrad(m->fusion2.w, E, E, NB, 32 * NB);
      pb.dgrad(m->fusion0.w, 2 * E, E, 2 * NB, 64 * NB, 32);
      for (int br = br_hi; br >= br_lo; --br) {
        if (br == 1) { pool_b(m->omic_pool); enc_b(m->omic_tr[1]); enc_b(m->omic_tr[0]); }
        else { pool_b(m->path_pool); enc_b(m->path_tr[1]); enc_b(m->path_tr[0]); }
      }
      if (br_lo == 0) {
        pb.dgrad(m->coattn_out.w, E, E, NB, 32 * NB); pb.dgrad(Wv, E, E, NB, 32 * NB);
        if (nac) {
          pb.dgrad(m->cag.fc_c.w, E, E, NB, 32 * NB);
          for (int nb = 0; nb < NB; ++nb) {
            pb.dgrad(m->cag.fc3.w, E, E, 1, 32 * NB, 32, nb * 32);
            pb.dgrad(m->cag.fc2.w, E, E, 1, 32 * NB, 32, nb * 32);
          }
          pb.dgrad(m->cag.fc1.w, E, E, NB, 32 * NB);
        }
      }
    }
    return pb.ok;
  };
  // Launch plan.  The omic and the path branch are independent, so two clusters per slide group (roles) run them
  // concurrently -- at 113 KB of shared memory two CTAs share an SM and hide each other's barrier latencies.  The
  // branches meet in the fusion layer: a kernel boundary, after which BOTH roles redo the (tiny) fusion + head + loss
  // + fusion backward and continue with their own branch's backward pass.
  const int S = slides_per_cluster(B), ncl = (B + S - 1) / S;
  P.off_dG2 = static_cast<int>(w.fz_dG2);
  int plan_flags[2] = {0, 0}, plan_roles[2] = {1, 1}, nlaunch = 1;
  if (flags & F_FWD) {
    plan_flags[0] = F_FWD; plan_roles[0] = 2;
    plan_flags[1] = F_HEAD | (flags & (F_LOSS | F_BWD)); plan_roles[1] = (flags & F_BWD) ? 2 : 1;
    nlaunch = 2;
  } else {
    plan_flags[0] = flags; plan_roles[0] = 2;
  }
  int rc = MPO_OK;
  // weight ring through 2-D tensor copies (default) or per-thread cp.async (MPO_TAIL_TMA=0, and whenever a program has a
  // chunk the TMA form does not cover)
  const char* tma_env = getenv("MPO_TAIL_TMA");       // read on every call: tests flip it to compare the two forms
  const int use_tma = (tma_env == nullptr || atoi(tma_env) != 0) ? 1 : 0;
  for (int l = 0; l < nlaunch; ++l) {
    P.flags = plan_flags[l];
    P.nroles = plan_roles[l];
    bool ok = false;
    for (int form = use_tma ? 1 : 0; form >= 0 && !ok; --form) {
      CUtensorMap* m0 = form ? &P.maps[0][0] : nullptr;
      CUtensorMap* m1 = form ? &P.maps[1][0] : nullptr;
      if (P.nroles == 2) { ok = build_program(P.prog[0], P.flags, 0, 0, m0) && build_program(P.prog[1], P.flags, 1, 1, m1); }
      else { ok = build_program(P.prog[0], P.flags, 1, 0, m0); P.prog[1].n = 0; P.prog[1].tma = 0; }
      if (form == 1 && !ok) {
        static bool told = false;
        if (!told && getenv("MPO_TAIL_VERBOSE") != nullptr) {
          told = true;
          fprintf(stderr, "[mpo] fused tail: a weight stream does not fit the TMA form, cp.async ring used (%s)\n", mpo_last_error());
        }
      }
    }
    if (!ok) return fail(MPO_E_CUDA, "%s", "fused tail: program table overflow");
    cudaError_t e = S == 2 ? launch_cluster(path_kernel<2>, P, ncl * P.nroles, PathSmem<2>::total * sizeof(float), st)
                           : launch_cluster(path_kernel<1>, P, ncl * P.nroles, PathSmem<1>::total * sizeof(float), st);
    rc = fin(e, "path_kernel (fused tail)");
    if (rc) return rc;
  }
  if (!(flags & F_BWD)) return rc;

  static WParams W;
  memset(&W, 0, sizeof(W));
  JobBuilder jb{W};
  float* ws = io->ws;
  for (int e4 = 0; e4 < 4; ++e4) {
    const mpo_encoder_layer& L = *enc[e4];
    const tailws::EncBuf& b = w.enc[e4];
    const Ws::FzEnc& f = w.fz_enc[e4];
    jb.lin(ws + f.dqkv, 3 * E, ws + enc_x[e4], E, L.in_proj, 3 * E, E, R);
    jb.lin(ws + f.dsa, E, ws + b.ctx, E, L.out_proj, E, E, R);
    jb.lin(ws + f.df, FF, ws + b.y1, E, L.linear1, FF, E, R);
    jb.lin(ws + f.df2, E, ws + b.f, FF, L.linear2, E, FF, R);
    jb.norm(ws + f.dy1, ws + b.xh1, L.norm1, R);
    jb.norm(ws + f.dy2, ws + b.xh2, L.norm2, R);
  }
  const long long tok[2] = {w.enc[1].y2, w.enc[3].y2};
  for (int p = 0; p < 2; ++p) {
    const mpo_pool_head& H = *pool[p];
    jb.lin(ws + w.fz_pool[p].da, E, ws + tok[p], E, H.att_a, E, E, R);
    jb.lin(ws + w.fz_pool[p].db, E, ws + tok[p], E, H.att_b, E, E, R);
    jb.lin(ws + w.fz_pool[p].dzr, E, ws + w.pool[p].hp, E, H.rho, E, E, B);
  }
  jb.lin(ws + w.dz1, E, ws + w.cat, 2 * E, m->fusion0, E, 2 * E, B);
  jb.lin(ws + w.dz2, E, ws + w.z1, E, m->fusion2, E, E, B);
  jb.lin(ws + w.dlogits, K, ws + w.z2, E, m->classifier, K, E, B);
  jb.lin(ws + w.dhc, E, ws + w.v, E, m->coattn_out, E, E, R);
  mpo_lin Lv = {nullptr, nullptr, m->coattn_in.gw ? m->coattn_in.gw + static_cast<size_t>(2 * E) * E : nullptr,
                m->coattn_in.gb ? m->coattn_in.gb + 2 * E : nullptr};
  // with attention dropout the value bias enters as (sum_n a'_in) b_v: its gradient is weighted by suma
  jb.push(0, ws + w.dv, E, io->pooled, E, Lv.gw, Lv.gb, E, E, R, 1.f, nac ? io->suma : nullptr);
  if (nac) {
    const mpo_cag& C = m->cag;
    jb.lin(ws + w.fz_cag[0], E, ws + w.cag_m, E, C.fc_c, E, E, R);
    jb.lin(ws + w.fz_cag[5], E, io->qp, E, C.fc3, E, E, R);
    jb.lin(ws + w.fz_cag[4], E, io->qp, E, C.fc2, E, E, R);
    jb.lin(ws + w.fz_cag[3], E, ws + w.G, E, C.fc1, E, E, R);
    jb.norm(ws + w.fz_cag[1], ws + w.cag_Gxh, C.G, R);
    jb.norm(ws + w.fz_cag[2], ws + w.cag_Exh, C.E, R);
  }
  if (!jb.ok) return fail(MPO_E_CUDA, "%s", "fused tail: job table overflow");
  SideStream& side = side_stream();
  if (side_wgrad && side.state == 1) {
    cudaError_t e2 = cudaEventRecord(side.ev_fork, st);
    if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(side.s, side.ev_fork, 0);
    pdl_bar_next(side.s);      // first kernel behind the fork: fully serialized
    if (e2 != cudaSuccess) return check_cuda(e2, "fused tail: fork of the weight-gradient stream");
    rc = launch_wgrad(W, side.s);
    if (rc) return rc;
    side.pending = true;
    return check_cuda(cudaEventRecord(side.ev_join, side.s), "fused tail: weight-gradient stream event");
  }
  return launch_wgrad(W, st);
}

}  // namespace fused
}  // namespace mpo

#ifdef MPO_TAIL_PROF
extern "C" int mpo_tail_prof_read(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, mpo::fused::g_prof, 32 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long z[32] = {};
    int two = 2;
    cudaMemcpyToSymbol(mpo::fused::g_prof, z, sizeof(z));
    cudaMemcpyToSymbol(mpo::fused::g_prof_pass, &two, sizeof(int));
  }
  return 0;
}
#endif
