// Orchestration of the slide tail (declared in include/mpo_b200.h): a fixed sequence of the small kernels in
// tail_kernels.cuh, batched over the B slides of a step, all on the caller's stream.  Mirrors, stage by stage,
// what the reference modules compute on the 6 omic tokens (file:line given at each stage).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/mpo_b200.h"
#include "launchers.h"
#include "tail_kernels.cuh"
#include "tail_ws.h"
#include "tail_fused.h"

namespace mpo {
using namespace tailws;
namespace {

// ---------------------------------------------------------------------------------------------- op helpers
// Auxiliary streams: the path and omic branches of the tail are independent, and weight-gradient GEMMs are off the
// critical path of the backward pass.  Everything forks from / joins back into the caller's stream through
// events, so the whole tail is still "stream-ordered on `stream`" for the caller and is CUDA-graph capturable.
struct StreamPool {
  cudaStream_t aux[7] = {};   // 0: second branch, 1: wgrads of main, 2: wgrads of second, 3..6: SNN chains 2..5
  cudaEvent_t ev[128] = {};
  int dummy_ = 0;
  int next = 0;
  int state = 0;   // 0 = not created, 1 = ready, -1 = disabled
};
StreamPool& stream_pool() {
  static StreamPool pools[32];
  int dev = 0;
  cudaGetDevice(&dev);
  StreamPool& p = pools[dev & 31];
  if (p.state == 0) {
    const char* env = getenv("MPO_TAIL_STREAMS");
    if (env && atoi(env) == 0) { p.state = -1; return p; }
    bool ok = true;
    for (auto& s : p.aux) ok = ok && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& e : p.ev) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    p.state = ok ? 1 : -1;
  }
  return p;
}

struct Ctx {
  cudaStream_t st;             // compute stream of this branch
  cudaStream_t wst = nullptr;  // weight-gradient stream of this branch, valid when async_w
  bool async_w = false;        // (a null stream handle is the legacy default stream, so it cannot double as "none")
  int sb = 0;                  // scratch set of this branch
  bool train = false;          // train mode: dropout sites draw masks (the rate of a site may still be 0)
  float drop_p = 0.f;          // the model's `dropout` rate; seed of the mask streams
  uint32_t seed = 0;
  const uint32_t* seed_dev = nullptr;
  // GE-NaCAGaT: scratch for the bf16 (hi, lo) operand pairs of the tensor-core weight-gradient GEMMs over tc_rows tokens
  float* tc_dz = nullptr;      // [2][tc_rows][768] bf16
  float* tc_x = nullptr;       // [2][tc_rows][512] bf16
  int tc_rows = 0;
  cudaError_t err = cudaSuccess;
  const char* where = "";
  void chk(cudaError_t e, const char* w) { if (err == cudaSuccess && e != cudaSuccess) { err = e; where = w; } }
};
// `to` waits for everything submitted to `from` so far
void dep(Ctx& c, cudaStream_t from, cudaStream_t to) {
  if (from == to) return;
  StreamPool& p = stream_pool();
  cudaEvent_t e = p.ev[p.next];
  p.next = (p.next + 1) & 127;
  c.chk(cudaEventRecord(e, from), "event record");
  c.chk(cudaStreamWaitEvent(to, e, 0), "stream wait");
  pdl_bar_next(to);
}
// the compute stream waits for this branch's pending weight-gradient GEMMs (before their inputs are overwritten)
void join_w(Ctx& c) { if (c.async_w) dep(c, c.wst, c.st); }

DropSpec no_drop() { DropSpec d = {}; return d; }
DropSpec mk_drop(const Ctx& c, float p, uint32_t site, bool alpha = false) {
  DropSpec d = {};
  if (!c.train || p <= 0.f) return d;              // eval, or a site whose rate is 0 (a model built with dropout = 0
                                                   // still drops at the hard-coded p = 0.25 sites, as in the reference)
  d.thr = static_cast<uint32_t>(p * 256.f + 0.5f);
  if (d.thr == 0) return d;
  const float pe = static_cast<float>(d.thr) / 256.f;     // the probability actually drawn
  d.site = site; d.seed = c.seed; d.seed_dev = c.seed_dev; d.alpha = alpha ? 1 : 0;
  if (alpha) {        // nn.AlphaDropout: a = ((1-p)(1 + p alpha'^2))^-1/2, b = -a alpha' p
    d.scale = 1.f / sqrtf((1.f - pe) * (1.f + pe * kAlphaPrime * kAlphaPrime));
    d.shift = -d.scale * kAlphaPrime * pe;
  } else {
    d.scale = 1.f / (1.f - pe);
    d.shift = 0.f;
  }
  return d;
}

inline unsigned nblk(long long n, int t = 256) { return static_cast<unsigned>((n + t - 1) / t); }
// g[c] += column sums of x (optionally of x * y); bag-long reductions are split over row blocks
void colsum(cudaStream_t st, const float* x, long long ldx, const float* y, long long ldy, float* g, int rows, int cols) {
  int ry = rows / 512;
  if (ry < 1) ry = 1;
  if (ry > 64) ry = 64;
  launch_k(colsum_kernel, dim3(nblk(cols, 32), ry), dim3(256), 0, st, x, ldx, y, ldy, g, rows, cols); count_launch();
}


mpo_lin sub(const mpo_lin& L, int row0, int in) {   // rows [row0, ...) of a packed projection
  mpo_lin s;
  s.w = L.w + (long long)row0 * in;
  s.b = L.b ? L.b + row0 : nullptr;
  s.gw = L.gw ? L.gw + (long long)row0 * in : nullptr;
  s.gb = L.gb ? L.gb + row0 : nullptr;
  return s;
}

// y[rows,out] = act(x[rows,in] W^T + b)
void lin_fwd(Ctx& c, const float* x, long long ldx, const mpo_lin& L, int out, int in, float* y, long long ldy, int rows,
             int act, const DropSpec drop = DropSpec{}) {
  GemmArgs g{x, ldx, 1, L.w, 1, in, y, ldy, L.b, rows, out, in, 1.f, 0, act, nullptr, 1, drop};
  c.chk(launch_gemm(g, c.st), "lin_fwd");
}
// dz [rows,out] is the gradient at the pre-activation.  dx (=|+=) dz W ; gw += dz^T x ; gb += colsum(dz)
// optional epilogues of the data gradient: + addend (residual-branch gradient), * act'(y) of the layer in front (with
// that layer's dropout undone), so that dx leaves the kernel as the gradient at that layer's pre-activation
struct DgradEpi {
  const float* addend = nullptr; long long ld_add = 0;
  const float* y = nullptr; long long ld_y = 0; int act = ACT_NONE; DropSpec drop = DropSpec{};
};
void lin_bwd(Ctx& c, const float* dz, long long lddz, const float* x, long long ldx, const mpo_lin& L, int out, int in,
             float* dx, long long lddx, int rows, bool acc_dx, const DgradEpi epi = DgradEpi{}) {
  if (dx != nullptr) {
    GemmArgs g{dz, lddz, 1, L.w, in, 1, dx, lddx, nullptr, rows, in, out, 1.f, acc_dx ? 1 : 0, ACT_NONE, nullptr, 1};
    g.addend = epi.addend; g.ld_add = epi.ld_add;
    g.bwd_y = epi.y; g.ld_bwd = epi.ld_y; g.bwd_act = epi.act; g.bwd_drop = epi.drop;
    c.chk(launch_gemm(g, c.st), "lin_bwd.dgrad");
  }
  if (L.gw != nullptr && c.tc_dz != nullptr && rows == c.tc_rows && out <= 768 && in <= 512) {
    // bag-long reduction (GE-NaCAGaT's N-token layers): gw += dz^T x on tcgen05 -- both operands read MN-major out of
    // their bf16 (hi, lo) pairs, split-K over the tokens, partial products added atomically; gb by the column-sum kernel
    const int po = (out + 63) / 64 * 64, pi = (in + 63) / 64 * 64;
    __nv_bfloat16* dzh = reinterpret_cast<__nv_bfloat16*>(c.tc_dz);
    __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(c.tc_x);
    __nv_bfloat16* dzl = dzh + static_cast<long long>(rows) * po;
    __nv_bfloat16* xl = xh + static_cast<long long>(rows) * pi;
    c.chk(launch_split_bf16(dz, lddz, rows, out, dzh, dzl, po, c.st), "lin_bwd.split dz");
    c.chk(launch_split_bf16(x, ldx, rows, in, xh, xl, pi, c.st), "lin_bwd.split x");
    if (c.err == cudaSuccess &&
        launch_tc_gemm(dzh, dzl, rows, po, true, xh, xl, rows, pi, true, L.gw, in, out, in, rows, 1.f, true, c.st) != 0)
      c.chk(cudaErrorUnknown, "lin_bwd.wgrad (tc)");
    if (L.gb != nullptr) colsum(c.st, dz, lddz, nullptr, 0, L.gb, rows, out);
  } else if (L.gw != nullptr) {     // gw += dz^T x, with gb += rowsum(dz^T) fused into the same kernel
    GemmArgs g{dz, 1, lddz, x, ldx, 1, L.gw, in, nullptr, out, in, rows, 1.f, 1, ACT_NONE, L.gb};
    cudaStream_t ws_ = c.async_w ? c.wst : c.st;
    if (c.async_w) dep(c, c.st, c.wst);      // dz is complete on the compute stream
    c.chk(launch_gemm(g, ws_), "lin_bwd.wgrad");
  } else if (L.gb != nullptr) {
    colsum(c.st, dz, lddz, nullptr, 0, L.gb, rows, out);
    c.chk(cudaGetLastError(), "lin_bwd.bgrad");
  }
}
void act_bwd(Ctx& c, const float* dy, long long lddy, const float* y, long long ldy, float* dz, long long lddz, int rows,
             int cols, int act, const DropSpec drop = DropSpec{}) {
  launch_k(act_bwd_kernel, dim3(nblk((long long)rows * cols)), dim3(256), 0, c.st, dy, lddy, y, ldy, dz, lddz, rows, cols, act,
           drop); count_launch();
  c.chk(cudaGetLastError(), "act_bwd");
}
void add(Ctx& c, const float* a, const float* b, float* out, long long n) {
  launch_k(add_kernel, dim3(nblk(n)), dim3(256), 0, c.st, a, b, out, n); count_launch();
  c.chk(cudaGetLastError(), "add");
}
void mul(Ctx& c, const float* a, const float* b, float* out, long long n) {
  launch_k(mul_kernel, dim3(nblk(n)), dim3(256), 0, c.st, a, b, out, n); count_launch();
  c.chk(cudaGetLastError(), "mul");
}
void act(Ctx& c, const float* x, float* y, long long n, int a) {
  launch_k(act_kernel, dim3(nblk(n)), dim3(256), 0, c.st, x, y, n, a); count_launch();
  c.chk(cudaGetLastError(), "act");
}
void ln_fwd(Ctx& c, const float* a, const float* b, const mpo_norm& N, float* y, float* xh, float* rs, int rows) {
  launch_k(layernorm_fwd_kernel, dim3(nblk(rows, 8)), dim3(256), 0, c.st, a, b, N.g, N.b, y, xh, rs, rows); count_launch();
  c.chk(cudaGetLastError(), "ln_fwd");
}
void ln_bwd(Ctx& c, const float* dy, const mpo_norm& N, const float* xh, const float* rs, float* dx, int rows,
            float* dx_drop = nullptr, const DropSpec drop = DropSpec{}) {
  launch_k(layernorm_bwd_kernel, dim3(nblk(rows, 8)), dim3(256), 0, c.st, dy, N.g, xh, rs, dx, rows, dx_drop, drop); count_launch();
  c.chk(cudaGetLastError(), "ln_bwd");
  if (N.gg != nullptr) {
    cudaStream_t ws_ = c.async_w ? c.wst : c.st;
    if (c.async_w) dep(c, c.st, c.wst);
    colsum(ws_, dy, E, xh, E, N.gg, rows, E);
    colsum(ws_, dy, E, nullptr, 0, N.gb, rows, E);
    c.chk(cudaGetLastError(), "ln_bwd.params");
  }
}

// ---------------------------------------------------------------------------------------------- encoder layer
// reference: nn.TransformerEncoderLayer (post-norm) as built at models/mcat/mcat.py:51-53
void enc_fwd(Ctx& c, const mpo_encoder_layer& P, const EncBuf& b, float* ws, const float* x, int B, int eidx) {
  const int R = 6 * B;
  const uint32_t s0 = SITE_ENC + 4 * eidx;      // attention probabilities, dropout1, feed-forward dropout, dropout2
  lin_fwd(c, x, E, P.in_proj, 3 * E, E, ws + b.qkv, 3 * E, R, ACT_NONE);
  launch_k(mha6_fwd_kernel, dim3(nblk((long long)B * 8, 8)), dim3(256), 0, c.st, ws + b.qkv, ws + b.probs, ws + b.ctx, B,
           mk_drop(c, c.drop_p, s0)); count_launch();
  c.chk(cudaGetLastError(), "mha6_fwd");
  lin_fwd(c, ws + b.ctx, E, P.out_proj, E, E, ws + b.sa, E, R, ACT_NONE, mk_drop(c, c.drop_p, s0 + 1));
  ln_fwd(c, x, ws + b.sa, P.norm1, ws + b.y1, ws + b.xh1, ws + b.rs1, R);
  lin_fwd(c, ws + b.y1, E, P.linear1, FF, E, ws + b.f, FF, R, ACT_RELU, mk_drop(c, c.drop_p, s0 + 2));
  lin_fwd(c, ws + b.f, FF, P.linear2, E, FF, ws + b.f2, E, R, ACT_NONE, mk_drop(c, c.drop_p, s0 + 3));
  ln_fwd(c, ws + b.y1, ws + b.f2, P.norm2, ws + b.y2, ws + b.xh2, ws + b.rs2, R);
}
// dy2 -> dx (written to dx_out).  Scratch of branch c.sb; every gradient that a pending weight-gradient GEMM still
// reads keeps its own buffer until the next join_w().
void enc_bwd(Ctx& c, const mpo_encoder_layer& P, const EncBuf& b, const Ws& w, float* ws, const float* x,
             const float* dy2, float* dx_out, int B, int eidx) {
  const int R = 6 * B;
  const uint32_t s0 = SITE_ENC + 4 * eidx;
  const bool train = c.train && c.drop_p > 0.f;
  join_w(c);
  float* dr2 = ws + w.s256a[c.sb];     // gradient of (y1 + dropout2(f2))
  const float* df2 = dr2;              // gradient of f2: through dropout2 in train mode (second output of the LN backward)
  if (train) {
    ln_bwd(c, dy2, P.norm2, ws + b.xh2, ws + b.rs2, dr2, R, ws + w.dmk2[c.sb], mk_drop(c, c.drop_p, s0 + 3));
    df2 = ws + w.dmk2[c.sb];
  } else {
    ln_bwd(c, dy2, P.norm2, ws + b.xh2, ws + b.rs2, dr2, R);
  }
  // linear2 data gradient with the ReLU (+ feed-forward dropout) derivative in its epilogue: df is at linear1's pre-activation
  float* df = ws + w.s512[c.sb];
  DgradEpi e2; e2.y = ws + b.f; e2.ld_y = FF; e2.act = ACT_RELU; e2.drop = mk_drop(c, c.drop_p, s0 + 2);
  lin_bwd(c, df2, E, ws + b.f, FF, P.linear2, E, FF, df, FF, R, false, e2);
  // linear1 data gradient + the residual branch: dy1 = df W1 + dr2
  float* dy1 = ws + w.s256b[c.sb];
  DgradEpi e1; e1.addend = dr2; e1.ld_add = E;
  lin_bwd(c, df, FF, ws + b.y1, E, P.linear1, FF, E, dy1, E, R, false, e1);
  float* dr1 = ws + w.s256c[c.sb];     // gradient of (x + dropout1(sa))
  const float* dsa = dr1;
  if (train) {
    ln_bwd(c, dy1, P.norm1, ws + b.xh1, ws + b.rs1, dr1, R, ws + w.dmk1[c.sb], mk_drop(c, c.drop_p, s0 + 1));
    dsa = ws + w.dmk1[c.sb];
  } else {
    ln_bwd(c, dy1, P.norm1, ws + b.xh1, ws + b.rs1, dr1, R);
  }
  float* dctx = ws + w.dxb[c.sb];
  lin_bwd(c, dsa, E, ws + b.ctx, E, P.out_proj, E, E, dctx, E, R, false);
  float* dqkv = ws + w.s768[c.sb];
  launch_k(mha6_bwd_kernel, dim3(nblk((long long)B * 8, 8)), dim3(256), 0, c.st, ws + b.qkv, ws + b.probs, dctx, dqkv, B,
           mk_drop(c, c.drop_p, s0)); count_launch();
  c.chk(cudaGetLastError(), "mha6_bwd");
  // in_proj data gradient + the residual branch: dx = dqkv W_in + dr1
  DgradEpi e0; e0.addend = dr1; e0.ld_add = E;
  lin_bwd(c, dqkv, 3 * E, x, E, P.in_proj, 3 * E, E, dx_out, E, R, false, e0);
}

// ---------------------------------------------------------------------------------------------- pooling + rho
// reference: models/blocks.py:42-48, models/mcat/mcat.py:105-109
void pool_fwd(Ctx& c, const mpo_pool_head& P, const PoolBuf& b, float* ws, const float* x, float* att_logits, int B,
              int pidx) {
  const int R = 6 * B;
  // AttentionNetGated hard-codes p = 0.25 for its two dropout layers (blocks.py:34-36); rho uses the model's dropout
  lin_fwd(c, x, E, P.att_a, E, E, ws + b.a, E, R, ACT_TANH, mk_drop(c, 0.25f, SITE_POOL + 2 * pidx));
  lin_fwd(c, x, E, P.att_b, E, E, ws + b.b, E, R, ACT_SIGMOID, mk_drop(c, 0.25f, SITE_POOL + 2 * pidx + 1));
  launch_k(pool_fwd_kernel, dim3(B), dim3(256), 0, c.st, x, ws + b.a, ws + b.b, P.att_c.w, P.att_c.b, att_logits, ws + b.w, ws + b.hp); count_launch();
  c.chk(cudaGetLastError(), "pool_fwd");
  lin_fwd(c, ws + b.hp, E, P.rho, E, E, ws + b.h, b.hld, B, ACT_RELU, mk_drop(c, c.drop_p, SITE_RHO + pidx));
}
// dh [B,256] (gradient of rho's output) -> dx_out [R,256]
void pool_bwd(Ctx& c, const mpo_pool_head& P, const PoolBuf& b, const Ws& w, float* ws, const float* x, const float* dh,
              long long lddh, float* dhp, float* dx_out, int B, int pidx) {
  const int R = 6 * B;
  join_w(c);
  float* dzr = ws + w.dzr[c.sb];
  act_bwd(c, dh, lddh, ws + b.h, b.hld, dzr, E, B, E, ACT_RELU, mk_drop(c, c.drop_p, SITE_RHO + pidx));
  lin_bwd(c, dzr, E, ws + b.hp, E, P.rho, E, E, dhp, E, B, false);
  launch_k(pool_bwd_kernel, dim3(B), dim3(256), 0, c.st, x, ws + b.a, ws + b.b, P.att_c.w, ws + b.w, dhp, dx_out, ws + w.dxa[c.sb],
           ws + w.dxb[c.sb], P.att_c.gw, P.att_c.gb, mk_drop(c, 0.25f, SITE_POOL + 2 * pidx),
           mk_drop(c, 0.25f, SITE_POOL + 2 * pidx + 1)); count_launch();
  c.chk(cudaGetLastError(), "pool_bwd");
  lin_bwd(c, ws + w.dxa[c.sb], E, x, E, P.att_a, E, E, dx_out, E, R, true);
  lin_bwd(c, ws + w.dxb[c.sb], E, x, E, P.att_b, E, E, dx_out, E, R, true);
  join_w(c);                 // dxa / dxb are reused right away by the encoder backward
}

// ---------------------------------------------------------------------------------------------- CAG (NaCAGaT)
// reference: models/blocks.py:247-253   C = fc_c( LN(ELU(fc1 Q + fc2 Qh)) * LN(ELU(fc3 Qh)) ), every fc = Linear+ELU
void cag_fwd(Ctx& c, const mpo_cag& P, const Ws& w, float* ws, const float* Q, const float* Qh, int R) {
  lin_fwd(c, Q, E, P.fc1, E, E, ws + w.cag_f1, E, R, ACT_ELU);
  lin_fwd(c, Qh, E, P.fc2, E, E, ws + w.cag_f2, E, R, ACT_ELU);
  lin_fwd(c, Qh, E, P.fc3, E, E, ws + w.cag_f3, E, R, ACT_ELU);
  add(c, ws + w.cag_f1, ws + w.cag_f2, ws + w.cag_s, (long long)R * E);
  act(c, ws + w.cag_s, ws + w.cag_u, (long long)R * E, ACT_ELU);
  ln_fwd(c, ws + w.cag_u, nullptr, P.G, ws + w.cag_Gg, ws + w.cag_Gxh, ws + w.cag_Grs, R);
  act(c, ws + w.cag_f3, ws + w.cag_w, (long long)R * E, ACT_ELU);
  ln_fwd(c, ws + w.cag_w, nullptr, P.E, ws + w.cag_Ee, ws + w.cag_Exh, ws + w.cag_Ers, R);
  mul(c, ws + w.cag_Gg, ws + w.cag_Ee, ws + w.cag_m, (long long)R * E);
  lin_fwd(c, ws + w.cag_m, E, P.fc_c, E, E, ws + w.cag_C, E, R, ACT_ELU);
}
// dC -> dQ accumulated into dQ_acc, dQh written to dQh_out
void cag_bwd(Ctx& c0, const mpo_cag& P, const Ws& w, float* ws, const float* Q, const float* Qh, const float* dC,
             float* dQ_acc, float* dQh_out, int R) {
  const long long n = (long long)R * E;
  join_w(c0);
  Ctx c = c0;                // scratch is recycled aggressively here: keep the weight gradients on the compute stream
  c.async_w = false;
  float* t0 = ws + w.s256a[0]; float* t1 = ws + w.s256b[0]; float* t2 = ws + w.s256c[0];
  act_bwd(c, dC, E, ws + w.cag_C, E, t0, E, R, E, ACT_ELU);
  lin_bwd(c, t0, E, ws + w.cag_m, E, P.fc_c, E, E, t1, E, R, false);          // t1 = dm
  mul(c, t1, ws + w.cag_Ee, t0, n);                                            // t0 = dGg
  mul(c, t1, ws + w.cag_Gg, t2, n);                                            // t2 = dEe
  ln_bwd(c, t0, P.G, ws + w.cag_Gxh, ws + w.cag_Grs, t1, R);                   // t1 = du
  act_bwd(c, t1, E, ws + w.cag_u, E, t1, E, R, E, ACT_ELU);                    // t1 = d(f1+f2)
  ln_bwd(c, t2, P.E, ws + w.cag_Exh, ws + w.cag_Ers, t0, R);                   // t0 = dw
  act_bwd(c, t0, E, ws + w.cag_w, E, t0, E, R, E, ACT_ELU);                    // t0 = df3 (post fc3's ELU)
  act_bwd(c, t0, E, ws + w.cag_f3, E, t0, E, R, E, ACT_ELU);                   // through fc3's own ELU
  lin_bwd(c, t0, E, Qh, E, P.fc3, E, E, dQh_out, E, R, false);
  act_bwd(c, t1, E, ws + w.cag_f2, E, t2, E, R, E, ACT_ELU);
  lin_bwd(c, t2, E, Qh, E, P.fc2, E, E, dQh_out, E, R, true);
  act_bwd(c, t1, E, ws + w.cag_f1, E, t2, E, R, E, ACT_ELU);
  lin_bwd(c, t2, E, Q, E, P.fc1, E, E, dQ_acc, E, R, true);
  c0.chk(c.err, c.where);
}

// ---------------------------------------------------------------------------------------------- bilinear fusion
// reference: models/fusion.py:81-113 (eval: dropouts are identity)
void bil_side_fwd(Ctx& c, const mpo_lin& Lh, const mpo_lin& Lz, const mpo_lin& Lo, const Ws& w, float* ws, int s,
                  const float* xa, const float* xb, int B) {
  lin_fwd(c, xa, E, Lh, BH, E, ws + w.bh[s], BH, B, ACT_RELU);
  // U[b][k*256+i] = sum_j W[k][i][j] xb[b][j]
  GemmArgs g{xb, E, 1, Lz.w, 1, E, ws + w.bU[s], (long long)BH * E, nullptr, B, BH * E, E, 1.f, 0, ACT_NONE, nullptr};
  c.chk(launch_gemm(g, c.st), "bil.U");
  launch_k(bil_gate_fwd_kernel, dim3(B), dim3(256), 0, c.st, xa, ws + w.bU[s], Lz.b, ws + w.bh[s], ws + w.bg[s], ws + w.bgh[s]); count_launch();
  c.chk(cudaGetLastError(), "bil_gate_fwd");
  lin_fwd(c, ws + w.bgh[s], BH, Lo, BH, BH, ws + w.bo[s], BH, B, ACT_RELU, mk_drop(c, 0.25f, SITE_BIL + s));   // fusion.py:58,62
}
// do (gradient of o_s) in w.bdo[s] -> dxa (=|+=), dxb (+=)
void bil_side_bwd(Ctx& c, const mpo_lin& Lh, const mpo_lin& Lz, const mpo_lin& Lo, const Ws& w, float* ws, int s,
                  const float* xa, const float* xb, float* dxa, bool acc_a, float* dxb, int B) {
  float* dpre = ws + w.bdo[s];
  act_bwd(c, dpre, BH, ws + w.bo[s], BH, dpre, BH, B, BH, ACT_RELU, mk_drop(c, 0.25f, SITE_BIL + s));
  lin_bwd(c, dpre, BH, ws + w.bgh[s], BH, Lo, BH, BH, ws + w.bdgh[s], BH, B, false);
  launch_k(bil_gate_bwd_kernel, dim3(B), dim3(256), 0, c.st, xa, ws + w.bU[s], ws + w.bh[s], ws + w.bg[s], ws + w.bdgh[s], ws + w.bdh[s],
                                           ws + w.bdz[s], ws + w.bV, dxa, acc_a ? 1 : 0); count_launch();
  c.chk(cudaGetLastError(), "bil_gate_bwd");
  // dW[(k,i)][j] += sum_b V[b][(k,i)] xb[b][j] ;  dxb[b][j] += sum_(k,i) V[b][(k,i)] W[(k,i)][j] ; db += colsum(dz)
  if (Lz.gw != nullptr) {
    GemmArgs g{ws + w.bV, 1, (long long)BH * E, xb, E, 1, Lz.gw, E, nullptr, BH * E, E, B, 1.f, 1, ACT_NONE, nullptr};
    c.chk(launch_gemm(g, c.st), "bil.dW");
    launch_k(colsum_kernel, dim3(1), dim3(256), 0, c.st, ws + w.bdz[s], BH, nullptr, 0, Lz.gb, B, BH); count_launch();
  }
  GemmArgs g2{ws + w.bV, (long long)BH * E, 1, Lz.w, E, 1, dxb, E, nullptr, B, E, BH * E, 1.f, 1, ACT_NONE, nullptr};
  c.chk(launch_gemm(g2, c.st), "bil.dxb");
  lin_bwd(c, ws + w.bdh[s], BH, xa, E, Lh, BH, E, dxa, E, B, true);
}

// main context + (optional) second-branch context sharing the caller's stream as the join point
struct Branches {
  Ctx main, second;
  bool par;
};
Branches make_branches(cudaStream_t stream, bool async_wgrad, const mpo_tail_io* io = nullptr) {
  StreamPool& p = stream_pool();
  Branches b;
  b.par = p.state == 1;
  b.main.st = stream;
  b.main.sb = 0;
  b.second.sb = 1;
  if (b.par) {
    b.second.st = p.aux[0];
    b.main.wst = p.aux[1];
    b.second.wst = p.aux[2];
    b.main.async_w = b.second.async_w = async_wgrad;
  } else {
    b.second.st = stream;     // everything in program order on the caller's stream
  }
  if (io != nullptr && io->train != 0) {
    b.main.train = b.second.train = true;
    b.main.drop_p = b.second.drop_p = io->drop_p;
    b.main.seed = b.second.seed = io->seed;
    b.main.seed_dev = b.second.seed_dev = io->seed_dev;
  }
  return b;
}
// second branch (and both weight-gradient streams) start after everything already on the caller's stream
void fork(Branches& b) {
  if (!b.par) return;
  dep(b.main, b.main.st, b.second.st);
  if (b.main.async_w) dep(b.main, b.main.st, b.main.wst);
  if (b.second.async_w) dep(b.main, b.main.st, b.second.wst);
}
// the caller's stream waits for the second branch and for every pending weight gradient
void join(Branches& b) {
  if (b.par) {
    if (b.second.async_w) dep(b.main, b.second.wst, b.main.st);
    if (b.main.async_w) dep(b.main, b.main.wst, b.main.st);
    dep(b.main, b.second.st, b.main.st);
  }
  b.main.chk(b.second.err, b.second.where);
}

// the six SNN chains on six branches: chain 0 on the caller's stream, 1 on the second-branch stream, 2..5 on their own
void snn_branches(Branches& b, Ctx (&chain)[MPO_Q]) {
  StreamPool& p = stream_pool();
  for (int i = 0; i < MPO_Q; ++i) {
    chain[i] = b.main;
    chain[i].async_w = false;          // a chain's weight gradients stay on the chain's own stream
    if (b.par && i > 0) {
      chain[i].st = i == 1 ? p.aux[0] : p.aux[1 + i];
      dep(b.main, b.main.st, chain[i].st);
    }
  }
}
void snn_join(Branches& b, Ctx (&chain)[MPO_Q]) {
  for (int i = 0; i < MPO_Q; ++i) {
    if (chain[i].st != b.main.st) dep(b.main, chain[i].st, b.main.st);
    b.main.chk(chain[i].err, chain[i].where);
  }
}

int finish(Ctx& c) {
  if (c.err == cudaSuccess) return MPO_OK;
  snprintf(g_err, sizeof(g_err), "tail (%s): %s", c.where, cudaGetErrorString(c.err));
  return MPO_E_CUDA;
}

int check_model(const mpo_model* m, const mpo_tail_io* io, const char* who) {
  if (!m || !io) return fail(MPO_E_ARG, "%s: NULL model/io", who);
  if (m->variant != MPO_VARIANT_MCAT && m->variant != MPO_VARIANT_NACAGAT) return fail(MPO_E_UNSUPPORTED, "%s: unknown variant", who);
  if (m->fusion != MPO_FUSION_CONCAT && m->fusion != MPO_FUSION_BILINEAR && m->fusion != MPO_FUSION_GATED_CONCAT)
    return fail(MPO_E_UNSUPPORTED, "%s: unsupported fusion", who);
  if (m->fusion == MPO_FUSION_GATED_CONCAT && (!m->gate[0].w || !m->gate[0].b || !m->gate[1].w || !m->gate[1].b))
    return fail(MPO_E_ARG, "%s: gated_concat needs the two gate layers", who);
  if (m->n_classes < 1 || m->n_classes > 16) return fail(MPO_E_ARG, "%s: n_classes out of range", who);
  if (io->num_slides <= 0) return fail(MPO_E_ARG, "%s: num_slides must be positive", who);
  if (!io->ws) return fail(MPO_E_ARG, "%s: workspace is NULL", who);
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s: no CUDA device (this library has no CPU fallback)", who);
  return MPO_OK;
}

// ---------------------------------------------------------------------------------------------- GE-NaCAGaT
// reference: models/ge_nacagat/ge_nacagat.py:41-72.  One slide per call; every stage is a bag-scale op over the N
// patch tokens.  First functional version: fp32 CUDA-core GEMMs with the N x N attention matrices materialised
// (the reference does the same and returns the 1-head map), tensor cores only in the H projection / dW_H.
struct GeEnc { long long qkv, probs, ctx, sa, y1, xh1, rs1, f, f2, y2, xh2, rs2; };
struct GeWs {
  long long H, sa_qkv, sa_ctx, sa_out;
  GeEnc enc[2];
  long long pa, pb, pab, pw, hp, h, logits;
  long long dlogits, dh, dzr, dhp, dw, dA, dab, t0, dx, dr2, df, dy1, dr1, dctx, dqkv, dP, dmid, dH, dzf;
  long long dP2, dmk1, dmk2;      // train mode: dropped probabilities of one head; gradients behind dropout1 / dropout2
  long long tcA, tcS[4];          // bf16 (hi, lo) operand pairs of the tensor-core GEMMs: one N x N operand, four N x 256 ones
  long long tcWdz, tcWx;          // operand pairs of the weight-gradient GEMMs: dz (<= 768 columns), x (<= 512 columns)
  long long total;
};
void ge_layout(long long N, GeWs& w) {
  long long off = 0;
  auto A = [&](long long n) { const long long o = off; off += (n + 63) / 64 * 64; return o; };
  w.H = A(N * E); w.sa_qkv = A(N * 3 * E); w.sa_ctx = A(N * E); w.sa_out = A(N * E);
  for (int l = 0; l < 2; ++l) {
    GeEnc& b = w.enc[l];
    b.qkv = A(N * 3 * E); b.probs = A(8 * N * N); b.ctx = A(N * E); b.sa = A(N * E); b.y1 = A(N * E); b.xh1 = A(N * E);
    b.rs1 = A(N); b.f = A(N * FF); b.f2 = A(N * E); b.y2 = A(N * E); b.xh2 = A(N * E); b.rs2 = A(N);
  }
  w.pa = A(N * E); w.pb = A(N * E); w.pab = A(N * E); w.pw = A(N); w.hp = A(E); w.h = A(E); w.logits = A(16);
  w.dlogits = A(16); w.dh = A(E); w.dzr = A(E); w.dhp = A(E); w.dw = A(N); w.dA = A(N); w.dab = A(N * E); w.t0 = A(N * E);
  w.dx = A(N * E); w.dr2 = A(N * E); w.df = A(N * FF); w.dy1 = A(N * E); w.dr1 = A(N * E); w.dctx = A(N * E);
  w.dqkv = A(N * 3 * E); w.dP = A(N * N); w.dmid = A(N * E); w.dH = A(N * E); w.dzf = A(N * E);
  w.dP2 = A(N * N); w.dmk1 = A(N * E); w.dmk2 = A(N * E);
  const long long Np = (N + 63) / 64 * 64;
  w.tcA = A(N * Np);                                   // 2 x [N][Np] bf16 = N * Np floats
  for (int i = 0; i < 4; ++i) w.tcS[i] = A(N * E);     // 2 x [N][256] bf16 = N * 256 floats
  w.tcWdz = A(N * 768); w.tcWx = A(N * 512);
  w.total = off;
}
// multi-head attention over N tokens from a packed [N, 3E] projection: probs [nh][N][N], ctx [N, E]
// The bag-scale products of the attention run on tcgen05 (tc_gemm.cu: bf16 hi/lo operand pairs, three MMAs per K step,
// fp32 accumulation); MPO_GE_TC=0 keeps the fp32 CUDA-core GEMMs (gemm_big_kernel) for comparison.
bool ge_use_tc() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MPO_GE_TC"); on = e ? atoi(e) : 1; }
  return on != 0;
}
struct TcOp { void* hi; void* lo; int rows, pitch; };
// splits a strided fp32 matrix [rows x cols] into the bf16 pair living at float offset `off` of the workspace
// drop / base: the split applies the attention-probability dropout on the fly (element index base + row * cols + col)
TcOp tc_split(Ctx& c, float* ws, long long off, const float* src, long long ld, int rows, int cols,
              const DropSpec* drop = nullptr, uint32_t base = 0) {
  TcOp o;
  o.rows = rows; o.pitch = (cols + 63) / 64 * 64;
  o.hi = ws + off;
  o.lo = reinterpret_cast<__nv_bfloat16*>(ws + off) + static_cast<long long>(rows) * o.pitch;
  c.chk(launch_split_bf16(src, ld, rows, cols, o.hi, o.lo, o.pitch, c.st, drop, base), "ge.split");
  return o;
}
void tc_gemm(Ctx& c, const TcOp& a, bool a_mn, const TcOp& b, bool b_mn, float* C, long long ldc, int M, int Nn, int K,
             float alpha) {
  if (c.err != cudaSuccess) return;
  if (launch_tc_gemm(a.hi, a.lo, a.rows, a.pitch, a_mn, b.hi, b.lo, b.rows, b.pitch, b_mn, C, ldc, M, Nn, K, alpha, false, c.st) != 0)
    c.chk(cudaErrorUnknown, "ge.tc_gemm");
}

// drop / pdrop: attention-probability dropout (train mode); the stored probabilities stay un-dropped (the soft-max
// backward needs them), the dropped copy of one head lives in the scratch `pdrop` while its context GEMM runs
void ge_attn_fwd(Ctx& c, const GeWs& w, float* ws, const float* qkv, float* probs, float* ctx, int N, int nh,
                 const DropSpec drop = DropSpec{}, float* pdrop = nullptr) {
  const int hd = E / nh;
  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  if (ge_use_tc()) {
    for (int h = 0; h < nh; ++h) {
      float* P = probs + (long long)h * N * N;
      const TcOp sQ = tc_split(c, ws, w.tcS[0], qkv + h * hd, 3 * E, N, hd);
      const TcOp sK = tc_split(c, ws, w.tcS[1], qkv + E + h * hd, 3 * E, N, hd);
      tc_gemm(c, sQ, false, sK, false, P, N, N, N, hd, scale);                      // S = Q K^T / sqrt(hd)
      launch_k(row_softmax_kernel, dim3(N), dim3(256), 0, c.st, P, (long long)N, N); count_launch();
      // train mode: the dropped probabilities exist only as the bf16 operand pair (mask regenerated inside the split)
      const TcOp sP = tc_split(c, ws, w.tcA, P, N, N, N, &drop,
                               static_cast<uint32_t>(h) * static_cast<uint32_t>(N) * static_cast<uint32_t>(N));
      const TcOp sV = tc_split(c, ws, w.tcS[2], qkv + 2 * E + h * hd, 3 * E, N, hd);   // [tokens][hd]: N-major B
      tc_gemm(c, sP, false, sV, true, ctx + h * hd, E, N, hd, N, 1.f);              // ctx_h = P V_h
    }
    c.chk(cudaGetLastError(), "ge_attn_fwd (tc)");
    return;
  }
  for (int h = 0; h < nh; ++h) {
    float* P = probs + (long long)h * N * N;
    GemmArgs s{qkv + h * hd, 3 * E, 1, qkv + E + h * hd, 1, 3 * E, P, N, nullptr, N, N, hd, scale, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(s, c.st), "ge.scores");
    launch_k(row_softmax_kernel, dim3(N), dim3(256), 0, c.st, P, (long long)N, N); count_launch();
    if (drop.thr != 0) {
      launch_k(probs_dropout_kernel, dim3(nblk((long long)N * N, 1024)), dim3(256), 0, c.st, (const float*)P, pdrop,
               (long long)N * N, static_cast<uint32_t>(h) * static_cast<uint32_t>(N) * static_cast<uint32_t>(N), drop); count_launch();
      P = pdrop;
    }
    GemmArgs o{P, N, 1, qkv + 2 * E + h * hd, 3 * E, 1, ctx + h * hd, E, nullptr, N, hd, N, 1.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(o, c.st), "ge.ctx");
  }
  c.chk(cudaGetLastError(), "ge_attn_fwd");
}
// dctx [N, E] -> dqkv [N, 3E]; dP is an [N, N] scratch
void ge_attn_bwd(Ctx& c, const GeWs& w, float* ws, const float* qkv, const float* probs, const float* dctx, float* dP,
                 float* dqkv, int N, int nh, const DropSpec drop = DropSpec{}, float* pdrop = nullptr) {
  const int hd = E / nh;
  const float scale = 1.f / sqrtf(static_cast<float>(hd));
  if (ge_use_tc()) {
    for (int h = 0; h < nh; ++h) {
      const float* P = probs + (long long)h * N * N;
      const uint32_t base = static_cast<uint32_t>(h) * static_cast<uint32_t>(N) * static_cast<uint32_t>(N);
      const TcOp sD = tc_split(c, ws, w.tcS[3], dctx + h * hd, E, N, hd);
      const TcOp sV = tc_split(c, ws, w.tcS[2], qkv + 2 * E + h * hd, 3 * E, N, hd);
      tc_gemm(c, sD, false, sV, false, dP, N, N, N, hd, 1.f);                       // dP' = dctx_h V_h^T
      const TcOp sP = tc_split(c, ws, w.tcA, P, N, N, N, &drop, base);              // P' = dropout(P), regenerated
      tc_gemm(c, sP, true, sD, true, dqkv + 2 * E + h * hd, 3 * E, N, hd, N, 1.f);  // dV_h = P'^T dctx_h (P' read M-major)
      // dS as the operand pair of the two products below, written by the soft-max backward itself (no fp32 dS, no split
      // pass); it reuses the N x N operand slot of P'
      TcOp sS;
      sS.rows = N; sS.pitch = (N + 63) / 64 * 64;
      sS.hi = ws + w.tcA;
      sS.lo = reinterpret_cast<__nv_bfloat16*>(ws + w.tcA) + static_cast<long long>(N) * sS.pitch;
      if (drop.thr != 0) {
        launch_k(row_softmax_bwd_pair_kernel<true>, dim3(N), dim3(256), 0, c.st, P, (long long)N, (const float*)dP, (long long)N, N,
                 scale, static_cast<__nv_bfloat16*>(sS.hi), static_cast<__nv_bfloat16*>(sS.lo), sS.pitch, base, drop); count_launch();
      } else {
        launch_k(row_softmax_bwd_pair_kernel<false>, dim3(N), dim3(256), 0, c.st, P, (long long)N, (const float*)dP, (long long)N, N,
                 scale, static_cast<__nv_bfloat16*>(sS.hi), static_cast<__nv_bfloat16*>(sS.lo), sS.pitch, base, DropSpec{}); count_launch();
      }
      const TcOp sK = tc_split(c, ws, w.tcS[1], qkv + E + h * hd, 3 * E, N, hd);
      const TcOp sQ = tc_split(c, ws, w.tcS[0], qkv + h * hd, 3 * E, N, hd);
      tc_gemm(c, sS, false, sK, true, dqkv + h * hd, 3 * E, N, hd, N, 1.f);         // dQ_h = dS K_h
      tc_gemm(c, sS, true, sQ, true, dqkv + E + h * hd, 3 * E, N, hd, N, 1.f);      // dK_h = dS^T Q_h (dS read M-major)
    }
    c.chk(cudaGetLastError(), "ge_attn_bwd (tc)");
    return;
  }
  for (int h = 0; h < nh; ++h) {
    const float* P = probs + (long long)h * N * N;
    const uint32_t base = static_cast<uint32_t>(h) * static_cast<uint32_t>(N) * static_cast<uint32_t>(N);
    // dP = dctx_h V_h^T ; dV_h = P'^T dctx_h  (P' = the dropped probabilities in train mode, regenerated)
    GemmArgs g1{dctx + h * hd, E, 1, qkv + 2 * E + h * hd, 1, 3 * E, dP, N, nullptr, N, N, hd, 1.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g1, c.st), "ge.dP");
    const float* Pv = P;
    if (drop.thr != 0) {
      launch_k(probs_dropout_kernel, dim3(nblk((long long)N * N, 1024)), dim3(256), 0, c.st, P, pdrop, (long long)N * N, base,
               drop); count_launch();
      Pv = pdrop;
    }
    GemmArgs g2{Pv, 1, N, dctx + h * hd, E, 1, dqkv + 2 * E + h * hd, 3 * E, nullptr, N, hd, N, 1.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g2, c.st), "ge.dV");
    if (drop.thr != 0) {
      launch_k(row_softmax_bwd_drop_kernel, dim3(N), dim3(256), 0, c.st, P, (long long)N, dP, (long long)N, N, scale, base,
               drop); count_launch();
    } else {
      launch_k(row_softmax_bwd_kernel, dim3(N), dim3(256), 0, c.st, P, (long long)N, dP, (long long)N, N, scale); count_launch();
    }
    // dQ_h = dS K_h ; dK_h = dS^T Q_h
    GemmArgs g3{dP, N, 1, qkv + E + h * hd, 3 * E, 1, dqkv + h * hd, 3 * E, nullptr, N, hd, N, 1.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g3, c.st), "ge.dQ");
    GemmArgs g4{dP, 1, N, qkv + h * hd, 3 * E, 1, dqkv + E + h * hd, 3 * E, nullptr, N, hd, N, 1.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g4, c.st), "ge.dK");
  }
  c.chk(cudaGetLastError(), "ge_attn_bwd");
}
// dropout sites of layer l as in enc_fwd: attention probabilities, dropout1, feed-forward dropout, dropout2
// (models/ge_nacagat/ge_nacagat.py:30-32: TransformerEncoderLayer(..., dropout=dropout))
void ge_enc_fwd(Ctx& c, const mpo_encoder_layer& P, const GeEnc& b, const GeWs& w, float* ws, const float* x, int N, int l) {
  const uint32_t s0 = SITE_ENC + 4 * l;
  lin_fwd(c, x, E, P.in_proj, 3 * E, E, ws + b.qkv, 3 * E, N, ACT_NONE);
  ge_attn_fwd(c, w, ws, ws + b.qkv, ws + b.probs, ws + b.ctx, N, 8, mk_drop(c, c.drop_p, s0), ws + w.dP2);
  lin_fwd(c, ws + b.ctx, E, P.out_proj, E, E, ws + b.sa, E, N, ACT_NONE, mk_drop(c, c.drop_p, s0 + 1));
  ln_fwd(c, x, ws + b.sa, P.norm1, ws + b.y1, ws + b.xh1, ws + b.rs1, N);
  lin_fwd(c, ws + b.y1, E, P.linear1, FF, E, ws + b.f, FF, N, ACT_RELU, mk_drop(c, c.drop_p, s0 + 2));
  lin_fwd(c, ws + b.f, FF, P.linear2, E, FF, ws + b.f2, E, N, ACT_NONE, mk_drop(c, c.drop_p, s0 + 3));
  ln_fwd(c, ws + b.y1, ws + b.f2, P.norm2, ws + b.y2, ws + b.xh2, ws + b.rs2, N);
}
void ge_enc_bwd(Ctx& c, const mpo_encoder_layer& P, const GeEnc& b, const GeWs& w, float* ws, const float* x,
                const float* dy2, float* dx_out, int N, int l) {
  const long long n = (long long)N * E;
  const uint32_t s0 = SITE_ENC + 4 * l;
  const bool train = c.train && c.drop_p > 0.f;
  // dr2 = gradient of (y1 + dropout2(f2)); df2 = gradient of f2 (through dropout2 in train mode)
  const float* df2 = ws + w.dr2;
  if (train) {
    ln_bwd(c, dy2, P.norm2, ws + b.xh2, ws + b.rs2, ws + w.dr2, N, ws + w.dmk2, mk_drop(c, c.drop_p, s0 + 3));
    df2 = ws + w.dmk2;
  } else {
    ln_bwd(c, dy2, P.norm2, ws + b.xh2, ws + b.rs2, ws + w.dr2, N);
  }
  lin_bwd(c, df2, E, ws + b.f, FF, P.linear2, E, FF, ws + w.df, FF, N, false);
  act_bwd(c, ws + w.df, FF, ws + b.f, FF, ws + w.df, FF, N, FF, ACT_RELU, mk_drop(c, c.drop_p, s0 + 2));
  lin_bwd(c, ws + w.df, FF, ws + b.y1, E, P.linear1, FF, E, ws + w.dy1, E, N, false);
  add(c, ws + w.dy1, ws + w.dr2, ws + w.dy1, n);
  const float* dsa = ws + w.dr1;
  if (train) {
    ln_bwd(c, ws + w.dy1, P.norm1, ws + b.xh1, ws + b.rs1, ws + w.dr1, N, ws + w.dmk1, mk_drop(c, c.drop_p, s0 + 1));
    dsa = ws + w.dmk1;
  } else {
    ln_bwd(c, ws + w.dy1, P.norm1, ws + b.xh1, ws + b.rs1, ws + w.dr1, N);
  }
  lin_bwd(c, dsa, E, ws + b.ctx, E, P.out_proj, E, E, ws + w.dctx, E, N, false);
  ge_attn_bwd(c, w, ws, ws + b.qkv, ws + b.probs, ws + w.dctx, ws + w.dP, ws + w.dqkv, N, 8, mk_drop(c, c.drop_p, s0), ws + w.dP2);
  lin_bwd(c, ws + w.dqkv, 3 * E, x, E, P.in_proj, 3 * E, E, dx_out, E, N, false);
  add(c, dx_out, ws + w.dr1, dx_out, n);
}

}  // namespace
}  // namespace mpo

using namespace mpo;

extern "C" {

int64_t mpo_sizeof(int32_t which) {
  switch (which) {
    case 0: return sizeof(mpo_bag);
    case 1: return sizeof(mpo_model);
    case 2: return sizeof(mpo_tail_io);
    case 3: return sizeof(mpo_nacagat_bwd);
    case 4: return sizeof(mpo_ge_model);
    case 5: return sizeof(mpo_peer_group);
    default: return -1;
  }
}

int64_t mpo_tail_ws_floats(const mpo_model* m, int32_t B) {
  if (!m || B <= 0) return -1;
  Ws w;
  build_layout(m, B, w);
  return w.lay.total;
}

int64_t mpo_tail_ws_lookup(const mpo_model* m, int32_t B, const char* name, int64_t* length) {
  if (!m || B <= 0 || !name) return -1;
  Ws w;
  build_layout(m, B, w);
  for (auto& it : w.lay.items)
    if (it.first == name) { if (length) *length = it.second.second; return it.second.first; }
  return -1;
}

int mpo_tail_pre_fwd(const mpo_model* m, const mpo_tail_io* io, void* stream) {
  int rc = check_model(m, io, "mpo_tail_pre_fwd");
  if (rc) return rc;
  if (!io->qp || !io->qk) return fail(MPO_E_ARG, "%s", "mpo_tail_pre_fwd: qp/qk are NULL");
  const int B = io->num_slides, R = 6 * B;
  Ws w; build_layout(m, B, w);
  float* ws = io->ws;
  Branches br = make_branches(static_cast<cudaStream_t>(stream), false, io);
  Ctx& c = br.main;
  // SNN encoders (mcat.py:32-45,90-92): G_bag row (b, i) = ELU(W2 ELU(W1 x_i + b1) + b2); six independent chains,
  // alternated over the two branch streams
  for (int i = 0; i < MPO_Q; ++i)
    if (!io->omics[i]) return fail(MPO_E_ARG, "%s", "mpo_tail_pre_fwd: omics pointer is NULL");
  if (fused::eligible(m, io)) return fused::pre_fwd(m, io, w, static_cast<cudaStream_t>(stream));
  // six independent chains: one branch each (the caller's stream, the second-branch stream, four SNN streams)
  Ctx chain[MPO_Q];
  snn_branches(br, chain);
  for (int i = 0; i < MPO_Q; ++i) {
    Ctx& ci = chain[i];
    const int d = m->omic_dims[i];
    // Linear + ELU + AlphaDropout, twice (mcat.py:34-44)
    lin_fwd(ci, io->omics[i], d, m->snn[i][0], E, d, ws + w.snn_h[i], E, B, ACT_ELU, mk_drop(ci, ci.drop_p, SITE_SNN + 2 * i, true));
    lin_fwd(ci, ws + w.snn_h[i], E, m->snn[i][1], E, E, ws + w.G + i * E, 6 * E, B, ACT_ELU,
            mk_drop(ci, ci.drop_p, SITE_SNN + 2 * i + 1, true));
  }
  snn_join(br, chain);
  // query in-projection (rows 0..255 of co_attention.in_proj): q = W_q g + b_q
  lin_fwd(c, ws + w.G, E, sub(m->coattn_in, 0, E), E, E, io->qp, E, R, ACT_NONE);
  // key fold: qk[r][d] = sum_e q[r][e] W_k[e][d] / sqrt(256)
  {
    const float* Wk = m->coattn_in.w + (long long)E * E;
    GemmArgs g{io->qp, E, 1, Wk, E, 1, io->qk, E, nullptr, R, E, E, 1.f / 16.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g, c.st), "fold");
  }
  if (m->variant == MPO_VARIANT_NACAGAT) {
    if (!io->kc) return fail(MPO_E_ARG, "%s", "mpo_tail_pre_fwd: kc is NULL (NaCAGaT)");
    // kc[r] = q[r] . b_k / 16
    const float* bk = m->coattn_in.b + E;
    GemmArgs g{io->qp, E, 1, bk, 1, 0, io->kc, 1, nullptr, R, 1, E, 1.f / 16.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g, c.st), "kc");
  }
  return finish(c);
}

int mpo_tail_post_fwd(const mpo_model* m, const mpo_tail_io* io, void* stream) {
  int rc = check_model(m, io, "mpo_tail_post_fwd");
  if (rc) return rc;
  if (!io->pooled || !io->hazards || !io->S || !io->Y || !io->att_path || !io->att_omic)
    return fail(MPO_E_ARG, "%s", "mpo_tail_post_fwd: NULL pointer");
  const int B = io->num_slides, R = 6 * B, K = m->n_classes;
  Ws w; build_layout(m, B, w);
  float* ws = io->ws;
  Branches br = make_branches(static_cast<cudaStream_t>(stream), false, io);
  Ctx& c = br.main;
  Ctx& co = br.second;
  if (m->variant == MPO_VARIANT_NACAGAT && !io->qp) return fail(MPO_E_ARG, "%s", "mpo_tail_post_fwd: qp is NULL (NaCAGaT)");
  if (fused::eligible(m, io))
    return fused::post(m, io, w, fused::F_FWD, nullptr, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream));
  fork(br);
  // omic branch (second stream): omic transformer + pooling (mcat.py:102,111-115) -- independent of the bag
  enc_fwd(co, m->omic_tr[0], w.enc[2], ws, ws + w.G, B, 2);
  enc_fwd(co, m->omic_tr[1], w.enc[3], ws, ws + w.enc[2].y2, B, 3);
  pool_fwd(co, m->omic_pool, w.pool[1], ws, ws + w.enc[3].y2, io->att_omic, B, 1);
  // path branch: value and output projections on the pooled vectors (folded form of mcat.py:97)
  if (io->suma == nullptr) {
    lin_fwd(c, io->pooled, E, sub(m->coattn_in, 2 * E, E), E, E, ws + w.v, E, R, ACT_NONE);
  } else {
    // attention dropout leaves sum_n a'_in != 1: v_i = W_v pooled_i + (sum_n a'_in) b_v   (blocks.py:189-192)
    mpo_lin Lv = sub(m->coattn_in, 2 * E, E);
    const float* bv = Lv.b;
    Lv.b = nullptr;
    lin_fwd(c, io->pooled, E, Lv, E, E, ws + w.v, E, R, ACT_NONE);
    GemmArgs g{io->suma, 1, 0, bv, 0, 1, ws + w.v, E, nullptr, R, E, 1, 1.f, 1, ACT_NONE, nullptr};
    c.chk(launch_gemm(g, c.st), "v.suma_bias");
  }
  lin_fwd(c, ws + w.v, E, m->coattn_out, E, E, ws + w.hc, E, R, ACT_NONE);
  if (m->variant == MPO_VARIANT_NACAGAT) {
    cag_fwd(c, m->cag, w, ws, ws + w.G, io->qp, R);                      // blocks.py:110
    add(c, ws + w.hc, ws + w.cag_C, ws + w.hc, (long long)R * E);        // blocks.py:111
  }
  // path transformer + pooling (mcat.py:101,105-109)
  enc_fwd(c, m->path_tr[0], w.enc[0], ws, ws + w.hc, B, 0);
  enc_fwd(c, m->path_tr[1], w.enc[1], ws, ws + w.enc[0].y2, B, 1);
  pool_fwd(c, m->path_pool, w.pool[0], ws, ws + w.enc[1].y2, io->att_path, B, 0);
  join(br);
  const float* hpath = ws + w.pool[0].h;
  const float* homic = ws + w.pool[1].h;
  const float* hfin;
  if (m->fusion != MPO_FUSION_BILINEAR) {          // fusion.py:17-19 (concat), :35-41 (gated concat)
    const float* cat = ws + w.cat;
    if (m->fusion == MPO_FUSION_GATED_CONCAT) {    // item_p * sigmoid(w_p . item_p + b_p) per input, then the same MLP
      launch_k(gate_concat_fwd_kernel, dim3(B), dim3(256), 0, c.st, ws + w.cat, m->gate[0].w, m->gate[0].b, m->gate[1].w,
               m->gate[1].b, ws + w.catg, ws + w.gateg); count_launch();
      c.chk(cudaGetLastError(), "gate_concat_fwd");
      cat = ws + w.catg;
    }
    lin_fwd(c, cat, 2 * E, m->fusion0, E, 2 * E, ws + w.z1, E, B, ACT_RELU);
    lin_fwd(c, ws + w.z1, E, m->fusion2, E, E, ws + w.z2, E, B, ACT_RELU);
    hfin = ws + w.z2;
  } else {                                          // fusion.py:81-113
    bil_side_fwd(c, m->bil.h1, m->bil.z1, m->bil.o1, w, ws, 0, hpath, homic, B);
    bil_side_fwd(c, m->bil.h2, m->bil.z2, m->bil.o2, w, ws, 1, homic, hpath, B);
    launch_k(bil_kron_fwd_kernel, dim3(B), dim3(256), 0, c.st, ws + w.bo[0], ws + w.bo[1], ws + w.kp, ws + w.cat130,
             mk_drop(c, 0.25f, SITE_BIL + 2)); count_launch();    // post_fusion_dropout, fusion.py:64,107
    c.chk(cudaGetLastError(), "bil_kron_fwd");
    lin_fwd(c, ws + w.kp, 1089, m->bil.fc1, BMM, 1089, ws + w.cat130, 130, B, ACT_RELU, mk_drop(c, 0.25f, SITE_BIL + 3));
    lin_fwd(c, ws + w.cat130, 130, m->bil.fc2, E, 130, ws + w.bf2, E, B, ACT_RELU, mk_drop(c, 0.25f, SITE_BIL + 4));
    hfin = ws + w.bf2;
  }
  lin_fwd(c, hfin, E, m->classifier, K, E, ws + w.logits, K, B, ACT_NONE);
  launch_k(surv_head_fwd_kernel, dim3(nblk(B, 128)), dim3(128), 0, c.st, ws + w.logits, io->hazards, io->S, io->Y, B, K); count_launch();
  c.chk(cudaGetLastError(), "surv_head_fwd");
  return finish(c);
}

int mpo_surv_loss(int32_t kind, const float* hazards, const float* S, const int64_t* label, const float* censor,
                  float alpha, float eps, float grad_scale, float* loss, float* dhaz, float* dS, int32_t B,
                  int32_t n_classes, void* stream) {
  if (kind != MPO_LOSS_NLL && kind != MPO_LOSS_CES) return fail(MPO_E_UNSUPPORTED, "%s", "mpo_surv_loss: unknown loss kind");
  if (!hazards || !S || !label || !censor || !loss || !dhaz || !dS || B <= 0)
    return fail(MPO_E_ARG, "%s", "mpo_surv_loss: bad arguments");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_surv_loss: no CUDA device (this library has no CPU fallback)");
  launch_k(surv_loss_kernel, dim3(nblk(B, 128)), dim3(128), 0, static_cast<cudaStream_t>(stream), kind, hazards, S, label, censor, alpha,
                                                                               eps, grad_scale, loss, dhaz, dS, B,
                                                                               n_classes); count_launch();
  return check_cuda(cudaGetLastError(), "surv_loss_kernel");
}

int mpo_tail_post_bwd(const mpo_model* m, const mpo_tail_io* io, const float* dhaz, const float* dS, const float* dY,
                      void* stream) {
  int rc = check_model(m, io, "mpo_tail_post_bwd");
  if (rc) return rc;
  if (!io->pooled || !io->dpooled || !io->hazards || !io->S || !io->Y)
    return fail(MPO_E_ARG, "%s", "mpo_tail_post_bwd: NULL pointer");
  const int B = io->num_slides, R = 6 * B, K = m->n_classes;
  Ws w; build_layout(m, B, w);
  float* ws = io->ws;
  if (fused::eligible(m, io))
    return fused::post(m, io, w, fused::F_BWD, nullptr, dhaz, dS, dY, static_cast<cudaStream_t>(stream));
  Branches br = make_branches(static_cast<cudaStream_t>(stream), true, io);
  Ctx c = br.main;            // fusion / head part: weight gradients on the side stream (joined at the end)
  launch_k(surv_head_bwd_kernel, dim3(nblk(B, 128)), dim3(128), 0, c.st, io->hazards, io->S, io->Y, dhaz, dS, dY, ws + w.dlogits, B, K); count_launch();
  c.chk(cudaGetLastError(), "surv_head_bwd");
  const float* hpath = ws + w.pool[0].h;
  const float* homic = ws + w.pool[1].h;
  // gradients of the two rho outputs: [B,256] blocks of `dcat` (bilinear), or the two column halves of the [B,512]
  // concat gradient (row stride 512)
  const bool concat = m->fusion != MPO_FUSION_BILINEAR;
  float* dhpath = ws + w.dcat;
  float* dhomic = concat ? ws + w.dcat + E : ws + w.dcat + (long long)B * E;
  const long long lddh = concat ? 2 * E : E;
  if (concat) {
    // each data gradient carries the ReLU derivative of the layer in front in its epilogue
    DgradEpi e2; e2.y = ws + w.z2; e2.ld_y = E; e2.act = ACT_RELU;
    lin_bwd(c, ws + w.dlogits, K, ws + w.z2, E, m->classifier, K, E, ws + w.dz2, E, B, false, e2);
    DgradEpi e1; e1.y = ws + w.z1; e1.ld_y = E; e1.act = ACT_RELU;
    lin_bwd(c, ws + w.dz2, E, ws + w.z1, E, m->fusion2, E, E, ws + w.dz1, E, B, false, e1);
    const bool gated = m->fusion == MPO_FUSION_GATED_CONCAT;
    lin_bwd(c, ws + w.dz1, E, ws + (gated ? w.catg : w.cat), 2 * E, m->fusion0, E, 2 * E, ws + w.dcat, 2 * E, B, false);
    if (gated) {     // gradient at the gated items -> gradient at the rho outputs (in place), gate gradients when asked for
      launch_k(gate_concat_bwd_kernel, dim3(B), dim3(256), 0, c.st, ws + w.cat, ws + w.gateg, m->gate[0].w, m->gate[1].w,
               ws + w.dcat, m->gate[0].gw, m->gate[0].gb, m->gate[1].gw, m->gate[1].gb); count_launch();
      c.chk(cudaGetLastError(), "gate_concat_bwd");
    }
  } else {
    lin_bwd(c, ws + w.dlogits, K, ws + w.bf2, E, m->classifier, K, E, ws + w.dh, E, B, false);
    act_bwd(c, ws + w.dh, E, ws + w.bf2, E, ws + w.dz2, E, B, E, ACT_RELU, mk_drop(c, 0.25f, SITE_BIL + 4));
    lin_bwd(c, ws + w.dz2, E, ws + w.cat130, 130, m->bil.fc2, E, 130, ws + w.bdcat130, 130, B, false);
    // fc1 (its output sits in cat130[:, :64])
    act_bwd(c, ws + w.bdcat130, 130, ws + w.cat130, 130, ws + w.dz1, BMM, B, BMM, ACT_RELU, mk_drop(c, 0.25f, SITE_BIL + 3));
    lin_bwd(c, ws + w.dz1, BMM, ws + w.kp, 1089, m->bil.fc1, BMM, 1089, ws + w.bdkp, 1089, B, false);
    launch_k(bil_kron_bwd_kernel, dim3(B), dim3(64), 0, c.st, ws + w.bo[0], ws + w.bo[1], ws + w.bdkp, ws + w.bdcat130, ws + w.bdo[0],
                                            ws + w.bdo[1], mk_drop(c, 0.25f, SITE_BIL + 2)); count_launch();
    c.chk(cudaGetLastError(), "bil_kron_bwd");
    // side 1: xa = h_path, xb = h_omic ; side 2: xa = h_omic, xb = h_path
    cudaMemsetAsync(dhomic, 0, (size_t)B * E * 4, c.st);
    bil_side_bwd(c, m->bil.h1, m->bil.z1, m->bil.o1, w, ws, 0, hpath, homic, dhpath, false, dhomic, B);
    bil_side_bwd(c, m->bil.h2, m->bil.z2, m->bil.o2, w, ws, 1, homic, hpath, dhomic, true, dhpath, B);
  }
  br.main.chk(c.err, c.where);
  // from here the path and omic branches are independent: two streams, weight gradients on two more
  fork(br);
  Ctx& cp = br.main;
  Ctx& co = br.second;
  pool_bwd(co, m->omic_pool, w.pool[1], w, ws, ws + w.enc[3].y2, dhomic, lddh, ws + w.dhp[1], ws + w.dtok[1], B, 1);
  enc_bwd(co, m->omic_tr[1], w.enc[3], w, ws, ws + w.enc[2].y2, ws + w.dtok[1], ws + w.dmid[1], B, 3);
  enc_bwd(co, m->omic_tr[0], w.enc[2], w, ws, ws + w.G, ws + w.dmid[1], ws + w.dG, B, 2);
  float* dhc = ws + w.dhc;
  pool_bwd(cp, m->path_pool, w.pool[0], w, ws, ws + w.enc[1].y2, dhpath, lddh, ws + w.dhp[0], ws + w.dtok[0], B, 0);
  enc_bwd(cp, m->path_tr[1], w.enc[1], w, ws, ws + w.enc[0].y2, ws + w.dtok[0], ws + w.dmid[0], B, 1);
  enc_bwd(cp, m->path_tr[0], w.enc[0], w, ws, ws + w.hc, ws + w.dmid[0], dhc, B, 0);
  // output and value projections back to the pooled vectors
  join_w(cp);
  lin_bwd(cp, dhc, E, ws + w.v, E, m->coattn_out, E, E, ws + w.dv, E, R, false);
  if (io->suma == nullptr) {
    lin_bwd(cp, ws + w.dv, E, io->pooled, E, sub(m->coattn_in, 2 * E, E), E, E, io->dpooled, E, R, false);
  } else {
    if (!io->dsuma) return fail(MPO_E_ARG, "%s", "mpo_tail_post_bwd: dsuma is NULL although suma is given");
    mpo_lin Lv = sub(m->coattn_in, 2 * E, E);
    float* gbv = Lv.gb;
    const float* bv = Lv.b;
    Lv.gb = nullptr;
    lin_bwd(cp, ws + w.dv, E, io->pooled, E, Lv, E, E, io->dpooled, E, R, false);
    // dsuma[r] = dv[r] . b_v ;  db_v[e] += sum_r suma[r] dv[r][e]
    GemmArgs g1{ws + w.dv, E, 1, bv, 1, 0, io->dsuma, 1, nullptr, R, 1, E, 1.f, 0, ACT_NONE, nullptr};
    cp.chk(launch_gemm(g1, cp.st), "v.dsuma");
    if (gbv) {
      GemmArgs g2{io->suma, 0, 1, ws + w.dv, E, 1, gbv, E, nullptr, 1, E, R, 1.f, 1, ACT_NONE, nullptr};
      cp.chk(launch_gemm(g2, cp.st), "v.dbv");
    }
  }
  join(br);
  if (m->variant == MPO_VARIANT_NACAGAT) {   // needs dG of the omic branch: after the join
    cag_bwd(br.main, m->cag, w, ws, ws + w.G, io->qp, dhc, ws + w.dG, ws + w.dqp, R);
  }
  Ctx& cfin = br.main;
  return finish(cfin);
}

int mpo_tail_post_step(const mpo_model* m, const mpo_tail_io* io, int32_t kind, const int64_t* label, const float* censor,
                       float alpha, float eps, float grad_scale, float* loss, float* dhaz, float* dS, int32_t flags,
                       void* stream) {
  int rc = check_model(m, io, "mpo_tail_post_step");
  if (rc) return rc;
  if (kind != MPO_LOSS_NLL && kind != MPO_LOSS_CES) return fail(MPO_E_UNSUPPORTED, "%s", "mpo_tail_post_step: unknown loss kind");
  if (!io->pooled || !io->dpooled || !io->hazards || !io->S || !io->Y || !io->att_path || !io->att_omic || !label ||
      !censor || !loss || !dhaz || !dS)
    return fail(MPO_E_ARG, "%s", "mpo_tail_post_step: NULL pointer");
  if (fused::eligible(m, io)) {
    Ws w; build_layout(m, io->num_slides, w);
    fused::LossArgs la{kind, label, censor, alpha, eps, grad_scale, loss, dhaz, dS};
    return fused::post(m, io, w, fused::F_FWD | fused::F_LOSS | fused::F_BWD, &la, nullptr, nullptr, nullptr,
                       static_cast<cudaStream_t>(stream), /*side_wgrad=*/(flags & MPO_POST_STEP_INLINE_WGRAD) == 0);
  }
  if ((rc = mpo_tail_post_fwd(m, io, stream))) return rc;
  if ((rc = mpo_surv_loss(kind, io->hazards, io->S, label, censor, alpha, eps, grad_scale, loss, dhaz, dS, io->num_slides,
                          m->n_classes, stream)))
    return rc;
  return mpo_tail_post_bwd(m, io, dhaz, dS, nullptr, stream);
}

int mpo_tail_side_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int32_t* step_dev, int32_t zero_grad, void* stream) {
  return fused::side_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step_dev,
                          (zero_grad & 1) != 0, static_cast<cudaStream_t>(stream));
}

int mpo_tail_pre_bwd(const mpo_model* m, const mpo_tail_io* io, void* stream) {
  int rc = check_model(m, io, "mpo_tail_pre_bwd");
  if (rc) return rc;
  if (!io->dqk || !io->qp) return fail(MPO_E_ARG, "%s", "mpo_tail_pre_bwd: NULL pointer");
  const int B = io->num_slides, R = 6 * B;
  Ws w; build_layout(m, B, w);
  float* ws = io->ws;
  if (fused::eligible(m, io)) return fused::pre_bwd(m, io, w, static_cast<cudaStream_t>(stream));
  Branches br = make_branches(static_cast<cudaStream_t>(stream), true, io);
  Ctx c = br.main;             // query / fold part: weight gradients on the side stream (joined at the end)
  const bool nac = m->variant == MPO_VARIANT_NACAGAT;
  if (nac && (!io->dkc || !io->dtq)) return fail(MPO_E_ARG, "%s", "mpo_tail_pre_bwd: dkc/dtq are NULL (NaCAGaT)");
  const float* Wk = m->coattn_in.w + (long long)E * E;
  float* gWk = m->coattn_in.gw ? m->coattn_in.gw + (long long)E * E : nullptr;
  // key fold backward: dq[r][e] (+)= sum_d dqk[r][d] W_k[e][d] / 16 ; dW_k[e][d] += sum_r q[r][e] dqk[r][d] / 16
  {
    GemmArgs g{io->dqk, E, 1, Wk, 1, E, ws + w.dqp, E, nullptr, R, E, E, 1.f / 16.f, nac ? 1 : 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g, c.st), "fold.dq");
    if (gWk) {
      GemmArgs g2{io->qp, 1, E, io->dqk, E, 1, gWk, E, nullptr, E, E, R, 1.f / 16.f, 1, ACT_NONE, nullptr};
      if (c.async_w) dep(c, c.st, c.wst);
      c.chk(launch_gemm(g2, c.async_w ? c.wst : c.st), "fold.dWk");
    }
  }
  if (nac) {
    const float* bk = m->coattn_in.b + E;
    float* gbk = m->coattn_in.gb ? m->coattn_in.gb + E : nullptr;
    // kc = q . b_k / 16 :  dq += dkc b_k / 16 ; db_k += sum_r dkc[r] q[r] / 16
    GemmArgs g{io->dkc, 1, 0, bk, 0, 1, ws + w.dqp, E, nullptr, R, E, 1, 1.f / 16.f, 1, ACT_NONE, nullptr};
    c.chk(launch_gemm(g, c.st), "kc.dq");
    if (gbk) {
      GemmArgs g2{io->dkc, 0, 1, io->qp, E, 1, gbk, E, nullptr, 1, E, R, 1.f / 16.f, 1, ACT_NONE, nullptr};
      if (c.async_w) dep(c, c.st, c.wst);
      c.chk(launch_gemm(g2, c.async_w ? c.wst : c.st), "kc.dbk");
    }
    // tanh(q) branch of the pre-gate: dq += dtq * (1 - tanh(q)^2)
    act(c, io->qp, ws + w.s256a[0], (long long)R * E, ACT_TANH);
    act_bwd(c, io->dtq, E, ws + w.s256a[0], E, ws + w.s256b[0], E, R, E, ACT_TANH);
    add(c, ws + w.dqp, ws + w.s256b[0], ws + w.dqp, (long long)R * E);
  }
  // query in-projection
  lin_bwd(c, ws + w.dqp, E, ws + w.G, E, sub(m->coattn_in, 0, E), E, E, ws + w.dG, E, R, true);
  br.main.chk(c.err, c.where);
  // SNN encoders: six independent chains on six branches
  fork(br);
  Ctx chain[MPO_Q];
  snn_branches(br, chain);
  for (int i = 0; i < MPO_Q; ++i) {
    Ctx& ci = chain[i];
    const int d = m->omic_dims[i];
    float* dz2 = ws + w.snn_dz2[i];   // [B,256] each, private to the chain
    float* dz1 = ws + w.snn_dz1[i];
    act_bwd(ci, ws + w.dG + i * E, 6 * E, ws + w.G + i * E, 6 * E, dz2, E, B, E, ACT_ELU,
            mk_drop(ci, ci.drop_p, SITE_SNN + 2 * i + 1, true));
    // second linear's data gradient with the first layer's ELU + AlphaDropout derivative in its epilogue
    DgradEpi e; e.y = ws + w.snn_h[i]; e.ld_y = E; e.act = ACT_ELU; e.drop = mk_drop(ci, ci.drop_p, SITE_SNN + 2 * i, true);
    lin_bwd(ci, dz2, E, ws + w.snn_h[i], E, m->snn[i][1], E, E, dz1, E, B, false, e);
    lin_bwd(ci, dz1, E, io->omics[i], d, m->snn[i][0], E, d, nullptr, 0, B, false);
  }
  snn_join(br, chain);
  join(br);
  return finish(br.main);
}

// ------------------------------------------------------------------------------------------------ GE-NaCAGaT
int64_t mpo_ge_ws_floats(int64_t N) {
  if (N <= 0) return -1;
  GeWs w; ge_layout(N, w);
  return w.total;
}

static int check_ge(const mpo_ge_model* m, int64_t N, const char* who) {
  if (!m) return fail(MPO_E_ARG, "%s: NULL model", who);
  if (N <= 0 || N > 46340) return fail(MPO_E_ARG, "%s: N out of range", who);
  if (m->n_classes < 1 || m->n_classes > 16) return fail(MPO_E_ARG, "%s: n_classes out of range", who);
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s: no CUDA device (this library has no CPU fallback)", who);
  return MPO_OK;
}

int mpo_ge_fwd(const mpo_ge_model* m, int64_t N64, const void* h_hi, const void* h_lo, float* ws, float* attn,
               float* path, float* Y, float drop_p, uint32_t seed, int32_t train, void* stream) {
  int rc = check_ge(m, N64, "mpo_ge_fwd");
  if (rc) return rc;
  if (!h_hi || !h_lo || !ws || !attn || !path || !Y) return fail(MPO_E_ARG, "%s", "mpo_ge_fwd: NULL pointer");
  const int N = static_cast<int>(N64);
  GeWs w; ge_layout(N, w);
  Ctx c; c.st = static_cast<cudaStream_t>(stream);
  c.train = train != 0; c.drop_p = drop_p; c.seed = seed;
  const long long n = (long long)N * E;
  launch_k(ge_h_kernel, dim3(nblk(n)), dim3(256), 0, c.st, static_cast<const __half*>(h_hi), static_cast<const __half*>(h_lo),
           ws + w.H, n); count_launch();
  // self-attention over the patches, one head (ge_nacagat.py:27,49); the averaged map IS the head's map
  lin_fwd(c, ws + w.H, E, m->sa_in, 3 * E, E, ws + w.sa_qkv, 3 * E, N, ACT_NONE);
  ge_attn_fwd(c, w, ws, ws + w.sa_qkv, attn, ws + w.sa_ctx, N, 1);
  lin_fwd(c, ws + w.sa_ctx, E, m->sa_out, E, E, ws + w.sa_out, E, N, ACT_NONE);
  // encoder over the N tokens (ge_nacagat.py:30-32,53)
  ge_enc_fwd(c, m->tr[0], w.enc[0], w, ws, ws + w.sa_out, N, 0);
  ge_enc_fwd(c, m->tr[1], w.enc[1], w, ws, ws + w.enc[0].y2, N, 1);
  const float* x = ws + w.enc[1].y2;
  // gated attention pooling over N (blocks.py:42-48, ge_nacagat.py:56-60); p = 0.25 hard-coded in the head (blocks.py:34-36)
  lin_fwd(c, x, E, m->pool.att_a, E, E, ws + w.pa, E, N, ACT_TANH, mk_drop(c, 0.25f, SITE_POOL));
  lin_fwd(c, x, E, m->pool.att_b, E, E, ws + w.pb, E, N, ACT_SIGMOID, mk_drop(c, 0.25f, SITE_POOL + 1));
  mul(c, ws + w.pa, ws + w.pb, ws + w.pab, n);
  lin_fwd(c, ws + w.pab, E, m->pool.att_c, 1, E, path, 1, N, ACT_NONE);                 // raw logits A [N] (returned)
  c.chk(cudaMemcpyAsync(ws + w.pw, path, (size_t)N * 4, cudaMemcpyDeviceToDevice, c.st), "ge.copy");
  launch_k(row_softmax_kernel, dim3(1), dim3(256), 0, c.st, ws + w.pw, (long long)N, N); count_launch();
  { GemmArgs g{ws + w.pw, N, 1, x, E, 1, ws + w.hp, E, nullptr, 1, E, N, 1.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g, c.st), "ge.pool"); }
  lin_fwd(c, ws + w.hp, E, m->pool.rho, E, E, ws + w.h, E, 1, ACT_RELU, mk_drop(c, c.drop_p, SITE_RHO));
  lin_fwd(c, ws + w.h, E, m->classifier, m->n_classes, E, ws + w.logits, m->n_classes, 1, ACT_NONE);
  c.chk(cudaMemcpyAsync(Y, ws + w.logits, (size_t)m->n_classes * 4, cudaMemcpyDeviceToDevice, c.st), "ge.copyY");
  launch_k(row_softmax_kernel, dim3(1), dim3(256), 0, c.st, Y, (long long)m->n_classes, m->n_classes); count_launch();
  c.chk(cudaGetLastError(), "mpo_ge_fwd");
  return finish(c);
}

int mpo_ge_ce_loss(const float* Y, const int64_t* label, int32_t n_classes, float grad_scale, float* loss, float* dY,
                   void* stream) {
  if (!Y || !label || !loss || !dY || n_classes < 1) return fail(MPO_E_ARG, "%s", "mpo_ge_ce_loss: bad arguments");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_ge_ce_loss: no CUDA device (this library has no CPU fallback)");
  launch_k(ge_ce_loss_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream), Y,
           reinterpret_cast<const long long*>(label), n_classes, grad_scale, loss, dY); count_launch();
  return check_cuda(cudaGetLastError(), "ge_ce_loss_kernel");
}

int mpo_ge_bwd(const mpo_ge_model* m, const mpo_bag* bag, const void* h_hi, float* ws, const float* attn, const float* path,
               const float* Y, const float* dY, void* dz_ws, float keep_scale, float drop_p, uint32_t seed, int32_t train,
               void* stream) {
  if (!bag) return fail(MPO_E_ARG, "%s", "mpo_ge_bwd: bag is NULL");
  int rc = check_ge(m, bag->total_rows, "mpo_ge_bwd");
  if (rc) return rc;
  if (!h_hi || !ws || !attn || !path || !Y || !dY || !dz_ws || !bag->x || bag->num_slides != 1)
    return fail(MPO_E_ARG, "%s", "mpo_ge_bwd: NULL pointer (or more than one slide)");
  if (!m->H.gw || !m->H.gb) return fail(MPO_E_ARG, "%s", "mpo_ge_bwd: the model carries no gradient buffers");
  const int N = static_cast<int>(bag->total_rows);
  GeWs w; ge_layout(N, w);
  Ctx c; c.st = static_cast<cudaStream_t>(stream);
  c.train = train != 0; c.drop_p = drop_p; c.seed = seed;
  if (ge_use_tc() && N >= 1024) { c.tc_dz = ws + w.tcWdz; c.tc_x = ws + w.tcWx; c.tc_rows = static_cast<int>(N); }
  const long long n = (long long)N * E;
  const int K = m->n_classes;
  const float* x = ws + w.enc[1].y2;
  // Y = softmax(logits)
  c.chk(cudaMemcpyAsync(ws + w.dlogits, dY, (size_t)K * 4, cudaMemcpyDeviceToDevice, c.st), "ge.copy");
  launch_k(row_softmax_bwd_kernel, dim3(1), dim3(256), 0, c.st, Y, (long long)K, ws + w.dlogits, (long long)K, K, 1.f); count_launch();
  lin_bwd(c, ws + w.dlogits, K, ws + w.h, E, m->classifier, K, E, ws + w.dh, E, 1, false);
  act_bwd(c, ws + w.dh, E, ws + w.h, E, ws + w.dzr, E, 1, E, ACT_RELU, mk_drop(c, c.drop_p, SITE_RHO));
  lin_bwd(c, ws + w.dzr, E, ws + w.hp, E, m->pool.rho, E, E, ws + w.dhp, E, 1, false);
  // pooled = softmax(A)^T x : dw = dhp x^T [N], dx = w dhp [N,E]
  { GemmArgs g{ws + w.dhp, E, 1, x, 1, E, ws + w.dw, N, nullptr, 1, N, E, 1.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g, c.st), "ge.dw"); }
  { GemmArgs g{ws + w.pw, 1, 0, ws + w.dhp, 0, 1, ws + w.dx, E, nullptr, N, E, 1, 1.f, 0, ACT_NONE, nullptr};
    c.chk(launch_gemm(g, c.st), "ge.dx"); }
  launch_k(row_softmax_bwd_kernel, dim3(1), dim3(256), 0, c.st, ws + w.pw, (long long)N, ws + w.dw, (long long)N, N, 1.f); count_launch();
  // A = (a * b) w_c + b_c
  lin_bwd(c, ws + w.dw, 1, ws + w.pab, E, m->pool.att_c, 1, E, ws + w.dab, E, N, false);
  mul(c, ws + w.dab, ws + w.pb, ws + w.t0, n);
  act_bwd(c, ws + w.t0, E, ws + w.pa, E, ws + w.t0, E, N, E, ACT_TANH, mk_drop(c, 0.25f, SITE_POOL));
  lin_bwd(c, ws + w.t0, E, x, E, m->pool.att_a, E, E, ws + w.dx, E, N, true);
  mul(c, ws + w.dab, ws + w.pa, ws + w.t0, n);
  act_bwd(c, ws + w.t0, E, ws + w.pb, E, ws + w.t0, E, N, E, ACT_SIGMOID, mk_drop(c, 0.25f, SITE_POOL + 1));
  lin_bwd(c, ws + w.t0, E, x, E, m->pool.att_b, E, E, ws + w.dx, E, N, true);
  // encoder
  ge_enc_bwd(c, m->tr[1], w.enc[1], w, ws, ws + w.enc[0].y2, ws + w.dx, ws + w.dmid, N, 1);
  ge_enc_bwd(c, m->tr[0], w.enc[0], w, ws, ws + w.sa_out, ws + w.dmid, ws + w.dx, N, 0);
  // self-attention
  lin_bwd(c, ws + w.dx, E, ws + w.sa_ctx, E, m->sa_out, E, E, ws + w.dctx, E, N, false);
  ge_attn_bwd(c, w, ws, ws + w.sa_qkv, attn, ws + w.dctx, ws + w.dP, ws + w.dqkv, N, 1);
  lin_bwd(c, ws + w.dqkv, 3 * E, ws + w.H, E, m->sa_in, 3 * E, E, ws + w.dH, E, N, false);
  // H projection: dz = dH * mask; db_H += colsum(dz); dW_H += dz^T X on the tensor cores
  launch_k(ge_dz_kernel, dim3(nblk(n)), dim3(256), 0, c.st, ws + w.dH, static_cast<const __half*>(h_hi), keep_scale,
           ws + w.dzf, static_cast<__nv_bfloat16*>(dz_ws), n); count_launch();
  colsum(c.st, ws + w.dzf, (long long)E, nullptr, (long long)0, m->H.gb, N, E);
  c.chk(cudaGetLastError(), "mpo_ge_bwd");
  rc = finish(c);
  if (rc) return rc;
  CUtensorMap tm_dz, tm_x;
  if ((rc = make_tmap_bf16_2d(&tm_dz, dz_ws, static_cast<uint64_t>(N), kD, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tm_x, bag->x, static_cast<uint64_t>(N), kDIn, 64, 64))) return rc;
  return check_cuda(launch_bag_bwd_dw(tm_dz, tm_x, m->H.gw, N, kDIn, kDIn, false, nullptr, num_sms(), c.st),
                    "bag_bwd_dw_kernel (GE)");
}


// ------------------------------------------------------------------------------------------------ stand-alone operators
// The reference's building blocks can also be called on their own (its unit tests do: models/blocks.py:304-325,
// models/fusion.py:116-170).  On this path they run fused inside a slide pass; these entry points expose the same device
// kernels one operator at a time, so that the stand-alone `forward` of the block / fusion classes is CUDA as well
// (inference: no gradients are produced).
static DropSpec op_drop(float p, uint32_t seed, uint32_t site) {
  Ctx c; c.train = p > 0.f; c.seed = seed;
  return mk_drop(c, p, site);
}
int mpo_op_linear(const float* x, int64_t ldx, const float* w, const float* b, float* y, int64_t ldy, int32_t rows,
                  int32_t in, int32_t out, int32_t act, float drop_p, uint32_t seed, uint32_t site, void* stream) {
  if (!x || !w || !y || rows < 0 || in <= 0 || out <= 0 || act < ACT_NONE || act > ACT_SIGMOID)
    return fail(MPO_E_ARG, "%s", "mpo_op_linear: bad arguments");
  if (num_sms() <= 0) return fail(MPO_E_CUDA, "%s", "mpo_op_linear: no CUDA device (this library has no CPU fallback)");
  GemmArgs g{x, ldx, 1, w, 1, in, y, ldy, b, rows, out, in, 1.f, 0, act, nullptr, 1, op_drop(drop_p, seed, site)};
  return check_cuda(launch_gemm(g, static_cast<cudaStream_t>(stream)), "mpo_op_linear");
}
int mpo_op_layernorm(const float* x, const float* gamma, const float* beta, float* y, int32_t rows, int32_t cols, float eps,
                     void* stream) {
  if (!x || !gamma || !beta || !y || rows < 0 || cols <= 0) return fail(MPO_E_ARG, "%s", "mpo_op_layernorm: bad arguments");
  if (rows == 0) return MPO_OK;
  op_layernorm_kernel<<<nblk(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, y, rows, cols, eps);
  count_launch();
  return check_cuda(cudaGetLastError(), "mpo_op_layernorm");
}
int mpo_op_ewise(int32_t op, const float* a, const float* b, float* y, int64_t n, void* stream) {
  if (!a || !y || n < 0 || ((op == MPO_OP_ADD || op == MPO_OP_MUL) && !b)) return fail(MPO_E_ARG, "%s", "mpo_op_ewise: bad arguments");
  if (n == 0) return MPO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (op == MPO_OP_ADD) add_kernel<<<nblk(n), 256, 0, st>>>(a, b, y, n);
  else if (op == MPO_OP_MUL) mul_kernel<<<nblk(n), 256, 0, st>>>(a, b, y, n);
  else if (op >= MPO_OP_ACT && op <= MPO_OP_ACT + ACT_SIGMOID) act_kernel<<<nblk(n), 256, 0, st>>>(a, y, n, op - MPO_OP_ACT);
  else return fail(MPO_E_ARG, "%s", "mpo_op_ewise: unknown operator");
  count_launch();
  return check_cuda(cudaGetLastError(), "mpo_op_ewise");
}
int mpo_op_rowscale(const float* x, const float* g, float* y, int32_t rows, int32_t cols, void* stream) {
  if (!x || !g || !y || rows < 0 || cols <= 0) return fail(MPO_E_ARG, "%s", "mpo_op_rowscale: bad arguments");
  if (rows == 0) return MPO_OK;
  op_rowscale_kernel<<<nblk(static_cast<long long>(rows) * cols), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, g, y, rows, cols);
  count_launch();
  return check_cuda(cudaGetLastError(), "mpo_op_rowscale");
}
int mpo_op_dropout(const float* x, float* y, int64_t n, float p, uint32_t seed, uint32_t site, void* stream) {
  if (!x || !y || n < 0) return fail(MPO_E_ARG, "%s", "mpo_op_dropout: bad arguments");
  if (n == 0) return MPO_OK;
  probs_dropout_kernel<<<nblk(n, 1024), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, 0u, op_drop(p, seed, site));
  count_launch();
  return check_cuda(cudaGetLastError(), "mpo_op_dropout");
}
int mpo_op_bil_gate(const float* x1, const float* U, const float* bias, const float* h, float* g, float* gh, int32_t rows,
                    void* stream) {
  if (!x1 || !U || !bias || !h || !g || !gh || rows < 0) return fail(MPO_E_ARG, "%s", "mpo_op_bil_gate: bad arguments");
  if (rows == 0) return MPO_OK;
  bil_gate_fwd_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(x1, U, bias, h, g, gh);
  count_launch();
  return check_cuda(cudaGetLastError(), "mpo_op_bil_gate");
}
int mpo_op_bil_kron(const float* o1, const float* o2, float* kp, float* cat, int32_t rows, float drop_p, uint32_t seed,
                    uint32_t site, void* stream) {
  if (!o1 || !o2 || !kp || !cat || rows < 0) return fail(MPO_E_ARG, "%s", "mpo_op_bil_kron: bad arguments");
  if (rows == 0) return MPO_OK;
  bil_kron_fwd_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(o1, o2, kp, cat, op_drop(drop_p, seed, site));
  count_launch();
  return check_cuda(cudaGetLastError(), "mpo_op_bil_kron");
}

}  // extern "C"
