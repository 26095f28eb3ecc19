// Workspace layout of the slide tail, shared by the per-op orchestration (tail.cu) and the fused cluster kernels
// (tail_fused.cu).  Every intermediate is a named, 256-byte-aligned slice of one caller-owned fp32 buffer.
#pragma once
#include <cstdio>
#include <string>
#include <utility>
#include <vector>
#include "../../include/mpo_b200.h"

namespace mpo {
namespace tailws {

constexpr int E = 256;       // model width
constexpr int FF = 512;      // transformer feed-forward width
constexpr int BH = 32;       // bilinear hidden
constexpr int BMM = 64;      // bilinear mm_hidden

// ---------------------------------------------------------------------------------------------- workspace layout
struct Layout {
  std::vector<std::pair<std::string, std::pair<long long, long long>>> items;
  long long total = 0;
  long long add(const char* name, long long len) {
    const long long off = total;
    items.push_back({name, {off, len}});
    total += (len + 63) / 64 * 64;   // keep every buffer 256-byte aligned
    return off;
  }
};

struct EncBuf { long long qkv, probs, ctx, sa, y1, xh1, rs1, f, f2, y2, xh2, rs2; };
struct PoolBuf { long long a, b, w, hp, h; long long hld; };   // h: rho output [B][hld] (a view into `cat` for concat fusion)

struct Ws {
  long long snn_h[MPO_Q], G, v, hc;
  long long cag_f1, cag_f2, cag_f3, cag_s, cag_u, cag_Gg, cag_Gxh, cag_Grs, cag_w, cag_Ee, cag_Exh, cag_Ers, cag_m, cag_C;
  EncBuf enc[4];      // path.0, path.1, omic.0, omic.1
  PoolBuf pool[2];    // path, omic
  long long cat, z1, z2;                                   // concat fusion
  long long catg = 0, gateg = 0;                            // gated concat: gated copy of `cat` [B][512], gate values [B][2]
  long long bh[2], bU[2], bg[2], bgh[2], bo[2], kp, cat130, bf2;   // bilinear fusion
  long long logits;
  // gradients / scratch
  long long dlogits, dh, dcat, dz1, dz2, dhp[2], dtok[2], dG, dqp, dhc, dv;
  // per-branch scratch (0 = path / main stream, 1 = omic / second stream): the branches run concurrently
  long long dxa[2], dxb[2], dzr[2], dmid[2], s768[2], s512[2], s256a[2], s256b[2], s256c[2], dmk1[2], dmk2[2];
  long long snn_dz1[MPO_Q], snn_dz2[MPO_Q], snn_dh[MPO_Q];
  long long bV, bdkp, bdcat130, bdo[2], bdgh[2], bdh[2], bdz[2], bdx[2];
  // fused cluster tail (tail_fused.cu): the gradient at every linear layer's pre-activation and at every LayerNorm
  // output keeps its own buffer until the grouped weight-gradient kernel has consumed it
  struct FzEnc { long long dy2, df2, df, dy1, dsa, dqkv; } fz_enc[4];
  struct FzPool { long long dzr, da, db; } fz_pool[2];
  long long fz_cag[6];     // NaCAGaT CAG: t0 (d fc_c pre-activation), dGg, dEe, df1, df2, df3
  long long fz_dG2 = 0;    // NaCAGaT CAG: dQ, added to dG by the pre-backward kernel
  Layout lay;
};

inline void build_layout(const mpo_model* m, int B, Ws& w) {
  const long long R = 6LL * B;
  Layout& L = w.lay;
  char nm[64];
  for (int i = 0; i < MPO_Q; ++i) { snprintf(nm, sizeof nm, "snn_h%d", i); w.snn_h[i] = L.add(nm, (long long)B * E); }
  w.G = L.add("G", R * E);
  w.v = L.add("v", R * E);
  w.hc = L.add("hc", R * E);
  if (m->variant == MPO_VARIANT_NACAGAT) {
    w.cag_f1 = L.add("cag_f1", R * E); w.cag_f2 = L.add("cag_f2", R * E); w.cag_f3 = L.add("cag_f3", R * E);
    w.cag_s = L.add("cag_s", R * E); w.cag_u = L.add("cag_u", R * E); w.cag_Gg = L.add("cag_Gg", R * E);
    w.cag_Gxh = L.add("cag_Gxh", R * E); w.cag_Grs = L.add("cag_Grs", R);
    w.cag_w = L.add("cag_w", R * E); w.cag_Ee = L.add("cag_Ee", R * E); w.cag_Exh = L.add("cag_Exh", R * E);
    w.cag_Ers = L.add("cag_Ers", R); w.cag_m = L.add("cag_m", R * E); w.cag_C = L.add("cag_C", R * E);
  }
  const char* en[4] = {"path0", "path1", "omic0", "omic1"};
  for (int e = 0; e < 4; ++e) {
    EncBuf& b = w.enc[e];
    auto A = [&](const char* s, long long n) { snprintf(nm, sizeof nm, "%s_%s", en[e], s); return L.add(nm, n); };
    b.qkv = A("qkv", R * 3 * E); b.probs = A("probs", (long long)B * 8 * 36); b.ctx = A("ctx", R * E);
    b.sa = A("sa", R * E); b.y1 = A("y1", R * E); b.xh1 = A("xh1", R * E); b.rs1 = A("rs1", R);
    b.f = A("f", R * FF); b.f2 = A("f2", R * E); b.y2 = A("y2", R * E); b.xh2 = A("xh2", R * E); b.rs2 = A("rs2", R);
  }
  const char* pn[2] = {"pathpool", "omicpool"};
  // concat fusion reads [h_path | h_omic] as one [B, 512] row block: the two rho outputs are written straight into it
  if (m->fusion != MPO_FUSION_BILINEAR) w.cat = L.add("cat", (long long)B * 2 * E);
  for (int p = 0; p < 2; ++p) {
    PoolBuf& b = w.pool[p];
    auto A = [&](const char* s, long long n) { snprintf(nm, sizeof nm, "%s_%s", pn[p], s); return L.add(nm, n); };
    b.a = A("a", R * E); b.b = A("b", R * E); b.w = A("w", (long long)B * 6); b.hp = A("hp", (long long)B * E);
    if (m->fusion != MPO_FUSION_BILINEAR) { b.h = w.cat + p * E; b.hld = 2 * E; }
    else { b.h = A("h", (long long)B * E); b.hld = E; }
  }
  if (m->fusion != MPO_FUSION_BILINEAR) {
    w.z1 = L.add("z1", (long long)B * E); w.z2 = L.add("z2", (long long)B * E);
    if (m->fusion == MPO_FUSION_GATED_CONCAT) {
      w.catg = L.add("catg", (long long)B * 2 * E); w.gateg = L.add("gateg", (long long)B * 2);
    }
  } else {
    for (int s = 0; s < 2; ++s) {
      auto A = [&](const char* t, long long n) { snprintf(nm, sizeof nm, "bil%d_%s", s + 1, t); return L.add(nm, n); };
      w.bh[s] = A("h", (long long)B * BH); w.bU[s] = A("U", (long long)B * BH * E); w.bg[s] = A("g", (long long)B * BH);
      w.bgh[s] = A("gh", (long long)B * BH); w.bo[s] = A("o", (long long)B * BH);
      w.bdo[s] = A("do", (long long)B * BH); w.bdgh[s] = A("dgh", (long long)B * BH); w.bdh[s] = A("dh", (long long)B * BH);
      w.bdz[s] = A("dz", (long long)B * BH); w.bdx[s] = A("dx", (long long)B * E);
    }
    w.kp = L.add("kp", (long long)B * 1089); w.cat130 = L.add("cat130", (long long)B * 130);
    w.bf2 = L.add("bf2", (long long)B * E);
    w.bV = L.add("bV", (long long)B * BH * E); w.bdkp = L.add("bdkp", (long long)B * 1089);
    w.bdcat130 = L.add("bdcat130", (long long)B * 130);
  }
  w.logits = L.add("logits", (long long)B * m->n_classes);
  w.dlogits = L.add("dlogits", (long long)B * m->n_classes);
  w.dh = L.add("dh", (long long)B * E);
  w.dcat = L.add("dcat", (long long)B * 2 * E); w.dz1 = L.add("dz1", (long long)B * E); w.dz2 = L.add("dz2", (long long)B * E);
  w.dhp[0] = L.add("dhp_path", (long long)B * E); w.dhp[1] = L.add("dhp_omic", (long long)B * E);
  w.dtok[0] = L.add("dtok_path", R * E); w.dtok[1] = L.add("dtok_omic", R * E);
  for (int s = 0; s < 2; ++s) {
    auto A = [&](const char* t, long long n) { snprintf(nm, sizeof nm, "scr%d_%s", s, t); return L.add(nm, n); };
    w.dxa[s] = A("dxa", R * E); w.dxb[s] = A("dxb", R * E); w.dzr[s] = A("dzr", (long long)B * E);
    w.dmid[s] = A("dmid", R * E); w.s768[s] = A("s768", R * 3 * E); w.s512[s] = A("s512", R * FF);
    w.s256a[s] = A("s256a", R * E); w.s256b[s] = A("s256b", R * E); w.s256c[s] = A("s256c", R * E);
    w.dmk1[s] = A("dmk1", R * E); w.dmk2[s] = A("dmk2", R * E);
  }
  for (int i = 0; i < MPO_Q; ++i) {
    snprintf(nm, sizeof nm, "snn_dz1_%d", i); w.snn_dz1[i] = L.add(nm, (long long)B * E);
    snprintf(nm, sizeof nm, "snn_dz2_%d", i); w.snn_dz2[i] = L.add(nm, (long long)B * E);
    snprintf(nm, sizeof nm, "snn_dh_%d", i); w.snn_dh[i] = L.add(nm, (long long)B * E);
  }
  w.dG = L.add("dG", R * E); w.dqp = L.add("dqp", R * E); w.dhc = L.add("dhc", R * E); w.dv = L.add("dv", R * E);
  for (int e = 0; e < 4; ++e) {
    auto A = [&](const char* s, long long n) { snprintf(nm, sizeof nm, "fz_%s_%s", en[e], s); return L.add(nm, n); };
    w.fz_enc[e].dy2 = A("dy2", R * E); w.fz_enc[e].df2 = A("df2", R * E); w.fz_enc[e].df = A("df", R * FF);
    w.fz_enc[e].dy1 = A("dy1", R * E); w.fz_enc[e].dsa = A("dsa", R * E); w.fz_enc[e].dqkv = A("dqkv", R * 3 * E);
  }
  if (m->variant == MPO_VARIANT_NACAGAT) {
    const char* cn[6] = {"t0", "dGg", "dEe", "df1", "df2", "df3"};
    for (int i = 0; i < 6; ++i) { snprintf(nm, sizeof nm, "fz_cag_%s", cn[i]); w.fz_cag[i] = L.add(nm, R * E); }
    w.fz_dG2 = L.add("fz_cag_dG2", R * E);
  }
  for (int p = 0; p < 2; ++p) {
    auto A = [&](const char* s, long long n) { snprintf(nm, sizeof nm, "fz_%s_%s", pn[p], s); return L.add(nm, n); };
    w.fz_pool[p].dzr = A("dzr", (long long)B * E); w.fz_pool[p].da = A("da", R * E); w.fz_pool[p].db = A("db", R * E);
  }
}

}  // namespace tailws
}  // namespace mpo
