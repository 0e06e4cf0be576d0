// Shared-memory plan and per-step record of the uniform-kernel path (host sizing + device carving).
#pragma once
#include "gpmpc_common.cuh"
#include "gpmpc_internal.h"

namespace gpmpc {

struct UniLayout {
  int rec, tail, tlen, rhot, pre, prelen, out, nOut, part, partlen, wp, wplen, S, cst, cstw, colred;
  int m, s, mu, A, Q, misc, M, V, acc, accN, am, r, rv, ints, small2, total;
};

// bwd=false: forward kernel; bwd=true: reverse-sweep kernel (its row / column sums live in a global per-CTA scratch)
// premat (reverse sweep only): per-step small matrices + stage-cost adjoints precomputed for all H steps in parallel
HD UniLayout make_uni_layout(int EV, bool bwd, int NP, int DP, int D, int H, int Na, bool premat = false) {
  UniLayout L;
  const int E = EV, P = E * (E + 1) / 2;
  int o = 0;
  // hot-loop record per training point j: { nu_j[0..EV), kap_j, beta_j[0..E), e_j } -- 2 EV + 2 doubles, a compile-time
  // stride, so that every lane fetches it with (EV + 1) broadcast LDS.128 at immediate offsets from one running pointer.
  // The action/time dims of nu, needed only by the O(N) reductions, live in a separate array (tlen per point).
  L.rhot = 2 * EV + 2;
  L.tlen = (D - EV + 1) & ~1;
  L.rec = o; o += NP * L.rhot;
  L.tail = o; o += NP * L.tlen;
  L.prelen = (4 * EV * EV + 2 + EV + Na + 1) & ~1;   // A, Q, Rinv, dS (E x E each), c, detR, dmu (E), da (Na)
  L.pre = o; if (bwd && premat) o += H * L.prelen;
  L.nOut = 1 + D;
  L.out = o; o += E * L.nOut;
  // per-warp accumulator rows of the forward sweep (P + 1 sums, padded to the 16-wide halving reduction)
  L.partlen = ((P + 1) + 15) & ~15;
  if (UNI_USE_MMA8(EV)) L.partlen += 64;   // + the warp's E x E scratch of the tensor-core sweep (uni_fwd_item_mma8)
  const int MW = UNI_MAXT(EV) / 32;   // warps per CTA at most
  L.part = o; if (!bwd) o += MW * L.partlen;
  // per-warp partial rows of the O(N) reductions (lane per output): forward E (1 + D) outputs, reverse sweep D + P
  L.wplen = bwd ? (D + P) : (E * L.nOut);
  L.wp = o; o += (bwd ? 2 * MW : MW) * L.wplen;   // reverse sweep: two sets (B1, B3)
  // reverse sweep, optional (UNI_BWD_COLRED_SMEM): per-warp scratch of col_reduce8s (COLRED_WARP doubles, <= 8 warps)
  o = (o + 1) & ~1;
  L.colred = o;
#if defined(UNI_BWD_COLRED_SMEM) && UNI_BWD_COLRED_SMEM
  if (bwd) o += 8 * 320;
#endif
  L.S = o; o += 2 * E * E;          // S and s V^T of the recurrence stage
  L.cstw = o; if (!bwd) o += E * E + GPMPC_MAX_D;   // forward: scratch of the warp-wide stage cost (stage_cost_warp)
  L.cst = o;                        // target, W, WT (read from shared memory by the forward kernel only)
  if (!bwd) o += GPMPC_MAX_D + GPMPC_MAX_D * GPMPC_MAX_D + GPMPC_MAX_EV * GPMPC_MAX_EV;
  L.m = o; o += GPMPC_MAX_D;
  L.s = o; o += EV * EV;
  L.mu = o; o += GPMPC_MAX_EV;
  L.A = o; o += EV * EV;
  L.Q = o; o += EV * EV;
  L.misc = o; o += 16;             // c, detR, detB, s2, ...
  L.M = o; o += GPMPC_MAX_EV;
  L.V = o; o += E * D;
  L.accN = bwd ? (1 + D + EV * EV + D + P) : (P + 1);
  L.acc = o; o += L.accN + 1;
  L.am = o; o += H * Na + 1;
  L.r = o; o += H + 1;
  L.rv = o; o += H + 1;
  L.ints = o; o += 4;
  o = (o + 1) & ~1;
  L.small2 = o; o += bwd ? (15 * EV * EV + 64) : 0;   // carve: uniform_bwd_kernel (small matrices + B0/B4 scratch)
  L.total = (o + 1) & ~1;
  return L;
}

struct UniRecLayout { int offM, offV, offH, offG, offS, size; };
HD UniRecLayout uni_rec_layout(int E) {
  UniRecLayout r;
  r.offM = 0; r.offV = E; r.offH = E + E * E; r.offG = r.offH + E; r.offS = r.offG + E * E;
  r.size = r.offS + E * (E + 1) / 2;
  return r;
}

}  // namespace gpmpc
