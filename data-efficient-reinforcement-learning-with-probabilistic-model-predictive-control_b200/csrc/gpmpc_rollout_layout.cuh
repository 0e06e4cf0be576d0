// Shared-memory plan of rollout_kernel, shared by host (sizing) and device (carving).
#pragma once
#include "gpmpc_common.cuh"
#include "gpmpc_internal.h"

namespace gpmpc {

// ---------------------------------------------------------------------------------------------
// shared-memory layout (offsets in doubles), computed identically on host and device
// ---------------------------------------------------------------------------------------------
struct SmemLayout {
  int nu, grp, kap, out, nOut, PV, colred;
  int m, s, mu, A, c, il2, s2, logs2, Q, Wd, detR, Sraw, M, V, pacc, paccN, am, r, rv, ints, cst, total;
};

// G = pairs per N^2 phase: the only per-pair shared-memory arrays are the column records {kap'_j, sign constant} (2 NP
// doubles per pair, one 16-byte load per column in the hot loop);
// the row sums of the sweeps stay in registers and the column sums live in a per-CTA global scratch (L2)
// lb_global: the per-GP weights lb[E][NP] of phases P1/P2 live in the CTA's global scratch instead of shared memory
// (large shapes: E = 8, N = 1000 with the gradient would not fit otherwise)
HD SmemLayout make_layout(int EV, bool grad, int NP, int DP, int D, int E, int G, int H, int Na, int nwarps, bool lb_global = false) {
  SmemLayout L;
  const int P = E * (E + 1) / 2;
  L.PV = EV * (EV + 1) / 2;
  int o = 0;
  L.nu = o; o += NP * DP;
  L.grp = o;
  int grp = 2 * G * NP;
  L.nOut = 1 + D + (grad ? (EV * D + EV * L.PV) : 0);
  // the lb[E][NP] array of phases P1/P2 and the moment outputs of P2/P2b alias the column terms of P3
  const int lbn = lb_global ? 0 : E * NP;
  if (grp < lbn + E * L.nOut) grp = lbn + E * L.nOut;
  L.kap = L.grp;
  L.out = L.grp + lbn;
  o += grp;
  o = (o + 1) & ~1;   // 16-byte aligned (col_reduce8s stores double2)
  L.colred = o; if (grad) o += nwarps * 320;   // per-warp scratch of col_reduce8s (COLRED_WARP doubles)
  L.m = o; o += GPMPC_MAX_D;
  L.s = o; o += EV * EV;
  L.mu = o; o += GPMPC_MAX_EV;
  L.A = o; o += E * EV * EV;
  L.c = o; o += GPMPC_MAX_EV;
  L.il2 = o; o += E * D;
  L.s2 = o; o += GPMPC_MAX_EV;
  L.logs2 = o; o += GPMPC_MAX_EV;
  L.Q = o; o += P * EV * EV;
  L.Wd = o; o += P * EV;
  L.detR = o; o += P;
  L.Sraw = o; o += P;
  L.M = o; o += GPMPC_MAX_EV;
  L.V = o; o += E * D;
  L.paccN = 1 + D + L.PV;   // S_raw, dS/dm (D), upper triangle of dS/dQ
  L.pacc = o; o += G * L.paccN;
  L.am = o; o += H * Na + 1;
  L.r = o; o += H + 1;
  L.rv = o; o += H + 1;
  L.ints = o; o += 2 + P;  // counter, bad flag, pair table (packed a*16+b)
  o = (o + 1) & ~1;
  L.cst = o; o += GPMPC_MAX_D + GPMPC_MAX_D * GPMPC_MAX_D + GPMPC_MAX_EV * GPMPC_MAX_EV;   // cost: target, W, WT
  L.total = (o + 1) & ~1;
  return L;
}

}  // namespace gpmpc
