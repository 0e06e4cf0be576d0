// Shared-memory plan of rollout_kernel, shared by host (sizing) and device (carving).
#pragma once
#include "gpmpc_common.cuh"
#include "gpmpc_internal.h"

namespace gpmpc {

// ---------------------------------------------------------------------------------------------
// shared-memory layout (offsets in doubles), computed identically on host and device
// ---------------------------------------------------------------------------------------------
struct SmemLayout {
  int nu, grp, kap, gam, rho, xi, out, nOut, PV;
  int m, s, mu, A, c, il2, s2, logs2, Q, Wd, detR, Sraw, M, V, pacc, paccN, am, r, rv, ints, tab, cst, total;
};

HD SmemLayout make_layout(int EV, bool grad, int NP, int DP, int D, int E, int G, int H, int Na) {
  SmemLayout L;
  const int P = E * (E + 1) / 2;
  L.PV = EV * (EV + 1) / 2;
  int o = 0;
  L.nu = o; o += NP * DP;
  L.grp = o;
  const int per_pair = grad ? (3 + EV) : 1;
  int grp = G * NP * per_pair;
  L.nOut = 1 + D + (grad ? (EV * D + EV * L.PV) : 0);
  // the lb[E][NP] array of phases P1/P2 and the moment outputs of P2/P2b alias the group arrays of P3
  if (grp < E * NP + E * L.nOut) grp = E * NP + E * L.nOut;
  L.kap = L.grp; L.gam = L.kap + G * NP; L.rho = L.gam + G * NP; L.xi = L.rho + G * NP;
  L.out = L.grp + E * NP;
  o += grp;
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
  L.paccN = 1 + D + EV * EV;
  L.pacc = o; o += G * L.paccN;
  L.am = o; o += H * Na + 1;
  L.r = o; o += H + 1;
  L.rv = o; o += H + 1;
  L.ints = o; o += 2 + P;  // counter, bad flag, pair table (packed a*16+b)
  o = (o + 1) & ~1;
  L.tab = o; o += EXP2S_N;   // 2^(j/2048) for exp2s
  L.cst = o; o += GPMPC_MAX_D + GPMPC_MAX_D * GPMPC_MAX_D + GPMPC_MAX_EV * GPMPC_MAX_EV;   // cost: target, W, WT
  L.total = (o + 1) & ~1;
  return L;
}

}  // namespace gpmpc
