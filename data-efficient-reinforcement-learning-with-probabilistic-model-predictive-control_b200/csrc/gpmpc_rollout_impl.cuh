// Batched moment-matching rollout + expected-cost scoring (the hot path).
//
//   predict_next_state_change   control_objects/models/gp_model.py:112-180
//   predict_trajectory          control_objects/models/gp_model.py:60-110
//   get_rewards_trajectory      control_objects/states_reward_mappers/setpoint_distance_reward_mapper.py:12-68,124-149
//   compute_mean_lcb_trajectory control_objects/controllers/gp_mpc_controller.py:229-285
//
// One persistent CTA per SM; each CTA takes candidates round-robin and runs all H steps of a
// candidate with the state (mu, Sigma) resident in shared memory.  Per step:
//   P0  small E x E algebra (A_a, c_a, Q_ab, det R_ab)            a few threads
//   P1  nu_i = x_i - m, per-GP exponent terms                     thread per training point
//   P2  O(N) moment sums (mean, V and their local Jacobians)      thread per output entry
//   P3  O(P N^2) covariance sums: the hot loop (see pair_item)    warp per 64-row x SEG-col item
//   P4  S assembly, (mu, Sigma) recurrence, stage cost            a few threads
// The training block (x, beta, iK) is identical for all candidates and stays L2-resident; iK is
// streamed through L1 with coalesced 256 B warp reads (row j of the symmetric matrix).
//
// Gradient mode (GRAD) additionally emits, for every N- and N^2-sum, its partial derivatives
// w.r.t. the small local parameters (m, A_a, Q_ab) into a per-step record; the reverse sweep
// (backward_kernel) is then pure small-matrix algebra.  tests/algo_spec.py is the executable spec.
#pragma once
#include "gpmpc_common.cuh"
#include "gpmpc_internal.h"
#include "gpmpc_rollout_layout.cuh"

namespace gpmpc {

// ---------------------------------------------------------------------------------------------
// cost (setpoint_distance_reward_mapper.py:12-68, :124-142)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_early(double v) {   // (warp_sum, defined with the other warp helpers below)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
struct CostView {
  const double* target; const double* W; const double* WT; const double* smin; const double* smax;
  double kappa; int use_constraints;
};

__device__ inline void stage_cost(const CostView& c, int E, int Na, const double* mu, const double* s,
                           const double* a, double& cmu, double& cvar) {
  const int Dc = E + Na;
  double e[GPMPC_MAX_D], We[GPMPC_MAX_D];
  for (int d = 0; d < Dc; d++) e[d] = (d < E ? mu[d] : a[d - E]) - c.target[d];
  for (int d = 0; d < Dc; d++) {
    double v = 0.0;
    for (int k = 0; k < Dc; k++) v += c.W[d * Dc + k] * e[k];
    We[d] = v;
  }
  double tr1 = 0.0, quad = 0.0;
  for (int i = 0; i < E; i++)
    for (int k = 0; k < E; k++) tr1 += s[i * E + k] * c.W[k * Dc + i];
  for (int d = 0; d < Dc; d++) quad += e[d] * We[d];
  cmu = tr1 + quad;
  // TS = Wss s ; tr(TS TS)
  double TS[GPMPC_MAX_EV * GPMPC_MAX_EV];
  for (int i = 0; i < E; i++)
    for (int k = 0; k < E; k++) {
      double v = 0.0;
      for (int l = 0; l < E; l++) v += c.W[i * Dc + l] * s[l * E + k];
      TS[i * E + k] = v;
    }
  double tr2 = 0.0, q2 = 0.0;
  for (int i = 0; i < E; i++)
    for (int k = 0; k < E; k++) tr2 += TS[i * E + k] * TS[k * E + i];
  for (int i = 0; i < E; i++)
    for (int k = 0; k < E; k++) q2 += We[i] * s[i * E + k] * We[k];
  cvar = 2.0 * tr2 + 4.0 * q2;
  if (c.use_constraints) {  // variance used as sigma: setpoint_distance_reward_mapper.py:60-64
    const double rt2 = 1.4142135623730951;
    for (int d = 0; d < E; d++) {
      double sig = s[d * E + d];
      double zmin = (c.smin[d] - mu[d]) / (sig * rt2), zmax = (c.smax[d] - mu[d]) / (sig * rt2);
      cmu += 0.5 * (1.0 + erf(zmin)) + (1.0 - 0.5 * (1.0 + erf(zmax)));
    }
  }
}

// stage_cost by one whole warp (uniform forward kernel, P0: the serial routine kept the CTA at the barrier for ~8 k clocks
// per step).  scr: E * E + Dc doubles of shared memory private to the warp.  Results in every lane.
__device__ inline void stage_cost_warp(const CostView& c, int E, int Na, const double* mu, const double* s, const double* a,
                                       int lane, double* scr, double& cmu, double& cvar) {
  const int Dc = E + Na, E2 = E * E;
  double* TS = scr;            // Wss s
  double* We = scr + E2;       // W (x - target)
  for (int d = lane; d < Dc; d += 32) {
    double v = 0.0;
    for (int k = 0; k < Dc; k++) v += c.W[d * Dc + k] * ((k < E ? mu[k] : a[k - E]) - c.target[k]);
    We[d] = v;
  }
  double tr1 = 0.0;
  for (int o = lane; o < E2; o += 32) {
    const int i = o / E, k = o - i * E;
    tr1 += s[o] * c.W[k * Dc + i];
    double v = 0.0;
    for (int l = 0; l < E; l++) v += c.W[i * Dc + l] * s[l * E + k];
    TS[o] = v;
  }
  __syncwarp();
  double quad = 0.0, tr2 = 0.0, q2 = 0.0, pen = 0.0;
  for (int d = lane; d < Dc; d += 32) quad += ((d < E ? mu[d] : a[d - E]) - c.target[d]) * We[d];
  for (int o = lane; o < E2; o += 32) {
    const int i = o / E, k = o - i * E;
    tr2 += TS[o] * TS[k * E + i];
    q2 += We[i] * s[o] * We[k];
  }
  if (c.use_constraints) {  // variance used as sigma: setpoint_distance_reward_mapper.py:60-64
    const double rt2 = 1.4142135623730951;
    for (int d = lane; d < E; d += 32) {
      double sig = s[d * E + d];
      double zmin = (c.smin[d] - mu[d]) / (sig * rt2), zmax = (c.smax[d] - mu[d]) / (sig * rt2);
      pen += 0.5 * (1.0 + erf(zmin)) + (1.0 - 0.5 * (1.0 + erf(zmax)));
    }
  }
  cmu = warp_sum_early(tr1 + quad + pen);
  cvar = warp_sum_early(2.0 * tr2 + 4.0 * q2);
  __syncwarp();
}

__device__ inline void terminal_cost(const CostView& c, int E, const double* mu, const double* s, double& cmu,
                              double& cvar) {
  double e[GPMPC_MAX_EV], We[GPMPC_MAX_EV], TS[GPMPC_MAX_EV * GPMPC_MAX_EV];
  for (int d = 0; d < E; d++) e[d] = mu[d] - c.target[d];
  for (int d = 0; d < E; d++) {
    double v = 0.0;
    for (int k = 0; k < E; k++) v += c.WT[d * E + k] * e[k];
    We[d] = v;
  }
  double tr1 = 0.0, quad = 0.0, tr2 = 0.0, q2 = 0.0;
  for (int i = 0; i < E; i++)
    for (int k = 0; k < E; k++) {
      tr1 += s[i * E + k] * c.WT[k * E + i];
      double v = 0.0;
      for (int l = 0; l < E; l++) v += c.WT[i * E + l] * s[l * E + k];
      TS[i * E + k] = v;
    }
  for (int d = 0; d < E; d++) quad += e[d] * We[d];
  for (int i = 0; i < E; i++)
    for (int k = 0; k < E; k++) {
      tr2 += TS[i * E + k] * TS[k * E + i];
      q2 += We[i] * s[i * E + k] * We[k];
    }
  cmu = tr1 + quad;
  cvar = 2.0 * tr2 + 4.0 * q2;
}

// ---------------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column sums over the 32 lanes for 8 columns at once (transpose-reduce: 9 shuffles instead of
// 40).  On return every lane of quad q = lane>>2 holds the total of column `col`.
__device__ __forceinline__ double col_reduce8(const double (&v)[8], int lane, int& col) {
  double a[4], b[2], c;
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    double send = u16 ? v[k] : v[k + 4];
    double keep = u16 ? v[k + 4] : v[k];
    a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 2; k++) {
    double send = u8 ? a[k] : a[k + 2];
    double keep = u8 ? a[k + 2] : a[k];
    b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    double send = u4 ? b[0] : b[1];
    double keep = u4 ? b[1] : b[0];
    c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  col = (u16 ? 4 : 0) + (u8 ? 2 : 0) + (u4 ? 1 : 0);
  return c;
}

// The same column sums through a per-warp shared-memory scratch (COLRED_WARP doubles): every lane stores its 8 values
// (rows of 80 bytes: conflict-free 16-byte stores), lane (g, c) = (lane >> 3, lane & 7) adds column c of source lanes
// 8 g .. 8 g + 7 (the groups of a half-warp read rows 12 apart: conflict-free), two shuffle rounds finish.  30 instructions
// instead of the ~105 of the transpose-reduce (18 SHFL + 28 FSEL + moves): the sweeps are issue-bound, every instruction
// counts (profiles/r02_micro_gen_loop.txt).  On return every lane holds the total of column lane & 7.
constexpr int COLRED_STRIDE = 10;
constexpr int COLRED_WARP = 32 * COLRED_STRIDE;
__device__ __forceinline__ double col_reduce8s(const double (&v)[8], int lane, double* __restrict__ scr) {
  double2* w = reinterpret_cast<double2*>(scr + lane * COLRED_STRIDE);
#pragma unroll
  for (int q = 0; q < 4; q++) w[q] = make_double2(v[2 * q], v[2 * q + 1]);
  __syncwarp();
  const int c = lane & 7, g = lane >> 3;
  const double* ra = scr + (8 * g + 4 * (g & 1)) * COLRED_STRIDE + c;
  const double* rb = scr + (8 * g + 4 - 4 * (g & 1)) * COLRED_STRIDE + c;
  double s0 = ra[0] + ra[COLRED_STRIDE], s1 = ra[2 * COLRED_STRIDE] + ra[3 * COLRED_STRIDE];
  s0 += rb[0] + rb[COLRED_STRIDE];
  s1 += rb[2 * COLRED_STRIDE] + rb[3 * COLRED_STRIDE];
  double s = s0 + s1;
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  __syncwarp();   // the scratch is rewritten by the next round
  return s;
}

// Sums of K2 (power of two, <= 32) per-lane values over the 32 lanes with K2 - 1 + log2(32 / K2) shuffles instead
// of 5 K2 (halving exchange: at offset 16, 8, ... every lane keeps one half of its values and sends the other half).
// On return value number `idx` is complete in every lane that shares `idx`; lanes with (lane & (32 / K2 - 1)) == 0
// are the designated writers (idx = lane / (32 / K2)).
template <int K2>
__device__ __forceinline__ double warp_reduce_multi(double (&v)[K2], int lane, int& idx) {
  static_assert(K2 >= 1 && K2 <= 32 && (K2 & (K2 - 1)) == 0, "K2 must be a power of two <= 32");
  int off = 16;
  idx = 0;
#pragma unroll
  for (int n = K2; n > 1; n >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < n / 2; k++) {
      const double send = up ? v[k] : v[k + n / 2];
      const double keep = up ? v[k + n / 2] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    idx += up ? n / 2 : 0;
  }
  double r = v[0];
#pragma unroll
  for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

// ---------------------------------------------------------------------------------------------
// float64 reduction at L2 without a return value (RED.E.ADD.F64: fire and forget); float64 atomics on shared memory
// are CAS spin loops, so sums that several warps contribute to per element leave the SM this way
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void uni_red_add(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// Row factor of the sweeps: the per-row term kap_i of the exponent never enters the element loops,
//   exp(kap_i + kap_j + u_i . nu_j) = e_i * exp(kap_j + u_i . nu_j) ,   e_i = exp(max(kap_i, kmin)),
// and e_i multiplies the finished row sums (or is folded into the row's coefficients).  kr (table units) -> e, kr := the
// residual shift kap_i - max(kap_i, kmin), non-zero only for rows far away from the input mean (their warps run the
// loop variant with the extra add, SH).  kmin = -600 in natural units: with the total exponent bounded above by
// log(s2_a s2_b) (+ log|beta|), the remaining factor cannot overflow.
__device__ __forceinline__ double uni_row_factor(double& kr) {
  const double c = fmax(kr, -600.0 * GPMPC_EXP2S_SCALE);   // NaN -> kmin, and the residual keeps the NaN
  kr -= c;
  return exp2s(c);
}

// Warp totals of PV "pair" values (k <= l, row-major) and up to GPMPC_MAX_D "single" values that every lane has summed
// over its own rows / columns: halving exchanges, 16 values per round (16 shuffles instead of 80), then ONE shared-memory
// atomicAdd per value and warp into acc[o]: o < D singles, o = D + pr pairs.  (A few hundred per step: the CAS loop of
// the float64 shared atomic does not matter here, unlike per-element sums.)
template <int PV>
__device__ __forceinline__ void gen_warp_sums_add(const double (&vp)[PV], const double (&vs)[GPMPC_MAX_D], int D, int lane,
                                                  double* __restrict__ acc) {
  constexpr int NCH = (PV + GPMPC_MAX_D + 15) / 16;
#pragma unroll
  for (int ch = 0; ch < NCH; ch++) {
    if (16 * ch < PV + D) {   // warp-uniform: rounds that hold only unused single slots are skipped
      double v16[16];
#pragma unroll
      for (int k = 0; k < 16; k++) {
        const int sl = 16 * ch + k;
        v16[k] = (sl < PV) ? vp[sl < PV ? sl : 0] : ((sl - PV < GPMPC_MAX_D) ? vs[(sl - PV >= 0 && sl - PV < GPMPC_MAX_D) ? sl - PV : 0] : 0.0);
      }
      int idx;
      const double tot = warp_reduce_multi<16>(v16, lane, idx);
      const int sl = 16 * ch + idx;
      if ((lane & 1) == 0) {
        if (sl < PV) atomicAdd(acc + D + sl, tot);
        else if (sl - PV < D) atomicAdd(acc + (sl - PV), tot);
      }
    }
  }
}

// nu_j[0 .. EV) of one training point as 16-byte shared-memory loads (rows of s_nu start 16-byte aligned: DP is even)
template <int EV>
__device__ __forceinline__ void gen_load_nu(const double* __restrict__ row, double (&nu)[EV]) {
  const double2* r2 = reinterpret_cast<const double2*>(row);
#pragma unroll
  for (int q = 0; q < (EV + 1) / 2; q++) {
    const double2 v = r2[q];
    nu[2 * q] = v.x;
    if (2 * q + 1 < EV) nu[2 * q + 1] = v.y;
  }
}

// ---------------------------------------------------------------------------------------------
// The hot loop of the general path (per-GP hyper-parameters): one warp, one output pair (a, b), lane = two ADJACENT rows
// i0 = 64 I + 2 lane, i0 + 1, columns [jbeg, jend) two at a time (4 independent chains per warp).
//   t_ij = kap_i + kap_j + u_i . nu_j ;  w_ij = (beta_a,i beta_b,j - [a==b] iK_a,ij) exp(t_ij)
// gp_model.py:161-175 (X, X2, Q, maha, k, L, beta L beta, iK * L) collapsed to one exponent per element.
//  * row factor e_i = exp(kap_i) outside the loop (uni_row_factor);
//  * off-diagonal pairs: log|beta_b,j| is part of the column term kap'_j and the sign of beta_b,j rides in its lowest
//    mantissa bit (one integer XOR on the result), so the element costs no multiplication by the coefficient:
//    rho'_i = sum_j +-exp(kap'_j + u_i . nu_j), xi'_i likewise, and the row factor be_i = beta_a,i e_i multiplies the
//    finished sums;  E + 7 + 1 float64 instructions per element (value), + E + 1.6 with the gradient sums;
//  * diagonal pairs: c_ij = be_i beta_j - e_i iK_ij (both iK values of a column with one 16-byte load), upper tile
//    triangle only, the diagonal tile in full with HALF weights (w is symmetric and every consumer of (rho, gam, xi) is
//    invariant under that swap, see uni_bwd_item) -- no element masks;
//  * gradient sums: rows rho_i, xi_i = sum_j w_ij nu_j stay in registers (folded into the pair's accumulators by
//    gen_item); column sums gam_j by an 8-column transpose-reduce and a float64 RED at L2 (per-CTA scratch).
// ---------------------------------------------------------------------------------------------
#ifndef GEN_IK_PREFETCH
#define GEN_IK_PREFETCH 1     // diagonal pairs: L1 prefetch of the next round's iK rows (+ L1-allocating loads)
#endif
#ifndef GEN_ROWS4
#define GEN_ROWS4 1        // off-diagonal pairs: four rows per lane (gen_cols4) when NP is a multiple of 128
#endif
template <int EV, bool GRAD, bool DIAG, bool SH>
__device__ __forceinline__ void gen_cols(const double* __restrict__ s_nu, int DP, const double* __restrict__ s_kp,
                                         const double* __restrict__ beta_b, const double* __restrict__ ik0, int NP,
                                         int jbeg, int jend, const double (&u0)[EV], const double (&u1)[EV],
                                         double kr0, double kr1, double cb0, double cb1, double ce0, double ce1,
                                         double& rho0, double& rho1, double (&xi0)[EV], double (&xi1)[EV], int lane,
                                         double* __restrict__ g_gam, double* __restrict__ scr) {
  const double* __restrict__ pn = s_nu + (size_t)jbeg * DP;
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    double v[8];
#if GEN_IK_PREFETCH
    // diagonal pairs: the iK rows of the NEXT round of 8 columns -> L1 (they come from L2, ~700 clocks away, and the three
    // warps of a sub-partition cannot hide that: a diagonal chunk took 1.42x an off-diagonal one instead of the 1.19x of
    // its instruction count)
    if (DIAG && j0 + 8 < jend) {
#pragma unroll
      for (int q = 0; q < 8; q++) asm volatile("prefetch.global.L1 [%0];" ::"l"(ik0 + (size_t)(8 + q) * NP));
    }
#endif
#pragma unroll
    for (int jp = 0; jp < 4; jp++) {
      const int j = j0 + 2 * jp;
      double na[EV], nb[EV];
      gen_load_nu<EV>(pn, na);
      gen_load_nu<EV>(pn + DP, nb);
      pn += 2 * DP;
      // column record {kap'_j, magic constant of the exp's range reduction carrying the sign of beta_b,j}
      const double2 ca = *reinterpret_cast<const double2*>(s_kp + 2 * j);
      const double2 cb = *reinterpret_cast<const double2*>(s_kp + 2 * j + 2);
      double t[4], w[4];
      if (SH) { t[0] = kr0 + ca.x; t[1] = kr1 + ca.x; t[2] = kr0 + cb.x; t[3] = kr1 + cb.x; }
      else { t[0] = ca.x; t[1] = ca.x; t[2] = cb.x; t[3] = cb.x; }
      // serpentine order: every FMA shares one register operand with its predecessor (operand-reuse cache; a DFMA with
      // three fresh register operands issues at 2/3 rate on B200, tools/micro/dfma_operands.cu)
#pragma unroll
      for (int e = 0; e < EV; e++) {
        t[0] = fma(u0[e], na[e], t[0]);
        t[1] = fma(u1[e], na[e], t[1]);
        t[3] = fma(u1[e], nb[e], t[3]);
        t[2] = fma(u0[e], nb[e], t[2]);
      }
      if (DIAG) {
        exp2s_x4(t, w);
      } else {   // sign of beta_b,j through the magic constant of the range reduction (exp2s_x4_signed): no per-element cost
        exp2s_x4_signed(t, w, ca.y, cb.y);
      }
      if (DIAG) {
        const double2 bb2 = make_double2(ca.y, cb.y);   // beta_b,j: second slot of the column record (diagonal pairs)
#if GEN_IK_PREFETCH      // the prefetched rows are taken from L1
        const double2 ika = __ldg(reinterpret_cast<const double2*>(ik0));
        const double2 ikb = __ldg(reinterpret_cast<const double2*>(ik0 + NP));
#else
        const double2 ika = ldg_stream2<true>(ik0);
        const double2 ikb = ldg_stream2<true>(ik0 + NP);
#endif
        ik0 += 2 * (size_t)NP;
        double c[4] = {-ce0 * ika.x, -ce1 * ika.y, -ce0 * ikb.x, -ce1 * ikb.y};
        c[0] = fma(cb0, bb2.x, c[0]);
        c[1] = fma(cb1, bb2.x, c[1]);
        c[3] = fma(cb1, bb2.y, c[3]);
        c[2] = fma(cb0, bb2.y, c[2]);
        if (GRAD) {
#pragma unroll
          for (int q = 0; q < 4; q++) w[q] *= c[q];
          rho0 += w[0] + w[2];
          rho1 += w[1] + w[3];
        } else {
          rho0 = fma(c[0], w[0], rho0);
          rho1 = fma(c[1], w[1], rho1);
          rho0 = fma(c[2], w[2], rho0);
          rho1 = fma(c[3], w[3], rho1);
        }
      } else {
        rho0 += w[0] + w[2];
        rho1 += w[1] + w[3];
      }
      if (GRAD) {
#pragma unroll
        for (int e = 0; e < EV; e++) {
          if (e & 1) { xi1[e] = fma(w[1], na[e], xi1[e]); xi0[e] = fma(w[0], na[e], xi0[e]); }
          else       { xi0[e] = fma(w[0], na[e], xi0[e]); xi1[e] = fma(w[1], na[e], xi1[e]); }
        }
#pragma unroll
        for (int e = 0; e < EV; e++) {
          if (e & 1) { xi1[e] = fma(w[3], nb[e], xi1[e]); xi0[e] = fma(w[2], nb[e], xi0[e]); }
          else       { xi0[e] = fma(w[2], nb[e], xi0[e]); xi1[e] = fma(w[3], nb[e], xi1[e]); }
        }
        if (DIAG) { v[2 * jp] = w[0] + w[1]; v[2 * jp + 1] = w[2] + w[3]; }
        else { v[2 * jp] = fma(cb0, w[0], cb1 * w[1]); v[2 * jp + 1] = fma(cb0, w[2], cb1 * w[3]); }
      }
    }
    if (GRAD) {
      const double tot = col_reduce8s(v, lane, scr);
      if (lane < 8) uni_red_add(g_gam + j0 + lane, tot);
    }
  }
}

// One run of columns [jbeg, jend) (multiples of 8; diagonal pairs: jbeg >= 64 I) of row block I of pair (a, b): row
// set-up, the column loops, and -- everything downstream being linear in (rho_i, xi_i) -- the fold of the lane's
// partial row sums straight into the pair's accumulators  acc[0] = S_raw, acc[1 .. 1+D) = dS/dm (rho part),
// acc[1+D ..) = upper triangle of dS/dQ (rho and xi parts); the gam parts follow after the sweep (rollout_kernel).
template <int EV, bool GRAD, bool DIAG>
__device__ __forceinline__ void gen_item(const RolloutParams& p, const double* __restrict__ s_nu,
                                         const double* __restrict__ s_kp, const double* __restrict__ Qm,
                                         const double* __restrict__ il2a, const double* __restrict__ il2b,
                                         const double* __restrict__ kka, const double* __restrict__ beta_a,
                                         const double* __restrict__ beta_b, const double* __restrict__ iKa, int I,
                                         int jbeg, int jend, int lane, double* __restrict__ g_gam,
                                         double* __restrict__ acc, double* __restrict__ scr) {
  constexpr int PV = EV * (EV + 1) / 2;
  const int NP = p.NP, DP = p.DP, D = p.D;
  const int i0 = 64 * I + 2 * lane, i1 = i0 + 1;
  const double* __restrict__ n0p = s_nu + (size_t)i0 * DP;
  const double* __restrict__ n1p = n0p + DP;
  double u0[EV], u1[EV], kr0, kr1;
  {
    double z0[EV], z1[EV];
#pragma unroll
    for (int e = 0; e < EV; e++) {
      z0[e] = n0p[e] * il2a[e];
      z1[e] = n1p[e] * il2a[e];
    }
    kr0 = kka[i0];
    kr1 = kka[i1];
#pragma unroll
    for (int e = 0; e < EV; e++) {
      double q0 = 0.0, q1 = 0.0;
#pragma unroll
      for (int f = 0; f < EV; f++) {
        q0 = fma(Qm[e * EV + f], z0[f], q0);
        q1 = fma(Qm[e * EV + f], z1[f], q1);
      }
      kr0 = fma(z0[e], q0, kr0);
      kr1 = fma(z1[e], q1, kr1);
      u0[e] = (2.0 * GPMPC_EXP2S_SCALE) * q0 * il2b[e];   // exponent in table units (exp2s; s_kp is scaled too)
      u1[e] = (2.0 * GPMPC_EXP2S_SCALE) * q1 * il2b[e];
    }
    kr0 *= GPMPC_EXP2S_SCALE;
    kr1 *= GPMPC_EXP2S_SCALE;
  }
  const double e0 = uni_row_factor(kr0), e1 = uni_row_factor(kr1);   // kr0, kr1 become residual shifts
  const bool far = __any_sync(0xffffffffu, kr0 != 0.0 || kr1 != 0.0);
  const double be0 = __ldg(beta_a + i0) * e0, be1 = __ldg(beta_a + i1) * e1;
  double rho0 = 0.0, rho1 = 0.0, xi0[EV], xi1[EV];
#pragma unroll
  for (int e = 0; e < EV; e++) { xi0[e] = 0.0; xi1[e] = 0.0; }
  jend = min(jend, (p.N + 7) & ~7);   // zero-padded columns contribute (numerically) nothing: skip them
  if (DIAG) {
    // two column segments: the diagonal tile (half weights) and the tiles above it; ONE inlined instance of the loops
    // (the sweeps' code has to stay resident in the instruction caches of all warps, whatever variant they run)
    const int jd1 = 64 * I + 64;
#pragma unroll 1
    for (int seg = 0; seg < 2; seg++) {
      const int jb = seg == 0 ? jbeg : max(jbeg, jd1), je = seg == 0 ? min(jend, jd1) : jend;
      if (jb >= je) continue;
      const double wg = seg == 0 ? 0.5 : 1.0;
      const double* ikp = iKa + (size_t)jb * NP + i0;   // row j of the symmetric iK, lane = columns i0, i0 + 1
      if (far) gen_cols<EV, GRAD, true, true>(s_nu, DP, s_kp, beta_b, ikp, NP, jb, je, u0, u1, kr0, kr1, wg * be0, wg * be1,
                                              wg * e0, wg * e1, rho0, rho1, xi0, xi1, lane, g_gam, scr);
      else gen_cols<EV, GRAD, true, false>(s_nu, DP, s_kp, beta_b, ikp, NP, jb, je, u0, u1, kr0, kr1, wg * be0, wg * be1,
                                           wg * e0, wg * e1, rho0, rho1, xi0, xi1, lane, g_gam, scr);
    }
  } else {
    if (far) gen_cols<EV, GRAD, false, true>(s_nu, DP, s_kp, nullptr, nullptr, NP, jbeg, jend, u0, u1, kr0, kr1, be0, be1, 0.0, 0.0,
                                             rho0, rho1, xi0, xi1, lane, g_gam, scr);
    else gen_cols<EV, GRAD, false, false>(s_nu, DP, s_kp, nullptr, nullptr, NP, jbeg, jend, u0, u1, kr0, kr1, be0, be1, 0.0, 0.0,
                                          rho0, rho1, xi0, xi1, lane, g_gam, scr);
    rho0 *= be0;
    rho1 *= be1;
    if (GRAD) {
#pragma unroll
      for (int e = 0; e < EV; e++) { xi0[e] *= be0; xi1[e] *= be1; }
    }
  }
  {
    const double tot = warp_sum(rho0 + rho1);
    if (lane == 0) atomicAdd(acc, tot);
  }
  if (GRAD) {
    double vs[GPMPC_MAX_D], vp[PV];
#pragma unroll
    for (int d = 0; d < GPMPC_MAX_D; d++) vs[d] = (d < D) ? (rho0 * n0p[d < D ? d : 0] + rho1 * n1p[d < D ? d : 0]) * il2a[d < D ? d : 0] : 0.0;
    double za0[EV], za1[EV], xs0[EV], xs1[EV];
#pragma unroll
    for (int e = 0; e < EV; e++) {
      za0[e] = n0p[e] * il2a[e];
      za1[e] = n1p[e] * il2a[e];
      xs0[e] = xi0[e] * il2b[e];
      xs1[e] = xi1[e] * il2b[e];
    }
    int pr = 0;
#pragma unroll
    for (int k = 0; k < EV; k++) {
      const double rk0 = rho0 * za0[k], rk1 = rho1 * za1[k];
#pragma unroll
      for (int l = k; l < EV; l++) {
        double a0 = fma(rk0, za0[l], za0[k] * xs0[l]);
        a0 = fma(za0[l], xs0[k], a0);
        double a1 = fma(rk1, za1[l], za1[k] * xs1[l]);
        a1 = fma(za1[l], xs1[k], a1);
        vp[pr++] = a0 + a1;
      }
    }
    gen_warp_sums_add<PV>(vp, vs, D, lane, acc + 1);
  }
}

// ---------------------------------------------------------------------------------------------
// Off-diagonal pairs, 128-row warp tiles: a lane owns FOUR adjacent rows (128 I + 4 lane + r) and takes the columns one at
// a time (4 independent chains per warp = the 4 rows).  Per element this halves the column-record loads and the column-sum
// reduction of the 2 x 2 variant (gen_cols) -- the shared-memory pipe is the co-limiter of this kernel (70 %) -- at the
// price of twice the row state in registers (fits the 168 budget).  Used when NP is a multiple of 128.
// ---------------------------------------------------------------------------------------------
template <int EV, bool GRAD, bool SH>
__device__ __forceinline__ void gen_cols4(const double* __restrict__ s_nu, int DP, const double* __restrict__ s_kp,
                                          int jbeg, int jend, const double (&u)[4][EV], const double (&kr)[4],
                                          const double (&be)[4], double (&rho)[4], double (&xi)[4][EV], int lane,
                                          double* __restrict__ g_gam, double* __restrict__ scr) {
  const double* __restrict__ pn = s_nu + (size_t)jbeg * DP;
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    double v[8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      double nj[EV];
      gen_load_nu<EV>(pn, nj);
      pn += DP;
      const double2 cj = *reinterpret_cast<const double2*>(s_kp + 2 * (j0 + jj));   // {kap'_j, magic constant with the sign}
      double t[4], w[4];
#pragma unroll
      for (int r = 0; r < 4; r++) t[r] = SH ? kr[r] + cj.x : cj.x;
#pragma unroll
      for (int e = 0; e < EV; e++) {
#pragma unroll
        for (int r = 0; r < 4; r++) t[r] = fma(u[r][e], nj[e], t[r]);     // nj[e] shared by four consecutive FMAs (operand reuse)
      }
      exp2s_x4_signed(t, w, cj.y, cj.y);
#pragma unroll
      for (int r = 0; r < 4; r++) rho[r] += w[r];
      if (GRAD) {
#pragma unroll
        for (int e = 0; e < EV; e++) {
#pragma unroll
          for (int r = 0; r < 4; r++) xi[r][e] = fma(w[r], nj[e], xi[r][e]);
        }
        double vv = be[0] * w[0];
        vv = fma(be[1], w[1], vv);
        vv = fma(be[2], w[2], vv);
        v[jj] = fma(be[3], w[3], vv);
      }
    }
    if (GRAD) {
      const double tot = col_reduce8s(v, lane, scr);
      if (lane < 8) uni_red_add(g_gam + j0 + lane, tot);
    }
  }
}

template <int EV, bool GRAD>
__device__ __forceinline__ void gen_item4(const RolloutParams& p, const double* __restrict__ s_nu,
                                          const double* __restrict__ s_kp, const double* __restrict__ Qm,
                                          const double* __restrict__ il2a, const double* __restrict__ il2b,
                                          const double* __restrict__ kka, const double* __restrict__ beta_a, int I4,
                                          int jbeg, int jend, int lane, double* __restrict__ g_gam,
                                          double* __restrict__ acc, double* __restrict__ scr) {
  constexpr int PV = EV * (EV + 1) / 2;
  const int DP = p.DP, D = p.D;
  const int i0 = 128 * I4 + 4 * lane;
  const double* __restrict__ np0 = s_nu + (size_t)i0 * DP;
  double u[4][EV], kr[4], be[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const double* nr = np0 + r * DP;
    double z[EV];
#pragma unroll
    for (int e = 0; e < EV; e++) z[e] = nr[e] * il2a[e];
    double k = kka[i0 + r];
#pragma unroll
    for (int e = 0; e < EV; e++) {
      double q = 0.0;
#pragma unroll
      for (int f = 0; f < EV; f++) q = fma(Qm[e * EV + f], z[f], q);
      k = fma(z[e], q, k);
      u[r][e] = (2.0 * GPMPC_EXP2S_SCALE) * q * il2b[e];
    }
    kr[r] = k * GPMPC_EXP2S_SCALE;
  }
  bool anyfar = false;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const double e = uni_row_factor(kr[r]);          // kr becomes the residual shift
    be[r] = __ldg(beta_a + i0 + r) * e;
    anyfar = anyfar || kr[r] != 0.0;
  }
  const bool far = __any_sync(0xffffffffu, anyfar);
  double rho[4], xi[4][EV];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    rho[r] = 0.0;
#pragma unroll
    for (int e = 0; e < EV; e++) xi[r][e] = 0.0;
  }
  jend = min(jend, (p.N + 7) & ~7);
  if (far) gen_cols4<EV, GRAD, true>(s_nu, DP, s_kp, jbeg, jend, u, kr, be, rho, xi, lane, g_gam, scr);
  else gen_cols4<EV, GRAD, false>(s_nu, DP, s_kp, jbeg, jend, u, kr, be, rho, xi, lane, g_gam, scr);
  double accS = 0.0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    rho[r] *= be[r];
    accS += rho[r];
    if (GRAD) {
#pragma unroll
      for (int e = 0; e < EV; e++) xi[r][e] *= be[r];
    }
  }
  accS = warp_sum(accS);
  if (lane == 0) atomicAdd(acc, accS);
  if (GRAD) {
    double vs[GPMPC_MAX_D], vp[PV];
#pragma unroll
    for (int d = 0; d < GPMPC_MAX_D; d++) vs[d] = 0.0;
#pragma unroll
    for (int q = 0; q < PV; q++) vp[q] = 0.0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double* nr = np0 + r * DP;
#pragma unroll
      for (int d = 0; d < GPMPC_MAX_D; d++)
        if (d < D) vs[d] = fma(rho[r] * il2a[d], nr[d], vs[d]);
      double za[EV], xs[EV];
#pragma unroll
      for (int e = 0; e < EV; e++) { za[e] = nr[e] * il2a[e]; xs[e] = xi[r][e] * il2b[e]; }
      int q = 0;
#pragma unroll
      for (int k = 0; k < EV; k++) {
        const double rk = rho[r] * za[k];
#pragma unroll
        for (int l = k; l < EV; l++) {
          double a0 = fma(rk, za[l], vp[q]);
          a0 = fma(za[k], xs[l], a0);
          vp[q] = fma(za[l], xs[k], a0);
          q++;
        }
      }
    }
    gen_warp_sums_add<PV>(vp, vs, D, lane, acc + 1);
  }
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
// Launch plans (host: gpmpc_api.cu; GEN_MAXT in gpmpc_internal.h): state dimensions <= 5: 384 threads per SM with <= 168 registers
// -- two CTAs of 192 threads when the shared memory allows two per SM (one CTA's serial phases then overlap the other's
// sweep), else one of 384; larger state dimensions: one CTA of 256 threads with the full register file.
template <int EV, bool GRAD>
__global__ void __launch_bounds__(GEN_MAXT(EV), 1) rollout_kernel(const RolloutParams p) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, NT = blockDim.x, warp = tid >> 5, nwarps = NT >> 5;
  const int E = p.E, D = p.D, N = p.N, NP = p.NP, DP = p.DP, Na = p.Na, H = p.H;
  const int P = E * (E + 1) / 2, G = p.group;
  const SmemLayout L = make_layout(EV, GRAD, NP, DP, D, E, G, (p.mode == 0) ? H : 0, Na, nwarps, p.lb_global != 0);
  double* s_nu = sm + L.nu;
  // lb[E][NP]: aliases the column records in shared memory (only live in P1/P2), or the CTA's global scratch (lb_global)
  double* s_lb = p.lb_global ? p.ws_kk + ((size_t)blockIdx.x * 2 + 1) * E * NP : sm + L.grp;
  double* s_kap = sm + L.kap;
  double* s_colred = sm + L.colred + warp * COLRED_WARP;   // this warp's scratch of the column-sum reduction (gradient mode)
  double* s_out = sm + L.out;
  double* s_m = sm + L.m;
  double* s_s = sm + L.s;
  double* s_mu = sm + L.mu;
  double* s_A = sm + L.A;
  double* s_c = sm + L.c;
  double* s_il2 = sm + L.il2;
  double* s_s2 = sm + L.s2;
  double* s_logs2 = sm + L.logs2;
  double* s_Q = sm + L.Q;
  double* s_Wd = sm + L.Wd;
  double* s_detR = sm + L.detR;
  double* s_Sraw = sm + L.Sraw;
  double* s_M = sm + L.M;
  double* s_V = sm + L.V;
  double* s_pacc = sm + L.pacc;
  double* s_am = sm + L.am;
  double* s_r = sm + L.r;
  double* s_rv = sm + L.rv;
  int* s_int = reinterpret_cast<int*>(sm + L.ints);  // [0] counter, [1] bad flag, [2..] pair table
  const int nOut = L.nOut, PV = L.PV;
  const RecLayout RL = rec_layout(E, D);
  // cost description staged in shared memory (the stage cost is on the serial path of every step)
  double* s_cst = sm + L.cst;
  if (p.mode == 0) {
    for (int i = tid; i < E + Na; i += NT) s_cst[i] = p.c_target[i];
    for (int i = tid; i < (E + Na) * (E + Na); i += NT) s_cst[GPMPC_MAX_D + i] = p.c_W[i];
    for (int i = tid; i < E * E; i += NT) s_cst[GPMPC_MAX_D + GPMPC_MAX_D * GPMPC_MAX_D + i] = p.c_WT[i];
  }
  const CostView cv{s_cst, s_cst + GPMPC_MAX_D, s_cst + GPMPC_MAX_D + GPMPC_MAX_D * GPMPC_MAX_D, p.c_smin, p.c_smax,
                    p.kappa, p.use_constraints};

  // ---- candidate-independent constants
  for (int o = tid; o < E * D; o += NT) s_il2[o] = p.il2[o];
  if (tid < E) { s_s2[tid] = p.s2[tid]; s_logs2[tid] = log(p.s2[tid]); }
  exp2s_fill(p.exp2tab, tid, NT);
  if (tid == 0) {
    int pr = 0;
    for (int a = 0; a < E; a++)
      for (int b = a; b < E; b++) s_int[2 + pr++] = a * 16 + b;
  }
  __syncthreads();
  for (int o = tid; o < P * EV; o += NT) {
    int pr = o / EV, e = o % EV;
    int ab = s_int[2 + pr];
    s_Wd[o] = s_il2[(ab >> 4) * D + e] + s_il2[(ab & 15) * D + e];
  }
  double* kk = p.ws_kk + (size_t)blockIdx.x * 2 * E * NP;
  // column sums gam_j of the sweeps (float64 RED at L2), one row per pair of the group; zero on entry and re-zeroed
  // by their consumer after every step
  double* g_gam = GRAD ? p.ws_gam + (size_t)blockIdx.x * G * NP : nullptr;
  if (GRAD)
    for (int o = tid; o < G * NP; o += NT) g_gam[o] = 0.0;
  __syncthreads();

  // candidates are drawn from a global counter when the host provides one (rollouts; SM speeds differ by up to ~25 %,
  // see the uniform kernels), else dealt round-robin (single steps)
  // Small batches: a thread-block cluster of C CTAs shares each candidate.  The output pairs (a, b) -- independent N^2
  // sweeps -- are dealt to the CTAs (one pair per group, balanced by their cost); everything else is repeated by every
  // CTA bitwise identically; the S_raw of the pairs are exchanged through L2 (two buffers alternating with the step)
  // around ONE cluster barrier per horizon step; records of a pair are written by its owner, all other outputs by rank 0.
  const int C = (p.mode == 0) ? p.cluster : 1;
  const int crank = C > 1 ? (int)uni_cluster_rank() : 0, cid = blockIdx.x / C;
  const bool lead = crank == 0;
  int gstep = 0;
  long long clk_ = clock64();
#define GEN_CLK(k) do { if (p.dbg_clk && blockIdx.x == 0 && tid == 0) { const long long c_ = clock64(); p.dbg_clk[k] += c_ - clk_; clk_ = c_; } } while (0)
  int& s_next = gpmpc_ss.next;
  double& s_one = gpmpc_ss.one;
  if (tid == 0) s_one = 1.0;
  unsigned char* s_owner = gpmpc_ss.owner;   // cluster rank that sweeps pair pr
  if (tid == 0) {   // longest-processing-time deal: off-diagonal pairs sweep N^2 elements, diagonal ones half of that
    int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int pass = 0; pass < 2; pass++) {
      int pr = 0;
      for (int a = 0; a < E; a++)
        for (int b = a; b < E; b++, pr++) {
          if ((a == b) != (pass == 1)) continue;
          int best = 0;
          for (int c = 1; c < C; c++)
            if (load[c] < load[best]) best = c;
          load[best] += (a == b) ? 1 : 2;
          s_owner[pr] = (unsigned char)best;
        }
    }
  }
  for (int cand = cid;; cand += gridDim.x / C) {
    if (p.queue && C == 1) {
      if (tid == 0) s_next = atomicAdd(p.queue + 2, 1);
      __syncthreads();
      cand = s_next;
      __syncthreads();
    }
    if (cand >= p.B) break;
    // ================================================================== candidate init
    if (p.mode == 0) {
      if (tid < E) {
        double v = p.obs_mu[(p.per_cand_init ? (size_t)cand * E : 0) + tid];
        s_mu[tid] = v;
        if (lead) p.states_mu[((size_t)cand * (H + 1)) * E + tid] = v;
      }
      if (tid < E * E) {
        double v = p.obs_var[(p.per_cand_init ? (size_t)cand * E * E : 0) + tid];
        s_s[tid] = v;
        if (lead) p.states_var[((size_t)cand * (H + 1)) * E * E + tid] = v;
      }
      if (tid < Na) {  // action mapping (normalization_action_mapper.py:21-23 / derivative_action_mapper.py:28-35)
        double cum = 0.0;
        for (int t = 0; t < H; t++) {
          double raw = p.actions_mpc[(size_t)cand * H * Na + t * Na + tid];
          double am;
          if (p.limit_change) {
            double mc = p.max_change[tid];
            raw = raw * 2.0 * mc - mc;
            if (t == 0) raw += p.action_prev[tid];
            cum += raw;
            am = fmin(fmax(cum, 0.0), 1.0);
          } else {
            am = raw;
          }
          s_am[t * Na + tid] = am;
          if (lead) p.actions_model[(size_t)cand * H * Na + t * Na + tid] = am;
        }
      }
    } else {
      if (tid < EV * EV) s_s[tid] = p.obs_var[(size_t)cand * EV * EV + tid];
    }
    __syncthreads();
    const int nsteps = (p.mode == 0) ? H : 1;
    for (int t = 1; t <= nsteps; t++) {
      // ================================================================ P0a: model input, stage cost
      if (tid < D) {
        double v;
        if (p.mode == 1) v = p.obs_mu[(size_t)cand * D + tid];
        else if (tid < E) v = s_mu[tid];
        else if (tid < E + Na) v = s_am[(t - 1) * Na + (tid - E)];
        else v = (double)(p.iter_ctrl + t - 1);   // gp_model.py:101-102 (un-normalised time index)
        s_m[tid] = v;
      }
      __syncthreads();
      GEN_CLK(0);
      // ================================================================ P0b: small matrices
      if (tid < E) {
        const int a = tid;
        double Ca[EV * EV], Ai[EV * EV], det, pl = 1.0;
        for (int e = 0; e < EV; e++)
          for (int f = 0; f < EV; f++) Ca[e * EV + f] = s_s[e * EV + f] + (e == f ? 1.0 / s_il2[a * D + e] : 0.0);
        spd_inv_det<EV>(Ca, Ai, det);
        for (int e = 0; e < EV; e++) pl *= s_il2[a * D + e];
        for (int e = 0; e < EV * EV; e++) s_A[a * EV * EV + e] = Ai[e];
        s_c[a] = s_s2[a] / sqrt(det * pl);          // gp_model.py:150 (det B = det(s+Lambda)/det Lambda)
      } else if (tid >= 32 && tid < 32 + P) {
        const int pr = tid - 32;
        double Rinv[EV * EV], Qm[EV * EV], detR;
        pair_matrices<EV>(s_s, s_Wd + pr * EV, Rinv, Qm, detR);
        for (int e = 0; e < EV * EV; e++) s_Q[pr * EV * EV + e] = Qm[e];
        s_detR[pr] = detR;
      } else if (tid == 96 && p.mode == 0) {   // stage cost of the current state (serial, overlaps the matrix work)
        double cmu, cvar;
        stage_cost(cv, E, Na, s_mu, s_s, s_am + (t - 1) * Na, cmu, cvar);
        s_r[t - 1] = -cmu;
        s_rv[t - 1] = cvar;
      } else if (tid == 100) {
        double chk = 0.0;
        for (int d = 0; d < D; d++) chk += s_m[d];
        for (int e = 0; e < EV * EV; e++) chk += s_s[e];
        s_int[1] = isfinite(chk) ? 0 : 1;
      }
      __syncthreads();
      GEN_CLK(1);
      // ================================================================ P1: nu, lb, kk  (gp_model.py:138-148,168)
      for (int o = tid; o < E * nOut; o += NT) s_out[o] = 0.0;
      for (int i = tid; i < NP; i += NT) {
        double nu[GPMPC_MAX_D];
#pragma unroll
        for (int d = 0; d < GPMPC_MAX_D; d++) {
          if (d < DP) {
            double v = (i < N && d < D) ? (p.x[(size_t)i * D + d] - s_m[d]) : 0.0;
            nu[d] = v;
            s_nu[i * DP + d] = v;
          } else {
            nu[d] = 0.0;
          }
        }
        for (int a = 0; a < E; a++) {
          const double* Aa = s_A + a * EV * EV;
          const double* la = s_il2 + a * D;
          double quad = 0.0, head = 0.0, tail = 0.0;
#pragma unroll
          for (int e = 0; e < EV; e++) {
            double r = 0.0;
#pragma unroll
            for (int f = 0; f < EV; f++) r = fma(Aa[e * EV + f], nu[f], r);
            quad = fma(nu[e], r, quad);
            head = fma(nu[e] * nu[e], la[e], head);
          }
#pragma unroll
          for (int d = EV; d < GPMPC_MAX_D; d++)
            if (d < D) tail = fma(nu[d] * nu[d], la[d], tail);
          double lb = 0.0, kv = 0.0;
          if (i < N) {
            lb = __ldg(p.beta + (size_t)a * NP + i) * exp2s((-0.5 * GPMPC_EXP2S_SCALE) * (quad + tail));
            kv = s_logs2[a] - 0.5 * (head + tail);
          }
          s_lb[a * NP + i] = lb;
          kk[a * NP + i] = kv;
        }
      }
      __syncthreads();
      GEN_CLK(2);
      // ================================================================ P2: O(N) moment sums
      // (round 2) lane per training POINT: a task = (GP a, state dimension e1) owns the sums that share the factor
      // lb_a,i nu_i,e1 -- Gam[e1][d] (D) and T[e1][k<=l] (PV) -- plus, for e1 = 0, h and g_d; the lane loads nu_i and
      // lb_a,i ONCE per point (4 shared-memory loads for ~25 FMAs; the lane-per-output form took 4 loads per FMA and was
      // bound by the shared-memory pipe: 42.6 k clocks per step), keeps the sums of its points in registers, and the warp
      // adds them up with halving exchanges.  Every output has exactly one owner: plain stores, no atomics.
      {
        constexpr int PVc = EV * (EV + 1) / 2;
        const int ntask = GRAD ? E * EV : E;   // (splitting a task's points over several warps was slower: 25 k vs 20.6 k clocks)
        for (int task = tid >> 5; task < ntask; task += NT >> 5) {
          const int a = GRAD ? task / EV : task, e1 = GRAD ? task - a * EV : 0;
          const double* lbp = s_lb + a * NP;
          double hs = 0.0, gs[GPMPC_MAX_D], gam[GPMPC_MAX_D], tt[PVc];
#pragma unroll
          for (int d = 0; d < GPMPC_MAX_D; d++) { gs[d] = 0.0; gam[d] = 0.0; }
#pragma unroll
          for (int k = 0; k < PVc; k++) tt[k] = 0.0;
          for (int i = lane; i < NP; i += 32) {       // padded points carry lb = 0
            const double lb = lbp[i];
            double nu[GPMPC_MAX_D];
#pragma unroll
            for (int d = 0; d < GPMPC_MAX_D; d++) nu[d] = (d < DP) ? s_nu[i * DP + d] : 0.0;
            if (e1 == 0) {
              hs += lb;
#pragma unroll
              for (int d = 0; d < GPMPC_MAX_D; d++) gs[d] = fma(lb, nu[d], gs[d]);   // nu = 0 beyond D
            }
            if (GRAD) {
              double ne = nu[0];
#pragma unroll
              for (int e = 1; e < EV; e++) ne = (e == e1) ? nu[e] : ne;
              const double w = lb * ne;
#pragma unroll
              for (int d = 0; d < GPMPC_MAX_D; d++) gam[d] = fma(w, nu[d], gam[d]);
              int pr = 0;
#pragma unroll
              for (int k = 0; k < EV; k++) {
                const double wk = w * nu[k];
#pragma unroll
                for (int l = k; l < EV; l++) { tt[pr] = fma(wk, nu[l], tt[pr]); pr++; }
              }
            }
          }
          double* out = s_out + a * nOut;
          // warp totals, 16 values per round of halving exchanges (warp_reduce_multi); value idx lands in the even lanes
          if (e1 == 0) {
            double v16[16];
#pragma unroll
            for (int k = 0; k < 16; k++) v16[k] = (k == 0) ? hs : gs[k - 1];      // h, g_0 .. g_14
            int idx;
            double tot = warp_reduce_multi<16>(v16, lane, idx);
            if ((lane & 1) == 0 && idx < 1 + D) out[idx] = tot;
            if (D == GPMPC_MAX_D) {                                                // g_15 (D = 16 only)
              tot = warp_sum(gs[GPMPC_MAX_D - 1]);
              if (lane == 0) out[D] = tot;
            }
          }
          if (GRAD) {
            {
              double v16[16];
#pragma unroll
              for (int k = 0; k < 16; k++) v16[k] = gam[k];
              int idx;
              const double tot = warp_reduce_multi<16>(v16, lane, idx);
              if ((lane & 1) == 0 && idx < D) out[1 + D + e1 * D + idx] = tot;
            }
            constexpr int NCH = (PVc + 15) / 16;
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
              double v16[16];
#pragma unroll
              for (int k = 0; k < 16; k++) v16[k] = (16 * ch + k < PVc) ? tt[(16 * ch + k < PVc) ? 16 * ch + k : 0] : 0.0;
              int idx;
              const double tot = warp_reduce_multi<16>(v16, lane, idx);
              if ((lane & 1) == 0 && 16 * ch + idx < PVc) out[1 + D + EV * D + e1 * PVc + 16 * ch + idx] = tot;
            }
          }
        }
      }
      __syncthreads();
      GEN_CLK(3);
      // ================================================================ P2b: mean / V per GP (gp_model.py:152-153)
      if (tid < E) {
        const int a = tid;
        const double* out = s_out + a * nOut;
        const double* Aa = s_A + a * EV * EV;
        const double* la = s_il2 + a * D;
        const double h = out[0], c = s_c[a];
        const double* g = out + 1;
        s_M[a] = c * h;
        double Ag[EV];
        for (int e = 0; e < EV; e++) {
          double v = 0.0;
          for (int f = 0; f < EV; f++) v += Aa[e * EV + f] * g[f];
          Ag[e] = v;
          s_V[a * D + e] = c * v;
        }
        for (int d = EV; d < D; d++) s_V[a * D + d] = c * g[d] * la[d];
        if (p.mode == 1) {
          if (p.stepM) p.stepM[(size_t)cand * E + a] = c * h;
          if (p.stepV)
            for (int d = 0; d < D; d++) p.stepV[((size_t)cand * D + d) * E + a] = s_V[a * D + d];
        }
        if (GRAD && lead) {
          double* rec = p.records + ((size_t)cand * H + (t - 1)) * RL.size + RL.offGp + a * RL.gpStride;
          const double* Gam = out + 1 + D;            // [e][d]
          const double* T = out + 1 + D + EV * D;     // [e][kl]
          int w = 0;
          rec[w++] = h;
          rec[w++] = c;
          for (int e = 0; e < EV; e++) rec[w++] = g[e];
          for (int d = 0; d < D; d++) rec[w++] = (d < EV) ? Ag[d] : g[d] * la[d];       // dh/dm
          for (int k = 0; k < EV; k++)
            for (int l = 0; l < EV; l++) rec[w++] = -0.5 * Gam[k * D + l];               // dh/dA
          for (int e = 0; e < EV; e++)
            for (int d = 0; d < D; d++) {                                                 // dg_e/dm_d
              double v;
              if (d < EV) {
                v = (e == d) ? -h : 0.0;
                for (int k = 0; k < EV; k++) v += Gam[e * D + k] * Aa[k * EV + d];
              } else {
                v = Gam[e * D + d] * la[d];
              }
              rec[w++] = v;
            }
          for (int e = 0; e < EV * PV; e++) rec[w++] = -0.5 * T[e];                       // dg_e/dA_kl
        }
      }
      __syncthreads();
      GEN_CLK(4);
      // ================================================================ P3: O(P N^2) covariance sums
      // Groups of G pairs (all of them when their column terms fit the shared memory).  Per group: column terms kap'_j
      // per pair -> ONE sweep phase over the chunks of all its pairs (static weighted split into contiguous runs per warp)
      // -> column-sum (gam) parts of the gradient -> per-pair finalisation.  Cluster mode: a CTA handles the pairs it owns.
      for (int g0 = 0; g0 < P; g0 += G) {
        const int gn = min(G, P - g0);
        for (int o = tid; o < gn * NP; o += NT) {
          const int pl = o / NP, j = o - pl * NP;
          const int pr = g0 + pl;
          if (C > 1 && s_owner[pr] != crank) continue;
          const int ab = s_int[2 + pr], a = ab >> 4, b = ab & 15;
          double kap = 0.0;
          if (j < N) {
            const double* Qm = s_Q + pr * EV * EV;
            double z[EV];
#pragma unroll
            for (int e = 0; e < EV; e++) z[e] = s_nu[j * DP + e] * s_il2[b * D + e];
            kap = kk[b * NP + j];
#pragma unroll
            for (int e = 0; e < EV; e++) {
              double r = 0.0;
#pragma unroll
              for (int f = 0; f < EV; f++) r = fma(Qm[e * EV + f], z[f], r);
              kap = fma(z[e], r, kap);
            }
          }
          if (a != b) {
            // off-diagonal pairs: the coefficient beta_b,j moves into the exponent (log|beta|) and its sign into the magic
            // constant of the exp's range reduction (gen_cols, exp2s_x4_signed); padded columns get exp(-huge) = 0
            // (colcoef = {SCALE log|beta_b,j|, magic constant with the sign}: candidate independent, formed once per launch)
            const double2 cc = __ldg(reinterpret_cast<const double2*>(p.colcoef) + (size_t)b * NP + j);
            kap = (j < N) ? fma(GPMPC_EXP2S_SCALE, kap, cc.x) : -1.0e300;
            s_kap[2 * o + 1] = cc.y;
          } else {
            kap *= GPMPC_EXP2S_SCALE;
            s_kap[2 * o + 1] = __ldg(p.beta + (size_t)b * NP + j);   // diagonal pairs: the coefficient itself rides in the record
          }
          s_kap[2 * o] = kap;
        }
        for (int o = tid; o < gn * L.paccN; o += NT) s_pacc[o] = 0.0;
        __syncthreads();
        GEN_CLK(5);
        const long long sw0_ = p.dbg_clk ? clock64() : 0;
        {
          // chunks of CH columns x 64 rows: an off-diagonal pair has nrb * (NP / CH) of them (row-major), a diagonal pair
          // cpt (nrb - I) per row block I (upper tile triangle).  A diagonal chunk takes ~1.4x as long as an off-diagonal
          // one (iK stream from L2 on top of 3 more float64 instructions per element; measured per warp,
          // profiles/r02c_general_sweep_per_warp.txt), so the two kinds are split SEPARATELY: every warp gets an equal
          // contiguous run of the diagonal pairs' chunks and one of the off-diagonal pairs' chunks -- balanced whatever the
          // cost ratio; even warps start with their diagonal run, odd warps with the off-diagonal one (mixed L2 traffic).
          const int CH = p.seg, nrb = NP / 64, cpr = NP / CH, cpt = 64 / CH;
          const bool rows4 = GEN_ROWS4 && (NP % 128 == 0);          // off-diagonal pairs: 128-row warp tiles (gen_item4)
          const int nOff = (rows4 ? nrb / 2 : nrb) * cpr, nDia = cpt * nrb * (nrb + 1) / 2;
#pragma unroll 1
          for (int ph = 0; ph < 2; ph++) {
            const bool dia = ((ph ^ warp) & 1) == 0;
            const int n = dia ? nDia : nOff;
            int cnt = 0;
            for (int pl = 0; pl < gn; pl++) {
              const int pr = g0 + pl, ab = s_int[2 + pr];
              if (C > 1 && s_owner[pr] != crank) continue;
              cnt += (((ab >> 4) == (ab & 15)) == dia) ? 1 : 0;
            }
            const int T = cnt * n;
            const int clo = (int)((long long)T * warp / nwarps), chi = (int)((long long)T * (warp + 1) / nwarps);
            int cw = 0;
            for (int pl = 0; pl < gn; pl++) {
              const int pr = g0 + pl, ab = s_int[2 + pr], a = ab >> 4, b = ab & 15;
              if ((C > 1 && s_owner[pr] != crank) || ((a == b) != dia)) continue;
              int c0 = max(clo, cw) - cw;
              const int c1 = min(chi, cw + n) - cw;
              cw += n;
              if (c0 >= c1) continue;
              const double* Qm = s_Q + pr * EV * EV;
              if (dia) {
                int I = 0, base = 0;
                while (c0 < c1) {
                  while (c0 >= base + cpt * (nrb - I)) { base += cpt * (nrb - I); I++; }
                  const int ce = min(c1, base + cpt * (nrb - I));
                  gen_item<EV, GRAD, true>(p, s_nu, s_kap + 2 * pl * NP, Qm, s_il2 + a * D, s_il2 + b * D, kk + a * NP,
                                           p.beta + (size_t)a * NP, p.beta + (size_t)b * NP, p.iK + (size_t)a * NP * NP, I,
                                           64 * I + CH * (c0 - base), 64 * I + CH * (ce - base), lane, g_gam + pl * NP,
                                           s_pacc + pl * L.paccN, s_colred);
                  c0 = ce;
                }
              } else if (rows4) {
                while (c0 < c1) {
                  const int I = c0 / cpr, ce = min(c1, (I + 1) * cpr);
                  gen_item4<EV, GRAD>(p, s_nu, s_kap + 2 * pl * NP, Qm, s_il2 + a * D, s_il2 + b * D, kk + a * NP,
                                      p.beta + (size_t)a * NP, I, CH * (c0 - I * cpr), CH * (ce - I * cpr), lane,
                                      g_gam + pl * NP, s_pacc + pl * L.paccN, s_colred);
                  c0 = ce;
                }
              } else {
                while (c0 < c1) {
                  const int I = c0 / cpr, ce = min(c1, (I + 1) * cpr);
                  gen_item<EV, GRAD, false>(p, s_nu, s_kap + 2 * pl * NP, Qm, s_il2 + a * D, s_il2 + b * D, kk + a * NP,
                                            p.beta + (size_t)a * NP, p.beta + (size_t)b * NP, nullptr, I, CH * (c0 - I * cpr),
                                            CH * (ce - I * cpr), lane, g_gam + pl * NP, s_pacc + pl * L.paccN, s_colred);
                  c0 = ce;
                }
              }
            }
          }
        }
        if (p.dbg_clk && blockIdx.x == 0 && lane == 0) p.dbg_clk[32 + warp] += clock64() - sw0_;   // tuning aid: this warp's share of the sweep
        if (GRAD) __threadfence();   // the column sums are reductions at L2: make them visible before they are loaded below
        __syncthreads();
        GEN_CLK(6);
        if (GRAD) {
          // column-sum parts: S_raw += sum_j gam_j (diagonal pairs), dS/dm += sum_j gam_j lb nu_j, dS/dQ += sum_j gam_j zb zb^T
          // (L2 loads: the sums were formed by reductions at L2; the scratch is re-zeroed for the next step)
          // one warp per pair (round-robin), lanes over the training points: ONE warp-level reduction per pair
          constexpr int PVc = EV * (EV + 1) / 2;
          for (int pl = warp; pl < gn; pl += nwarps) {
            const int pr = g0 + pl, ab = s_int[2 + pr], a = ab >> 4, b = ab & 15;
            if (C > 1 && s_owner[pr] != crank) continue;
            const double* lbv = s_il2 + b * D;
            double accG = 0.0, vs[GPMPC_MAX_D], vp[PVc];
#pragma unroll
            for (int d = 0; d < GPMPC_MAX_D; d++) vs[d] = 0.0;
#pragma unroll
            for (int e = 0; e < PVc; e++) vp[e] = 0.0;
            double* gp = g_gam + pl * NP;
            for (int j0 = lane; j0 < NP; j0 += 128) {
              double gam4[4];
#pragma unroll
              for (int q = 0; q < 4; q++) gam4[q] = (j0 + 32 * q < NP) ? __ldcg(gp + j0 + 32 * q) : 0.0;   // NP is a multiple of 64
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const int j = j0 + 32 * q;
                if (j >= NP) continue;
                gp[j] = 0.0;
                const double gam = gam4[q];
                accG += gam;
                const double* nj = s_nu + j * DP;
                double zb[EV];
#pragma unroll
                for (int d = 0; d < GPMPC_MAX_D; d++)
                  if (d < D) vs[d] = fma(gam * lbv[d], nj[d], vs[d]);
#pragma unroll
                for (int e = 0; e < EV; e++) zb[e] = nj[e] * lbv[e];
                int q2 = 0;
#pragma unroll
                for (int k = 0; k < EV; k++) {
                  const double gk = gam * zb[k];
#pragma unroll
                  for (int l = k; l < EV; l++) { vp[q2] = fma(gk, zb[l], vp[q2]); q2++; }
                }
              }
            }
            if (a == b) {
              accG = warp_sum(accG);
              if (lane == 0) atomicAdd(s_pacc + pl * L.paccN, accG);
            }
            gen_warp_sums_add<PVc>(vp, vs, D, lane, s_pacc + pl * L.paccN + 1);
          }
          __syncthreads();
        }
        GEN_CLK(7);
        if (tid < gn && (C == 1 || s_owner[g0 + tid] == crank)) {
          const int pl = tid, pr = g0 + pl, ab = s_int[2 + pr], a = ab >> 4, b = ab & 15;
          double* acc = s_pacc + pl * L.paccN;
          double Sr = acc[0];
          if (!GRAD && a == b) Sr *= 2.0;
          s_Sraw[pr] = Sr;
          if (GRAD) {
            // dS/dm[:EV] -= 2 W (Q ybar), ybar = sum_ij w (z_a,i + z_b,j) = gm[:EV] before the correction
            const double* Qm = s_Q + pr * EV * EV;
            double yb[EV];
            if (a == b)  // upper-triangle sweep: (rho_up + gam_up) is the full row sum only once
              for (int e = 1; e < L.paccN; e++) acc[e] *= 2.0;
            for (int e = 0; e < EV; e++) yb[e] = acc[1 + e];
            for (int e = 0; e < EV; e++) {
              double v = 0.0;
              for (int f = 0; f < EV; f++) v += Qm[e * EV + f] * yb[f];
              acc[1 + e] -= 2.0 * s_Wd[pr * EV + e] * v;
            }
            double* rec = p.records + ((size_t)cand * H + (t - 1)) * RL.size + RL.offPair + pr * RL.pairStride;
            rec[0] = Sr;
            rec[1] = s_detR[pr];
            for (int d = 0; d < D; d++) rec[2 + d] = acc[1 + d];
            int q = 0;                                   // dS/dQ is accumulated as its upper triangle (it is symmetric)
            for (int k = 0; k < EV; k++)
              for (int l = k; l < EV; l++) {
                const double v = acc[1 + D + q];
                q++;
                rec[2 + D + k * EV + l] = v;
                rec[2 + D + l * EV + k] = v;
              }
          }
        }
        __syncthreads();
      }
      if (C > 1) {   // exchange the S_raw of the pairs: owners publish, cluster barrier, everybody reads all of them
        double* ex = p.ws_cl + ((size_t)cid * 2 + (gstep & 1)) * 64;
        for (int pr = tid; pr < P; pr += NT)
          if (s_owner[pr] == crank) ex[pr] = s_Sraw[pr];
        __threadfence();
        uni_cluster_sync();
        for (int pr = tid; pr < P; pr += NT) s_Sraw[pr] = __ldcg(ex + pr);
        __syncthreads();
      }
      gstep++;
      GEN_CLK(8);
      // ================================================================ P4: S, recurrence (gp_model.py:176-178, :105-108)
      if (tid == 0) {
        double S[GPMPC_MAX_EV * GPMPC_MAX_EV];
        const bool bad = s_int[1] != 0;
        for (int a = 0; a < E; a++)
          for (int b = a; b < E; b++) {
            const int pr = pair_index(a, b, E);
            double v = s_Sraw[pr] / sqrt(s_detR[pr]) - s_M[a] * s_M[b] + (a == b ? s_s2[a] : 0.0);
            if (bad) v = nan("");
            S[a * E + b] = v;
            S[b * E + a] = v;
          }
        if (p.mode == 1) {
          if (p.stepS)
            for (int e = 0; e < E * E; e++) p.stepS[(size_t)cand * E * E + e] = S[e];
        } else {
          if (GRAD && lead) {
            double* rec = p.records + ((size_t)cand * H + (t - 1)) * RL.size;
            for (int a = 0; a < E; a++) rec[RL.offM + a] = s_M[a];
            for (int a = 0; a < E; a++)
              for (int e = 0; e < E; e++) rec[RL.offV + a * E + e] = s_V[a * D + e];
          }
          double sv[GPMPC_MAX_EV * GPMPC_MAX_EV], sn[GPMPC_MAX_EV * GPMPC_MAX_EV];
          for (int e = 0; e < E; e++)
            for (int a = 0; a < E; a++) {
              double v = 0.0;
              for (int k = 0; k < E; k++) v += s_s[e * E + k] * s_V[a * D + k];
              sv[e * E + a] = v;
            }
          for (int e = 0; e < E; e++)
            for (int f = 0; f < E; f++) sn[e * E + f] = S[e * E + f] + s_s[e * E + f] + sv[e * E + f] + sv[f * E + e];
          for (int e = 0; e < E; e++) {
            double v = s_mu[e] + s_M[e];
            if (bad) v = nan("");
            s_mu[e] = v;
            if (lead) p.states_mu[((size_t)cand * (H + 1) + t) * E + e] = v;
          }
          for (int e = 0; e < E * E; e++) {
            s_s[e] = sn[e];
            if (lead) p.states_var[((size_t)cand * (H + 1) + t) * E * E + e] = sn[e];
          }
        }
      }
      __syncthreads();
      GEN_CLK(9);
    }  // steps
    // ================================================================== terminal cost + LCB (controller :270-276)
    if (p.mode == 0 && tid == 0 && lead) {
      double cmu, cvar;
      terminal_cost(cv, E, s_mu, s_s, cmu, cvar);
      s_r[H] = -cmu;
      s_rv[H] = cvar;
      double acc = 0.0;
      for (int t = 0; t <= H; t++) {
        double ucb = s_r[t] + p.kappa * sqrt(s_rv[t]);
        if (p.clip) ucb = fmin(ucb, 0.0);
        acc += ucb;
        p.rewards[(size_t)cand * (H + 1) + t] = s_r[t];
        p.rewards_var[(size_t)cand * (H + 1) + t] = s_rv[t];
      }
      p.cost[cand] = -acc / (double)(H + 1);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// reverse sweep: one thread per candidate, small-matrix algebra only (tests/algo_spec.py
// rollout()/step_backward() is the executable spec of this kernel).
// ---------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(128) backward_kernel(const BackwardParams p) {
  const int cand = blockIdx.x * blockDim.x + threadIdx.x;
  if (cand >= p.B) return;
  const int D = p.D, Na = p.Na, H = p.H, Dc = E + Na;
  const RecLayout RL = rec_layout(E, D);
  const int P = RL.P;
  const double wmu = 1.0 / (double)(H + 1);
  double mu_bar[E], s_bar[E * E];
  const double* mus = p.states_mu + (size_t)cand * (H + 1) * E;
  const double* vars = p.states_var + (size_t)cand * (H + 1) * E * E;
  const double* rvs = p.rewards_var + (size_t)cand * (H + 1);
  const double* ams = p.actions_model + (size_t)cand * H * Na;
  double* gout = p.grad + (size_t)cand * H * Na;
  {  // terminal (setpoint_distance_reward_mapper.py:124-142)
    const double* mu = mus + (size_t)H * E;
    const double* s = vars + (size_t)H * E * E;
    const double wv = -p.kappa * wmu * 0.5 / sqrt(rvs[H]);
    double e[E], We[E], sWe[E];
    for (int d = 0; d < E; d++) e[d] = mu[d] - p.c_target[d];
    for (int d = 0; d < E; d++) {
      double v = 0.0;
      for (int k = 0; k < E; k++) v += p.c_WT[d * E + k] * e[k];
      We[d] = v;
    }
    for (int d = 0; d < E; d++) {
      double v = 0.0;
      for (int k = 0; k < E; k++) v += s[d * E + k] * We[k];
      sWe[d] = v;
    }
    for (int d = 0; d < E; d++) {
      double v = 0.0;
      for (int k = 0; k < E; k++) v += p.c_WT[k * E + d] * sWe[k];
      mu_bar[d] = wmu * 2.0 * We[d] + wv * 8.0 * v;
    }
    // s_bar = wmu WT^T + wv (4 WT^T s^T WT^T + 4 We We^T)
    double t1[E * E];
    for (int i = 0; i < E; i++)
      for (int k = 0; k < E; k++) {
        double v = 0.0;
        for (int l = 0; l < E; l++) v += p.c_WT[l * E + i] * s[k * E + l];
        t1[i * E + k] = v;  // (WT^T s^T)[i][k]
      }
    for (int i = 0; i < E; i++)
      for (int k = 0; k < E; k++) {
        double v = 0.0;
        for (int l = 0; l < E; l++) v += t1[i * E + l] * p.c_WT[k * E + l];
        s_bar[i * E + k] = wmu * p.c_WT[k * E + i] + wv * (4.0 * v + 4.0 * We[i] * We[k]);
      }
  }
  for (int t = H; t >= 1; t--) {
    const double* rec = p.records + ((size_t)cand * H + (t - 1)) * RL.size;
    const double* sp = vars + (size_t)(t - 1) * E * E;
    const double* mup = mus + (size_t)(t - 1) * E;
    const double* am = ams + (size_t)(t - 1) * Na;
    const double* Mrec = rec + RL.offM;
    const double* Vrec = rec + RL.offV;  // [a][e]
    double U[E * E], Vb[E * E];
    for (int i = 0; i < E; i++)
      for (int k = 0; k < E; k++) U[i * E + k] = s_bar[i * E + k] + s_bar[k * E + i];
    // V_bar[a][e] = sum_k sp[k][e] U[k][a]
    for (int a = 0; a < E; a++)
      for (int e = 0; e < E; e++) {
        double v = 0.0;
        for (int k = 0; k < E; k++) v += sp[k * E + e] * U[k * E + a];
        Vb[a * E + e] = v;
      }
    double m_bar[GPMPC_MAX_D], sp_bar[E * E], M_bar[E];
    for (int d = 0; d < D; d++) m_bar[d] = 0.0;
    for (int e = 0; e < E * E; e++) sp_bar[e] = 0.0;
    for (int a = 0; a < E; a++) M_bar[a] = mu_bar[a];
    // ---- pairs
    for (int a = 0; a < E; a++)
      for (int b = a; b < E; b++) {
        const int pr = pair_index(a, b, E);
        const double* pe = rec + RL.offPair + pr * RL.pairStride;
        const double Sraw = pe[0], detR = pe[1];
        const double* gm = pe + 2;
        const double* gQ = pe + 2 + D;
        const double sb = s_bar[a * E + b] + (a != b ? s_bar[b * E + a] : 0.0);
        M_bar[a] -= sb * Mrec[b];
        M_bar[b] -= sb * Mrec[a];
        double Wd[E], Rinv[E * E], Q[E * E], dR;
        for (int e = 0; e < E; e++) Wd[e] = p.il2[a * D + e] + p.il2[b * D + e];
        pair_matrices<E>(sp, Wd, Rinv, Q, dR);
        const double rs = 1.0 / sqrt(detR);
        const double Sraw_bar = sb * rs;
        const double detR_bar = -0.5 * sb * Sraw * rs / detR;
        for (int d = 0; d < D; d++) m_bar[d] += Sraw_bar * gm[d];
        // RitQb = Rinv^T (Sraw_bar gQ)
        double RQ[E * E];
        for (int i = 0; i < E; i++)
          for (int k = 0; k < E; k++) {
            double v = 0.0;
            for (int l = 0; l < E; l++) v += Rinv[l * E + i] * gQ[l * E + k];
            RQ[i * E + k] = Sraw_bar * v;
          }
        for (int i = 0; i < E; i++)
          for (int k = 0; k < E; k++) {
            double v = 0.0;
            for (int l = 0; l < E; l++) v += RQ[i * E + l] * Q[k * E + l];   // (RQ Q^T)[i][k]
            sp_bar[i * E + k] += 0.5 * RQ[i * E + k] - v * Wd[k] + detR_bar * detR * Rinv[k * E + i] * Wd[k];
          }
      }
    // ---- per GP
    for (int a = 0; a < E; a++) {
      const double* ge = rec + RL.offGp + a * RL.gpStride;
      const double h = ge[0], c = ge[1];
      const double* gE = ge + 2;
      const double* dh_dm = gE + E;
      const double* dh_dA = dh_dm + D;
      const double* dg_dm = dh_dA + E * E;
      const double* dg_dA = dg_dm + E * D;
      double Ca[E * E], A[E * E], det;
      for (int e = 0; e < E; e++)
        for (int f = 0; f < E; f++) Ca[e * E + f] = sp[e * E + f] + (e == f ? 1.0 / p.il2[a * D + e] : 0.0);
      spd_inv_det<E>(Ca, A, det);
      double Ag[E], g_bar[E], A_bar[E * E];
      double c_bar = M_bar[a] * h;
      for (int e = 0; e < E; e++) {
        double v = 0.0;
        for (int f = 0; f < E; f++) v += A[e * E + f] * gE[f];
        Ag[e] = v;
        c_bar += Vb[a * E + e] * v;
      }
      const double h_bar = M_bar[a] * c;
      for (int e = 0; e < E; e++) {
        double v = 0.0;
        for (int f = 0; f < E; f++) v += A[f * E + e] * Vb[a * E + f];
        g_bar[e] = c * v;
      }
      for (int k = 0; k < E; k++)
        for (int l = 0; l < E; l++) A_bar[k * E + l] = c * Vb[a * E + k] * gE[l] + h_bar * dh_dA[k * E + l];
      for (int e = 0; e < E; e++) {
        int w = 0;
        for (int k = 0; k < E; k++)
          for (int l = k; l < E; l++) {
            double v = g_bar[e] * dg_dA[e * P + w];
            w++;
            A_bar[k * E + l] += v;
            if (l != k) A_bar[l * E + k] += v;
          }
      }
      for (int d = 0; d < D; d++) {
        double v = h_bar * dh_dm[d];
        for (int e = 0; e < E; e++) v += g_bar[e] * dg_dm[e * D + d];
        m_bar[d] += v;
      }
      // s_bar += -1/2 c_bar c A^T - A^T A_bar A^T
      double t1[E * E];
      for (int i = 0; i < E; i++)
        for (int k = 0; k < E; k++) {
          double v = 0.0;
          for (int l = 0; l < E; l++) v += A[l * E + i] * A_bar[l * E + k];
          t1[i * E + k] = v;
        }
      for (int i = 0; i < E; i++)
        for (int k = 0; k < E; k++) {
          double v = 0.0;
          for (int l = 0; l < E; l++) v += t1[i * E + l] * A[k * E + l];
          sp_bar[i * E + k] += -0.5 * c_bar * c * A[k * E + i] - v;
        }
    }
    // symmetrise the step adjoint, add the recurrence terms
    double X[E * E];
    for (int i = 0; i < E; i++)
      for (int k = 0; k < E; k++) {
        double v = 0.0;
        for (int a = 0; a < E; a++) v += U[i * E + a] * Vrec[a * E + k];
        X[i * E + k] = v;
      }
    double nsb[E * E], nmu[E];
    for (int i = 0; i < E; i++)
      for (int k = 0; k < E; k++)
        nsb[i * E + k] = 0.5 * (sp_bar[i * E + k] + sp_bar[k * E + i]) + 0.5 * (s_bar[i * E + k] + s_bar[k * E + i]) +
                         0.5 * (X[i * E + k] + X[k * E + i]);
    for (int e = 0; e < E; e++) nmu[e] = mu_bar[e] + m_bar[e];
    double a_bar[GPMPC_MAX_D];
    for (int k = 0; k < Na; k++) a_bar[k] = m_bar[E + k];
    // ---- stage cost at t-1 (setpoint_distance_reward_mapper.py:12-68)
    {
      const double wv = -p.kappa * wmu * 0.5 / sqrt(rvs[t - 1]);
      double e[GPMPC_MAX_D], We[GPMPC_MAX_D], sWe[E];
      for (int d = 0; d < Dc; d++) e[d] = (d < E ? mup[d] : am[d - E]) - p.c_target[d];
      for (int d = 0; d < Dc; d++) {
        double v = 0.0;
        for (int k = 0; k < Dc; k++) v += p.c_W[d * Dc + k] * e[k];
        We[d] = v;
      }
      for (int i = 0; i < E; i++) {
        double v = 0.0;
        for (int k = 0; k < E; k++) v += sp[i * E + k] * We[k];
        sWe[i] = v;
      }
      for (int d = 0; d < Dc; d++) {
        double v = 0.0;
        for (int i = 0; i < E; i++) v += p.c_W[i * Dc + d] * sWe[i];   // (W[:E,:]^T (s Wse))[d]
        double gd = wmu * 2.0 * We[d] + wv * 8.0 * v;
        if (d < E) nmu[d] += gd; else a_bar[d - E] += gd;
      }
      double t1[E * E];
      for (int i = 0; i < E; i++)
        for (int k = 0; k < E; k++) {
          double v = 0.0;
          for (int l = 0; l < E; l++) v += p.c_W[l * Dc + i] * sp[k * E + l];
          t1[i * E + k] = v;
        }
      for (int i = 0; i < E; i++)
        for (int k = 0; k < E; k++) {
          double v = 0.0;
          for (int l = 0; l < E; l++) v += t1[i * E + l] * p.c_W[k * Dc + l];
          const double full_ik = wmu * p.c_W[k * Dc + i] + wv * (4.0 * v + 4.0 * We[i] * We[k]);
          nsb[i * E + k] += 0.5 * full_ik;   // symmetrised: (F + F^T) / 2
          nsb[k * E + i] += 0.5 * full_ik;
        }
      if (p.use_constraints) {
        const double rt2 = 1.4142135623730951, ispi = 0.5641895835477563;
        for (int d = 0; d < E; d++) {
          double sig = sp[d * E + d];
          double zmin = (p.c_smin[d] - mup[d]) / (sig * rt2), zmax = (p.c_smax[d] - mup[d]) / (sig * rt2);
          double pmin = exp(-zmin * zmin) * ispi, pmax = exp(-zmax * zmax) * ispi;
          nmu[d] += wmu * (pmin - pmax) * (-1.0 / (sig * rt2));
          nsb[d * E + d] += wmu * (pmin * (-zmin / sig) - pmax * (-zmax / sig));
        }
      }
    }
    for (int k = 0; k < Na; k++) gout[(size_t)(t - 1) * Na + k] = a_bar[k];
    for (int e = 0; e < E; e++) mu_bar[e] = nmu[e];
    for (int e = 0; e < E * E; e++) s_bar[e] = nsb[e];
  }
  if (p.limit_change) {  // derivative_action_mapper.py:28-35: reverse cumsum, straight-through clamp
    for (int k = 0; k < Na; k++) {
      double cum = 0.0;
      for (int t = H - 1; t >= 0; t--) {
        cum += gout[(size_t)t * Na + k];
        gout[(size_t)t * Na + k] = cum * 2.0 * p.max_change[k];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-EV launchers; each gpmpc_inst_evN.cu instantiates one EV so that the build parallelises
// ---------------------------------------------------------------------------------------------
template <int EV>
cudaError_t launch_rollout_inst(bool grad, const RolloutParams& p, int grid, int threads, size_t smem, cudaStream_t st) {
  cudaError_t e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (p.mode == 0 && p.cluster > 1) {   // p.cluster consecutive CTAs form a thread-block cluster
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  if (grad) {
    e = cudaFuncSetAttribute(rollout_kernel<EV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaLaunchKernelEx(&cfg, rollout_kernel<EV, true>, p);
  } else {
    e = cudaFuncSetAttribute(rollout_kernel<EV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaLaunchKernelEx(&cfg, rollout_kernel<EV, false>, p);
  }
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

// Clusters of `cluster` CTAs of the general kernel the device holds at once (cudaOccupancyMaxActiveClusters: a cluster
// must fit one GPC, so this is less than SMs / cluster -- 16 clusters of 8 single-CTA-per-SM kernels do NOT fit 148 SMs).
template <int EV>
cudaError_t max_clusters_rollout_inst(bool grad, int cluster, int threads, size_t smem, int* nclusters) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster * 64);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
  if (grad) {
    e = cudaFuncSetAttribute(rollout_kernel<EV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveClusters(nclusters, rollout_kernel<EV, true>, &cfg);
  }
  e = cudaFuncSetAttribute(rollout_kernel<EV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveClusters(nclusters, rollout_kernel<EV, false>, &cfg);
}

template <int E>
cudaError_t launch_backward_inst(const BackwardParams& p, cudaStream_t st) {
  const int blk = 128, grid = (p.B + blk - 1) / blk;
  backward_kernel<E><<<grid, blk, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace gpmpc
