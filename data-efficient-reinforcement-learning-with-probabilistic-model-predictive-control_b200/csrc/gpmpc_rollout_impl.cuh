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
// The hot loop: one warp, rows {64 I + lane, 64 I + 32 + lane}, columns [jbeg, jend).
//   t_ij = kap_i + kap_j + u_i . nu_j ;  w_ij = (beta_a,i beta_b,j - [a==b] iK_a,ij) exp(t_ij)
// gp_model.py:161-175 (X, X2, Q, maha, k, L, beta L beta, iK * L) collapsed to one exponent.
// Diagonal pairs (a == b) sweep only the tiles on or above the diagonal (diagonal tile: half weights); the caller doubles.
// ---------------------------------------------------------------------------------------------
template <int EV, bool GRAD, bool DIAG>
__device__ __forceinline__ void pair_item(const RolloutParams& p, const double* __restrict__ s_nu,
                                          const double* __restrict__ s_kapj, const double* __restrict__ Qm,
                                          const double* __restrict__ il2a, const double* __restrict__ il2b,
                                          const double* __restrict__ kka, const double* __restrict__ beta_a,
                                          const double* __restrict__ beta_b, const double* __restrict__ iKa,
                                          int I, int jbeg, int jend, int lane, double* s_gam, double* s_rho,
                                          double* s_xi, double* s_acc, unsigned s_tab) {
  const int NP = p.NP, DP = p.DP;
  const int i0 = 64 * I + lane, i1 = i0 + 32;
  double u0[EV], u1[EV], kr0, kr1;
  {
    double z0[EV], z1[EV];
#pragma unroll
    for (int e = 0; e < EV; e++) {
      z0[e] = s_nu[i0 * DP + e] * il2a[e];
      z1[e] = s_nu[i1 * DP + e] * il2a[e];
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
      u0[e] = (2.0 * GPMPC_EXP2S_SCALE) * q0 * il2b[e];   // exponent in table units (exp2s; s_kapj is scaled too)
      u1[e] = (2.0 * GPMPC_EXP2S_SCALE) * q1 * il2b[e];
    }
    kr0 *= GPMPC_EXP2S_SCALE;
    kr1 *= GPMPC_EXP2S_SCALE;
  }
  const double bi0 = __ldg(beta_a + i0), bi1 = __ldg(beta_a + i1);
  double rho0 = 0.0, rho1 = 0.0;
  double xi0[EV], xi1[EV];
#pragma unroll
  for (int e = 0; e < EV; e++) { xi0[e] = 0.0; xi1[e] = 0.0; }

  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    // diagonal pairs: the diagonal tile is swept in full with HALF weights instead of its upper triangle -- w is symmetric
    // and every consumer of (rho, gam, xi) is invariant under that swap (see uni_bwd_item), so no element masks
    const double wgt = (DIAG && j0 < 64 * I + 64) ? 0.5 : 1.0;
    double v[8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const int j = j0 + jj;
      double nj[EV];
#pragma unroll
      for (int e = 0; e < EV; e++) nj[e] = s_nu[j * DP + e];
      const double kj = s_kapj[j];
      const double bj = __ldg(beta_b + j);
      double c0, c1;
      if (DIAG) {
        c0 = fma(bi0, bj, -__ldg(iKa + (size_t)j * NP + i0));
        c1 = fma(bi1, bj, -__ldg(iKa + (size_t)j * NP + i1));
      } else {
        c0 = bi0 * bj;
        c1 = bi1 * bj;
      }
      double t0 = kr0 + kj, t1 = kr1 + kj;
#pragma unroll
      for (int e = 0; e < EV; e++) {
        t0 = fma(u0[e], nj[e], t0);
        t1 = fma(u1[e], nj[e], t1);
      }
      if (DIAG) { c0 *= wgt; c1 *= wgt; }
      double w0 = c0 * exp2s(t0, s_tab);
      double w1 = c1 * exp2s(t1, s_tab);
      rho0 += w0;
      rho1 += w1;
      if (GRAD) {
#pragma unroll
        for (int e = 0; e < EV; e++) {
          xi0[e] = fma(w0, nj[e], xi0[e]);
          xi1[e] = fma(w1, nj[e], xi1[e]);
        }
        v[jj] = w0 + w1;
      }
    }
    if (GRAD) {
      int col;
      double tot = col_reduce8(v, lane, col);
      if ((lane & 3) == 0) atomicAdd(s_gam + j0 + col, tot);
    }
  }
  if (GRAD) {
    atomicAdd(s_rho + i0, rho0);
    atomicAdd(s_rho + i1, rho1);
#pragma unroll
    for (int e = 0; e < EV; e++) {
      atomicAdd(s_xi + i0 * EV + e, xi0[e]);
      atomicAdd(s_xi + i1 * EV + e, xi1[e]);
    }
  } else {
    double tot = warp_sum(rho0 + rho1);
    if (lane == 0) atomicAdd(s_acc, tot);
  }
}

// ---------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------
template <int EV, bool GRAD>
__global__ void __launch_bounds__(ROLLOUT_THREADS, 1) rollout_kernel(const RolloutParams p) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, NT = blockDim.x;
  const int E = p.E, D = p.D, N = p.N, NP = p.NP, DP = p.DP, Na = p.Na, H = p.H;
  const int P = E * (E + 1) / 2, G = p.group;
  const SmemLayout L = make_layout(EV, GRAD, NP, DP, D, E, G, (p.mode == 0) ? H : 0, Na);
  double* s_nu = sm + L.nu;
  double* s_lb = sm + L.grp;  // aliases the group arrays (only live in P1/P2)
  double* s_kap = sm + L.kap;
  double* s_gam = sm + L.gam;
  double* s_rho = sm + L.rho;
  double* s_xi = sm + L.xi;
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
  double* s_tabp = sm + L.tab;
  const unsigned s_tab = exp2s_table_addr(s_tabp);
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
  for (int i = tid; i < EXP2S_N; i += NT) s_tabp[i] = p.exp2tab[i];
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
  double* kk = p.ws_kk + (size_t)blockIdx.x * E * NP;
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
  __shared__ int s_next;
  __shared__ double s_one;
  if (tid == 0) s_one = 1.0;
  __shared__ unsigned char s_owner[GPMPC_MAX_EV * (GPMPC_MAX_EV + 1) / 2];   // cluster rank that sweeps pair pr
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
      } else if (tid == 128 && p.mode == 0) {   // stage cost of the current state (serial, overlaps the matrix work)
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
            lb = __ldg(p.beta + (size_t)a * NP + i) * exp2s((-0.5 * GPMPC_EXP2S_SCALE) * (quad + tail), s_tab);
            kv = s_logs2[a] - 0.5 * (head + tail);
          }
          s_lb[a * NP + i] = lb;
          kk[a * NP + i] = kv;
        }
      }
      __syncthreads();
      GEN_CLK(2);
      // ================================================================ P2: O(N) moment sums
      // lane per output (all lanes read the SAME training point: multicast LDS, no bank conflicts), the 32-output groups
      // x 4 slices of the training points dealt to the warps; the slice sums meet in s_out (zeroed in P1)
      {
        const int nTot = E * nOut, ngrp = (nTot + 31) >> 5, nsl = 4, slen = (N + nsl - 1) / nsl;
        for (int task = tid >> 5; task < ngrp * nsl; task += NT >> 5) {
          const int o = (task / nsl) * 32 + lane, sl = task % nsl;
          const int ibeg = sl * slen, iend = min(N, ibeg + slen);
          if (o < nTot) {
            const int a = o / nOut, q = o - a * nOut;
            const double* lb = s_lb + a * NP;
            // value = lb_i * f1 * f2 * f3 with up to three factors nu_i[.]; an absent factor reads the constant 1.0
            // (stride 0), so that all lanes run the same instruction stream whatever their output is
            const double* b1 = &s_one; const double* b2 = &s_one; const double* b3 = &s_one;
            int st1 = 0, st2 = 0, st3 = 0;
            if (q == 0) {
            } else if (q <= D) {
              b1 = s_nu + (q - 1); st1 = DP;
            } else if (q < 1 + D + EV * D) {
              const int qq = q - 1 - D, e1 = qq / D;
              b1 = s_nu + e1; st1 = DP;
              b2 = s_nu + (qq - e1 * D); st2 = DP;
            } else {
              const int qq = q - 1 - D - EV * D, e1 = qq / PV;
              int kl = qq - e1 * PV, k1 = 0;
              while (kl >= EV - k1) { kl -= EV - k1; k1++; }
              b1 = s_nu + e1; st1 = DP;
              b2 = s_nu + k1; st2 = DP;
              b3 = s_nu + (k1 + kl); st3 = DP;
            }
            double acc0 = 0.0, acc1 = 0.0;
            int i = ibeg;
            for (; i + 1 < iend; i += 2) {
              const double f0 = lb[i] * b1[i * st1], f1 = lb[i + 1] * b1[(i + 1) * st1];
              const double g0 = b2[i * st2] * b3[i * st3], g1 = b2[(i + 1) * st2] * b3[(i + 1) * st3];
              acc0 = fma(f0, g0, acc0);
              acc1 = fma(f1, g1, acc1);
            }
            if (i < iend) acc0 = fma(lb[i] * b1[i * st1], b2[i * st2] * b3[i * st3], acc0);
            atomicAdd(s_out + o, acc0 + acc1);
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
      for (int g0 = 0; g0 < P; g0 += G) {
        if (C > 1 && s_owner[g0] != crank) continue;    // another CTA of the cluster owns this pair (G == 1 in cluster mode)
        const int gn = min(G, P - g0);
        for (int o = tid; o < gn * NP; o += NT) {
          const int pl = o / NP, j = o - pl * NP;
          const int pr = g0 + pl, b = s_int[2 + pr] & 15;
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
          s_kap[o] = GPMPC_EXP2S_SCALE * kap;
          if (GRAD) {
            s_gam[o] = 0.0;
            s_rho[o] = 0.0;
#pragma unroll
            for (int e = 0; e < EV; e++) s_xi[(size_t)o * EV + e] = 0.0;
          }
        }
        for (int o = tid; o < gn * L.paccN; o += NT) s_pacc[o] = 0.0;
        if (tid == 0) s_int[0] = 0;
        __syncthreads();
        GEN_CLK(5);
        {
          const int nrb = NP / 64, nseg = (NP + p.seg - 1) / p.seg;
          const int nitems = gn * nrb * nseg;
          for (;;) {
            int item = 0;
            if (lane == 0) item = atomicAdd(&s_int[0], 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= nitems) break;
            // longest row sweeps (small I) first within each pair
            const int pl = item / (nrb * nseg);
            const int rem = item - pl * nrb * nseg;
            const int I = rem / nseg, js = rem - I * nseg;
            const int pr = g0 + pl, ab = s_int[2 + pr], a = ab >> 4, b = ab & 15;
            int jbeg = js * p.seg;
            const int jend = min(NP, jbeg + p.seg);
            const double* Qm = s_Q + pr * EV * EV;
            if (a == b) {
              if (jend <= 64 * I) continue;
              jbeg = max(jbeg, 64 * I);
              pair_item<EV, GRAD, true>(p, s_nu, s_kap + pl * NP, Qm, s_il2 + a * D, s_il2 + b * D, kk + a * NP,
                                        p.beta + (size_t)a * NP, p.beta + (size_t)b * NP,
                                        p.iK + (size_t)a * NP * NP, I, jbeg, jend, lane, s_gam + pl * NP,
                                        s_rho + pl * NP, s_xi + (size_t)pl * NP * EV, s_pacc + pl * L.paccN, s_tab);
            } else {
              pair_item<EV, GRAD, false>(p, s_nu, s_kap + pl * NP, Qm, s_il2 + a * D, s_il2 + b * D, kk + a * NP,
                                         p.beta + (size_t)a * NP, p.beta + (size_t)b * NP, nullptr, I, jbeg, jend,
                                         lane, s_gam + pl * NP, s_rho + pl * NP, s_xi + (size_t)pl * NP * EV,
                                         s_pacc + pl * L.paccN, s_tab);
            }
          }
        }
        __syncthreads();
        GEN_CLK(6);
        if (GRAD) {
          // reduce (rho, gam, xi) over the training points into S_raw, dS/dm (D), dS/dQ (EV x EV)
          for (int pl = 0; pl < gn; pl++) {
            const int pr = g0 + pl, ab = s_int[2 + pr], a = ab >> 4, b = ab & 15;
            const double* la = s_il2 + a * D;
            const double* lbv = s_il2 + b * D;
            double accS = 0.0, gm[GPMPC_MAX_D], gQ[EV * EV];
#pragma unroll
            for (int d = 0; d < GPMPC_MAX_D; d++) gm[d] = 0.0;
#pragma unroll
            for (int e = 0; e < EV * EV; e++) gQ[e] = 0.0;
            for (int i = tid; i < N; i += NT) {
              const double rho = s_rho[pl * NP + i], gam = s_gam[pl * NP + i];
              accS += (a == b) ? (rho + gam) : rho;
              double za[EV], zb[EV], xs[EV];
#pragma unroll
              for (int d = 0; d < GPMPC_MAX_D; d++)
                if (d < D) {
                  double nud = s_nu[i * DP + d];
                  gm[d] = fma(rho * la[d] + gam * lbv[d], nud, gm[d]);
                }
#pragma unroll
              for (int e = 0; e < EV; e++) {
                double nue = s_nu[i * DP + e];
                za[e] = nue * la[e];
                zb[e] = nue * lbv[e];
                xs[e] = s_xi[((size_t)pl * NP + i) * EV + e] * lbv[e];
              }
#pragma unroll
              for (int e = 0; e < EV; e++)
#pragma unroll
                for (int f = 0; f < EV; f++)
                  gQ[e * EV + f] += rho * za[e] * za[f] + gam * zb[e] * zb[f] + za[e] * xs[f] + za[f] * xs[e];
            }
            accS = warp_sum(accS);
            if (lane == 0) atomicAdd(s_pacc + pl * L.paccN, accS);
            for (int d = 0; d < D; d++) {
              double v = warp_sum(gm[d]);
              if (lane == 0) atomicAdd(s_pacc + pl * L.paccN + 1 + d, v);
            }
#pragma unroll
            for (int e = 0; e < EV * EV; e++) {
              double v = warp_sum(gQ[e]);
              if (lane == 0) atomicAdd(s_pacc + pl * L.paccN + 1 + D + e, v);
            }
          }
          __syncthreads();
        }
        GEN_CLK(7);
        if (tid < gn) {
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
            for (int e = 0; e < EV * EV; e++) rec[2 + D + e] = acc[1 + D + e];
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
cudaError_t launch_rollout_inst(bool grad, const RolloutParams& p, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(ROLLOUT_THREADS);
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

template <int E>
cudaError_t launch_backward_inst(const BackwardParams& p, cudaStream_t st) {
  const int blk = 128, grid = (p.B + blk - 1) / blk;
  backward_kernel<E><<<grid, blk, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace gpmpc
