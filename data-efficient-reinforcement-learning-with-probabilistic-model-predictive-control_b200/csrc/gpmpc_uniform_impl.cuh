// "Uniform-kernel" fast path: all E GPs share one hyper-parameter set (ARD lengthscales, outputscale, noise)
// -- the reference's configuration before hyper-parameter training (examples/*/config_*.py:41-45 give every GP
// the same values).  Then K_a = K for all a, hence ONE iK, and in gp_model.py:156-169 R_ab, Q_ab, z_a,i and the
// whole exponent are independent of (a, b):
//       L^{ab}_ij = s2^2 * Eh_ij ,   Eh_ij = exp(kap_i + kap_j + u_i . nu_j)
//       S^raw_ab  = s2^2 * ( beta_a^T Eh beta_b  -  [a==b] tr(iK Eh) )
// so one sweep with ONE exp per element serves all E(E+1)/2 output pairs -- and, Eh and iK being symmetric, only
// over the 64 x 64 tiles on or above the diagonal (uni_fwd_item): ~N^2/2 exps per prediction instead of
// (E/2 + E(E-1)/2) N^2.  The mean part shares A = (s + Lambda)^-1 and exp(-q_i/2) as well.
//
// Gradient: reverse mode.  With P = E(E+1)/2 scalar outputs per exp, emitting forward-mode Jacobians (general
// path) would cost ~4x more than a second sweep with the adjoint-weighted coefficient
//       W_ij = (Omega beta_i) . beta_j - wbar * iK_ij ,    w_ij = W_ij Eh_ij   (symmetric, same tile triangle)
// from which dS/dm and dS/dQ follow exactly as for a diagonal pair of the general path (rho, gamma, xi sums).
// uniform_fwd_kernel stores only (M, V^E, h, g^E, S^raw) per step; uniform_bwd_kernel runs the reverse sweep
// with one CTA per candidate.  tests/algo_spec.py remains the executable spec of the mathematics.
//
// Round 2: both kernels exist in two builds for E <= 5 (256 threads at <= 128 registers, 128 threads at <= 168 -- three
// CTAs per SM; the host plan picks, gpmpc_api.cu); for E >= 6 the sweeps run on the float64 tensor cores (uni_*_mma8:
// exponent, coefficient and the E-wide row sums as DMMA m8n8k4 tile products on 32 x 32 tiles); the E = 4 DMMA variants
// that were measured and lost stay behind -DUNI_MMA=1/2 (profiles/r02s_dmma_experiments.txt).
#pragma once
#include "gpmpc_rollout_impl.cuh"
#include "gpmpc_uniform_layout.cuh"

namespace gpmpc {

// tuning aid: thread 0 of CTA 0 adds the cycles since the previous mark to slot k (after a __syncthreads)
#define UNI_CLK(k) do { if (p.dbg_clk && blockIdx.x == 0 && tid == 0) { const long long c_ = clock64(); p.dbg_clk[k] += c_ - clk_; clk_ = c_; } } while (0)

__device__ __forceinline__ unsigned long long uni_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned uni_smid() { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
// tuning aid: every CTA reports {cycles, wall ns, SM id, start ns} of its whole life to dbg_clk[16 + 4 * (off + blockIdx.x)]
#define UNI_CTA_BEGIN() const long long cta_c0_ = clock64(); const unsigned long long cta_t0_ = uni_globaltimer()
#define UNI_CTA_END(off) do { if (p.dbg_clk && tid == 0) { long long* d_ = p.dbg_clk + 16 + 4 * ((off) + (size_t)blockIdx.x); \
    d_[0] = clock64() - cta_c0_; d_[1] = (long long)(uni_globaltimer() - cta_t0_); d_[2] = uni_smid(); d_[3] = (long long)cta_t0_; } } while (0)

// Dynamic scheduling: CTAs draw candidates from a global counter (kernel k = 0 forward, 1 reverse sweep).  SMs differ
// in speed by up to ~25 % on this workload (distance to the L2 slices), so a static round-robin split leaves the fast
// SMs idle at the end.  Ends with a barrier; s_int[3] is the hand-over slot.
__device__ __forceinline__ int uni_next_candidate(const RolloutParams& p, int k, int* s_int, int tid) {
  if (tid == 0) s_int[3] = atomicAdd(p.queue + k, 1);
  __syncthreads();
  const int c = s_int[3];
  __syncthreads();
  return c;
}
// Thread-block clusters for small batches: with fewer candidates than SMs, p.cluster (2, 4 or 8) CTAs on neighbouring
// SMs share ONE candidate.  Every CTA repeats the cheap O(N) / O(E^3) phases (bitwise identically), the tile-triangle
// sweep is split over all their warps, and the partial sums meet in L2 (float64 RED into a per-cluster accumulator,
// rotating over three buffers so that zeroing never races with the next step) followed by ONE hardware cluster barrier
// per horizon step.  Only the cluster's rank-0 CTA writes outputs.
// (uni_red_add, the float64 reduction at L2 without a return value, lives in gpmpc_rollout_impl.cuh)
constexpr int UNI_CL_ACC = 64;   // doubles per accumulator buffer of the forward sweep (P + 1 <= 37 used)

// column sums of the reverse sweep: 0 = shuffle transpose-reduce (col_reduce8), 1 = through a per-warp shared-memory
// scratch (col_reduce8s).  The scratch version wins in the general kernel (3 warps per sub-partition, shared-memory pipe at
// ~50 %) and loses here (27.5 -> 30.4 ms at B=2368: the records of this kernel already load the pipe, and the 20 KB of
// scratch cost the precomputed per-step matrices their place): profiles/r02c_*.  NB: a -D must reach the host code too.
#ifndef UNI_BWD_COLRED_SMEM
#define UNI_BWD_COLRED_SMEM 0
#endif
#ifndef UNI_P1_UNROLL
#define UNI_P1_UNROLL 2     // forward record phase: training points in flight per thread
#endif
#ifndef UNI_MINB
#define UNI_MINB(EV) ((EV) <= 5 ? 2 : 1)
#endif
// The forward kernel is built twice for state dimensions <= 5: MAXT = 256 (two CTAs per SM, <= 128 registers; also the
// thread-block-cluster launches) and MAXT = 128 with three CTAs per SM (<= 168 registers): with the larger register
// budget ptxas hoists the record / iK loads of a loop body to its top and interleaves the stages of both column pairs
// (forward 60.6 -> 57.9 ms at the headline shape).  The host picks it when three CTAs fit the shared memory.
#define UNI_FWD_MINCTAS(EV, MAXT) ((MAXT) == 128 && (EV) <= 5 ? 3 : UNI_MINB(EV))
// The reverse sweep has the same two builds: 3 CTAs x 128 threads (<= 168 registers) for E <= 5 when three fit the shared
// memory -- they do at the headline shape once the precomputed per-step records live in the global scratch (premat = 2)
// -- else 2 x 256 (128 registers).  Measured: 81.6 vs 82.5 ms (profiles/r02s_dmma_experiments.txt).
#define UNI_BWD_MINCTAS(EV, MAXT) ((MAXT) == 128 && (EV) <= 5 ? 3 : UNI_MINB(EV))

// ---------------------------------------------------------------------------------------------
// forward hot loop: full sweep, rows {64 I + lane, +32}, columns [jbeg, jend)
//   r_b,i += Eh_ij beta_b,j  (E FMAs),  tr += Eh_ij iK_ij  (1 FMA)   [gp_model.py:169-175 for all (a,b) at once]
// ---------------------------------------------------------------------------------------------
// training inputs x (read by every CTA at every step, N D doubles): with UNI_X_EVICT_LAST the loads ask L1 to keep them
// while the iK stream of the sweeps passes through
#ifndef UNI_X_EVICT_LAST
#define UNI_X_EVICT_LAST 0
#endif
#ifndef UNI_FWD_IK_EVICT_FIRST
#define UNI_FWD_IK_EVICT_FIRST 0
#endif
__device__ __forceinline__ double uni_ldg_x(const double* p) {
#if UNI_X_EVICT_LAST
  double v;
  asm("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ double2 uni_ldg_ik_fwd(const double* p) {
#if UNI_FWD_IK_EVICT_FIRST
  double2 v;
  asm("ld.global.nc.L1::evict_first.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#else
  return __ldg(reinterpret_cast<const double2*>(p));
#endif
}

template <int EV>
__device__ __forceinline__ void uni_load_rec(const double* __restrict__ s_rec, int j, double (&nu)[EV],
                                             double& kap, double (&beta)[EV]) {
  constexpr int RLEN = 2 * EV + 2;   // record stride (UniLayout::rhot)
  const double2* r2 = reinterpret_cast<const double2*>(s_rec + j * RLEN);
  double buf[RLEN];
#pragma unroll
  for (int q = 0; q < RLEN / 2; q++) { const double2 v = r2[q]; buf[2 * q] = v.x; buf[2 * q + 1] = v.y; }
#pragma unroll
  for (int e = 0; e < EV; e++) nu[e] = buf[e];
  kap = buf[EV];
#pragma unroll
  for (int b = 0; b < EV; b++) beta[b] = buf[EV + 1 + b];
}

// columns [jbeg, jend) of the forward sweep for the lane's rows i0, i0 + 1:  r_b,i += Eh'_ij beta_b,j ; tr_i += Eh'_ij iK_ij
// with the ROW factor of the exponential taken out of the loop:  Eh_ij = e_i Eh'_ij ,  Eh'_ij = exp(kap_j + u_i . nu_j + d_i),
// e_i = exp(max(kap_i, UNI_KAP_MIN)) applied by the caller to the finished row sums, d_i = kap_i - max(kap_i, UNI_KAP_MIN) the
// residual shift of far-away rows (SH: some lane of the warp has d_i != 0; otherwise the add is not even issued).
template <int EV, bool SH>
__device__ __forceinline__ void uni_fwd_cols(const RolloutParams& p, const double* __restrict__ s_rec, int i0,
                                             int jbeg, int jend, const double (&u0)[EV], const double (&u1)[EV],
                                             double kr0, double kr1, double (&r0)[EV], double (&r1)[EV], double& trOut0,
                                             double& trOut1) {
  constexpr int E = EV;
  const int NP = p.NP;
  const double* __restrict__ ik0 = p.iK + (size_t)jbeg * NP + i0;   // row j of the symmetric iK, lane = columns i0, i0 + 1
  double tr = 0.0, tr2 = 0.0;
#pragma unroll 2
  for (int j = jbeg; j < jend; j += 2) {   // 2 columns x 2 rows = 4 independent chains per warp
    double na[EV], nb[EV], ba[E], bb[E], ka, kb;
    uni_load_rec<EV>(s_rec, j, na, ka, ba);
    uni_load_rec<EV>(s_rec, j + 1, nb, kb, bb);
    const double2 ika = uni_ldg_ik_fwd(ik0);
    const double2 ikb = uni_ldg_ik_fwd(ik0 + NP);
    ik0 += 2 * (size_t)NP;
    double t[4], ex[4];
    if (SH) { t[0] = kr0 + ka; t[1] = kr1 + ka; t[2] = kr0 + kb; t[3] = kr1 + kb; }
    else { t[0] = ka; t[1] = ka; t[2] = kb; t[3] = kb; }
    // serpentine order: every FMA shares one register operand with its predecessor (operand-reuse cache; a DFMA
    // with three fresh register operands issues at 2/3 rate on B200, tools/micro/dfma_operands.cu)
#pragma unroll
    for (int e = 0; e < EV; e++) {
      t[0] = fma(u0[e], na[e], t[0]);
      t[1] = fma(u1[e], na[e], t[1]);
      t[3] = fma(u1[e], nb[e], t[3]);
      t[2] = fma(u0[e], nb[e], t[2]);
    }
    exp2s_x4(t, ex);
#pragma unroll
    for (int b = 0; b < E; b++) {
      if (b & 1) { r1[b] = fma(ex[1], ba[b], r1[b]); r0[b] = fma(ex[0], ba[b], r0[b]); }
      else       { r0[b] = fma(ex[0], ba[b], r0[b]); r1[b] = fma(ex[1], ba[b], r1[b]); }
    }
#pragma unroll
    for (int b = 0; b < E; b++) {
      if (b & 1) { r1[b] = fma(ex[3], bb[b], r1[b]); r0[b] = fma(ex[2], bb[b], r0[b]); }
      else       { r0[b] = fma(ex[2], bb[b], r0[b]); r1[b] = fma(ex[3], bb[b], r1[b]); }
    }
    tr = fma(ex[0], ika.x, tr);
    tr2 = fma(ex[1], ika.y, tr2);
    tr = fma(ex[2], ikb.x, tr);
    tr2 = fma(ex[3], ikb.y, tr2);
  }
  trOut0 += tr;
  trOut1 += tr2;
}

// Row factor of the sweeps (see uni_fwd_cols): uni_row_factor in gpmpc_rollout_impl.cuh.
template <int EV>
__device__ __forceinline__ void uni_fwd_item(const RolloutParams& p, const double* __restrict__ s_rec,
                                             const double* __restrict__ Qm, const double* __restrict__ il2, int I,
                                             int jbeg, int jend, int lane, double* s_part) {
  constexpr int E = EV;
  const int i0 = 64 * I + 2 * lane, i1 = i0 + 1;   // adjacent rows: one 16-byte load fetches both iK values of a column
  double u0[EV], u1[EV], bi0[E], bi1[E], kr0, kr1;
  {
    double n0[EV], n1[EV];
    uni_load_rec<EV>(s_rec, i0, n0, kr0, bi0);
    uni_load_rec<EV>(s_rec, i1, n1, kr1, bi1);
#pragma unroll
    for (int e = 0; e < EV; e++) {
      double q0 = 0.0, q1 = 0.0;
#pragma unroll
      for (int f = 0; f < EV; f++) {
        q0 = fma(Qm[e * EV + f], n0[f] * il2[f], q0);
        q1 = fma(Qm[e * EV + f], n1[f] * il2[f], q1);
      }
      u0[e] = (2.0 * GPMPC_EXP2S_SCALE) * q0 * il2[e];   // exponent in table units (exp2s)
      u1[e] = (2.0 * GPMPC_EXP2S_SCALE) * q1 * il2[e];
    }
  }
  double r0[E], r1[E];
#pragma unroll
  for (int b = 0; b < E; b++) { r0[b] = 0.0; r1[b] = 0.0; }
  // Eh and iK are symmetric, so only tiles on or above the diagonal are swept (columns j >= 64 I):
  //   S_ab = sum_{tiles above} (X_ab + X_ba) + sum_{diagonal tiles} (X_ab + X_ba) / 2 ,  X_ab = sum_ij beta_a,i Eh_ij beta_b,j
  //   tr(iK Eh) = 2 sum_{above} + sum_{diagonal}
  // -- the tile below the diagonal contributes X_ba of its mirror image.  The row sums r_b,i of the diagonal-tile
  // columns are halved before the columns above it are added.
  const int jd1 = 64 * I + 64;
  jend = min(jend, (p.N + 1) & ~1);   // zero-padded columns (beta = 0, iK = 0) contribute exact zeros: skip them
  const double e0 = uni_row_factor(kr0), e1 = uni_row_factor(kr1);   // kr0, kr1 become residual shifts
  const bool far = __any_sync(0xffffffffu, kr0 != 0.0 || kr1 != 0.0);
  double trD0 = 0.0, trD1 = 0.0, trU0 = 0.0, trU1 = 0.0;
  if (jbeg < jd1) {
    if (far) uni_fwd_cols<EV, true>(p, s_rec, i0, jbeg, min(jend, jd1), u0, u1, kr0, kr1, r0, r1, trD0, trD1);
    else uni_fwd_cols<EV, false>(p, s_rec, i0, jbeg, min(jend, jd1), u0, u1, kr0, kr1, r0, r1, trD0, trD1);
#pragma unroll
    for (int b = 0; b < E; b++) { r0[b] *= 0.5; r1[b] *= 0.5; }
  }
  if (far) uni_fwd_cols<EV, true>(p, s_rec, i0, max(jbeg, jd1), jend, u0, u1, kr0, kr1, r0, r1, trU0, trU1);
  else uni_fwd_cols<EV, false>(p, s_rec, i0, max(jbeg, jd1), jend, u0, u1, kr0, kr1, r0, r1, trU0, trU1);
  const double tr = e0 * fma(2.0, trU0, trD0) + e1 * fma(2.0, trU1, trD1);
#pragma unroll
  for (int b = 0; b < E; b++) { bi0[b] *= e0; bi1[b] *= e1; }   // the row factor, applied once to the finished row sums
  // S_ab += sum_i beta_a,i r_b,i  (a <= b), trace: halving reduction of the P + 1 lane partials, 16 at a time, into
  // this warp's private accumulator row (no shared-memory float64 atomics: those are CAS spin loops)
  constexpr int P1 = E * (E + 1) / 2 + 1, NCH = (P1 + 15) / 16;
  double vals[16 * NCH];
  {
    int pr = 0;
#pragma unroll
    for (int a = 0; a < E; a++)
#pragma unroll
      for (int b = a; b < E; b++) {
        vals[pr] = (a == b) ? 2.0 * (bi0[a] * r0[a] + bi1[a] * r1[a])
                            : (bi0[a] * r0[b] + bi0[b] * r0[a]) + (bi1[a] * r1[b] + bi1[b] * r1[a]);
        pr++;
      }
    vals[pr++] = tr;
#pragma unroll
    for (; pr < 16 * NCH; pr++) vals[pr] = 0.0;
  }
#pragma unroll
  for (int ch = 0; ch < NCH; ch++) {
    double v16[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v16[k] = vals[16 * ch + k];
    int idx;
    const double tot = warp_reduce_multi<16>(v16, lane, idx);
    if ((lane & 1) == 0 && 16 * ch + idx < P1) s_part[16 * ch + idx] += tot;
  }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core variant of the forward sweep (UNI_USE_MMA): the exponent tile  t_ij = kap_j + u_i . nu_j  of 8 rows x 8
// columns is ONE float64 DMMA (m8n8k4: k = state dimension, zero-padded / two steps for E != 4) instead of E DFMAs per
// element.  A warp tile is 32 rows x 8 columns per round: lane (g, q) = (lane >> 2, lane & 3) owns rows ib + 8 m + g
// (m = 0..3) and columns j0 + 2 q, j0 + 2 q + 1 -- the DMMA accumulator layout, 8 independent elements per lane.
//   A fragment  u_{ib + 8 m + g, 4 ks + q}   (loop invariant, registers)
//   B fragment  nu_{j0 + g, 4 ks + q}         (one LDS.64 per round)
//   C           kap_{j0 + 2 q}, kap_{j0 + 2 q + 1}  (+ the residual row shift of far-away rows, SH)
// Everything after the exponent is per element as before (exp2s, r_b,i += Eh beta_b,j, tr += Eh iK_ij); a row's sums are
// spread over the 4 lanes of its quad, which the linear warp reduction at the end of the item does not care about.
// 32-row tiles also halve the redundant part of the diagonal tiles (32 x 32 swept in full with half weights).
// ---------------------------------------------------------------------------------------------
#ifndef UNI_MMA
#define UNI_MMA 0          // 1: DMMA exponent / coefficient tiles in both sweeps; 2: forward sweep on the quad layout with DFMAs
#endif
#define UNI_USE_MMA(EV) (UNI_MMA == 1 && (EV) == 4)
#define UNI_USE_QUAD_FWD(EV) ((UNI_MMA == 1 || UNI_MMA == 2) && (EV) == 4)

template <int EV, bool SH>
__device__ __forceinline__ void uni_fwd_cols_mma(const RolloutParams& p, const double* __restrict__ s_rec, int ib,
                                                 int jbeg, int jend, int lane, const double (&ua)[4][(EV + 3) / 4],
                                                 const double (&uf)[4][EV], const double (&kr)[4], double (&r)[4][EV],
                                                 double (&tr)[4]) {
  constexpr int E = EV, RLEN = 2 * EV + 2, KS = (EV + 3) / 4;
  constexpr bool MMA = UNI_MMA == 1;   // else: the exponent with DFMAs on the same layout (uf = the rows' full u vectors)
  const int g = lane >> 2, q = lane & 3;
  const size_t rs = 8 * (size_t)p.NP;
  const double* __restrict__ ik = p.iK + (size_t)(ib + g) * p.NP + jbeg + 2 * q;   // rows of the symmetric iK, two adjacent columns
  const double* __restrict__ rb = s_rec + (jbeg + g) * RLEN + q;                    // B fragment source
  const double* __restrict__ rc = s_rec + (jbeg + 2 * q) * RLEN;                    // the lane's two columns
#pragma unroll 1
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    double bf[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) bf[ks] = MMA ? rb[4 * ks] : 0.0;   // slots >= EV hold finite record data; their A entries are zero
    const double ka0 = rc[EV], ka1 = rc[RLEN + EV];
    double be0[E], be1[E];
#pragma unroll
    for (int b = 0; b < E; b++) { be0[b] = rc[EV + 1 + b]; be1[b] = rc[RLEN + EV + 1 + b]; }
    double2 ikv[4];
#pragma unroll
    for (int m = 0; m < 4; m++) ikv[m] = ldg_stream2<false>(ik + m * rs);
    double t[8], ex[8];
#pragma unroll
    for (int m = 0; m < 4; m++) {
      double d0 = SH ? ka0 + kr[m] : ka0, d1 = SH ? ka1 + kr[m] : ka1;
      if (MMA) {
#pragma unroll
        for (int ks = 0; ks < KS; ks++) dmma_m8n8k4(d0, d1, ua[m][ks], bf[ks]);
      }
      t[2 * m] = d0;
      t[2 * m + 1] = d1;
    }
    if (!MMA) {
      double n0[EV], n1[EV];
#pragma unroll
      for (int e = 0; e < EV; e++) { n0[e] = rc[e]; n1[e] = rc[RLEN + e]; }
#pragma unroll
      for (int e = 0; e < EV; e++) {
#pragma unroll
        for (int m = 0; m < 4; m++) {   // serpentine: every FMA shares a register operand with its predecessor
          if (m & 1) { t[2 * m + 1] = fma(uf[m][e], n1[e], t[2 * m + 1]); t[2 * m] = fma(uf[m][e], n0[e], t[2 * m]); }
          else       { t[2 * m] = fma(uf[m][e], n0[e], t[2 * m]); t[2 * m + 1] = fma(uf[m][e], n1[e], t[2 * m + 1]); }
        }
      }
    }
    exp2s_xn<8>(t, ex);
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int b = 0; b < E; b++) r[m][b] = fma(ex[2 * m], be0[b], r[m][b]);
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int b = 0; b < E; b++) r[m][b] = fma(ex[2 * m + 1], be1[b], r[m][b]);
#pragma unroll
    for (int m = 0; m < 4; m++) {
      tr[m] = fma(ex[2 * m], ikv[m].x, tr[m]);
      tr[m] = fma(ex[2 * m + 1], ikv[m].y, tr[m]);
    }
    ik += 8;
    rb += 8 * RLEN;
    rc += 8 * RLEN;
  }
}

// One run of columns [jbeg, jend) (multiples of 8, jbeg >= 32 I) of the 32-row block I: see uni_fwd_item for the algebra.
template <int EV>
__device__ __forceinline__ void uni_fwd_item_mma(const RolloutParams& p, const double* __restrict__ s_rec,
                                                 const double* __restrict__ Qm, const double* __restrict__ il2, int I,
                                                 int jbeg, int jend, int lane, double* s_part) {
  constexpr int E = EV, RLEN = 2 * EV + 2, KS = (EV + 3) / 4;
  const int g = lane >> 2, q = lane & 3, ib = 32 * I;
  double ua[4][KS], uf[4][EV], kr[4], ei[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const double* rec = s_rec + (ib + 8 * m + g) * RLEN;
    kr[m] = rec[EV];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const int ee = 4 * ks + q;
      double acc = 0.0;
      if (UNI_MMA == 1 && ee < EV) {
#pragma unroll
        for (int f = 0; f < EV; f++) acc = fma(Qm[ee * EV + f], rec[f] * il2[f], acc);
        acc *= (2.0 * GPMPC_EXP2S_SCALE) * il2[ee];   // exponent in table units (exp2s)
      }
      ua[m][ks] = acc;
    }
#pragma unroll
    for (int e = 0; e < EV; e++) {
      double acc = 0.0;
      if (UNI_MMA != 1) {
#pragma unroll
        for (int f = 0; f < EV; f++) acc = fma(Qm[e * EV + f], rec[f] * il2[f], acc);
        acc *= (2.0 * GPMPC_EXP2S_SCALE) * il2[e];
      }
      uf[m][e] = acc;
    }
  }
  double r[4][E], trD[4], trU[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    trD[m] = 0.0;
    trU[m] = 0.0;
#pragma unroll
    for (int b = 0; b < E; b++) r[m][b] = 0.0;
  }
  const int jd1 = ib + 32;
  jend = min(jend, (p.N + 7) & ~7);   // zero-padded columns (beta = 0, iK = 0) contribute exact zeros: skip them
#pragma unroll
  for (int m = 0; m < 4; m++) ei[m] = uni_row_factor(kr[m]);   // kr becomes the residual shift
  const bool far = __any_sync(0xffffffffu, kr[0] != 0.0 || kr[1] != 0.0 || kr[2] != 0.0 || kr[3] != 0.0);
  if (jbeg < jd1) {
    if (far) uni_fwd_cols_mma<EV, true>(p, s_rec, ib, jbeg, min(jend, jd1), lane, ua, uf, kr, r, trD);
    else uni_fwd_cols_mma<EV, false>(p, s_rec, ib, jbeg, min(jend, jd1), lane, ua, uf, kr, r, trD);
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int b = 0; b < E; b++) r[m][b] *= 0.5;
  }
  if (far) uni_fwd_cols_mma<EV, true>(p, s_rec, ib, max(jbeg, jd1), jend, lane, ua, uf, kr, r, trU);
  else uni_fwd_cols_mma<EV, false>(p, s_rec, ib, max(jbeg, jd1), jend, lane, ua, uf, kr, r, trU);
  double tr = 0.0;
#pragma unroll
  for (int m = 0; m < 4; m++) tr = fma(ei[m], fma(2.0, trU[m], trD[m]), tr);
  constexpr int P1 = E * (E + 1) / 2 + 1, NCH = (P1 + 15) / 16;
  double vals[16 * NCH];
#pragma unroll
  for (int k = 0; k < 16 * NCH; k++) vals[k] = 0.0;
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const double* rec = s_rec + (ib + 8 * m + g) * RLEN + EV + 1;
    double bi[E];
#pragma unroll
    for (int a = 0; a < E; a++) bi[a] = rec[a] * ei[m];   // the row factor, applied once to the finished row sums
    int pr = 0;
#pragma unroll
    for (int a = 0; a < E; a++)
#pragma unroll
      for (int b = a; b < E; b++) {
        vals[pr] += (a == b) ? 2.0 * (bi[a] * r[m][a]) : fma(bi[a], r[m][b], bi[b] * r[m][a]);
        pr++;
      }
  }
  vals[P1 - 1] = tr;
#pragma unroll
  for (int ch = 0; ch < NCH; ch++) {
    double v16[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v16[k] = vals[16 * ch + k];
    int idx;
    const double tot = warp_reduce_multi<16>(v16, lane, idx);
    if ((lane & 1) == 0 && 16 * ch + idx < P1) s_part[16 * ch + idx] += tot;
  }
}

// ---------------------------------------------------------------------------------------------
// Tensor-core sweeps for the large state dimensions (UNI_USE_MMA8: E = 6..8).  Same tile layout as uni_fwd_cols_mma
// (32 rows x 8 columns per round, lane (g, q) = rows ib + 8 m + g, columns j0 + 2 q, j0 + 2 q + 1), and here the products
// AFTER the exponential are DMMAs too: with E = 8 output dimensions a row-sum update
//   r_b,i += sum_j Eh_ij beta_b,j      (forward)        xi_i,e += sum_j w_ij nu_j,e     (reverse sweep)
// is an (8 rows x 8 columns) x (8 columns x 8 outputs) tile product.  The lane's own values Eh / w ARE the A fragments --
// the k slot q of the first DMMA stands for column j0 + 2 q, of the second for j0 + 2 q + 1, and the B fragments
// (beta_{g, column} / nu_{column, g}: one LDS.64 each) are picked accordingly -- so nothing moves between lanes, and a
// row's E sums shrink from E registers per row to the 2 accumulator entries (outputs 2 q, 2 q + 1) of its lane.  An
// element costs 8 float64 instructions + 2 DMMA / 8 (forward) or 11 + 3 DMMA / 8 (reverse) instead of 24 / 39, the
// per-column record loads all but vanish, and the kernels need ~150 registers instead of 255 + spills.
// Slots beyond E (zero-padded k steps, unused outputs) read finite record data against zero A entries / into ignored
// accumulator columns.
// ---------------------------------------------------------------------------------------------
// sums of v[16] (index 2 a + k) over the 8 row groups g of the warp; on return lane (g, q) holds index 2 g, 2 g + 1
__device__ __forceinline__ void uni_reduce_over_g16(const double (&v)[16], int lane, double& o0, double& o1) {
  const bool u16 = (lane & 16) != 0, u8 = (lane & 8) != 0, u4 = (lane & 4) != 0;
  double a[8], b[4];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const double send = u16 ? v[k] : v[k + 8], keep = u16 ? v[k + 8] : v[k];
    a[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double send = u8 ? a[k] : a[k + 4], keep = u8 ? a[k + 4] : a[k];
    b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const double s0 = u4 ? b[0] : b[2], k0 = u4 ? b[2] : b[0], s1 = u4 ? b[1] : b[3], k1 = u4 ? b[3] : b[1];
    o0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
    o1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 4);
  }
}

template <int EV, bool SH>
__device__ __forceinline__ void uni_fwd_cols_mma8(const RolloutParams& p, const double* __restrict__ s_rec, int ib,
                                                  int jbeg, int jend, int lane, const double (&ua)[4][2],
                                                  const double (&kr)[4], double (&r)[4][2], double (&tr)[4]) {
  constexpr int RLEN = 2 * EV + 2;
  const int g = lane >> 2, q = lane & 3;
  const size_t rs = 8 * (size_t)p.NP;
  const double* __restrict__ ik = p.iK + (size_t)(ib + g) * p.NP + jbeg + 2 * q;   // rows of the symmetric iK, two adjacent columns
  const double* __restrict__ rb = s_rec + (jbeg + g) * RLEN + q;                    // exponent B fragments: nu_{j0 + g, q}, nu_{j0 + g, 4 + q}
  const double* __restrict__ rc = s_rec + (jbeg + 2 * q) * RLEN;                    // the lane's two columns
#pragma unroll 1
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    const double bf0 = rb[0], bf1 = rb[4];
    const double ka0 = rc[EV], ka1 = rc[RLEN + EV];
    const double bb0 = rc[EV + 1 + g], bb1 = rc[RLEN + EV + 1 + g];   // row-sum B fragments: beta_{g, column}
    double2 ikv[4];
#pragma unroll
    for (int m = 0; m < 4; m++) ikv[m] = ldg_stream2<false>(ik + m * rs);
    double t[8], ex[8];
#pragma unroll
    for (int m = 0; m < 4; m++) {
      double d0 = SH ? ka0 + kr[m] : ka0, d1 = SH ? ka1 + kr[m] : ka1;
      dmma_m8n8k4(d0, d1, ua[m][0], bf0);
      dmma_m8n8k4(d0, d1, ua[m][1], bf1);
      t[2 * m] = d0;
      t[2 * m + 1] = d1;
    }
    exp2s_xn<8>(t, ex);
#pragma unroll
    for (int m = 0; m < 4; m++) dmma_m8n8k4(r[m][0], r[m][1], ex[2 * m], bb0);
#pragma unroll
    for (int m = 0; m < 4; m++) dmma_m8n8k4(r[m][0], r[m][1], ex[2 * m + 1], bb1);
#pragma unroll
    for (int m = 0; m < 4; m++) {
      tr[m] = fma(ex[2 * m], ikv[m].x, tr[m]);
      tr[m] = fma(ex[2 * m + 1], ikv[m].y, tr[m]);
    }
    ik += 8;
    rb += 8 * RLEN;
    rc += 8 * RLEN;
  }
}

// One run of columns [jbeg, jend) (multiples of 8, jbeg >= 32 I) of the 32-row block I (algebra: uni_fwd_item).
// s_part: the warp's accumulator row (P + 1 sums, padded) followed by its 64-double scratch.
template <int EV>
__device__ __forceinline__ void uni_fwd_item_mma8(const RolloutParams& p, const double* __restrict__ s_rec,
                                                  const double* __restrict__ Qm, const double* __restrict__ il2, int I,
                                                  int jbeg, int jend, int lane, double* s_part) {
  constexpr int E = EV, RLEN = 2 * EV + 2, P = E * (E + 1) / 2, PL = ((P + 1) + 15) & ~15;
  const int g = lane >> 2, q = lane & 3, ib = 32 * I;
  double ua[4][2], kr[4], ei[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const double* rec = s_rec + (ib + 8 * m + g) * RLEN;
    kr[m] = rec[EV];
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
      const int ee = 4 * ks + q;
      double acc = 0.0;
      if (ee < EV) {
#pragma unroll
        for (int f = 0; f < EV; f++) acc = fma(Qm[ee * EV + f], rec[f] * il2[f], acc);
        acc *= (2.0 * GPMPC_EXP2S_SCALE) * il2[ee];   // exponent in table units (exp2s)
      }
      ua[m][ks] = acc;
    }
  }
  double r[4][2], trD[4], trU[4];
#pragma unroll
  for (int m = 0; m < 4; m++) { trD[m] = 0.0; trU[m] = 0.0; r[m][0] = 0.0; r[m][1] = 0.0; }
  const int jd1 = ib + 32;
  jend = min(jend, (p.N + 7) & ~7);   // zero-padded columns (beta = 0, iK = 0) contribute exact zeros: skip them
#pragma unroll
  for (int m = 0; m < 4; m++) ei[m] = uni_row_factor(kr[m]);   // kr becomes the residual shift
  const bool far = __any_sync(0xffffffffu, kr[0] != 0.0 || kr[1] != 0.0 || kr[2] != 0.0 || kr[3] != 0.0);
  if (jbeg < jd1) {
    if (far) uni_fwd_cols_mma8<EV, true>(p, s_rec, ib, jbeg, min(jend, jd1), lane, ua, kr, r, trD);
    else uni_fwd_cols_mma8<EV, false>(p, s_rec, ib, jbeg, min(jend, jd1), lane, ua, kr, r, trD);
#pragma unroll
    for (int m = 0; m < 4; m++) { r[m][0] *= 0.5; r[m][1] *= 0.5; }
  }
  if (far) uni_fwd_cols_mma8<EV, true>(p, s_rec, ib, max(jbeg, jd1), jend, lane, ua, kr, r, trU);
  else uni_fwd_cols_mma8<EV, false>(p, s_rec, ib, max(jbeg, jd1), jend, lane, ua, kr, r, trU);
  double tr = 0.0;
#pragma unroll
  for (int m = 0; m < 4; m++) tr = fma(ei[m], fma(2.0, trU[m], trD[m]), tr);
  tr = warp_sum(tr);
  // X_ab = sum_i beta_a,i e_i r_b,i : the lane holds r for b = 2 q, 2 q + 1 of its rows -> 2 E partial sums, added up over
  // the row groups; lane (g, q) ends up with X[g][2 q], X[g][2 q + 1]; S_ab += X_ab + X_ba through the warp's scratch
  double v[16];
#pragma unroll
  for (int k = 0; k < 16; k++) v[k] = 0.0;
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const double* rec = s_rec + (ib + 8 * m + g) * RLEN + EV + 1;
#pragma unroll
    for (int a = 0; a < E; a++) {
      const double bi = rec[a] * ei[m];   // the row factor, applied once to the finished row sums
      v[2 * a] = fma(bi, r[m][0], v[2 * a]);
      v[2 * a + 1] = fma(bi, r[m][1], v[2 * a + 1]);
    }
  }
  double x0, x1;
  uni_reduce_over_g16(v, lane, x0, x1);
  double* xs = s_part + PL;
  xs[8 * g + 2 * q] = x0;
  xs[8 * g + 2 * q + 1] = x1;
  __syncwarp();
  for (int pr = lane; pr < P; pr += 32) {
    int a = 0, rem = pr;
    while (rem >= E - a) { rem -= E - a; a++; }
    const int b = a + rem;
    s_part[pr] += xs[8 * a + b] + xs[8 * b + a];   // (a == b: 2 X_aa, as in uni_fwd_item)
  }
  if (lane == 0) s_part[P] += tr;
  __syncwarp();
}

// Mean-part moments (gp_model.py:138-153), lane-per-output: lane o = a * nOut + q of every warp sums its output over the
// warp's slice of training points,
//   q = 0: h_a = sum_i e_i beta_a,i ;  q = 1 + d: g_a,d = sum_i e_i beta_a,i nu_i,d ,
// reading the hot-loop records as multicast LDS (all lanes: the same record, different slots); e_i sits in the
// record's spare slot, nu of the action / time dimensions in its tail.  No shuffles, no atomics: the per-warp
// partial rows s_wp[warp][o] are added up by the consumer (P4).
template <int EV>
__device__ __forceinline__ void uni_moments_slice(const double* __restrict__ s_rec, const double* __restrict__ s_tail,
                                                  int tlen, int nOut, int ibeg, int iend, int lane,
                                                  double* __restrict__ s_wp_row) {
  constexpr int RLEN = 2 * EV + 2;
  const int nTot = EV * nOut;
  for (int ob = 0; ob < nTot; ob += 32) {
    const int o = ob + lane;
    const bool act = o < nTot;
    const int a = act ? o / nOut : 0, q = act ? o - a * nOut : 0, d = q - 1;
    const int sb = EV + 1 + a;
    // nu_d of this lane's output: state dims in the hot record, action / time dims in the tail array
    const double* fb = (d < EV) ? s_rec + (d < 0 ? 0 : d) : s_tail + (d - EV);
    const int fs = (d < EV) ? RLEN : tlen;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i0 = ibeg; i0 < iend; i0 += 4) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double* r = s_rec + (i0 + k) * RLEN;
        const double lb = r[sb] * r[2 * EV + 1];
        const double f = (q == 0) ? 1.0 : fb[(i0 + k) * fs];
        acc[k] = fma(lb, f, acc[k]);
      }
    }
    if (act) s_wp_row[o] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// uniform forward kernel (value + small per-step record when p.records != NULL)
// ---------------------------------------------------------------------------------------------
template <int EV, int MAXT>
__global__ void __launch_bounds__(MAXT, UNI_FWD_MINCTAS(EV, MAXT)) uniform_fwd_kernel(const RolloutParams p) {
  extern __shared__ __align__(16) double sm[];
  constexpr int E = EV;
  const int tid = threadIdx.x, lane = tid & 31, NT = blockDim.x;
  const int D = p.D, N = p.N, NP = p.NP, DP = p.DP, Na = p.Na, H = p.H;
  constexpr int P = E * (E + 1) / 2;
  const UniLayout L = make_uni_layout(EV, false, NP, DP, D, H, Na);
  double* s_rec = sm + L.rec; double* s_tail = sm + L.tail; double* s_out = sm + L.out;
  double* s_m = sm + L.m; double* s_s = sm + L.s; double* s_mu = sm + L.mu; double* s_A = sm + L.A;
  double* s_Q = sm + L.Q; double* s_misc = sm + L.misc; double* s_M = sm + L.M; double* s_V = sm + L.V;
  double* s_acc = sm + L.acc; double* s_am = sm + L.am; double* s_r = sm + L.r; double* s_rv = sm + L.rv;
  int* s_int = reinterpret_cast<int*>(sm + L.ints);
  double* s_part = sm + L.part; double* s_wp = sm + L.wp; double* s_S = sm + L.S; double* s_cst = sm + L.cst;
  const int nOut = L.nOut, warp = tid >> 5, nwarps = NT >> 5;
  const UniRecLayout RL = uni_rec_layout(E);
  // cost description staged in shared memory (the stage cost is on the serial path of every step)
  for (int i = tid; i < E + Na; i += NT) s_cst[i] = p.c_target[i];
  for (int i = tid; i < (E + Na) * (E + Na); i += NT) s_cst[GPMPC_MAX_D + i] = p.c_W[i];
  for (int i = tid; i < E * E; i += NT) s_cst[GPMPC_MAX_D + GPMPC_MAX_D * GPMPC_MAX_D + i] = p.c_WT[i];
  const CostView cv{s_cst, s_cst + GPMPC_MAX_D, s_cst + GPMPC_MAX_D + GPMPC_MAX_D * GPMPC_MAX_D, p.c_smin, p.c_smax,
                    p.kappa, p.use_constraints};
  const double* il2 = p.il2;            // row 0 (all rows equal)
  const double s2 = p.s2[0];
  exp2s_fill(p.exp2tab, tid, NT);
  // beta_a,j of the hot-loop records does not depend on the candidate or the step: written once per CTA
  for (int o = tid; o < NP * E; o += NT) s_rec[(o / E) * (2 * EV + 2) + EV + 1 + (o % E)] = __ldg(p.betaT + o);
  __syncthreads();

  long long clk_ = clock64();
  UNI_CTA_BEGIN();
  const int C = p.cluster;                                  // CTAs per candidate (1: no cluster)
  const int crank = C > 1 ? (int)uni_cluster_rank() : 0, cid = blockIdx.x / C;
  const bool lead = crank == 0;                             // writes the outputs
  double* cl_acc = C > 1 ? p.ws_cl + (size_t)cid * 3 * UNI_CL_ACC : nullptr;
  int gstep = 0;                                            // running step count: accumulator buffer = gstep % 3
  for (int cand_static = cid;; cand_static += gridDim.x / C) {
    // clusters walk the candidates in lock step (their barriers must match); single CTAs draw from the global counter
    const int cand = C > 1 ? cand_static : uni_next_candidate(p, 0, s_int, tid);
    if (cand >= p.B) break;
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
    if (tid < Na) {
      double cum = 0.0;
      for (int t = 0; t < H; t++) {
        double raw = p.actions_mpc[(size_t)cand * H * Na + t * Na + tid], am;
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
    __syncthreads();
    for (int t = 1; t <= H; t++) {
      // ---- P0: model input, stage cost, shared small matrices
      if (tid < D) {
        double v;
        if (tid < E) v = s_mu[tid];
        else if (tid < E + Na) v = s_am[(t - 1) * Na + (tid - E)];
        else v = (double)(p.iter_ctrl + t - 1);
        s_m[tid] = v;
      }
      if (warp == 1) {
        double cmu, cvar;
        stage_cost_warp(cv, E, Na, s_mu, s_s, s_am + (t - 1) * Na, lane, sm + L.cstw, cmu, cvar);
        if (lane == 0) { s_r[t - 1] = -cmu; s_rv[t - 1] = cvar; }
      }
      // the two E x E inverses of the step, one warp each (warp_spd_inv_det): these used to be serial single-thread
      // routines with the whole CTA waiting at the barrier below
      if (warp == 2) {   // A = (s + diag(1 / il2))^-1, c = s2 / sqrt(det(.) prod il2)   (same for all GPs)
        for (int o = lane; o < EV * EV; o += 32) s_A[o] = s_s[o] + ((o / EV == o % EV) ? 1.0 / il2[o / EV] : 0.0);
        __syncwarp();
        const double det = warp_spd_inv_det<EV>(s_A, lane);
        if (lane == 0) {
          double pl = 1.0;
          for (int e = 0; e < EV; e++) pl *= il2[e];
          s_misc[0] = s2 / sqrt(det * pl);
        }
      }
      if (warp == 3) {   // pair matrices (pair_matrices in gpmpc_common.cuh): T = I + sq s sq, Rinv = sq^-1 T^-1 sq, Q = Rinv s / 2
        double* s_sq = s_S + EV * EV;   // s_S (2 E^2 doubles) is free until P4
        if (lane < EV) s_sq[lane] = sqrt(2.0 * il2[lane]);
        __syncwarp();
        for (int o = lane; o < EV * EV; o += 32) {
          const int e = o / EV, f = o - e * EV;
          s_Q[o] = s_sq[e] * s_s[o] * s_sq[f] + (e == f ? 1.0 : 0.0);
        }
        __syncwarp();
        const double detR = warp_spd_inv_det<EV>(s_Q, lane);
        for (int o = lane; o < EV * EV; o += 32) {
          const int e = o / EV, f = o - e * EV;
          s_S[o] = s_Q[o] * s_sq[f] / s_sq[e];
        }
        __syncwarp();
        for (int o = lane; o < EV * EV; o += 32) {
          const int e = o / EV, f = o - e * EV;
          double v = 0.0;
#pragma unroll
          for (int k = 0; k < EV; k++) v += s_S[e * EV + k] * s_s[k * EV + f];
          s_Q[o] = 0.5 * v;
        }
        if (lane == 0) s_misc[1] = detR;
      }
      for (int o = tid; o < nwarps * L.partlen; o += NT) s_part[o] = 0.0;
      if (tid == 0) s_int[0] = 0;
      __syncthreads();
      UNI_CLK(0);
      if (tid == 100) {
        double chk = 0.0;
        for (int d = 0; d < D; d++) chk += s_m[d];
        for (int e = 0; e < EV * EV; e++) chk += s_s[e];
        s_int[1] = isfinite(chk) ? 0 : 1;
      }
      // ---- P1: nu, shared exponent terms, hot-loop record (thread per training point)
      // two training points per thread in flight: this phase is load / dependent-FMA latency (P1a 16.6 k -> 13.8 k clocks)
      constexpr int P1U = (UNI_P1_UNROLL == 2 && EV <= 5) ? 2 : 1;   // (the 255-register kernels spill with two in flight)
#pragma unroll P1U
      for (int i = tid; i < NP; i += NT) {
        double nu[GPMPC_MAX_D];
#pragma unroll
        for (int d = 0; d < GPMPC_MAX_D; d++) nu[d] = (i < N && d < D) ? (uni_ldg_x(p.x + (size_t)i * D + d) - s_m[d]) : 0.0;
        double quad = 0.0, head = 0.0, tail = 0.0, zqz = 0.0;
#pragma unroll
        for (int e = 0; e < EV; e++) {
          double r = 0.0, q = 0.0;
#pragma unroll
          for (int f = 0; f < EV; f++) {
            r = fma(s_A[e * EV + f], nu[f], r);
            q = fma(s_Q[e * EV + f], nu[f] * il2[f], q);
          }
          quad = fma(nu[e], r, quad);
          head = fma(nu[e] * nu[e], il2[e], head);
          zqz = fma(nu[e] * il2[e], q, zqz);
        }
#pragma unroll
        for (int d = EV; d < GPMPC_MAX_D; d++)
          if (d < D) tail = fma(nu[d] * nu[d], il2[d], tail);
        const double ei = (i < N) ? exp2s((-0.5 * GPMPC_EXP2S_SCALE) * (quad + tail)) : 0.0;
        double* rec = s_rec + i * (2 * EV + 2);
#pragma unroll
        for (int e = 0; e < EV; e++) rec[e] = nu[e];
        rec[EV] = (i < N) ? GPMPC_EXP2S_SCALE * (-0.5 * (head + tail) + zqz) : 0.0;   // table units; log s2 factored out (s2^2 applied at the end)
        rec[2 * EV + 1] = ei;   // spare slot of the (even-length) record
#pragma unroll
        for (int d = EV; d < GPMPC_MAX_D; d++)
          if (d < D) s_tail[i * L.tlen + d - EV] = nu[d];
      }
      __syncthreads();
      UNI_CLK(4);
      // ---- P1b: mean-part moments h_a, g_a (lane per output, warp per slice of points; summed over warps in P4)
      {
        const int per = ((NP + nwarps - 1) / nwarps + 3) & ~3;   // slices of whole groups of 4 points (NP is a multiple of 64)
        const int ibeg = min(NP, warp * per);
        uni_moments_slice<EV>(s_rec, s_tail, L.tlen, nOut, ibeg, min(NP, ibeg + per), lane, s_wp + warp * L.wplen);
      }
      UNI_CLK(1);
      // ---- P3: one sweep over the upper tile triangle for all pairs.  Static balanced split: row block I (64 rows)
      //      holds (64 / p.seg) (nrb - I) chunks of p.seg columns from its diagonal on; the chunks of all row blocks,
      //      in row-major order, are dealt to the warps in equal contiguous runs (cut at row-block boundaries).
      {
        constexpr int TR = (UNI_USE_QUAD_FWD(EV) || UNI_USE_MMA8(EV)) ? 32 : 64;        // rows (and columns) of a tile
        const int CH = min(p.seg, TR), nrb = NP / TR, cpt = TR / CH;       // chunks per tile
        const int T = cpt * nrb * (nrb + 1) / 2, per = (T + nwarps * C - 1) / (nwarps * C);
        int c0 = (crank * nwarps + warp) * per;
        const int c1 = min(T, c0 + per);
        int I = 0, base = 0;
        while (c0 < c1) {
          while (c0 >= base + cpt * (nrb - I)) { base += cpt * (nrb - I); I++; }
          const int ce = min(c1, base + cpt * (nrb - I));
          if (UNI_USE_MMA8(EV))
            uni_fwd_item_mma8<EV>(p, s_rec, s_Q, il2, I, TR * I + CH * (c0 - base), TR * I + CH * (ce - base), lane,
                                  s_part + warp * L.partlen);
          else if (UNI_USE_QUAD_FWD(EV))
            uni_fwd_item_mma<EV>(p, s_rec, s_Q, il2, I, TR * I + CH * (c0 - base), TR * I + CH * (ce - base), lane,
                                 s_part + warp * L.partlen);
          else
            uni_fwd_item<EV>(p, s_rec, s_Q, il2, I, TR * I + CH * (c0 - base), TR * I + CH * (ce - base), lane,
                             s_part + warp * L.partlen);
          c0 = ce;
        }
      }
      __syncthreads();
      if (C > 1) {   // cluster: this CTA's share of the sweep sums -> L2 accumulator of the step, then the cluster barrier
        double* acc = cl_acc + (gstep % 3) * UNI_CL_ACC;
        if (warp == 0)
          for (int k = lane; k <= P; k += 32) {
            double v = 0.0;
            for (int w = 0; w < nwarps; w++) v += s_part[w * L.partlen + k];
            uni_red_add(acc + k, v);
          }
        if (lead && warp == 1)
          for (int k = lane; k < UNI_CL_ACC; k += 32) cl_acc[((gstep + 1) % 3) * UNI_CL_ACC + k] = 0.0;   // next step's buffer
        __threadfence();
        uni_cluster_sync();
      }
      UNI_CLK(2);
      // ---- P4: mean, V, S, recurrence (gp_model.py:150-180, :92-99): warp 0, one small stage per __syncwarp
      if (warp == 0) {
        const double c = s_misc[0], detR = s_misc[1], rs = 1.0 / sqrt(detR);
        const bool bad = s_int[1] != 0;
        double* rec = (p.records && lead) ? p.records + ((size_t)cand * H + (t - 1)) * RL.size : nullptr;
        // A: add up the per-warp partial rows
        for (int o = lane; o < E * nOut; o += 32) {
          double v = 0.0;
          for (int w = 0; w < nwarps; w++) v += s_wp[w * L.wplen + o];
          s_out[o] = v;
        }
        for (int k = lane; k <= P; k += 32) {
          double v = 0.0;
          if (C > 1) v = __ldcg(cl_acc + (gstep % 3) * UNI_CL_ACC + k);
          else for (int w = 0; w < nwarps; w++) v += s_part[w * L.partlen + k];
          s_acc[k] = v;
        }
        __syncwarp();
        // B: M_a = c h_a ; V_a = c A g_a (state block), c g_a il2 (action / time dims)
        for (int o = lane; o < E * D; o += 32) {
          const int a = o / D, d = o - a * D;
          const double* g = s_out + a * nOut + 1;
          double v;
          if (d < EV) {
            v = 0.0;
#pragma unroll
            for (int f = 0; f < EV; f++) v = fma(s_A[d * EV + f], g[f], v);
          } else {
            v = g[d] * il2[d];
          }
          v *= c;
          s_V[o] = v;
          if (rec && d < E) { rec[RL.offV + a * E + d] = v; rec[RL.offG + a * E + d] = g[d]; }
        }
        if (lane < E) {
          const double h = s_out[lane * nOut];
          s_M[lane] = c * h;
          if (rec) { rec[RL.offM + lane] = c * h; rec[RL.offH + lane] = h; }
        }
        __syncwarp();
        // C: S_ab (pairs) and s V^T
        const double trc = s_acc[P];
        for (int pr = lane; pr < P; pr += 32) {
          int a = 0, w = pr;
          while (w >= E - a) { w -= E - a; a++; }
          const int b = a + w;
          const double Sraw = s2 * s2 * (s_acc[pr] - (a == b ? trc : 0.0));
          if (rec) rec[RL.offS + pr] = Sraw;
          double v = Sraw * rs - s_M[a] * s_M[b] + (a == b ? s2 : 0.0);
          if (bad) v = nan("");
          s_S[a * E + b] = v;
          s_S[b * E + a] = v;
        }
        for (int o = lane; o < E * E; o += 32) {
          const int e = o / E, a = o - e * E;
          double v = 0.0;
#pragma unroll
          for (int k = 0; k < E; k++) v = fma(s_s[e * E + k], s_V[a * D + k], v);
          s_S[E * E + o] = v;            // sv[e][a]
        }
        __syncwarp();
        // D: recurrence
        constexpr int NR = (E * E + 31) / 32;
        double sn[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) {
          const int o = lane + 32 * r;
          sn[r] = 0.0;
          if (o < E * E) {
            const int e = o / E, f = o - e * E;
            sn[r] = s_S[o] + s_s[o] + s_S[E * E + o] + s_S[E * E + f * E + e];
          }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < NR; r++) {
          const int o = lane + 32 * r;
          if (o < E * E) {
            s_s[o] = sn[r];
            if (lead) p.states_var[((size_t)cand * (H + 1) + t) * E * E + o] = sn[r];
          }
        }
        if (lane < E) {
          double v = s_mu[lane] + s_M[lane];
          if (bad) v = nan("");
          s_mu[lane] = v;
          if (lead) p.states_mu[((size_t)cand * (H + 1) + t) * E + lane] = v;
        }
      }
      gstep++;
      __syncthreads();
      UNI_CLK(3);
    }
    if (tid == 0 && lead) {
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
  UNI_CTA_END(0);
}

// ---------------------------------------------------------------------------------------------
// reverse sweep hot loop: upper-triangle sweep with the adjoint-weighted coefficient
//   W_ij = p_i . beta_j - wbar iK_ij ,  w_ij = W_ij Eh_ij ; rows: rho_i, xi_i ; columns: gam_j
// Row and column sums leave the warp as float64 reductions into the CTA's global scratch (native RED.ADD.F64 at L2,
// fire and forget); float64 atomics on shared memory are CAS spin loops and cost ~15 % of an item.
// ---------------------------------------------------------------------------------------------
// columns [jbeg, jend) (a multiple of 8) of the reverse sweep for the lane's rows i0, i0 + 1, with coefficient row
// vectors p0, p1 and trace weights wb0, wb1 as given: the caller folds the row factor e_i of the exponential into them
// (and halves them on the diagonal tile), so w_ij = (p_i . beta_j - wb_i iK_ij) Eh'_ij needs no per-element row term
// (SH: residual shifts kr0, kr1 of far-away rows, see uni_fwd_cols)
template <int EV, bool SH>
__device__ __forceinline__ void uni_bwd_cols(const RolloutParams& p, const double* __restrict__ s_rec, int i0,
                                             int jbeg, int jend, const double (&u0)[EV], const double (&u1)[EV],
                                             double kr0, double kr1, const double (&p0)[EV], const double (&p1)[EV],
                                             double wb0, double wb1, double& rho0, double& rho1, double (&xi0)[EV],
                                             double (&xi1)[EV], int lane, double* __restrict__ g_gam,
                                             double* __restrict__ scr) {
  constexpr int E = EV;
  const int NP = p.NP;
  const double* __restrict__ ik0 = p.iK + (size_t)jbeg * NP + i0;
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    double v[8];
#pragma unroll
    for (int jp = 0; jp < 4; jp++) {   // pairs of columns: 2 columns x 2 rows = 4 independent chains per warp
      const int j = j0 + 2 * jp;
      double na[EV], nb[EV], ba[E], bb[E], ka, kb;
      uni_load_rec<EV>(s_rec, j, na, ka, ba);
      uni_load_rec<EV>(s_rec, j + 1, nb, kb, bb);
      const double2 ika = ldg_stream2<(EV <= 5)>(ik0);
      const double2 ikb = ldg_stream2<(EV <= 5)>(ik0 + NP);
      double c[4] = {-wb0 * ika.x, -wb1 * ika.y, -wb0 * ikb.x, -wb1 * ikb.y};
      ik0 += 2 * (size_t)NP;
#pragma unroll
      for (int b = 0; b < E; b++) {   // serpentine order (operand-reuse cache, see uni_fwd_cols)
        c[0] = fma(p0[b], ba[b], c[0]);
        c[1] = fma(p1[b], ba[b], c[1]);
        c[3] = fma(p1[b], bb[b], c[3]);
        c[2] = fma(p0[b], bb[b], c[2]);
      }
      double t[4], w[4];
      if (SH) { t[0] = kr0 + ka; t[1] = kr1 + ka; t[2] = kr0 + kb; t[3] = kr1 + kb; }
      else { t[0] = ka; t[1] = ka; t[2] = kb; t[3] = kb; }
#pragma unroll
      for (int e = 0; e < EV; e++) {
        t[0] = fma(u0[e], na[e], t[0]);
        t[1] = fma(u1[e], na[e], t[1]);
        t[3] = fma(u1[e], nb[e], t[3]);
        t[2] = fma(u0[e], nb[e], t[2]);
      }
      exp2s_x4<false>(t, w);   // run-time table base: 2 % faster here (common.cuh, exp2s_entry)
#pragma unroll
      for (int q = 0; q < 4; q++) w[q] *= c[q];
      rho0 += w[0] + w[2];
      rho1 += w[1] + w[3];
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
      v[2 * jp] = w[0] + w[1];
      v[2 * jp + 1] = w[2] + w[3];
    }
#if UNI_BWD_COLRED_SMEM
    const double tot = col_reduce8s(v, lane, scr);   // column sums through the warp's shared-memory scratch
    if (lane < 8) uni_red_add(g_gam + j0 + lane, tot);
#else
    int col;
    const double tot = col_reduce8(v, lane, col);
    if ((lane & 3) == 0) uni_red_add(g_gam + j0 + col, tot);
#endif
  }
}

// One run of columns [jbeg, jend), jbeg >= 64 I, of row block I.  w is symmetric, and the O(N) reductions that consume
// the sums (B3b) only need  g_i = rho_i + gam_i  and  sum_i z_i,k xi_i,l + xi_i,k z_i,l : both come out the same if the
// diagonal tile is swept in full with HALF weights instead of its upper triangle -- no element masks anywhere.
template <int EV>
__device__ __forceinline__ void uni_bwd_item(const RolloutParams& p, const double* __restrict__ s_rec,
                                             const double* __restrict__ Qm, const double* __restrict__ il2,
                                             const double* __restrict__ Om, double wbar, int I, int jbeg, int jend,
                                             int lane, double* __restrict__ g_gam, double* __restrict__ g_rho,
                                             double* __restrict__ g_xi, double* __restrict__ scr) {
  constexpr int E = EV;
  const int NP = p.NP;
  const int i0 = 64 * I + 2 * lane, i1 = i0 + 1;   // adjacent rows (16-byte iK loads), as in the forward sweep
  double u0[EV], u1[EV], p0[E], p1[E], kr0, kr1;
  {
    double n0[EV], n1[EV], b0[E], b1[E];
    uni_load_rec<EV>(s_rec, i0, n0, kr0, b0);
    uni_load_rec<EV>(s_rec, i1, n1, kr1, b1);
#pragma unroll
    for (int e = 0; e < EV; e++) {
      double q0 = 0.0, q1 = 0.0;
#pragma unroll
      for (int f = 0; f < EV; f++) {
        q0 = fma(Qm[e * EV + f], n0[f] * il2[f], q0);
        q1 = fma(Qm[e * EV + f], n1[f] * il2[f], q1);
      }
      u0[e] = (2.0 * GPMPC_EXP2S_SCALE) * q0 * il2[e];   // exponent in table units (exp2s)
      u1[e] = (2.0 * GPMPC_EXP2S_SCALE) * q1 * il2[e];
    }
#pragma unroll
    for (int a = 0; a < E; a++) {
      double v0 = 0.0, v1 = 0.0;
#pragma unroll
      for (int b = 0; b < E; b++) { v0 = fma(Om[a * E + b], b0[b], v0); v1 = fma(Om[a * E + b], b1[b], v1); }
      p0[a] = v0;
      p1[a] = v1;
    }
  }
  double rho0 = 0.0, rho1 = 0.0, xi0[EV], xi1[EV];
#pragma unroll
  for (int e = 0; e < EV; e++) { xi0[e] = 0.0; xi1[e] = 0.0; }
  const int jd1 = 64 * I + 64;
  jend = min(jend, (p.N + 7) & ~7);   // zero-padded columns contribute exact zeros: skip them (8 columns per round)
  // row factor e_i of the exponential (uni_fwd_cols) folded into the coefficient row vectors and the trace weights
  const double e0 = uni_row_factor(kr0), e1 = uni_row_factor(kr1);   // kr0, kr1 become residual shifts
  const bool far = __any_sync(0xffffffffu, kr0 != 0.0 || kr1 != 0.0);
  const double wb0 = wbar * e0, wb1 = wbar * e1;
#pragma unroll
  for (int a = 0; a < E; a++) { p0[a] *= e0; p1[a] *= e1; }
  if (jbeg < jd1) {
#pragma unroll
    for (int a = 0; a < E; a++) { p0[a] *= 0.5; p1[a] *= 0.5; }
    if (far) uni_bwd_cols<EV, true>(p, s_rec, i0, jbeg, min(jend, jd1), u0, u1, kr0, kr1, p0, p1, 0.5 * wb0, 0.5 * wb1, rho0, rho1,
                                    xi0, xi1, lane, g_gam, scr);
    else uni_bwd_cols<EV, false>(p, s_rec, i0, jbeg, min(jend, jd1), u0, u1, kr0, kr1, p0, p1, 0.5 * wb0, 0.5 * wb1, rho0, rho1,
                                 xi0, xi1, lane, g_gam, scr);
#pragma unroll
    for (int a = 0; a < E; a++) { p0[a] *= 2.0; p1[a] *= 2.0; }
  }
  if (far) uni_bwd_cols<EV, true>(p, s_rec, i0, max(jbeg, jd1), jend, u0, u1, kr0, kr1, p0, p1, wb0, wb1, rho0, rho1, xi0, xi1,
                                  lane, g_gam, scr);
  else uni_bwd_cols<EV, false>(p, s_rec, i0, max(jbeg, jd1), jend, u0, u1, kr0, kr1, p0, p1, wb0, wb1, rho0, rho1, xi0, xi1,
                               lane, g_gam, scr);
  uni_red_add(g_rho + i0, rho0);
  uni_red_add(g_rho + i1, rho1);
#pragma unroll
  for (int e = 0; e < EV; e++) { uni_red_add(g_xi + (size_t)e * NP + i0, xi0[e]); uni_red_add(g_xi + (size_t)e * NP + i1, xi1[e]); }
}

// Tensor-core variant of the reverse sweep (UNI_USE_MMA, see uni_fwd_cols_mma for the tile layout): BOTH bilinear forms of
// an element are DMMAs -- the exponent  t_ij = kap_j + u_i . nu_j  and the coefficient  c_ij = p_i . beta_j - wb_i iK_ij
// (the trace part is the accumulator's initial value) -- so an element costs 15 float64 instructions + 2/8 DMMA instead
// of 23.  Column sums: the lane's two columns summed over its 4 rows, then over the 8 row groups of the warp (three
// exchanges, the first one halving); row sums stay in registers until the end of the item.
template <int EV, bool SH>
__device__ __forceinline__ void uni_bwd_cols_mma(const RolloutParams& p, const double* __restrict__ s_rec, int ib,
                                                 int jbeg, int jend, int lane, const double (&ua)[4][(EV + 3) / 4],
                                                 const double (&pa)[4][(EV + 3) / 4], const double (&kr)[4],
                                                 const double (&wb)[4], double (&rho)[4], double (&xi)[4][EV],
                                                 double* __restrict__ g_gam) {
  constexpr int RLEN = 2 * EV + 2, KS = (EV + 3) / 4;
  const int g = lane >> 2, q = lane & 3;
  const size_t rs = 8 * (size_t)p.NP;
  const double* __restrict__ ik = p.iK + (size_t)(ib + g) * p.NP + jbeg + 2 * q;
  const double* __restrict__ rb = s_rec + (jbeg + g) * RLEN + q;     // B fragment sources: nu (slot q), beta (slot EV + 1 + q)
  const double* __restrict__ rc = s_rec + (jbeg + 2 * q) * RLEN;     // the lane's two columns
  const bool up = (lane & 16) != 0;
  double* __restrict__ gcol = g_gam + jbeg + 2 * q + (up ? 1 : 0);
#pragma unroll 1
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    double bfn[KS], bfb[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      bfn[ks] = rb[4 * ks];
      bfb[ks] = (EV % 4 == 0 || 4 * ks + q < EV) ? rb[EV + 1 + 4 * ks] : 0.0;
    }
    const double ka0 = rc[EV], ka1 = rc[RLEN + EV];
    double n0[EV], n1[EV];
#pragma unroll
    for (int e = 0; e < EV; e++) { n0[e] = rc[e]; n1[e] = rc[RLEN + e]; }
    double2 ikv[4];
#pragma unroll
    for (int m = 0; m < 4; m++) ikv[m] = ldg_stream2<(EV <= 5)>(ik + m * rs);
    double t[8], c[8], w[8];
#pragma unroll
    for (int m = 0; m < 4; m++) {
      double c0 = -wb[m] * ikv[m].x, c1 = -wb[m] * ikv[m].y;
      double d0 = SH ? ka0 + kr[m] : ka0, d1 = SH ? ka1 + kr[m] : ka1;
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        dmma_m8n8k4(c0, c1, pa[m][ks], bfb[ks]);
        dmma_m8n8k4(d0, d1, ua[m][ks], bfn[ks]);
      }
      c[2 * m] = c0; c[2 * m + 1] = c1;
      t[2 * m] = d0; t[2 * m + 1] = d1;
    }
    exp2s_xn<8, false>(t, w);
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] *= c[k];
#pragma unroll
    for (int m = 0; m < 4; m++) rho[m] += w[2 * m] + w[2 * m + 1];
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int e = 0; e < EV; e++) xi[m][e] = fma(w[2 * m], n0[e], xi[m][e]);
#pragma unroll
    for (int m = 0; m < 4; m++)
#pragma unroll
      for (int e = 0; e < EV; e++) xi[m][e] = fma(w[2 * m + 1], n1[e], xi[m][e]);
    const double v0 = (w[0] + w[2]) + (w[4] + w[6]), v1 = (w[1] + w[3]) + (w[5] + w[7]);
    // lanes 0-15 keep column 2 q, lanes 16-31 column 2 q + 1; then the sum over the row groups g
    double a = (up ? v1 : v0) + __shfl_xor_sync(0xffffffffu, up ? v0 : v1, 16);
    a += __shfl_xor_sync(0xffffffffu, a, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    if ((lane & 12) == 0) uni_red_add(gcol, a);
    gcol += 8;
    ik += 8;
    rb += 8 * RLEN;
    rc += 8 * RLEN;
  }
}

// One run of columns [jbeg, jend) (multiples of 8, jbeg >= 32 I) of the 32-row block I: see uni_bwd_item for the algebra.
template <int EV>
__device__ __forceinline__ void uni_bwd_item_mma(const RolloutParams& p, const double* __restrict__ s_rec,
                                                 const double* __restrict__ Qm, const double* __restrict__ il2,
                                                 const double* __restrict__ Om, double wbar, int I, int jbeg, int jend,
                                                 int lane, double* __restrict__ g_gam, double* __restrict__ g_rho,
                                                 double* __restrict__ g_xi) {
  constexpr int E = EV, RLEN = 2 * EV + 2, KS = (EV + 3) / 4;
  const int NP = p.NP;
  const int g = lane >> 2, q = lane & 3, ib = 32 * I;
  double ua[4][KS], pa[4][KS], kr[4], wb[4], ei[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const double* rec = s_rec + (ib + 8 * m + g) * RLEN;
    kr[m] = rec[EV];
    ei[m] = uni_row_factor(kr[m]);   // kr becomes the residual shift
    wb[m] = wbar * ei[m];
#pragma unroll
    for (int ks = 0; ks < KS; ks++) {
      const int ee = 4 * ks + q;
      double acc = 0.0, pv = 0.0;
      if (ee < EV) {
#pragma unroll
        for (int f = 0; f < EV; f++) acc = fma(Qm[ee * EV + f], rec[f] * il2[f], acc);
        acc *= (2.0 * GPMPC_EXP2S_SCALE) * il2[ee];   // exponent in table units (exp2s)
#pragma unroll
        for (int b = 0; b < E; b++) pv = fma(Om[ee * E + b], rec[EV + 1 + b], pv);
        pv *= ei[m];                                  // row factor of the exponential folded into the coefficients
      }
      ua[m][ks] = acc;
      pa[m][ks] = pv;
    }
  }
  double rho[4], xi[4][EV];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    rho[m] = 0.0;
#pragma unroll
    for (int e = 0; e < EV; e++) xi[m][e] = 0.0;
  }
  const int jd1 = ib + 32;
  jend = min(jend, (p.N + 7) & ~7);   // zero-padded columns contribute exact zeros: skip them
  const bool far = __any_sync(0xffffffffu, kr[0] != 0.0 || kr[1] != 0.0 || kr[2] != 0.0 || kr[3] != 0.0);
  if (jbeg < jd1) {   // diagonal tile: swept in full with HALF weights
    double ph[4][KS], wh[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
      wh[m] = 0.5 * wb[m];
#pragma unroll
      for (int ks = 0; ks < KS; ks++) ph[m][ks] = 0.5 * pa[m][ks];
    }
    if (far) uni_bwd_cols_mma<EV, true>(p, s_rec, ib, jbeg, min(jend, jd1), lane, ua, ph, kr, wh, rho, xi, g_gam);
    else uni_bwd_cols_mma<EV, false>(p, s_rec, ib, jbeg, min(jend, jd1), lane, ua, ph, kr, wh, rho, xi, g_gam);
  }
  if (far) uni_bwd_cols_mma<EV, true>(p, s_rec, ib, max(jbeg, jd1), jend, lane, ua, pa, kr, wb, rho, xi, g_gam);
  else uni_bwd_cols_mma<EV, false>(p, s_rec, ib, max(jbeg, jd1), jend, lane, ua, pa, kr, wb, rho, xi, g_gam);
  // a row's sums are spread over the 4 lanes of its quad: add them up, lane q = 0 sends them to the scratch
#pragma unroll
  for (int m = 0; m < 4; m++) {
    double v = rho[m];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (q == 0) uni_red_add(g_rho + ib + 8 * m + g, v);
#pragma unroll
    for (int e = 0; e < EV; e++) {
      double x = xi[m][e];
      x += __shfl_xor_sync(0xffffffffu, x, 1);
      x += __shfl_xor_sync(0xffffffffu, x, 2);
      if (q == (e & 3)) uni_red_add(g_xi + (size_t)e * NP + ib + 8 * m + g, x);
    }
  }
}

// Reverse sweep, large state dimensions (see uni_fwd_cols_mma8): exponent, coefficient and the xi row sums are DMMAs.
template <int EV, bool SH>
__device__ __forceinline__ void uni_bwd_cols_mma8(const RolloutParams& p, const double* __restrict__ s_rec, int ib,
                                                  int jbeg, int jend, int lane, const double (&ua)[4][2],
                                                  const double (&pa)[4][2], const double (&kr)[4], const double (&wb)[4],
                                                  double (&rho)[4], double (&xi)[4][2], double* __restrict__ g_gam) {
  constexpr int RLEN = 2 * EV + 2;
  const int g = lane >> 2, q = lane & 3;
  const size_t rs = 8 * (size_t)p.NP;
  const double* __restrict__ ik = p.iK + (size_t)(ib + g) * p.NP + jbeg + 2 * q;
  const double* __restrict__ rb = s_rec + (jbeg + g) * RLEN + q;     // B fragments of column j0 + g: nu (slots q, 4 + q), beta (EV + 1 + ...)
  const double* __restrict__ rc = s_rec + (jbeg + 2 * q) * RLEN;     // the lane's two columns
  const bool up = (lane & 16) != 0;
  double* __restrict__ gcol = g_gam + jbeg + 2 * q + (up ? 1 : 0);
#pragma unroll 1
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    const double bfn0 = rb[0], bfn1 = rb[4], bfb0 = rb[EV + 1], bfb1 = rb[EV + 5];
    const double ka0 = rc[EV], ka1 = rc[RLEN + EV];
    const double bx0 = rc[g], bx1 = rc[RLEN + g];   // xi B fragments: nu_{column, g}
    double2 ikv[4];
#pragma unroll
    for (int m = 0; m < 4; m++) ikv[m] = ldg_stream2<false>(ik + m * rs);
    double t[8], c[8], w[8];
#pragma unroll
    for (int m = 0; m < 4; m++) {
      double c0 = -wb[m] * ikv[m].x, c1 = -wb[m] * ikv[m].y;
      double d0 = SH ? ka0 + kr[m] : ka0, d1 = SH ? ka1 + kr[m] : ka1;
      dmma_m8n8k4(c0, c1, pa[m][0], bfb0);
      dmma_m8n8k4(d0, d1, ua[m][0], bfn0);
      dmma_m8n8k4(c0, c1, pa[m][1], bfb1);
      dmma_m8n8k4(d0, d1, ua[m][1], bfn1);
      c[2 * m] = c0; c[2 * m + 1] = c1;
      t[2 * m] = d0; t[2 * m + 1] = d1;
    }
    exp2s_xn<8>(t, w);
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] *= c[k];
#pragma unroll
    for (int m = 0; m < 4; m++) dmma_m8n8k4(xi[m][0], xi[m][1], w[2 * m], bx0);
#pragma unroll
    for (int m = 0; m < 4; m++) dmma_m8n8k4(xi[m][0], xi[m][1], w[2 * m + 1], bx1);
#pragma unroll
    for (int m = 0; m < 4; m++) rho[m] += w[2 * m] + w[2 * m + 1];
    const double v0 = (w[0] + w[2]) + (w[4] + w[6]), v1 = (w[1] + w[3]) + (w[5] + w[7]);
    // lanes 0-15 keep column 2 q, lanes 16-31 column 2 q + 1; then the sum over the row groups g
    double a = (up ? v1 : v0) + __shfl_xor_sync(0xffffffffu, up ? v0 : v1, 16);
    a += __shfl_xor_sync(0xffffffffu, a, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    if ((lane & 12) == 0) uni_red_add(gcol, a);
    gcol += 8;
    ik += 8;
    rb += 8 * RLEN;
    rc += 8 * RLEN;
  }
}

template <int EV>
__device__ __forceinline__ void uni_bwd_item_mma8(const RolloutParams& p, const double* __restrict__ s_rec,
                                                  const double* __restrict__ Qm, const double* __restrict__ il2,
                                                  const double* __restrict__ Om, double wbar, int I, int jbeg, int jend,
                                                  int lane, double* __restrict__ g_gam, double* __restrict__ g_rho,
                                                  double* __restrict__ g_xi) {
  constexpr int E = EV, RLEN = 2 * EV + 2;
  const int NP = p.NP;
  const int g = lane >> 2, q = lane & 3, ib = 32 * I;
  double ua[4][2], pa[4][2], kr[4], wb[4], ei[4];
#pragma unroll
  for (int m = 0; m < 4; m++) {
    const double* rec = s_rec + (ib + 8 * m + g) * RLEN;
    kr[m] = rec[EV];
    ei[m] = uni_row_factor(kr[m]);   // kr becomes the residual shift
    wb[m] = wbar * ei[m];
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
      const int ee = 4 * ks + q;
      double acc = 0.0, pv = 0.0;
      if (ee < EV) {
#pragma unroll
        for (int f = 0; f < EV; f++) acc = fma(Qm[ee * EV + f], rec[f] * il2[f], acc);
        acc *= (2.0 * GPMPC_EXP2S_SCALE) * il2[ee];   // exponent in table units (exp2s)
#pragma unroll
        for (int b = 0; b < E; b++) pv = fma(Om[ee * E + b], rec[EV + 1 + b], pv);
        pv *= ei[m];                                  // row factor of the exponential folded into the coefficients
      }
      ua[m][ks] = acc;
      pa[m][ks] = pv;
    }
  }
  double rho[4], xi[4][2];
#pragma unroll
  for (int m = 0; m < 4; m++) { rho[m] = 0.0; xi[m][0] = 0.0; xi[m][1] = 0.0; }
  const int jd1 = ib + 32;
  jend = min(jend, (p.N + 7) & ~7);   // zero-padded columns contribute exact zeros: skip them
  const bool far = __any_sync(0xffffffffu, kr[0] != 0.0 || kr[1] != 0.0 || kr[2] != 0.0 || kr[3] != 0.0);
  if (jbeg < jd1) {   // diagonal tile: swept in full with HALF weights
    double ph[4][2], wh[4];
#pragma unroll
    for (int m = 0; m < 4; m++) { wh[m] = 0.5 * wb[m]; ph[m][0] = 0.5 * pa[m][0]; ph[m][1] = 0.5 * pa[m][1]; }
    if (far) uni_bwd_cols_mma8<EV, true>(p, s_rec, ib, jbeg, min(jend, jd1), lane, ua, ph, kr, wh, rho, xi, g_gam);
    else uni_bwd_cols_mma8<EV, false>(p, s_rec, ib, jbeg, min(jend, jd1), lane, ua, ph, kr, wh, rho, xi, g_gam);
  }
  if (far) uni_bwd_cols_mma8<EV, true>(p, s_rec, ib, max(jbeg, jd1), jend, lane, ua, pa, kr, wb, rho, xi, g_gam);
  else uni_bwd_cols_mma8<EV, false>(p, s_rec, ib, max(jbeg, jd1), jend, lane, ua, pa, kr, wb, rho, xi, g_gam);
  // xi: the accumulator entries are complete (outputs 2 q, 2 q + 1 of the lane's rows); rho: spread over the quad
#pragma unroll
  for (int m = 0; m < 4; m++) {
    double v = rho[m];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (q == 0) uni_red_add(g_rho + ib + 8 * m + g, v);
    if (2 * q < EV) uni_red_add(g_xi + (size_t)(2 * q) * NP + ib + 8 * m + g, xi[m][0]);
    if (2 * q + 1 < EV) uni_red_add(g_xi + (size_t)(2 * q + 1) * NP + ib + 8 * m + g, xi[m][1]);
  }
}

// O(N) reductions of the reverse sweep.  Every thread has summed, over its own training points, P "pair" values
// (k <= l, row-major) and up to GPMPC_MAX_D "single" values; the warp adds them up with halving exchanges (16 values
// per round, 16 shuffles instead of 80) and stores its totals as row[o]: o < D singles, o = D + pr pairs.  The per-warp
// rows are added up by the consumer.  Two uses:
//  B1 (mean part):  singles sum_i phi_i nu_i,d ;  pairs sum_i phi_i nu_i,k nu_i,l
//  B3 (pair part):  singles sum_i g_i nu_i,d   ;  pairs sum_i g_i z_k z_l + z_l x_k + z_k x_l   (z = nu * il2, state dims)
template <int EV>
__device__ __forceinline__ void uni_warp_point_sums(const double (&vp)[EV * (EV + 1) / 2], const double (&vs)[GPMPC_MAX_D],
                                                    int D, int lane, double* __restrict__ row) {
  constexpr int P = EV * (EV + 1) / 2, NCH = (P + GPMPC_MAX_D + 15) / 16;
#pragma unroll
  for (int ch = 0; ch < NCH; ch++) {
    if (16 * ch < P + D) {   // warp-uniform: rounds that hold only unused single slots are skipped
      double v16[16];
#pragma unroll
      for (int k = 0; k < 16; k++) {
        constexpr int dummy = 0; (void)dummy;
        const int sl = 16 * ch + k;
        v16[k] = (sl < P) ? vp[sl < P ? sl : 0] : ((sl - P < GPMPC_MAX_D) ? vs[(sl - P >= 0 && sl - P < GPMPC_MAX_D) ? sl - P : 0] : 0.0);
      }
      int idx;
      const double tot = warp_reduce_multi<16>(v16, lane, idx);
      const int sl = 16 * ch + idx;
      if ((lane & 1) == 0) {
        if (sl < P) row[D + sl] = tot;
        else if (sl - P < D) row[sl - P] = tot;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Step quantities of the reverse sweep that depend only on the FORWARD trajectory (not on the adjoints flowing
// back): the shared small matrices and the stage-cost adjoint.  With `premat` they are computed for all H steps in
// parallel (one thread per step) before the serial sweep, which shortens its single-thread critical path.
// ---------------------------------------------------------------------------------------------
template <int EV>
__device__ inline void uni_step_matrices(const double* sp, const double* il2, double s2, double* A, double* Q,
                                         double* Rinv, double& c, double& detR) {
  double Ca[EV * EV], det, pl = 1.0, Wd[EV];
  for (int e = 0; e < EV; e++)
    for (int f = 0; f < EV; f++) Ca[e * EV + f] = sp[e * EV + f] + (e == f ? 1.0 / il2[e] : 0.0);
  spd_inv_det<EV>(Ca, A, det);
  for (int e = 0; e < EV; e++) { pl *= il2[e]; Wd[e] = 2.0 * il2[e]; }
  c = s2 / sqrt(det * pl);
  pair_matrices<EV>(sp, Wd, Rinv, Q, detR);
}

// adjoint of the stage cost at (mup, sp, am) (setpoint_distance_reward_mapper.py:12-68): ADDS into dmu (E), da (Na),
// dS (E x E, symmetrised)
template <int EV>
__device__ inline void uni_stage_adjoint(const RolloutParams& p, int Na, double wmu, double wv, const double* mup,
                                         const double* sp, const double* am, double* dmu, double* da, double* dS) {
  constexpr int E = EV;
  const int Dc = E + Na;
  double e[GPMPC_MAX_D], We[GPMPC_MAX_D], sWe[EV], t1[EV * EV];
  for (int d = 0; d < Dc; d++) e[d] = (d < E ? mup[d] : am[d - E]) - p.c_target[d];
  for (int d = 0; d < Dc; d++) { double v = 0.0; for (int k = 0; k < Dc; k++) v += p.c_W[d * Dc + k] * e[k]; We[d] = v; }
  for (int i = 0; i < E; i++) { double v = 0.0; for (int k = 0; k < E; k++) v += sp[i * E + k] * We[k]; sWe[i] = v; }
  for (int d = 0; d < Dc; d++) {
    double v = 0.0;
    for (int i = 0; i < E; i++) v += p.c_W[i * Dc + d] * sWe[i];
    const double gd = wmu * 2.0 * We[d] + wv * 8.0 * v;
    if (d < E) dmu[d] += gd; else da[d - E] += gd;
  }
  for (int i = 0; i < E; i++)
    for (int k = 0; k < E; k++) { double v = 0.0; for (int l = 0; l < E; l++) v += p.c_W[l * Dc + i] * sp[k * E + l]; t1[i * E + k] = v; }
  for (int i = 0; i < E; i++)
    for (int k = 0; k < E; k++) {
      double v = 0.0;
      for (int l = 0; l < E; l++) v += t1[i * E + l] * p.c_W[k * Dc + l];
      const double full_ik = wmu * p.c_W[k * Dc + i] + wv * (4.0 * v + 4.0 * We[i] * We[k]);
      dS[i * E + k] += 0.5 * full_ik;
      dS[k * E + i] += 0.5 * full_ik;
    }
  if (p.use_constraints) {
    const double rt2 = 1.4142135623730951, ispi = 0.5641895835477563;
    for (int d = 0; d < E; d++) {
      double sig = sp[d * E + d];
      double zmin = (p.c_smin[d] - mup[d]) / (sig * rt2), zmax = (p.c_smax[d] - mup[d]) / (sig * rt2);
      double pmin = exp(-zmin * zmin) * ispi, pmax = exp(-zmax * zmax) * ispi;
      dmu[d] += wmu * (pmin - pmax) * (-1.0 / (sig * rt2));
      dS[d * E + d] += wmu * (pmin * (-zmin / sig) - pmax * (-zmax / sig));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// uniform reverse-sweep kernel: one CTA per candidate, t = H .. 1
// ---------------------------------------------------------------------------------------------
template <int EV, int MAXT>
__global__ void __launch_bounds__(MAXT, UNI_BWD_MINCTAS(EV, MAXT)) uniform_bwd_kernel(const RolloutParams p, double* __restrict__ grad) {
  extern __shared__ __align__(16) double sm[];
  constexpr int E = EV, P = E * (E + 1) / 2;
  const int tid = threadIdx.x, lane = tid & 31, NT = blockDim.x;
  const int D = p.D, N = p.N, NP = p.NP, DP = p.DP, Na = p.Na, H = p.H;
  const bool premat = p.premat != 0;
  const UniLayout L = make_uni_layout(EV, true, NP, DP, D, H, Na, p.premat == 1);
  // premat == 2: the per-step records live in this CTA's global scratch (written and read by this CTA only, between barriers)
  double* s_pre = p.premat == 2 ? p.ws_pre + (size_t)blockIdx.x * H * L.prelen : sm + L.pre;
  const int oA = 0, oQ = EV * EV, oRi = 2 * EV * EV, odS = 3 * EV * EV, oc = 4 * EV * EV, odet = oc + 1, odmu = oc + 2,
            oda = odmu + EV;
  double* s_rec = sm + L.rec; double* s_tail = sm + L.tail;
  // global scratch of the sweep's row / column sums (L2 resident): gam[NP], rho[NP], xi[EV][NP] -- one per CTA, or three
  // per cluster (rotating with the step, see uni_cluster_sync) when p.cluster CTAs share a candidate
  const int C = p.cluster;
  const int crank = C > 1 ? (int)uni_cluster_rank() : 0, cid = blockIdx.x / C;
  const bool lead = crank == 0;
  const size_t scr = (size_t)NP * (2 + EV);
  double* g_base = p.ws_uni + (C > 1 ? (size_t)cid * 3 : (size_t)blockIdx.x) * scr;
  double* g_gam = g_base;
  double* g_rho = g_gam + NP;
  double* g_xi = g_rho + NP;
  int gstep = 0;
  const int warp = tid >> 5, nwarps = NT >> 5;
  double* s_wp = sm + L.wp; double* s_wp2 = s_wp + (UNI_MAXT(EV) / 32) * L.wplen;   // per-warp rows of the B1 / B3 point sums
  double* s_m = sm + L.m; double* s_A = sm + L.A; double* s_Q = sm + L.Q;
  double* s_acc = sm + L.acc; int* s_int = reinterpret_cast<int*>(sm + L.ints);
  double* s2p = sm + L.small2;
  // small2 carve: mu_bar[E], s_bar[E2], Om[E2], Rinv[E2], Ub[E2], Vb[E2], hg[E + E2] (h_bar, g_bar), scal[16]
  double* s_mubar = s2p; double* s_sbar = s_mubar + GPMPC_MAX_EV; double* s_Om = s_sbar + EV * EV;
  double* s_Rinv = s_Om + EV * EV; double* s_U = s_Rinv + EV * EV; double* s_Vb = s_U + EV * EV;
  double* s_hbar = s_Vb + EV * EV; double* s_gbar = s_hbar + GPMPC_MAX_EV; double* s_scal = s_gbar + EV * EV;
  // scratch of the staged small algebra (B0 / B4, warp 0): the forward record of the step, then E x E temporaries
  double* s_rc = s_scal + 8 + EV * EV; double* s_Mb = s_rc + uni_rec_layout(EV).size; double* s_mb = s_Mb + GPMPC_MAX_EV;
  double* s_sp = s_mb + GPMPC_MAX_D; double* s_Ag = s_sp + EV * EV; double* s_t1 = s_Ag + EV * EV; double* s_X = s_t1 + EV * EV;
  double* s_RQ = s_sp; double* s_spb = s_Ag;    // B4 reuses the B0 temporaries
  const UniRecLayout RL = uni_rec_layout(E);
  const double* il2 = p.il2;
  const double s2 = p.s2[0];
  const double wmu = 1.0 / (double)(H + 1);
  exp2s_fill(p.exp2tab, tid, NT);
  // beta_a,j of the hot-loop records does not depend on the candidate or the step: written once per CTA
  for (int o = tid; o < NP * E; o += NT) s_rec[(o / E) * (2 * EV + 2) + EV + 1 + (o % E)] = __ldg(p.betaT + o);
  if (C == 1)
    for (int i = tid; i < NP * (2 + EV); i += NT) g_gam[i] = 0.0;   // B3 re-zeroes after every step (clusters: host memset)
  __syncthreads();
  // accumulator layout: [0] unused, [1 .. D] G_m, [1+D .. 1+D+E2) G_Q, then N-pass: Phi_m[D], Phi_A[P]
  const int accGm = 1, accGQ = 1 + D, accPm = 1 + D + EV * EV, accPA = accPm + D;

  long long clk_ = clock64();
  UNI_CTA_BEGIN();
  for (int cand_static = cid;; cand_static += gridDim.x / C) {
    const int cand = C > 1 ? cand_static : uni_next_candidate(p, 1, s_int, tid);
    if (cand >= p.B) break;
    const double* mus = p.states_mu + (size_t)cand * (H + 1) * E;
    const double* vars = p.states_var + (size_t)cand * (H + 1) * E * E;
    const double* rvs = p.rewards_var + (size_t)cand * (H + 1);
    const double* ams = p.actions_model + (size_t)cand * H * Na;
    double* gout = grad + (size_t)cand * H * Na;
    if (tid == 0) {  // terminal adjoint (setpoint_distance_reward_mapper.py:124-142)
      const double* mu = mus + (size_t)H * E;
      const double* s = vars + (size_t)H * E * E;
      const double wv = -p.kappa * wmu * 0.5 / sqrt(rvs[H]);
      double e[E], We[E], sWe[E], t1[E * E];
      for (int d = 0; d < E; d++) e[d] = mu[d] - p.c_target[d];
      for (int d = 0; d < E; d++) { double v = 0.0; for (int k = 0; k < E; k++) v += p.c_WT[d * E + k] * e[k]; We[d] = v; }
      for (int d = 0; d < E; d++) { double v = 0.0; for (int k = 0; k < E; k++) v += s[d * E + k] * We[k]; sWe[d] = v; }
      for (int d = 0; d < E; d++) {
        double v = 0.0;
        for (int k = 0; k < E; k++) v += p.c_WT[k * E + d] * sWe[k];
        s_mubar[d] = wmu * 2.0 * We[d] + wv * 8.0 * v;
      }
      for (int i = 0; i < E; i++)
        for (int k = 0; k < E; k++) { double v = 0.0; for (int l = 0; l < E; l++) v += p.c_WT[l * E + i] * s[k * E + l]; t1[i * E + k] = v; }
      for (int i = 0; i < E; i++)
        for (int k = 0; k < E; k++) {
          double v = 0.0;
          for (int l = 0; l < E; l++) v += t1[i * E + l] * p.c_WT[k * E + l];
          s_sbar[i * E + k] = wmu * p.c_WT[k * E + i] + wv * (4.0 * v + 4.0 * We[i] * We[k]);
        }
    }
    if (premat) {
      for (int tt = tid; tt < H; tt += NT) {   // step t = tt + 1 of the sweep uses the state at index tt
        double* pr = s_pre + (size_t)tt * L.prelen;
        double A[EV * EV], Q[EV * EV], Ri[EV * EV], c, detR, dmu[EV], da[GPMPC_MAX_D], dS[EV * EV];
        uni_step_matrices<EV>(vars + (size_t)tt * E * E, il2, s2, A, Q, Ri, c, detR);
        for (int e = 0; e < EV; e++) dmu[e] = 0.0;
        for (int k = 0; k < Na; k++) da[k] = 0.0;
        for (int e = 0; e < EV * EV; e++) dS[e] = 0.0;
        uni_stage_adjoint<EV>(p, Na, wmu, -p.kappa * wmu * 0.5 / sqrt(rvs[tt]), mus + (size_t)tt * E,
                              vars + (size_t)tt * E * E, ams + (size_t)tt * Na, dmu, da, dS);
        for (int e = 0; e < EV * EV; e++) { pr[oA + e] = A[e]; pr[oQ + e] = Q[e]; pr[oRi + e] = Ri[e]; pr[odS + e] = dS[e]; }
        pr[oc] = c; pr[odet] = detR;
        for (int e = 0; e < EV; e++) pr[odmu + e] = dmu[e];
        for (int k = 0; k < Na; k++) pr[oda + k] = da[k];
      }
    }
    __syncthreads();
    UNI_CLK(8);
    for (int t = H; t >= 1; t--) {
      const double* rec = p.records + ((size_t)cand * H + (t - 1)) * RL.size;
      const double* sp = vars + (size_t)(t - 1) * E * E;
      const double* mup = mus + (size_t)(t - 1) * E;
      const double* am = ams + (size_t)(t - 1) * Na;
      // ---- B0: model input, shared matrices, adjoint coefficients (thread 0: O(E^3))
      if (tid < D) s_m[tid] = (tid < E) ? mup[tid] : (tid < E + Na ? am[tid - E] : (double)(p.iter_ctrl + t - 1));
      for (int o = tid; o < L.accN + 1; o += NT) s_acc[o] = 0.0;
      // warp 0, one small stage per __syncwarp: every lane owns entries of the E x E matrices (a product is E FMAs deep per
      // lane instead of E^3 on one thread: these serial phases are latency, and they hold the CTA's other warps up)
      if (warp == 0) {
        if (lane == 0) s_int[0] = 0;
        if (premat) {
          const double* pr = s_pre + (size_t)(t - 1) * L.prelen;
          for (int o = lane; o < EV * EV; o += 32) { s_A[o] = pr[oA + o]; s_Q[o] = pr[oQ + o]; s_Rinv[o] = pr[oRi + o]; }
          if (lane == 0) { s_scal[0] = pr[oc]; s_scal[1] = pr[odet]; }
        } else if (lane == 0) {
          double Ai[EV * EV], Rinv[EV * EV], Qm[EV * EV], c0, detR0;
          uni_step_matrices<EV>(sp, il2, s2, Ai, Qm, Rinv, c0, detR0);
          for (int e = 0; e < EV * EV; e++) { s_A[e] = Ai[e]; s_Q[e] = Qm[e]; s_Rinv[e] = Rinv[e]; }
          s_scal[0] = c0; s_scal[1] = detR0;
        }
        for (int o = lane; o < E * E; o += 32) {
          s_sp[o] = sp[o];
          s_U[o] = s_sbar[o] + s_sbar[(o % E) * E + o / E];              // U = s_bar + s_bar^T
        }
        for (int o = lane; o < RL.size; o += 32) s_rc[o] = rec[o];       // the step's forward record (M, V, h, g, S_raw)
        __syncwarp();
        const double c = s_scal[0], detR = s_scal[1], rs = 1.0 / sqrt(detR);
        // V_bar[a][e] = sum_k sp[k][e] U[k][a] ;  M_bar = mu_bar - U M ;  pairs: Omega, wbar, detR_bar
        for (int o = lane; o < E * E; o += 32) {
          const int a = o / E, e = o - a * E;
          double v = 0.0;
#pragma unroll
          for (int k = 0; k < E; k++) v = fma(s_sp[k * E + e], s_U[k * E + a], v);
          s_Vb[o] = v;
        }
        if (lane < E) {
          double v = s_mubar[lane];
#pragma unroll
          for (int b = 0; b < E; b++) v -= s_U[lane * E + b] * s_rc[RL.offM + b];
          s_Mb[lane] = v;
        }
        {
          double dpart = 0.0, wpart = 0.0;
          for (int pr = lane; pr < P; pr += 32) {
            int a = 0, w = pr;
            while (w >= E - a) { w -= E - a; a++; }
            const int b = a + w;
            const double sb = s_sbar[a * E + b] + (a != b ? s_sbar[b * E + a] : 0.0);
            dpart += -0.5 * sb * s_rc[RL.offS + pr] * rs / detR;
            const double om = sb * rs * s2 * s2;          // d L / d Shat_ab
            if (a == b) { s_Om[a * E + a] = om; wpart += om; }
            else { s_Om[a * E + b] = 0.5 * om; s_Om[b * E + a] = 0.5 * om; }
          }
          dpart = warp_sum(dpart);
          wpart = warp_sum(wpart);
          if (lane == 0) { s_scal[2] = dpart; s_scal[3] = wpart; }
        }
        __syncwarp();
        // mean part adjoints: (A g_a), g_bar_a = c A^T V_bar_a, A_bar (direct part) = c sum_a V_bar_a g_a^T, h_bar_a = c M_bar_a
        for (int o = lane; o < E * E; o += 32) {
          const int a = o / E, e = o - a * E;
          double ag = 0.0, gb = 0.0, ab = 0.0;
#pragma unroll
          for (int f = 0; f < E; f++) {
            ag = fma(s_A[e * E + f], s_rc[RL.offG + a * E + f], ag);
            gb = fma(s_A[f * E + e], s_Vb[a * E + f], gb);
            ab = fma(s_Vb[f * E + a], s_rc[RL.offG + f * E + e], ab);     // (k, l) = (a, e): sum over the GPs f
          }
          s_Ag[o] = ag;
          s_gbar[o] = c * gb;
          s_scal[8 + o] = c * ab;                                          // needs 8 + E2 <= 72 doubles
        }
        if (lane < E) s_hbar[lane] = s_Mb[lane] * c;
        __syncwarp();
        {   // c_bar (summed) * c = c sum_a (M_bar_a h_a + V_bar_a . (A g_a))
          double part = 0.0;
          for (int o = lane; o < E * E; o += 32) part = fma(s_Vb[o], s_Ag[o], part);
          if (lane < E) part = fma(s_Mb[lane], s_rc[RL.offH + lane], part);
          part = warp_sum(part);
          if (lane == 0) s_scal[4] = part * c;
        }
      }
      __syncthreads();
      UNI_CLK(9);
      // ---- B1: nu, exponent terms, hot-loop record (as in the forward) + N pass of the mean part
      double vp[P], vs[GPMPC_MAX_D];
#pragma unroll
      for (int k = 0; k < P; k++) vp[k] = 0.0;
#pragma unroll
      for (int d = 0; d < GPMPC_MAX_D; d++) vs[d] = 0.0;
      for (int i = tid; i < NP; i += NT) {
        double nu[GPMPC_MAX_D];
#pragma unroll
        for (int d = 0; d < GPMPC_MAX_D; d++) nu[d] = (i < N && d < D) ? (uni_ldg_x(p.x + (size_t)i * D + d) - s_m[d]) : 0.0;
        double quad = 0.0, head = 0.0, tail = 0.0, zqz = 0.0, an[GPMPC_MAX_D];
#pragma unroll
        for (int e = 0; e < EV; e++) {
          double r = 0.0, q = 0.0;
#pragma unroll
          for (int f = 0; f < EV; f++) {
            r = fma(s_A[e * EV + f], nu[f], r);
            q = fma(s_Q[e * EV + f], nu[f] * il2[f], q);
          }
          an[e] = r;
          quad = fma(nu[e], r, quad);
          head = fma(nu[e] * nu[e], il2[e], head);
          zqz = fma(nu[e] * il2[e], q, zqz);
        }
#pragma unroll
        for (int d = EV; d < GPMPC_MAX_D; d++) {
          an[d] = 0.0;
          if (d < D) { tail = fma(nu[d] * nu[d], il2[d], tail); an[d] = nu[d] * il2[d]; }
        }
        const double ei = (i < N) ? exp2s((-0.5 * GPMPC_EXP2S_SCALE) * (quad + tail)) : 0.0;
        double* rcd = s_rec + i * (2 * EV + 2);
#pragma unroll
        for (int e = 0; e < EV; e++) rcd[e] = nu[e];
        rcd[EV] = (i < N) ? GPMPC_EXP2S_SCALE * (-0.5 * (head + tail) + zqz) : 0.0;
#pragma unroll
        for (int d = EV; d < GPMPC_MAX_D; d++)
          if (d < D) s_tail[i * L.tlen + d - EV] = nu[d];
        // mean part weight phi_i = sum_a e_i beta_a,i (h_bar_a + g_bar_a . nu_i^E), kept in the record's spare slot
        double phi = 0.0;
#pragma unroll
        for (int a = 0; a < E; a++) {
          double w = s_hbar[a];
#pragma unroll
          for (int e = 0; e < EV; e++) w = fma(s_gbar[a * E + e], nu[e], w);
          phi = fma(ei * rcd[EV + 1 + a], w, phi);
        }
        // raw moments of the mean part: sum_i phi_i nu_i,d (D) and sum_i phi_i nu_i,k nu_i,l (P)
#pragma unroll
        for (int d = 0; d < GPMPC_MAX_D; d++) vs[d] = fma(phi, nu[d], vs[d]);   // nu = 0 beyond D
        {
          int pr = 0;
#pragma unroll
          for (int k = 0; k < EV; k++) {
            const double pk = phi * nu[k];
#pragma unroll
            for (int l = k; l < EV; l++) { vp[pr] = fma(pk, nu[l], vp[pr]); pr++; }
          }
        }
      }
      UNI_CLK(14);
      // ---- B1b: the warp's totals of the raw moments -> s_wp row (added up over warps after the sweep)
      uni_warp_point_sums<EV>(vp, vs, D, lane, s_wp + warp * L.wplen);
      __syncthreads();
      UNI_CLK(10);
      // ---- B2: adjoint-weighted sweep over the upper tile triangle (static balanced split as in the forward, P3)
      {
        const double wbar = s_scal[3];
        constexpr int TR = (UNI_USE_MMA(EV) || UNI_USE_MMA8(EV)) ? 32 : 64;             // rows (and columns) of a tile
        const int CH = min(p.seg_bwd, TR), nrb = NP / TR, cpt = TR / CH;
        const int T = cpt * nrb * (nrb + 1) / 2, per = (T + nwarps * C - 1) / (nwarps * C);
        int c0 = (crank * nwarps + warp) * per;
        const int c1 = min(T, c0 + per);
        int I = 0, base = 0;
        if (C > 1) {   // this step's buffer of the cluster scratch
          g_gam = g_base + (size_t)(gstep % 3) * scr;
          g_rho = g_gam + NP;
          g_xi = g_rho + NP;
        }
        while (c0 < c1) {
          while (c0 >= base + cpt * (nrb - I)) { base += cpt * (nrb - I); I++; }
          const int ce = min(c1, base + cpt * (nrb - I));
          if (UNI_USE_MMA8(EV))
            uni_bwd_item_mma8<EV>(p, s_rec, s_Q, il2, s_Om, wbar, I, TR * I + CH * (c0 - base), TR * I + CH * (ce - base),
                                  lane, g_gam, g_rho, g_xi);
          else if (UNI_USE_MMA(EV))
            uni_bwd_item_mma<EV>(p, s_rec, s_Q, il2, s_Om, wbar, I, TR * I + CH * (c0 - base), TR * I + CH * (ce - base),
                                 lane, g_gam, g_rho, g_xi);
          else
            uni_bwd_item<EV>(p, s_rec, s_Q, il2, s_Om, wbar, I, TR * I + CH * (c0 - base), TR * I + CH * (ce - base),
                             lane, g_gam, g_rho, g_xi, sm + L.colred + warp * COLRED_WARP);
          c0 = ce;
        }
      }
      __threadfence();   // the sweep's row / column sums are reductions at L2: make them visible before B3 loads them
      __syncthreads();
      if (C > 1) {
        uni_cluster_sync();
        // the buffer of step gstep + 2 (last read at step gstep - 1) is cleared now, a slice per CTA: every CTA passes the
        // next barrier only after its slice is done, and the reductions into that buffer start after that barrier
        double* nz = g_base + (size_t)((gstep + 2) % 3) * scr;
        for (int i = crank * NT + tid; i < (int)scr; i += C * NT) nz[i] = 0.0;
      }
      UNI_CLK(11);
      for (int o = tid; o < D + P; o += NT) {   // B1b's per-warp rows -> raw moments (accPA = accPm + D: pairs follow)
        double v = 0.0;
        for (int w = 0; w < nwarps; w++) v += s_wp[w * L.wplen + o];
        s_acc[accPm + o] = v;
      }
      // ---- B3: per training point, g_i = gam_i + rho_i and x_i = xi_i * il2 from the scratch (L2 loads: the sums were
      //          formed by reductions at L2; scratch re-zeroed), folded straight into the thread's share of
      //          G_m (D singles: sum_i g_i nu_i,d) and the upper triangle of G_Q (P pairs), then reduced per warp
#pragma unroll
      for (int k = 0; k < P; k++) vp[k] = 0.0;
#pragma unroll
      for (int d = 0; d < GPMPC_MAX_D; d++) vs[d] = 0.0;
      for (int i = tid; i < NP; i += NT) {
        const double* rc = s_rec + i * (2 * EV + 2);
        double xv[EV], z[EV];
        const double gv = __ldcg(g_gam + i), rv = __ldcg(g_rho + i);   // all loads first (the stores below may alias)
#pragma unroll
        for (int e = 0; e < EV; e++) xv[e] = __ldcg(g_xi + (size_t)e * NP + i);
        if (C == 1) {
          g_gam[i] = 0.0;
          g_rho[i] = 0.0;
#pragma unroll
          for (int e = 0; e < EV; e++) g_xi[(size_t)e * NP + i] = 0.0;
        }
        const double g = gv + rv;
#pragma unroll
        for (int e = 0; e < EV; e++) {
          const double ne = rc[e];
          vs[e] = fma(g, ne, vs[e]);
          z[e] = ne * il2[e];
          xv[e] *= il2[e];
        }
#pragma unroll
        for (int d = EV; d < GPMPC_MAX_D; d++)
          if (d < D) vs[d] = fma(g, s_tail[i * L.tlen + d - EV], vs[d]);
        {
          int pr = 0;
#pragma unroll
          for (int k = 0; k < EV; k++) {
            const double gk = g * z[k];
#pragma unroll
            for (int l = k; l < EV; l++) {
              double acc = fma(gk, z[l], vp[pr]);
              acc = fma(z[l], xv[k], acc);
              vp[pr] = fma(z[k], xv[l], acc);
              pr++;
            }
          }
        }
      }
      UNI_CLK(15);
      uni_warp_point_sums<EV>(vp, vs, D, lane, s_wp2 + warp * L.wplen);
      __syncthreads();
      for (int o = tid; o < D + P; o += NT) {
        double v = 0.0;
        for (int w = 0; w < nwarps; w++) v += s_wp2[w * L.wplen + o];
        if (o < D) s_acc[accGm + o] = v * il2[o];
        else {
          int k = 0, w = o - D;
          while (w >= EV - k) { w -= EV - k; k++; }
          const int l = k + w;
          s_acc[accGQ + k * EV + l] = v;
          s_acc[accGQ + l * EV + k] = v;
        }
      }
      __syncthreads();
      UNI_CLK(12);
      // ---- B4: small algebra (warp 0, staged like B0): assemble m_bar, s_prev_bar, stage-cost adjoints
      if (warp == 0) {
        const double detR = s_scal[1], detR_bar = s_scal[2], cbar_c = s_scal[4];
        // pair part: x2 for the upper-triangle sweep, ybar correction (dS/dm[:EV] -= 2 W (Q ybar)), R^-T G_Q
        for (int d = lane; d < D; d += 32) {
          double g = 2.0 * s_acc[accGm + d];
          if (d < EV) {
            double v = 0.0;
#pragma unroll
            for (int f = 0; f < EV; f++) v = fma(s_Q[d * EV + f], 2.0 * s_acc[accGm + f], v);
            g -= 2.0 * (2.0 * il2[d]) * v;
          }
          // mean part: m_bar += Phi_m - sum_a h_a g_bar_a (state dims); Phi_m = A (sum_i phi_i nu_i) on the state block
          double v;
          if (d < EV) {
            v = 0.0;
#pragma unroll
            for (int f = 0; f < EV; f++) v = fma(s_A[d * EV + f], s_acc[accPm + f], v);
#pragma unroll
            for (int a = 0; a < E; a++) v -= s_rc[RL.offH + a] * s_gbar[a * E + d];
          } else {
            v = il2[d] * s_acc[accPm + d];
          }
          s_mb[d] = g + v;
        }
        for (int o = lane; o < E * E; o += 32) {
          const int i = o / E, k = o - i * E;
          double v = 0.0;
#pragma unroll
          for (int l = 0; l < E; l++) v = fma(s_Rinv[l * E + i], 2.0 * s_acc[accGQ + l * E + k], v);
          s_RQ[o] = v;
          // A_bar += -1/2 Phi_A (symmetric, stored as its upper triangle)
          const int kk = i < k ? i : k, ll = i < k ? k : i;
          s_scal[8 + o] += -0.5 * s_acc[accPA + kk * E - (kk * (kk - 1)) / 2 + (ll - kk)];
        }
        __syncwarp();
        for (int o = lane; o < E * E; o += 32) {
          const int i = o / E, k = o - i * E;
          const double Wk = 2.0 * il2[k];
          double v = 0.0, t1 = 0.0, x = 0.0;
#pragma unroll
          for (int l = 0; l < E; l++) {
            v = fma(s_RQ[i * E + l], s_Q[k * E + l], v);
            t1 = fma(s_A[l * E + i], s_scal[8 + l * E + k], t1);          // (A^T A_bar)[i][k]
            x = fma(s_U[i * E + l], s_rc[RL.offV + l * E + k], x);       // recurrence term U V
          }
          s_spb[o] = 0.5 * s_RQ[o] - v * Wk + detR_bar * detR * s_Rinv[k * E + i] * Wk;
          s_t1[o] = t1;
          s_X[o] = x;
        }
        __syncwarp();
        for (int o = lane; o < E * E; o += 32) {
          const int i = o / E, k = o - i * E;
          double v = 0.0;
#pragma unroll
          for (int l = 0; l < E; l++) v = fma(s_t1[i * E + l], s_A[k * E + l], v);
          s_spb[o] += -0.5 * cbar_c * s_A[k * E + i] - v;
        }
        __syncwarp();
        // symmetrised step adjoint + recurrence + stage-cost adjoint at t-1 -> the adjoints the next (earlier) step starts from
        constexpr int NR = (E * E + 31) / 32;
        double nsb[NR], nmu = 0.0, abar = 0.0;
        const double* pr = premat ? s_pre + (size_t)(t - 1) * L.prelen : nullptr;
#pragma unroll
        for (int r = 0; r < NR; r++) {
          const int o = lane + 32 * r;
          nsb[r] = 0.0;
          if (o < E * E) {
            const int i = o / E, k = o - i * E, ot = k * E + i;
            nsb[r] = 0.5 * (s_spb[o] + s_spb[ot]) + 0.5 * (s_sbar[o] + s_sbar[ot]) + 0.5 * (s_X[o] + s_X[ot]);
            if (premat) nsb[r] += pr[odS + o];
          }
        }
        if (lane < E) nmu = s_mubar[lane] + s_mb[lane] + (premat ? pr[odmu + lane] : 0.0);
        if (lane < Na) abar = s_mb[E + lane] + (premat ? pr[oda + lane] : 0.0);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < NR; r++) {
          const int o = lane + 32 * r;
          if (o < E * E) s_sbar[o] = nsb[r];
        }
        if (lane < E) s_mubar[lane] = nmu;
        if (premat) {
          if (lane < Na && lead) gout[(size_t)(t - 1) * Na + lane] = abar;
        } else {
          if (lane < Na) s_mb[lane] = abar;      // a_bar staged for the serial stage-cost adjoint
          __syncwarp();
          if (lane == 0) {
            double nmu1[EV], a_bar[GPMPC_MAX_D], nsb1[EV * EV];
            for (int e = 0; e < E; e++) nmu1[e] = s_mubar[e];
            for (int k = 0; k < Na; k++) a_bar[k] = s_mb[k];
            for (int e = 0; e < E * E; e++) nsb1[e] = s_sbar[e];
            uni_stage_adjoint<EV>(p, Na, wmu, -p.kappa * wmu * 0.5 / sqrt(rvs[t - 1]), mup, sp, am, nmu1, a_bar, nsb1);
            if (lead)
              for (int k = 0; k < Na; k++) gout[(size_t)(t - 1) * Na + k] = a_bar[k];
            for (int e = 0; e < E; e++) s_mubar[e] = nmu1[e];
            for (int e = 0; e < E * E; e++) s_sbar[e] = nsb1[e];
          }
        }
      }
      gstep++;
      __syncthreads();
      UNI_CLK(13);
    }
    if (p.limit_change && tid < Na && lead) {
      double cum = 0.0;
      for (int t = H - 1; t >= 0; t--) {
        cum += gout[(size_t)t * Na + tid];
        gout[(size_t)t * Na + tid] = cum * 2.0 * p.max_change[tid];
      }
    }
    __syncthreads();
  }
  UNI_CTA_END(1024);
}

template <int EV>
cudaError_t launch_uniform_inst(bool bwd, const RolloutParams& p, double* grad, int grid, int threads, size_t smem,
                                cudaStream_t st) {
  cudaError_t e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (p.cluster > 1) {   // p.cluster consecutive CTAs form a thread-block cluster (grid is a multiple of it)
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  if (bwd) {
    if (EV <= 5 && threads <= 128) {   // three-CTAs-per-SM build (see UNI_BWD_MINCTAS)
      constexpr int T = EV <= 5 ? 128 : UNI_MAXT(EV);   // (no extra instantiation for the larger state dims)
      e = cudaFuncSetAttribute(uniform_bwd_kernel<EV, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      e = cudaLaunchKernelEx(&cfg, uniform_bwd_kernel<EV, T>, p, grad);
    } else {
      e = cudaFuncSetAttribute(uniform_bwd_kernel<EV, UNI_MAXT(EV)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      e = cudaLaunchKernelEx(&cfg, uniform_bwd_kernel<EV, UNI_MAXT(EV)>, p, grad);
    }
  } else {
    if (EV <= 5 && threads <= 128) {   // three-CTAs-per-SM build (see UNI_FWD_MINCTAS)
      constexpr int T = EV <= 5 ? 128 : UNI_MAXT(EV);   // (no extra instantiation for the larger state dims)
      e = cudaFuncSetAttribute(uniform_fwd_kernel<EV, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      e = cudaLaunchKernelEx(&cfg, uniform_fwd_kernel<EV, T>, p);
    } else {
      e = cudaFuncSetAttribute(uniform_fwd_kernel<EV, UNI_MAXT(EV)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      e = cudaLaunchKernelEx(&cfg, uniform_fwd_kernel<EV, UNI_MAXT(EV)>, p);
    }
  }
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

// Thread-block clusters of `cluster` CTAs (threads, smem each) of the uniform kernels that the device can hold at once
// (cudaOccupancyMaxActiveClusters: GPC boundaries make this less than SMs x CTAs-per-SM / cluster).
template <int EV>
cudaError_t max_clusters_uniform_inst(bool bwd, int cluster, int threads, size_t smem, int* nclusters) {
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
  if (bwd) {
    e = cudaFuncSetAttribute(uniform_bwd_kernel<EV, UNI_MAXT(EV)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveClusters(nclusters, uniform_bwd_kernel<EV, UNI_MAXT(EV)>, &cfg);
  }
  e = cudaFuncSetAttribute(uniform_fwd_kernel<EV, UNI_MAXT(EV)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveClusters(nclusters, uniform_fwd_kernel<EV, UNI_MAXT(EV)>, &cfg);
}

}  // namespace gpmpc
