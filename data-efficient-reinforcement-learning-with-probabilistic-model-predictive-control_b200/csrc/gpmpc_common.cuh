// Shared device helpers for the GP-MPC kernels (sm_100a).  All arithmetic is float64: the
// reference computes in float64 (config_classes/total_config.py:11) and the covariance sums
// cancel by ~1e8 (hard part 1), so fp32 is not an option for this path.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define GPMPC_MAX_EV 8
#define GPMPC_MAX_D 16
#define GPMPC_MAX_PAIRS 36

#define HD __host__ __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// Hot-loop exponential, 7 float64 operations + 5 integer/LDS instructions.  The caller passes the exponent already
// in table units, t2 = x * 2048 / ln 2 (the scale is folded into the per-row / per-column terms when they are built,
// so it costs nothing per element):
//   n = rint(t2) (magic-number add), f = t2 - n in [-1/2, 1/2]  (exact: no Cody-Waite constants needed),
//   exp(x) = 2^(n >> 11) * T[n & 2047] * (1 + f (c1 + f (c2 + f c3))),   T[j] = 2^(j/2048) in shared memory (16 KB,
//   correctly rounded on the host).  c2 carries the minimax correction of the even quartic remainder
//   (tools/gen/exp2_coeffs.py: truncation error 5.9e-18; total error <= ~1.1 ulp from the table and final roundings).
//   The table entries are stored pre-biased (exp2s_entry), so 2^(n >> 11) costs one integer add on the high word.
// Deep underflow: t2 is clamped to >= -1022 * 2048 - 1 by ONE unsigned integer min on its high word (negative doubles
// order like unsigned integers), so n always fits and the result collapses to <= ~2e-308 (possibly a tiny denormal
// pattern, never NaN/negative) for x < -708.  A NaN exponent also collapses to ~0: NaN inputs are screened per step
// by the callers (s_int[1]), as before.  t2 > 0 is limited by the callers' bound x <= log(s2^2).
constexpr int EXP2S_LOG = 11;
constexpr int EXP2S_N = 1 << EXP2S_LOG;
#define GPMPC_EXP2S_SCALE 2.95463944374059701659e+03   /* 2048 / ln 2 */
#define GPMPC_EXP2S_C1 3.38450771757785784290e-04
#define GPMPC_EXP2S_C2 5.72744625649513507015e-08
#define GPMPC_EXP2S_C3 6.46152867293236580665e-12
#define GPMPC_EXP2S_HI_MIN 0xC13FF000u                 /* high word of -(1022 * 2048).0 */

// The table lives in STATIC shared memory, as the first member of the only static shared object of the kernels.  On
// sm_100a user shared memory starts GPMPC_SS_OFFSET = 1 KB into the CTA's window (the first KB is reserved by the system)
// and, in a thread-block cluster, the window of CTA rank r starts at r << 24; with both known a lookup is
//   SHL (n << 3), LOP3 ((. & mask) | rank bits), LDS.64 [R + 0x400]
// -- two integer instructions per exp instead of three with a run-time table base (the sweeps are issue-bound, every
// instruction counts: profiles/r02_micro_gen_loop.txt).  exp2s_fill checks the assumed offset at kernel start (trap).
#define GPMPC_SS_OFFSET 1024
#ifndef GPMPC_EXP2S_IMM_OFFSET
#define GPMPC_EXP2S_IMM_OFFSET 1
#endif
struct __align__(16) GpmpcStaticSmem {
  double tab[EXP2S_N];      // 2^(j/2048), pre-biased (gpmpc_api.cu)
  double one;               // the constant 1.0 (rollout_kernel: stride-0 factor of the moment sums)
  int next;                 // hand-over slot of the dynamic candidate queue
  int pad;
  unsigned char owner[GPMPC_MAX_PAIRS + 4];   // cluster rank that sweeps pair pr (rollout_kernel)
};
__shared__ GpmpcStaticSmem gpmpc_ss;
__device__ __forceinline__ void exp2s_fill(const double* __restrict__ tab_global, int tid, int nthreads) {   // + __syncthreads()
  if (tid == 0 && ((unsigned)__cvta_generic_to_shared(&gpmpc_ss) & 0xffffffu) != GPMPC_SS_OFFSET) __trap();   // see exp2s_entry
  for (int i = tid; i < EXP2S_N; i += nthreads) gpmpc_ss.tab[i] = tab_global[i];
}
__device__ __forceinline__ double exp2s_clamp(double t2) {
  const unsigned h = min((unsigned)__double2hiint(t2), GPMPC_EXP2S_HI_MIN);
  return __hiloint2double((int)h, __double2loint(t2));
}
// The table is stored PRE-BIASED: entry j holds the bit pattern of 2^(j/2048) with (j << 9) subtracted from its high
// word (gpmpc_api.cu).  Adding (n << 9) = ((n >> 11) << 20) + ((n & 2047) << 9) to the high word of entry n & 2047 then
// restores the mantissa AND applies 2^(n >> 11) through the exponent field in ONE integer instruction (no masks).
template <bool IMM = true>
__device__ __forceinline__ double exp2s_entry(int n) {   // 2^(n / 2048)
  double raw;
  if (IMM && GPMPC_EXP2S_IMM_OFFSET) {
    const unsigned rank_bits = (unsigned)__cvta_generic_to_shared(&gpmpc_ss) - GPMPC_SS_OFFSET;   // loop invariant (r << 24)
    asm("ld.shared.f64 %0, [%1 + 1024];" : "=d"(raw) : "r"(((n << 3) & (8 * EXP2S_N - 8)) | rank_bits));
  } else {   // run-time base, three integer instructions (measured 2 % faster in the uniform reverse sweep, 3 % slower elsewhere:
             // profiles/r02e_table_addressing.txt)
    asm("ld.shared.f64 %0, [%1];" : "=d"(raw) : "r"((unsigned)__cvta_generic_to_shared(&gpmpc_ss) + ((n << 3) & (8 * EXP2S_N - 8))));
  }
  return __hiloint2double(__double2hiint(raw) + (n << (20 - EXP2S_LOG)), __double2loint(raw));
}

__device__ __forceinline__ double exp2s(double t2) {
  const double SHIFT = 6755399441055744.0;
  t2 = exp2s_clamp(t2);
  double kd = t2 + SHIFT;
  const int n = __double2loint(kd);
  kd -= SHIFT;
  const double f = t2 - kd;
  const double t = exp2s_entry(n);
  double p = __fma_rn(GPMPC_EXP2S_C3, f, GPMPC_EXP2S_C2);
  p = __fma_rn(p, f, GPMPC_EXP2S_C1);
  p *= f;
  return __fma_rn(t, p, t);
}

// Four at once, stage by stage (4 independent float64 operations per stage).
template <bool IMM = true>
__device__ __forceinline__ void exp2s_x4(const double (&xin)[4], double (&res)[4]) {
  const double SHIFT = 6755399441055744.0;
  double x[4], kd[4], f[4], p[4], t[4];
  int n[4];
#pragma unroll
  for (int c = 0; c < 4; c++) x[c] = exp2s_clamp(xin[c]);
#pragma unroll
  for (int c = 0; c < 4; c++) kd[c] = x[c] + SHIFT;
#pragma unroll
  for (int c = 0; c < 4; c++) { n[c] = __double2loint(kd[c]); kd[c] -= SHIFT; }   // (an I2F.F64 instead of this add is slower)
#pragma unroll
  for (int c = 0; c < 4; c++) { f[c] = x[c] - kd[c]; t[c] = exp2s_entry<IMM>(n[c]); }
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] = __fma_rn(GPMPC_EXP2S_C3, f[c], GPMPC_EXP2S_C2);
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] = __fma_rn(p[c], f[c], GPMPC_EXP2S_C1);
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] *= f[c];
#pragma unroll
  for (int c = 0; c < 4; c++) res[c] = __fma_rn(t[c], p[c], t[c]);
}

// N at once (N independent chains), same stages as exp2s_x4.
template <int N, bool IMM = true>
__device__ __forceinline__ void exp2s_xn(const double (&xin)[N], double (&res)[N]) {
  const double SHIFT = 6755399441055744.0;
  double x[N], kd[N], f[N], p[N], t[N];
  int n[N];
#pragma unroll
  for (int c = 0; c < N; c++) x[c] = exp2s_clamp(xin[c]);
#pragma unroll
  for (int c = 0; c < N; c++) kd[c] = x[c] + SHIFT;
#pragma unroll
  for (int c = 0; c < N; c++) { n[c] = __double2loint(kd[c]); kd[c] -= SHIFT; }
#pragma unroll
  for (int c = 0; c < N; c++) { f[c] = x[c] - kd[c]; t[c] = exp2s_entry<IMM>(n[c]); }
#pragma unroll
  for (int c = 0; c < N; c++) p[c] = __fma_rn(GPMPC_EXP2S_C3, f[c], GPMPC_EXP2S_C2);
#pragma unroll
  for (int c = 0; c < N; c++) p[c] = __fma_rn(p[c], f[c], GPMPC_EXP2S_C1);
#pragma unroll
  for (int c = 0; c < N; c++) p[c] *= f[c];
#pragma unroll
  for (int c = 0; c < N; c++) res[c] = __fma_rn(t[c], p[c], t[c]);
}

// float64 tensor-core tile product D (8 x 8) = A (8 x 4) B (4 x 8) + D, one warp.  Fragments (g = lane >> 2, q = lane & 3):
//   a = A[g][q] ,  b = B[q][g] ,  d0, d1 = D[g][2 q], D[g][2 q + 1].
// DMMA runs on the float64 pipe (no extra flops: profiles/r01_micro_dmma_mix.txt) but takes ONE issue slot for 8 warp-DFMAs
// worth of work and does not pay the three-register-operand penalty of a DFMA (profiles/r02_micro_dfma_operands.txt).
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// exp2s_x4 with a SIGN per pair of results: res[0], res[1] = +-exp(xin[0]), +-exp(xin[1]) (sign a), res[2], res[3] (sign b).
// The sign costs nothing per element: the caller passes the magic constant of the range reduction as
//   sh = SHIFT + (negative ? 2^22 : 0)      (exp2s_shift: high word 0x43380000, low word 0 or 0x400000)
// so that n = lo(x + sh) carries bit 22; the table index (n & 2047) ignores it, and in the one integer add that applies
// 2^(n >> 11) to the table entry, (n << 9) moves it to bit 31 -- the sign bit of the result.  x itself is not touched.
__device__ __forceinline__ double exp2s_shift(int lo_word) { return __hiloint2double(0x43380000, lo_word); }
#define GPMPC_EXP2S_NEG_LO 0x00400000
__device__ __forceinline__ void exp2s_x4_signed(const double (&xin)[4], double (&res)[4], double sha, double shb) {
  double x[4], kd[4], f[4], p[4], t[4];
  int n[4];
#pragma unroll
  for (int c = 0; c < 4; c++) x[c] = exp2s_clamp(xin[c]);
#pragma unroll
  for (int c = 0; c < 4; c++) kd[c] = x[c] + (c < 2 ? sha : shb);
#pragma unroll
  for (int c = 0; c < 4; c++) { n[c] = __double2loint(kd[c]); kd[c] -= (c < 2 ? sha : shb); }
#pragma unroll
  for (int c = 0; c < 4; c++) { f[c] = x[c] - kd[c]; t[c] = exp2s_entry(n[c]); }
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] = __fma_rn(GPMPC_EXP2S_C3, f[c], GPMPC_EXP2S_C2);
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] = __fma_rn(p[c], f[c], GPMPC_EXP2S_C1);
#pragma unroll
  for (int c = 0; c < 4; c++) p[c] *= f[c];
#pragma unroll
  for (int c = 0; c < 4; c++) res[c] = __fma_rn(t[c], p[c], t[c]);
}

// iK loads of the sweeps (16 bytes: the lane's two adjacent rows of one column).  NO_L1: read-only load that does not
// allocate in L1 -- every value is used once per prediction by THIS CTA.  Measured (profiles/r02q_ik_l1_policy_and_unroll.txt):
// the reverse sweep at E <= 5 gains 2 % from it, the forward sweep and the E = 8 kernels LOSE 4-8 % (co-resident CTAs walk
// the same tile order, one candidate behind the other: L1 serves the second one), so only that kernel streams.
template <bool NO_L1>
__device__ __forceinline__ double2 ldg_stream2(const double* p) {
  if (NO_L1) {
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
  }
  return __ldg(reinterpret_cast<const double2*>(p));
}

// ---------------------------------------------------------------------------------------------
// Thread-block cluster helpers (small batches: several CTAs on neighbouring SMs share one candidate).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned uni_cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void uni_cluster_sync() {   // all threads of all CTAs of the cluster; orders global memory too
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Small dense SPD helpers (n <= 8), used once per step per GP / pair.
// ---------------------------------------------------------------------------------------------
// The same for a matrix in SHARED memory, by one whole warp, in place: n Gauss-Jordan sweeps (no pivoting: the pivots of an
// SPD matrix are its positive Schur complements, their product is the determinant), every lane owning the entries
// lane, lane + 32 -- O(n) dependent steps of ~150 clocks instead of the O(n^3) dependent operations of the serial
// routine (33 k clocks per step at n = 8, 8 k at n = 4, with the CTA's other warps waiting at the barrier).
// A non-positive pivot yields NaN, as in spd_inv_det.  Returns det(a) in every lane.
template <int n>
__device__ __forceinline__ double warp_spd_inv_det(double* a, int lane) {
  constexpr int NR = (n * n + 31) / 32;
  double det = 1.0;
#pragma unroll 1
  for (int k = 0; k < n; k++) {
    double pv = a[k * n + k];
    if (!(pv > 0.0)) pv = nan("");
    const double ip = 1.0 / pv;
    det *= pv;
    double nv[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const int o = lane + 32 * r;
      nv[r] = 0.0;
      if (o < n * n) {
        const int i = o / n, j = o - i * n;
        const double aik = a[i * n + k] * ip, akj = a[k * n + j];
        if (i == k) nv[r] = (j == k) ? ip : akj * ip;
        else nv[r] = (j == k) ? -aik : fma(-aik, akj, a[o]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const int o = lane + 32 * r;
      if (o < n * n) a[o] = nv[r];
    }
    __syncwarp();
  }
  return det;
}

// inv = a^-1, det = det(a) for symmetric positive definite a (n x n, row-major).  A non-positive
// pivot yields NaN, which then propagates like the reference's det/solve would.
template <int n>
HD void spd_inv_det(const double* a, double* inv, double& det) {
  double L[n * n];
  double Li[n * n];
  det = 1.0;
#pragma unroll
  for (int j = 0; j < n; j++) {
    double d = a[j * n + j];
  #pragma unroll
  for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
    det *= d;
    d = sqrt(d);
    L[j * n + j] = d;
    double id = 1.0 / d;
  #pragma unroll
  for (int i = j + 1; i < n; i++) {
      double v = a[i * n + j];
    #pragma unroll
  for (int k = 0; k < j; k++) v -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = v * id;
    }
  }
#pragma unroll
  for (int j = 0; j < n; j++) {
    Li[j * n + j] = 1.0 / L[j * n + j];
  #pragma unroll
  for (int i = j + 1; i < n; i++) {
      double v = 0.0;
    #pragma unroll
  for (int k = j; k < i; k++) v -= L[i * n + k] * Li[k * n + j];
      Li[i * n + j] = v / L[i * n + i];
    }
  }
#pragma unroll
  for (int i = 0; i < n; i++)
  #pragma unroll
  for (int j = 0; j <= i; j++) {
      double v = 0.0;
    #pragma unroll
  for (int k = i; k < n; k++) v += Li[k * n + i] * Li[k * n + j];
      inv[i * n + j] = v;
      inv[j * n + i] = v;
    }
}

// Pair matrices of gp_model.py:156-163 restricted to the EV x EV block that carries variance:
//   R = s W + I,  Rinv = R^-1,  Q = 1/2 R^-1 s,  detR = det R        (W = diag(Wd))
// computed through the symmetric form  s~ = W^1/2 s W^1/2,  R = W^-1/2 (s~ + I) W^1/2.
template <int n>
HD void pair_matrices(const double* s, const double* Wd, double* Rinv, double* Q, double& detR) {
  double st[n * n], Ti[n * n], sq[n];
#pragma unroll
  for (int e = 0; e < n; e++) sq[e] = sqrt(Wd[e]);
#pragma unroll
  for (int e = 0; e < n; e++)
  #pragma unroll
  for (int f = 0; f < n; f++) st[e * n + f] = sq[e] * s[e * n + f] * sq[f] + (e == f ? 1.0 : 0.0);
  spd_inv_det<n>(st, Ti, detR);
#pragma unroll
  for (int e = 0; e < n; e++)
  #pragma unroll
  for (int f = 0; f < n; f++) Rinv[e * n + f] = Ti[e * n + f] * sq[f] / sq[e];
#pragma unroll
  for (int e = 0; e < n; e++)
  #pragma unroll
  for (int f = 0; f < n; f++) {
      double v = 0.0;
    #pragma unroll
  for (int k = 0; k < n; k++) v += Rinv[e * n + k] * s[k * n + f];
      Q[e * n + f] = 0.5 * v;
    }
}

HD int pair_index(int a, int b, int E) {  // a <= b, row-major upper triangle
  return a * E - (a * (a - 1)) / 2 + (b - a);
}

// ---------------------------------------------------------------------------------------------
// Record written per (candidate, step) by the forward kernel in gradient mode and consumed by
// the reverse sweep (tests/algo_spec.py step_forward/step_backward is the executable spec).
// ---------------------------------------------------------------------------------------------
struct RecLayout {
  int E, D, P, offM, offV, offGp, gpStride, offPair, pairStride, size;
  // per-GP block: h, c, gE[E], dh_dm[D], dh_dA[E*E], dg_dm[E*D], dg_dA[E*P]
  // per-pair block: Sraw, detR, gm[D], gQ[E*E]
};
HD RecLayout rec_layout(int E, int D) {
  RecLayout r;
  r.E = E; r.D = D; r.P = E * (E + 1) / 2;
  r.offM = 0;
  r.offV = E;
  r.offGp = E + E * E;
  r.gpStride = 2 + E + D + E * E + E * D + E * r.P;
  r.offPair = r.offGp + E * r.gpStride;
  r.pairStride = 2 + D + E * E;
  r.size = r.offPair + r.P * r.pairStride;
  return r;
}

struct CostParams {
  const double* target;  // (E+Na)
  const double* W;       // (E+Na)^2
  const double* WT;      // E^2
  const double* smin;    // E
  const double* smax;    // E
  double kappa;
  int use_constraints;
  int clip;
};
