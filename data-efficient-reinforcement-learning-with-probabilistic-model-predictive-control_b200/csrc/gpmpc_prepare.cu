// prepare_inference / calculate_factorizations on the device (gp_model.py:182-191, :400-431):
//   K_a = s2_a exp(-1/2 |(x_i-x_j)/l_a|^2) + noise_a I   ->  Cholesky  ->  iK_a = K_a^-1,
//   beta_a = K_a^-1 y_a.        float64 throughout (cond(K) ~ 1e6 in the reference's regime).
//
// Layout: every N x N matrix lives in an NP x NP buffer (NP = N rounded up to 64, zero padded) so
// that the rollout kernel needs no bounds checks; iK is stored symmetrised.
//
// Multi-CTA blocked algorithms (all E matrices in the same launches); float64 throughout.  At the reference
// factorisation is ~1 % of a control step, so the design goal here is
// float64 correctness; the N^2 rollout loop is where the time goes.
#include "gpmpc_common.cuh"
#include "gpmpc_internal.h"

namespace gpmpc {

// ------------------------------------------------------------------ Gram matrix (+ noise on diag)
__global__ void gram_kernel(const double* __restrict__ x, const double* __restrict__ ls,
                            const double* __restrict__ s2, const double* __restrict__ noise,
                            double* __restrict__ K, int N, int NP, int D) {
  const int a = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= NP || j >= NP) return;
  double v = 0.0;
  if (i < N && j < N) {
    double d2 = 0.0;
    for (int d = 0; d < D; d++) {
      double t = (x[i * D + d] - x[j * D + d]) / ls[a * D + d];
      d2 = fma(t, t, d2);
    }
    v = s2[a] * exp(-0.5 * d2);          // exact s2 on the diagonal (d2 == 0)
    if (i == j) v += noise[a];
  }
  K[((size_t)a * NP + i) * NP + j] = v;
}

// ------------------------------------------------------------------ blocked right-looking Cholesky, in place
// Lower factor written into the lower triangle of A (ld = NP).  Per 32-column panel: (1) factor the diagonal block
// (one warp per GP), (2) panel solve, one thread per row below, (3) trailing symmetric update in 64 x 64 tiles spread
// over the whole GPU (this is ~all of the N^3/3 flops).  All E matrices are processed by the same launches.
constexpr int CH_NB = 32;

__global__ void __launch_bounds__(32) chol_diag_kernel(double* __restrict__ Aall, int k0, int N, int NP,
                                                       int* __restrict__ info) {
  double* A = Aall + (size_t)blockIdx.x * NP * NP;
  __shared__ double Ld[CH_NB][CH_NB + 1];
  const int lane = threadIdx.x;
  const int kb = min(CH_NB, N - k0);
  for (int c = 0; c < CH_NB; c++)
    Ld[lane][c] = (lane < kb && c < kb) ? A[(size_t)(k0 + lane) * NP + k0 + c] : (lane == c ? 1.0 : 0.0);
  __syncwarp();
  for (int c = 0; c < kb; c++) {
    const double piv = Ld[c][c];
    if (!(piv > 0.0) && lane == 0) atomicExch(info + blockIdx.x, k0 + c + 1);
    const double d = sqrt(piv);
    __syncwarp();
    if (lane == c) Ld[c][c] = d;
    if (lane > c && lane < kb) Ld[lane][c] /= d;
    __syncwarp();
    if (lane > c && lane < kb) {
      const double lrc = Ld[lane][c];
      for (int c2 = c + 1; c2 <= lane; c2++) Ld[lane][c2] -= lrc * Ld[c2][c];
    }
    __syncwarp();
  }
  if (lane < kb)
    for (int c = 0; c <= lane; c++) A[(size_t)(k0 + lane) * NP + k0 + c] = Ld[lane][c];
}

__global__ void __launch_bounds__(128) chol_trsm_kernel(double* __restrict__ Aall, int k0, int N, int NP) {
  double* A = Aall + (size_t)blockIdx.y * NP * NP;
  __shared__ double Ld[CH_NB][CH_NB + 1];
  const int kb = min(CH_NB, N - k0);
  for (int e = threadIdx.x; e < CH_NB * CH_NB; e += blockDim.x) {
    const int r = e / CH_NB, c = e % CH_NB;
    Ld[r][c] = (r < kb && c <= r) ? A[(size_t)(k0 + r) * NP + k0 + c] : (r == c ? 1.0 : 0.0);
  }
  __syncthreads();
  const int r = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  double xr[CH_NB];
  double* row = A + (size_t)r * NP + k0;
#pragma unroll
  for (int c = 0; c < CH_NB; c++) xr[c] = (c < kb) ? row[c] : 0.0;
#pragma unroll
  for (int c = 0; c < CH_NB; c++) {
    double v = xr[c];
#pragma unroll
    for (int c2 = 0; c2 < c; c2++) v -= xr[c2] * Ld[c][c2];
    xr[c] = v / Ld[c][c];
  }
#pragma unroll
  for (int c = 0; c < CH_NB; c++)
    if (c < kb) row[c] = xr[c];
}

// A[i][j] -= sum_c P[i][c] P[j][c] for i >= j >= r0, P = panel columns [k0, k0+kb); one 64 x 64 tile per CTA
__global__ void __launch_bounds__(256) chol_syrk_kernel(double* __restrict__ Aall, int k0, int N, int NP) {
  double* A = Aall + (size_t)blockIdx.y * NP * NP;
  __shared__ double PI[64][CH_NB + 1], PJ[64][CH_NB + 1];
  const int kb = min(CH_NB, N - k0), r0 = k0 + kb;
  int ti = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((ti + 1) * (ti + 2) / 2 <= (int)blockIdx.x) ti++;
  while (ti * (ti + 1) / 2 > (int)blockIdx.x) ti--;
  const int tj = blockIdx.x - ti * (ti + 1) / 2;
  const int tid = threadIdx.x;
  for (int e = tid; e < 64 * CH_NB; e += blockDim.x) {
    const int r = e / CH_NB, c = e % CH_NB;
    const int gi = r0 + ti * 64 + r, gj = r0 + tj * 64 + r;
    PI[r][c] = (gi < N && c < kb) ? A[(size_t)gi * NP + k0 + c] : 0.0;
    PJ[r][c] = (gj < N && c < kb) ? A[(size_t)gj * NP + k0 + c] : 0.0;
  }
  __syncthreads();
  const int ty = tid >> 4, tx = tid & 15;
  double acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; u++)
#pragma unroll
    for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
  for (int c = 0; c < CH_NB; c++) {
    double pi[4], pj[4];
#pragma unroll
    for (int u = 0; u < 4; u++) pi[u] = PI[ty + 16 * u][c];
#pragma unroll
    for (int v = 0; v < 4; v++) pj[v] = PJ[tx + 16 * v][c];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int v = 0; v < 4; v++) acc[u][v] = fma(pi[u], pj[v], acc[u][v]);
  }
#pragma unroll
  for (int u = 0; u < 4; u++)
#pragma unroll
    for (int v = 0; v < 4; v++) {
      const int gi = r0 + ti * 64 + ty + 16 * u, gj = r0 + tj * 64 + tx + 16 * v;
      if (gi < N && gj <= gi) A[(size_t)gi * NP + gj] -= acc[u][v];
    }
}

// ------------------------------------------------------------------ iK = (L L^T)^-1 and beta = iK y
// One CTA per block of 32 right-hand-side columns of [ I | y_a ]: forward substitution L Y = RHS then backward
// L^T X = Y, block row by block row (32 x 32 tile products through shared memory + a 32-step triangular solve by one
// warp).  Identity columns are zero above their diagonal block, so the forward sweep starts there.  Z (ld = NC) holds
// the columns; X overwrites Y in place.
__global__ void __launch_bounds__(256) chol_inverse_kernel(const double* __restrict__ Lall, const double* __restrict__ y,
                                                           double* __restrict__ Zall, int N, int NP, int E) {
  const int a = blockIdx.y, cb = blockIdx.x;
  const double* L = Lall + (size_t)a * NP * NP;
  const int NC = NP + 64;
  double* Z = Zall + (size_t)a * NP * NC;
  __shared__ double Lt[32][33], Yt[32][33], Acc[32][33];
  const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;   // rows ty, ty+8, ty+16, ty+24 ; column tx
  const int c0 = 32 * cb, c = c0 + tx;
  const int nrb = (N + 31) / 32;
  const int rb0 = (c0 + 31 < N) ? cb : 0;                        // the block holding the y column starts at the top
  // ---- forward
  for (int rb = rb0; rb < nrb; rb++) {
    const int r0 = 32 * rb;
    double acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int r = r0 + ty + 8 * u;
      acc[u] = (r < N) ? ((c < N) ? (r == c ? 1.0 : 0.0) : (c == N ? y[(size_t)r * E + a] : 0.0)) : 0.0;
    }
    for (int kb = rb0; kb < rb; kb++) {
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int rr = ty + 8 * u;
        Lt[rr][tx] = (r0 + rr < N) ? L[(size_t)(r0 + rr) * NP + 32 * kb + tx] : 0.0;
        Yt[rr][tx] = Z[(size_t)(32 * kb + rr) * NC + c];
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < 32; k++) {
        const double yv = Yt[k][tx];
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] = fma(-Lt[ty + 8 * u][k], yv, acc[u]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int rr = ty + 8 * u;
      Acc[rr][tx] = acc[u];
      Lt[rr][tx] = (r0 + rr < N && tx <= rr) ? L[(size_t)(r0 + rr) * NP + r0 + tx] : (rr == tx ? 1.0 : 0.0);
    }
    __syncthreads();
    if (ty == 0) {   // one warp: column tx, 32 sequential rows
      double yv[32];
#pragma unroll
      for (int i = 0; i < 32; i++) {
        double v = Acc[i][tx];
#pragma unroll
        for (int k = 0; k < i; k++) v = fma(-Lt[i][k], yv[k], v);
        yv[i] = v / Lt[i][i];
      }
#pragma unroll
      for (int i = 0; i < 32; i++) Z[(size_t)(r0 + i) * NC + c] = (r0 + i < N) ? yv[i] : 0.0;
    }
    __syncthreads();
  }
  // rows above the first processed block of an identity column block are zero
  for (int rb = 0; rb < rb0; rb++)
#pragma unroll
    for (int u = 0; u < 4; u++) Z[(size_t)(32 * rb + ty + 8 * u) * NC + c] = 0.0;
  __syncthreads();
  // ---- backward: X[rb] = Lrr^-T (Y[rb] - sum_{kb > rb} L[kb, rb]^T X[kb])
  for (int rb = nrb - 1; rb >= 0; rb--) {
    const int r0 = 32 * rb;
    double acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u] = Z[(size_t)(r0 + ty + 8 * u) * NC + c];
    for (int kb = rb + 1; kb < nrb; kb++) {
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int rr = ty + 8 * u;    // Lt[k][i] = L[32 kb + k][r0 + i]
        Lt[rr][tx] = (32 * kb + rr < N) ? L[(size_t)(32 * kb + rr) * NP + r0 + tx] : 0.0;
        Yt[rr][tx] = Z[(size_t)(32 * kb + rr) * NC + c];
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < 32; k++) {
        const double xv = Yt[k][tx];
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] = fma(-Lt[k][ty + 8 * u], xv, acc[u]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int rr = ty + 8 * u;
      Acc[rr][tx] = acc[u];
      Lt[rr][tx] = (r0 + rr < N && tx <= rr) ? L[(size_t)(r0 + rr) * NP + r0 + tx] : (rr == tx ? 1.0 : 0.0);
    }
    __syncthreads();
    if (ty == 0) {
      double xv[32];
#pragma unroll
      for (int i = 31; i >= 0; i--) {
        double v = Acc[i][tx];
#pragma unroll
        for (int k = i + 1; k < 32; k++) v = fma(-Lt[k][i], xv[k], v);
        xv[i] = v / Lt[i][i];
      }
#pragma unroll
      for (int i = 0; i < 32; i++) Z[(size_t)(r0 + i) * NC + c] = (r0 + i < N) ? xv[i] : 0.0;
    }
    __syncthreads();
  }
}

// iK[a][i][j] = 1/2 (X[i][j] + X[j][i]) (zero padded to NP x NP), beta[a][i] = X[i][N]
__global__ void finalize_kernel(const double* __restrict__ Zall, double* __restrict__ iK,
                                double* __restrict__ beta, double* __restrict__ betaT, int N, int NP, int E) {
  const int a = blockIdx.z;
  const int NC = NP + 64;
  const double* Z = Zall + (size_t)a * NP * NC;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= NP || j >= NP) return;
  double v = 0.0;
  if (i < N && j < N) v = 0.5 * (Z[(size_t)i * NC + j] + Z[(size_t)j * NC + i]);
  iK[((size_t)a * NP + i) * NP + j] = v;
  if (j == 0) {
    const double b = (i < N) ? Z[(size_t)i * NC + N] : 0.0;
    beta[(size_t)a * NP + i] = b;
    betaT[(size_t)i * E + a] = b;
  }
}

// ------------------------------------------------------------------ exact-GP log marginal likelihood + gradient
// (reference: gpytorch ExactMarginalLogLikelihood inside GpStateTransitionModel.train, gp_model.py:193-306)
//   LML_a = -1/2 y^T alpha - sum_i log L_ii - N/2 log(2 pi),   alpha = K^-1 y,
//   dLML/dtheta = 1/2 sum_ij (alpha_i alpha_j - iK_ij) dK_ij/dtheta      for theta in {lengthscale_d, s2, noise}.
// out[a] = { LML, dLML/ds2, dLML/dnoise, dLML/dl_0 .. dLML/dl_{D-1} }  (MLL_STRIDE doubles per GP, zeroed first)
__global__ void mll_logdet_kernel(const double* __restrict__ Lall, const double* __restrict__ beta,
                                  const double* __restrict__ y, double* __restrict__ out, int N, int NP, int E, int stride) {
  const int a = blockIdx.x;
  const double* L = Lall + (size_t)a * NP * NP;
  double acc = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    acc += -0.5 * y[(size_t)i * E + a] * beta[(size_t)a * NP + i] - log(L[(size_t)i * NP + i]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out + (size_t)a * stride, acc);
  if (threadIdx.x == 0) atomicAdd(out + (size_t)a * stride, -0.5 * N * 1.8378770664093453);   // log(2 pi)
}

__global__ void __launch_bounds__(256) mll_grad_kernel(const double* __restrict__ x, const double* __restrict__ ls,
                                                       const double* __restrict__ s2, const double* __restrict__ iK,
                                                       const double* __restrict__ beta, double* __restrict__ out,
                                                       int N, int NP, int D, int stride) {
  const int a = blockIdx.z;
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int i0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
  double g[2 + GPMPC_MAX_D];
#pragma unroll
  for (int q = 0; q < 2 + GPMPC_MAX_D; q++) g[q] = 0.0;
  if (j < N) {
    const double bj = beta[(size_t)a * NP + j];
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u;
      if (i >= N) break;
      const double w = 0.5 * (beta[(size_t)a * NP + i] * bj - iK[((size_t)a * NP + i) * NP + j]);
      double d2 = 0.0, dd[GPMPC_MAX_D];
#pragma unroll
      for (int d = 0; d < GPMPC_MAX_D; d++) {
        dd[d] = 0.0;
        if (d < D) {
          const double l = ls[a * D + d], t = (x[(size_t)i * D + d] - x[(size_t)j * D + d]) / l;
          d2 = fma(t, t, d2);
          dd[d] = t * t / l;               // (x_i - x_j)^2 / l^3 * l^0 ... times k below
        }
      }
      const double k = exp(-0.5 * d2);     // dK/ds2
      g[0] = fma(w, k, g[0]);
      if (i == j) g[1] += w;               // dK/dnoise = I
      const double wk = w * s2[a] * k;
#pragma unroll
      for (int d = 0; d < GPMPC_MAX_D; d++)
        if (d < D) g[2 + d] = fma(wk, dd[d], g[2 + d]);
    }
  }
  for (int q = 0; q < 2 + D; q++) {
    double v = g[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + (size_t)a * stride + 1 + q, v);
  }
}

// ------------------------------------------------------------------ one-point append (O(N^2) per GP)
// The training set grows by one point per control step (gp_memory.py:31-64 -> gp_mpc_controller.py:114-118, where the
// reference refactorises from scratch).  With K' = [[K, k], [k^T, kap]]:
//   L l = k,  sig = kap - l.l  (exactly the row a full Cholesky of K' would append),  L^T v = l  (v = iK k),
//   iK' = [[iK + v v^T / sig, -v / sig], [-v^T / sig, 1 / sig]],   gam = (y_new - k.beta) / sig,   beta' = [beta - gam v ; gam].
// append_solve_kernel: one CTA per GP, blocked (32) forward / backward substitution; writes the new row of L, and v,
// sig, gam into the workspace.  append_update_kernel: the rank-one update of iK, beta, betaT over the whole GPU.
__global__ void __launch_bounds__(256) append_solve_kernel(const double* __restrict__ x, const double* __restrict__ xnew,
                                                           const double* __restrict__ ynew, const double* __restrict__ ls,
                                                           const double* __restrict__ s2, const double* __restrict__ noise,
                                                           double* __restrict__ Lall, const double* __restrict__ beta,
                                                           double* __restrict__ ws, int N, int NP, int D, int E,
                                                           int* __restrict__ info) {
  extern __shared__ double sh[];
  double* s_k = sh;              // k, then l   (NP)
  double* s_v = sh + NP;         // v           (NP)
  __shared__ double s_blk[32][33], s_part[8][33], s_red[8];
  const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* L = Lall + (size_t)a * NP * NP;
  // ---- k_i = s2 exp(-1/2 |(x_i - x_new) / l|^2)  and  k . beta
  double dot = 0.0;
  for (int i = tid; i < N; i += 256) {
    double d2 = 0.0;
    for (int d = 0; d < D; d++) {
      const double t = (x[(size_t)i * D + d] - xnew[d]) / ls[a * D + d];
      d2 = fma(t, t, d2);
    }
    const double k = s2[a] * exp(-0.5 * d2);
    s_k[i] = k;
    dot = fma(k, beta[(size_t)a * NP + i], dot);
  }
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if (lane == 0) s_red[warp] = dot;
  __syncthreads();
  double kbeta = 0.0;
  for (int w = 0; w < 8; w++) kbeta += s_red[w];
  const int nb = (N + 31) / 32;
  // ---- forward substitution L l = k
  for (int b = 0; b < nb; b++) {
    const int r0 = 32 * b, rows = min(32, N - r0);
    for (int rr = warp; rr < rows; rr += 8) {
      const double* Lr = L + (size_t)(r0 + rr) * NP;
      double acc = 0.0;
      for (int c = lane; c < r0; c += 32) acc = fma(Lr[c], s_k[c], acc);
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) s_k[r0 + rr] -= acc;
    }
    for (int rr = warp; rr < 32; rr += 8)
      s_blk[rr][lane] = (rr < rows && lane < rows) ? L[(size_t)(r0 + rr) * NP + r0 + lane] : (rr == lane ? 1.0 : 0.0);
    __syncthreads();
    if (warp == 0) {
      double kj = (lane < rows) ? s_k[r0 + lane] : 0.0;
      for (int c = 0; c < rows; c++) {
        const double lc = __shfl_sync(0xffffffffu, kj, c) / s_blk[c][c];
        if (lane == c) kj = lc;
        else if (lane > c) kj = fma(-s_blk[lane][c], lc, kj);
      }
      if (lane < rows) s_k[r0 + lane] = kj;
    }
    __syncthreads();
  }
  // ---- sig = kap - l.l
  double ll = 0.0;
  for (int i = tid; i < N; i += 256) ll = fma(s_k[i], s_k[i], ll);
  for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = ll;
  __syncthreads();
  ll = 0.0;
  for (int w = 0; w < 8; w++) ll += s_red[w];
  const double sig = (s2[a] + noise[a]) - ll;
  // ---- backward substitution L^T v = l
  for (int b = nb - 1; b >= 0; b--) {
    const int r0 = 32 * b, rows = min(32, N - r0);
    double acc = 0.0;
    if (lane < rows)
      for (int j = r0 + 32 + warp; j < N; j += 8) acc = fma(L[(size_t)j * NP + r0 + lane], s_v[j], acc);
    s_part[warp][lane] = acc;
    for (int rr = warp; rr < 32; rr += 8)
      s_blk[rr][lane] = (rr < rows && lane < rows) ? L[(size_t)(r0 + rr) * NP + r0 + lane] : (rr == lane ? 1.0 : 0.0);
    __syncthreads();
    if (warp == 0) {
      double vj = 0.0;
      if (lane < rows) {
        vj = s_k[r0 + lane];
        for (int w = 0; w < 8; w++) vj -= s_part[w][lane];
      }
      for (int c = rows - 1; c >= 0; c--) {
        const double vc = __shfl_sync(0xffffffffu, vj, c) / s_blk[c][c];
        if (lane == c) vj = vc;
        else if (lane < c) vj = fma(-s_blk[c][lane], vc, vj);
      }
      if (lane < rows) s_v[r0 + lane] = vj;
    }
    __syncthreads();
  }
  // ---- results: new row of L, v, sig, gam
  if (!(sig > 0.0)) {
    if (tid == 0) atomicExch(info + a, N + 1);
    return;
  }
  double* Lrow = L + (size_t)N * NP;
  for (int i = tid; i < NP; i += 256) {
    Lrow[i] = (i < N) ? s_k[i] : (i == N ? sqrt(sig) : 0.0);
    ws[(size_t)a * NP + i] = (i < N) ? s_v[i] : 0.0;
  }
  if (tid == 0) {
    ws[(size_t)E * NP + 2 * a] = sig;
    ws[(size_t)E * NP + 2 * a + 1] = (ynew[a] - kbeta) / sig;
  }
}

__global__ void append_update_kernel(const double* __restrict__ ws, double* __restrict__ iK, double* __restrict__ beta,
                                     double* __restrict__ betaT, const int* __restrict__ info, int N, int NP, int E) {
  const int a = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  if (i > N || j > N) return;
  for (int q = 0; q < E; q++)
    if (info[q] != 0) return;          // a failed append leaves the whole factorisation untouched
  const double* v = ws + (size_t)a * NP;
  const double sig = ws[(size_t)E * NP + 2 * a], gam = ws[(size_t)E * NP + 2 * a + 1];
  double* out = iK + ((size_t)a * NP + i) * NP + j;
  if (i < N && j < N) *out = fma(v[i] / sig, v[j], *out);
  else if (i == N && j == N) *out = 1.0 / sig;
  else *out = -v[i < N ? i : j] / sig;
  if (j == 0) {
    const double b = (i < N) ? fma(-gam, v[i], beta[(size_t)a * NP + i]) : gam;
    beta[(size_t)a * NP + i] = b;
    betaT[(size_t)i * E + a] = b;
  }
}

cudaError_t launch_append(const double* x, const double* xnew, const double* ynew, const double* ls, const double* s2,
                          const double* noise, double* Lbuf, double* ws, double* iK, double* beta, double* betaT,
                          int* info, int N, int NP, int D, int E, cudaStream_t st, long long* launches) {
  cudaMemsetAsync(info, 0, sizeof(int) * E, st);
  append_solve_kernel<<<E, 256, sizeof(double) * 2 * NP, st>>>(x, xnew, ynew, ls, s2, noise, Lbuf, beta, ws, N, NP, D, E, info);
  dim3 blk(32, 8);
  append_update_kernel<<<dim3(N / 32 + 1, N / 8 + 1, E), blk, 0, st>>>(ws, iK, beta, betaT, info, N, NP, E);
  *launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_mll(const double* x, const double* y, const double* ls, const double* s2, const double* Lbuf,
                       const double* iK, const double* beta, double* out, int N, int NP, int D, int E, int stride,
                       cudaStream_t st, long long* launches) {
  cudaMemsetAsync(out, 0, sizeof(double) * E * stride, st);
  mll_logdet_kernel<<<E, 256, 0, st>>>(Lbuf, beta, y, out, N, NP, E, stride);
  mll_grad_kernel<<<dim3((N + 31) / 32, (N + 31) / 32, E), 256, 0, st>>>(x, ls, s2, iK, beta, out, N, NP, D, stride);
  *launches += 2;
  return cudaGetLastError();
}

cudaError_t launch_prepare(const double* x, const double* y, const double* ls, const double* s2,
                           const double* noise, int N, int NP, int D, int E, double* Kbuf, double* Zbuf,
                           double* iK, double* beta, double* betaT, int* info, cudaStream_t st, long long* launches) {
  dim3 blk(32, 8);
  dim3 grd((NP + 31) / 32, (NP + 7) / 8, E);
  cudaMemsetAsync(info, 0, sizeof(int) * E, st);
  gram_kernel<<<grd, blk, 0, st>>>(x, ls, s2, noise, Kbuf, N, NP, D);
  long long n = 1;
  for (int k0 = 0; k0 < N; k0 += CH_NB) {
    chol_diag_kernel<<<E, 32, 0, st>>>(Kbuf, k0, N, NP, info);
    n++;
    const int rem = N - k0 - CH_NB;
    if (rem > 0) {
      chol_trsm_kernel<<<dim3((rem + 127) / 128, E), 128, 0, st>>>(Kbuf, k0, N, NP);
      const int nt = (rem + 63) / 64;
      chol_syrk_kernel<<<dim3(nt * (nt + 1) / 2, E), 256, 0, st>>>(Kbuf, k0, N, NP);
      n += 2;
    }
  }
  chol_inverse_kernel<<<dim3(N / 32 + 1, E), 256, 0, st>>>(Kbuf, y, Zbuf, N, NP, E);
  finalize_kernel<<<grd, blk, 0, st>>>(Zbuf, iK, beta, betaT, N, NP, E);
  *launches += n + 2;
  return cudaGetLastError();
}

}  // namespace gpmpc
