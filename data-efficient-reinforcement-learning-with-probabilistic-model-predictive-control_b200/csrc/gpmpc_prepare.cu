// prepare_inference / calculate_factorizations on the device (gp_model.py:182-191, :400-431):
//   K_a = s2_a exp(-1/2 |(x_i-x_j)/l_a|^2) + noise_a I   ->  Cholesky  ->  iK_a = K_a^-1,
//   beta_a = K_a^-1 y_a.        float64 throughout (cond(K) ~ 1e6 in the reference's regime).
//
// Layout: every N x N matrix lives in an NP x NP buffer (NP = N rounded up to 64, zero padded) so
// that the rollout kernel needs no bounds checks; iK is stored symmetrised.
//
// One CTA per GP.  At the reference's sizes (N <= ~1500, gp_memory.py points_batch_memory) the
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

// ------------------------------------------------------------------ blocked Cholesky, in place
// Lower factor written into the lower triangle of A (ld = NP); one CTA (1024 threads) per GP.
constexpr int CH_NB = 32;
constexpr int CH_T = 128;  // trailing-update macro tile

__global__ void __launch_bounds__(1024, 1)
cholesky_kernel(double* __restrict__ Aall, int N, int NP, int* __restrict__ info) {
  double* A = Aall + (size_t)blockIdx.x * NP * NP;
  __shared__ double Ld[CH_NB][CH_NB + 1];
  extern __shared__ double dyn[];  // two CH_T x CH_NB panels
  double* PI = dyn;
  double* PJ = dyn + CH_T * (CH_NB + 1);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  for (int k0 = 0; k0 < N; k0 += CH_NB) {
    const int kb = min(CH_NB, N - k0);
    // (a) diagonal block -> smem, factor with one warp (lane = row)
    for (int e = tid; e < CH_NB * CH_NB; e += blockDim.x) {
      int r = e / CH_NB, c = e % CH_NB;
      Ld[r][c] = (r < kb && c < kb) ? A[(size_t)(k0 + r) * NP + k0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      for (int c = 0; c < kb; c++) {
        double piv = Ld[c][c];
        if (!(piv > 0.0) && lane == 0) atomicExch(info + blockIdx.x, k0 + c + 1);
        double d = sqrt(piv);
        __syncwarp();
        if (lane == c) Ld[c][c] = d;
        if (lane > c && lane < kb) Ld[lane][c] /= d;
        __syncwarp();
        // rank-1 update of the remaining columns: lane = row r, loop over columns c2 in (c, r]
        if (lane > c && lane < kb) {
          double lrc = Ld[lane][c];
          for (int c2 = c + 1; c2 <= lane; c2++) Ld[lane][c2] -= lrc * Ld[c2][c];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int e = tid; e < kb * kb; e += blockDim.x) {
      int r = e / kb, c = e % kb;
      if (c <= r) A[(size_t)(k0 + r) * NP + k0 + c] = Ld[r][c];
    }
    // (b) panel: rows below the block, one thread per row: solve x Ld^T = a
    const int r0 = k0 + kb;
    for (int r = r0 + tid; r < N; r += blockDim.x) {
      double xr[CH_NB];
      double* row = A + (size_t)r * NP + k0;
#pragma unroll
      for (int c = 0; c < CH_NB; c++) xr[c] = (c < kb) ? row[c] : 0.0;
#pragma unroll
      for (int c = 0; c < CH_NB; c++) {
        if (c < kb) {
          double v = xr[c];
#pragma unroll
          for (int c2 = 0; c2 < c; c2++) v -= xr[c2] * Ld[c][c2];
          xr[c] = v / Ld[c][c];
        }
      }
#pragma unroll
      for (int c = 0; c < CH_NB; c++)
        if (c < kb) row[c] = xr[c];
    }
    __syncthreads();
    // (c) trailing update A[i][j] -= sum_c P[i][c] P[j][c], i >= j >= r0, in CH_T x CH_T macro tiles
    const int nrem = N - r0;
    if (nrem <= 0) break;
    const int nt = (nrem + CH_T - 1) / CH_T;
    for (int ti = 0; ti < nt; ti++) {
      for (int tj = 0; tj <= ti; tj++) {
        __syncthreads();
        for (int e = tid; e < CH_T * CH_NB; e += blockDim.x) {
          int r = e / CH_NB, c = e % CH_NB;
          int gi = r0 + ti * CH_T + r, gj = r0 + tj * CH_T + r;
          PI[r * (CH_NB + 1) + c] = (gi < N && c < kb) ? A[(size_t)gi * NP + k0 + c] : 0.0;
          PJ[r * (CH_NB + 1) + c] = (gj < N && c < kb) ? A[(size_t)gj * NP + k0 + c] : 0.0;
        }
        __syncthreads();
        // 1024 threads -> 32 x 32 grid of 4 x 4 register tiles
        const int ty = tid >> 5, tx = tid & 31;
        double acc[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
        for (int c = 0; c < CH_NB; c++) {
          double pi[4], pj[4];
#pragma unroll
          for (int u = 0; u < 4; u++) pi[u] = PI[(ty + 32 * u) * (CH_NB + 1) + c];
#pragma unroll
          for (int v = 0; v < 4; v++) pj[v] = PJ[(tx + 32 * v) * (CH_NB + 1) + c];
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int v = 0; v < 4; v++) acc[u][v] = fma(pi[u], pj[v], acc[u][v]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int v = 0; v < 4; v++) {
            int gi = r0 + ti * CH_T + ty + 32 * u, gj = r0 + tj * CH_T + tx + 32 * v;
            if (gi < N && gj <= gi) A[(size_t)gi * NP + gj] -= acc[u][v];
          }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ solves: iK = (L L^T)^-1, beta
// One thread per right-hand-side column c of [ I | y_a ] (N + 1 columns).  Z (ld = NC) holds the
// columns; forward substitution L Z = RHS then backward L^T X = Z, rows of L broadcast from smem.
__global__ void __launch_bounds__(1024, 1)
chol_solve_kernel(const double* __restrict__ Lall, const double* __restrict__ y, double* __restrict__ Zall,
                  int N, int NP, int E) {
  const int a = blockIdx.x;
  const double* L = Lall + (size_t)a * NP * NP;
  const int NC = NP + 64;
  double* Z = Zall + (size_t)a * NP * NC;
  extern __shared__ double rowbuf[];  // NP doubles
  const int tid = threadIdx.x;
  // init RHS
  for (int c = tid; c <= N; c += blockDim.x)
    for (int r = 0; r < N; r++) Z[(size_t)r * NC + c] = (c < N) ? (r == c ? 1.0 : 0.0) : y[r * E + a];
  __syncthreads();
  // forward: for r: z[r] = (rhs[r] - sum_{k<r} L[r][k] z[k]) / L[r][r]
  for (int r = 0; r < N; r++) {
    for (int k = tid; k <= r; k += blockDim.x) rowbuf[k] = L[(size_t)r * NP + k];
    __syncthreads();
    for (int c = tid; c <= N; c += blockDim.x) {
      int kstart = (c < N) ? c : 0;  // identity columns are zero above their diagonal
      if (kstart <= r) {
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
        int k = kstart;
        for (; k + 3 < r; k += 4) {
          v0 = fma(rowbuf[k], Z[(size_t)k * NC + c], v0);
          v1 = fma(rowbuf[k + 1], Z[(size_t)(k + 1) * NC + c], v1);
          v2 = fma(rowbuf[k + 2], Z[(size_t)(k + 2) * NC + c], v2);
          v3 = fma(rowbuf[k + 3], Z[(size_t)(k + 3) * NC + c], v3);
        }
        for (; k < r; k++) v0 = fma(rowbuf[k], Z[(size_t)k * NC + c], v0);
        double v = (v0 + v1) + (v2 + v3);
        Z[(size_t)r * NC + c] = (Z[(size_t)r * NC + c] - v) / rowbuf[r];
      }
    }
    __syncthreads();
  }
  // backward: L^T X = Z : x[r] = (z[r] - sum_{k>r} L[k][r] x[k]) / L[r][r]
  for (int r = N - 1; r >= 0; r--) {
    for (int k = r + tid; k < N; k += blockDim.x) rowbuf[k] = L[(size_t)k * NP + r];  // column r of L
    __syncthreads();
    for (int c = tid; c <= N; c += blockDim.x) {
      double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
      int k = r + 1;
      for (; k + 3 < N; k += 4) {
        v0 = fma(rowbuf[k], Z[(size_t)k * NC + c], v0);
        v1 = fma(rowbuf[k + 1], Z[(size_t)(k + 1) * NC + c], v1);
        v2 = fma(rowbuf[k + 2], Z[(size_t)(k + 2) * NC + c], v2);
        v3 = fma(rowbuf[k + 3], Z[(size_t)(k + 3) * NC + c], v3);
      }
      for (; k < N; k++) v0 = fma(rowbuf[k], Z[(size_t)k * NC + c], v0);
      double v = (v0 + v1) + (v2 + v3);
      Z[(size_t)r * NC + c] = (Z[(size_t)r * NC + c] - v) / rowbuf[r];
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

cudaError_t launch_prepare(const double* x, const double* y, const double* ls, const double* s2,
                           const double* noise, int N, int NP, int D, int E, double* Kbuf, double* Zbuf,
                           double* iK, double* beta, double* betaT, int* info, cudaStream_t st, long long* launches) {
  dim3 blk(32, 8);
  dim3 grd((NP + 31) / 32, (NP + 7) / 8, E);
  cudaMemsetAsync(info, 0, sizeof(int) * E, st);
  gram_kernel<<<grd, blk, 0, st>>>(x, ls, s2, noise, Kbuf, N, NP, D);
  size_t sm1 = 2 * CH_T * (CH_NB + 1) * sizeof(double);
  cudaFuncSetAttribute(cholesky_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
  cholesky_kernel<<<E, 1024, sm1, st>>>(Kbuf, N, NP, info);
  size_t sm2 = NP * sizeof(double);
  chol_solve_kernel<<<E, 1024, sm2, st>>>(Kbuf, y, Zbuf, N, NP, E);
  finalize_kernel<<<grd, blk, 0, st>>>(Zbuf, iK, beta, betaT, N, NP, E);
  *launches += 4;
  return cudaGetLastError();
}

}  // namespace gpmpc
