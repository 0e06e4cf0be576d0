// Batched projected L-BFGS update on the device: one kernel launch per optimiser iteration instead of
// ~100 small tensor operations.
//
// The reference optimises ONE action sequence at a time with scipy's L-BFGS-B inside a serial restart loop
// (control_objects/controllers/gp_mpc_controller.py:125-148, bounds [(0, 1)] * H * Na from
// actions_mappers/normalization_action_mapper.py:13).  Here B candidates advance in lock step: every iteration is ONE
// batched rollout (gpmpc_rollout: cost + gradient of all candidates) followed by ONE call of gpmpc_lbfgs_update, which
// for every candidate (one CTA each)
//   1. judges the trial point evaluated last (Armijo test on the projected step, directional term capped at 0), moves
//      there if it passes, and appends the curvature pair (s, y) to the candidate's own history (ring of `history` slots);
//   2. freezes the variables sitting on a bound with the gradient pushing outward, forms the quasi-Newton direction of the
//      free ones with the two-loop recursion, falls back to projected steepest descent when that is no descent direction
//      or the step keeps failing, and writes the next trial point clamp(x + alpha d, 0, 1).
// rl_gp_mpc/control_objects/controllers/batched_optim.py::minimize_box_lbfgs is the executable specification (plain
// torch, also the path for CPU tensors); tests/test_gpu_parity.py compares the two iterate by iterate.
#include <math.h>

#include "../../include/gpmpc.h"
#include "gpmpc_internal.h"

namespace gpmpc {

constexpr int LBFGS_THREADS = 128;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();                       // s_red may still be read by the previous reduction
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int k = 0; k < nw; k++) t += s_red[k];
  return t;
}
__device__ __forceinline__ double block_max(double v, double* s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[w] = v;
  __syncthreads();
  double t = s_red[0];
  for (int k = 1; k < nw; k++) t = fmax(t, s_red[k]);
  return t;
}
__device__ __forceinline__ double clean(double v) { return isfinite(v) ? v : 0.0; }   // non-finite gradient entries count as 0

struct LbfgsParams {
  int B, n, history, head, have_trial;
  double c1, shrink, max_first_move;
  double *x, *g, *f, *S, *Y, *rho, *alpha, *xt;
  int *fails, *first;
  const double *ft, *gt;
};

__global__ void __launch_bounds__(LBFGS_THREADS) lbfgs_update_kernel(const LbfgsParams p) {
  extern __shared__ double sm[];
  const int b = blockIdx.x, tid = threadIdx.x, n = p.n, h = p.history;
  double* s_q = sm;            // n: work vector of the two-loop recursion
  double* s_gf = sm + n;       // n: projected gradient
  double* s_a = sm + 2 * n;    // history
  double* s_red = s_a + h;     // 8
  double* x = p.x + (size_t)b * n;
  double* g = p.g + (size_t)b * n;
  double* xt = p.xt + (size_t)b * n;
  const size_t slot = (size_t)p.B * n;
  int head = p.head;                                    // slot the next pair goes to (same for all candidates)
  // ---------------------------------------------------------------- 1. judge the trial point of the previous call
  if (p.have_trial) {
    const double* gt = p.gt + (size_t)b * n;
    const double f0 = p.f[b], ft = p.ft[b];
    double dot = 0.0, mx = 0.0, sy = 0.0, ss = 0.0, yy = 0.0;
    for (int i = tid; i < n; i += blockDim.x) {
      const double xi = x[i], gi = g[i];
      const bool frozen = (xi <= 0.0 && gi > 0.0) || (xi >= 1.0 && gi < 0.0);
      const double gf = frozen ? 0.0 : gi;
      const double st = xt[i] - xi;
      const double y = clean(gt[i]) - gi;
      dot = fma(gf, st, dot);
      mx = fmax(mx, fabs(st));
      sy = fma(st, y, sy);
      ss = fma(st, st, ss);
      yy = fma(y, y, yy);
    }
    dot = block_sum(dot, s_red);
    mx = block_max(mx, s_red);
    sy = block_sum(sy, s_red);
    ss = block_sum(ss, s_red);
    yy = block_sum(yy, s_red);
    const bool ok = isfinite(ft) && (ft <= f0 + p.c1 * fmin(dot, 0.0)) && mx > 0.0;
    const bool keep = ok && (sy > 1e-10 * sqrt(ss) * sqrt(yy));
    double* Sh = p.S + (size_t)head * slot + (size_t)b * n;
    double* Yh = p.Y + (size_t)head * slot + (size_t)b * n;
    for (int i = tid; i < n; i += blockDim.x) {
      const double xi = x[i], gi = g[i], gti = clean(gt[i]);
      Sh[i] = keep ? xt[i] - xi : 0.0;          // candidates without a new pair keep their history aligned: empty slot
      Yh[i] = keep ? gti - gi : 0.0;
      if (ok) { x[i] = xt[i]; g[i] = gti; }
    }
    if (tid == 0) {
      p.rho[(size_t)head * p.B + b] = keep ? 1.0 / fmax(sy, 1e-300) : 0.0;
      if (ok) { p.f[b] = ft; p.first[b] = 0; p.fails[b] = 0; p.alpha[b] = 1.0; }
      else { p.fails[b] += 1; p.alpha[b] *= p.shrink; }
    }
    head = (head + 1) % h;
    __syncthreads();
    __threadfence_block();
  }
  // ---------------------------------------------------------------- 2. next direction and trial point
  for (int i = tid; i < n; i += blockDim.x) {
    const double xi = x[i], gi = g[i];
    const bool frozen = (xi <= 0.0 && gi > 0.0) || (xi >= 1.0 && gi < 0.0);
    const double gf = frozen ? 0.0 : gi;
    s_gf[i] = gf;
    s_q[i] = gf;
  }
  __syncthreads();
  // two-loop recursion, newest pair first; slots with rho = 0 drop out of both loops
  double gamma = 1.0;
  bool found = false;
  for (int k = 0; k < h; k++) {
    const int sl = ((head - 1 - k) % h + h) % h;
    const double rk = p.rho[(size_t)sl * p.B + b];
    const double* Sk = p.S + (size_t)sl * slot + (size_t)b * n;
    const double* Yk = p.Y + (size_t)sl * slot + (size_t)b * n;
    double d1 = 0.0, d2 = 0.0;
    for (int i = tid; i < n; i += blockDim.x) { d1 = fma(Sk[i], s_q[i], d1); d2 = fma(Yk[i], Yk[i], d2); }
    d1 = block_sum(d1, s_red);
    d2 = block_sum(d2, s_red);
    const double ak = rk * d1;
    if (tid == 0) s_a[k] = ak;
    for (int i = tid; i < n; i += blockDim.x) s_q[i] -= ak * Yk[i];
    if (rk > 0.0 && !found) { gamma = 1.0 / fmax(rk * d2, 1e-300); found = true; }   // s.y / y.y of the newest pair
    __syncthreads();
  }
  for (int i = tid; i < n; i += blockDim.x) s_q[i] *= gamma;
  __syncthreads();
  for (int k = h - 1; k >= 0; k--) {
    const int sl = ((head - 1 - k) % h + h) % h;
    const double rk = p.rho[(size_t)sl * p.B + b];
    const double* Sk = p.S + (size_t)sl * slot + (size_t)b * n;
    const double* Yk = p.Y + (size_t)sl * slot + (size_t)b * n;
    double d1 = 0.0;
    for (int i = tid; i < n; i += blockDim.x) d1 = fma(Yk[i], s_q[i], d1);
    d1 = block_sum(d1, s_red);
    const double c = s_a[k] - rk * d1;
    for (int i = tid; i < n; i += blockDim.x) s_q[i] += c * Sk[i];
    __syncthreads();
  }
  // d = -r on the free variables; not a descent direction, or the step keeps failing: projected steepest descent
  double slope = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    const double xi = x[i], gi = g[i];
    const bool frozen = (xi <= 0.0 && gi > 0.0) || (xi >= 1.0 && gi < 0.0);
    const double d = frozen ? 0.0 : -s_q[i];
    s_q[i] = d;
    slope = fma(d, s_gf[i], slope);
  }
  slope = block_sum(slope, s_red);
  const bool sd = !(slope < 0.0) || p.fails[b] >= 2 || !isfinite(slope);
  double dmax = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    const double d = sd ? -s_gf[i] : s_q[i];
    s_q[i] = d;
    dmax = fmax(dmax, fabs(d));
  }
  dmax = fmax(block_max(dmax, s_red), 1e-300);
  // moves without curvature information (first move, steepest descent): `max_first_move` in the largest component
  const double al = p.alpha[b];
  const double a_eff = (p.first[b] != 0 || sd) ? al * p.max_first_move / dmax : al;
  for (int i = tid; i < n; i += blockDim.x) xt[i] = fmin(fmax(fma(a_eff, s_q[i], x[i]), 0.0), 1.0);
}

}  // namespace gpmpc

extern "C" int gpmpc_lbfgs_update(int B, int n, int history, int head, int have_trial, double c1, double shrink,
                                  double max_first_move, double* x, double* g, double* f, double* S, double* Y,
                                  double* rho, double* alpha, int* fails, int* first, double* xt, const double* ft,
                                  const double* gt, void* stream) {
  using namespace gpmpc;
  if (B < 1 || n < 1 || history < 1 || history > 64 || head < 0 || head >= history) return GPMPC_ERR_BAD_ARG;
  if (!x || !g || !f || !S || !Y || !rho || !alpha || !fails || !first || !xt) return GPMPC_ERR_BAD_ARG;
  if (have_trial && (!ft || !gt)) return GPMPC_ERR_BAD_ARG;
  const size_t smem = sizeof(double) * (2 * (size_t)n + history + 8);
  if (smem > 200 * 1024) return GPMPC_ERR_UNSUPPORTED;
  LbfgsParams p;
  p.B = B; p.n = n; p.history = history; p.head = head; p.have_trial = have_trial;
  p.c1 = c1; p.shrink = shrink; p.max_first_move = max_first_move;
  p.x = x; p.g = g; p.f = f; p.S = S; p.Y = Y; p.rho = rho; p.alpha = alpha; p.xt = xt; p.fails = fails; p.first = first;
  p.ft = ft; p.gt = gt;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(lbfgs_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return GPMPC_ERR_CUDA;
  }
  lbfgs_update_kernel<<<B, LBFGS_THREADS, smem, st>>>(p);
  return cudaGetLastError() == cudaSuccess ? GPMPC_OK : GPMPC_ERR_CUDA;
}
