// Internal declarations shared by the translation units of libgpmpc.so.
#pragma once
#include <cuda_runtime.h>

namespace gpmpc {

struct RolloutParams {
  // ---- training block (resident in the handle; identical for every candidate -> L2 resident)
  const double* x;      // (N, D)
  const double* beta;   // (E, NP)   zero padded
  const double* betaT;  // (NP, E)   transposed copy (uniform-kernel path)
  const double* iK;     // (E, NP, NP) symmetric, zero padded
  const double* il2;    // (E, D)  1 / lengthscale^2
  const double* s2;     // (E)     outputscale
  const double* exp2tab; // (2048)  2^(j/2048), correctly rounded, pre-biased (general path, exp2s)
  const double* colcoef; // (E, NP, 2) general path: { 2048/ln 2 * log|beta_b,j| , exp2s magic constant carrying sign(beta_b,j) }
  int N, NP, D, DP, E, Na;
  // ---- candidates
  int B, H, mode;            // mode 0: rollout, 1: single moment-matching step on given inputs
  int include_time, iter_ctrl, per_cand_init, limit_change;
  const double* actions_mpc; // (B, H*Na)
  const double* obs_mu;      // (E) or (B,E)            [mode 1: input_mu (B, D)]
  const double* obs_var;     // (E,E) or (B,E,E)        [mode 1: input_var (B, EV, EV)]
  const double* max_change;  // (Na)
  const double* action_prev; // (Na)
  // ---- cost
  const double* c_target; const double* c_W; const double* c_WT; const double* c_smin; const double* c_smax;
  double kappa; int use_constraints, clip;
  // ---- outputs (never NULL inside the kernel: the host substitutes workspace buffers)
  double* cost;        // (B)
  double* states_mu;   // (B, H+1, E)
  double* states_var;  // (B, H+1, E, E)
  double* rewards;     // (B, H+1)
  double* rewards_var; // (B, H+1)
  double* actions_model; // (B, H, Na)
  double* stepM; double* stepS; double* stepV;   // mode 1 outputs (may be NULL)
  double* records;     // (B, H, rec.size) gradient-mode records (NULL in value mode)
  double* ws_kk;       // (gridDim.x, 2, E, NP) per-CTA scratch: kk, and lb when lb_global
  int lb_global;       // general kernel: lb[E][NP] of phases P1/P2 in the global scratch (shapes whose shared-memory plan is too tight)
  double* ws_gam;      // general kernel, gradient mode: (gridDim.x, group, NP) column sums of the sweeps (RED at L2)
  // ---- launch geometry
  int group;           // pairs per N^2 phase
  int seg;             // columns per work item
  int premat;          // uniform reverse sweep: per-step matrices / stage-cost adjoints precomputed for all steps of a candidate
                       // (1: kept in shared memory, 2: in the per-CTA global scratch ws_pre when the shared-memory plan is tight)
  double* ws_pre;      // (gridDim.x, H, prelen) -- premat == 2
  double* ws_uni;      // uniform reverse sweep: (grid, NP * (2 + EV)) row / column sums of the sweep, reduced at L2
  int seg_bwd;         // columns per work item of the uniform reverse sweep (triangular: finer for balance)
  int* queue;          // candidate counters of the dynamic scheduling (SMs differ in speed by up to ~25 % on this workload:
                       // L2 distance): [0] uniform forward, [1] uniform reverse sweep, [2] general kernel; NULL = round-robin
  int cluster;         // uniform kernels: CTAs (thread-block cluster) sharing one candidate, 1 = none (small batches only)
  double* ws_cl;       // uniform forward, cluster mode: (clusters, 3, 64) accumulators of the sweep sums (zeroed by the host)
  long long* dbg_clk;  // tuning aid (GPMPC_DEBUG_CLOCKS): CTA 0 accumulates clock64() deltas per phase here, else NULL
};

struct BackwardParams {
  const double* il2; const double* s2;
  int D, E, Na, B, H, limit_change, include_time;
  const double* max_change;
  const double* c_target; const double* c_W; const double* c_WT; const double* c_smin; const double* c_smax;
  double kappa; int use_constraints;
  const double* states_mu; const double* states_var; const double* rewards_var; const double* actions_model;
  const double* records;
  double* grad;        // (B, H*Na)
};

size_t rollout_smem_bytes(int EV, bool grad, int NP, int DP, int D, int E, int group, int H, int Na, int nwarps, bool lb_global);
int rollout_pick_group(int EV, bool grad, int NP, int DP, int D, int E, int H, int Na, int nwarps, size_t smem_limit, bool lb_global);
cudaError_t launch_rollout(int EV, bool grad, const RolloutParams& p, int grid, int threads, size_t smem, cudaStream_t st);
// Launch bounds of rollout_kernel<EV, .> = threads per SM of its launch plans: state dimensions <= 5 are built for 384
// threads (<= 168 registers: two CTAs of 192 threads, or one of 384), larger ones for 256 threads (full register file).
#ifndef GEN_MAXT
#define GEN_MAXT(EV) ((EV) <= 5 ? 384 : 256)
#endif
inline int rollout_max_threads(int EV) { return GEN_MAXT(EV); }
cudaError_t launch_backward(int E, const BackwardParams& p, cudaStream_t st);
size_t uniform_smem_bytes(int EV, bool bwd, int NP, int DP, int D, int H, int Na, bool premat);
int uniform_premat_len(int EV, int NP, int DP, int D, int H, int Na);
cudaError_t rollout_max_clusters(int EV, bool grad, int cluster, int threads, size_t smem, int* nclusters);
cudaError_t launch_uniform(int EV, bool bwd, const RolloutParams& p, double* grad, int grid, int threads, size_t smem, cudaStream_t st);
cudaError_t uniform_max_clusters(int EV, bool bwd, int cluster, int threads, size_t smem, int* nclusters);
cudaError_t launch_prepare(const double* x, const double* y, const double* ls, const double* s2,
                           const double* noise, int N, int NP, int D, int E, double* Kbuf, double* Zbuf,
                           double* iK, double* beta, double* betaT, int* info, cudaStream_t st, long long* launches);
cudaError_t launch_append(const double* x, const double* xnew, const double* ynew, const double* ls, const double* s2,
                          const double* noise, double* Lbuf, double* ws, double* iK, double* beta, double* betaT,
                          int* info, int N, int NP, int D, int E, cudaStream_t st, long long* launches);
cudaError_t launch_mll(const double* x, const double* y, const double* ls, const double* s2, const double* Lbuf,
                       const double* iK, const double* beta, double* out, int N, int NP, int D, int E, int stride,
                       cudaStream_t st, long long* launches);
cudaError_t launch_il2(const double* ls, double* il2, int n, cudaStream_t st);
cudaError_t launch_colcoef(const double* beta, double* colcoef, int n, cudaStream_t st);

// static shared memory of the rollout kernels: the 16 KB table of exp2s (gpmpc_common.cuh) + a few scalars
constexpr size_t GPMPC_STATIC_SMEM = 16 * 1024 + 128;   // sizeof(GpmpcStaticSmem) rounded up
// Large state dimensions (E >= 6): float64 tensor-core sweeps (uni_*_mma8 in gpmpc_uniform_impl.cuh), one CTA of 256
// threads per SM (UNI_MMA8_THREADS; 384 threads at <= 168 registers measured no faster).
#ifndef UNI_MMA8
#define UNI_MMA8 1
#endif
#define UNI_USE_MMA8(EV) (UNI_MMA8 && (EV) >= 6)
#ifndef UNI_MMA8_THREADS
#define UNI_MMA8_THREADS 256   // (384 threads at <= 168 registers: same sweep time, slower small-matrix phases)
#endif
#define UNI_MAXT(EV) (UNI_USE_MMA8(EV) ? UNI_MMA8_THREADS : 256)
constexpr int UNIFORM_MAX_THREADS = 256;   // uniform kernels: __launch_bounds__(256, 2) -> <= 128 registers

}  // namespace gpmpc
