// C ABI of libgpmpc.so (include/gpmpc.h): handle, workspace and kernel orchestration.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <string>

#include "../../include/gpmpc.h"
#include "gpmpc_common.cuh"
#include "gpmpc_internal.h"
#include "gpmpc_uniform_layout.cuh"

using namespace gpmpc;

struct DevBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&ptr, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <typename T> T* as() const { return static_cast<T*>(ptr); }
};

struct gpmpc_handle {
  int device = 0;
  int num_sms = 0;
  size_t smem_optin = 0;
  std::string err;
  bool prepared = false, cost_set = false;
  int N = 0, NP = 0, D = 0, DP = 0, E = 0, Na = 0;
  DevBuf x, il2, s2, ls, noise, beta, betaT, iK, Kbuf, Zbuf, info, exp2tab;
  bool uniform = false;      // all GPs share one hyper-parameter set -> uniform-kernel fast path
  int path_mode = 0;         // 0 auto, 1 force the general path
  DevBuf c_target, c_W, c_WT, c_smin, c_smax;
  double kappa = 0.0;
  int use_constraints = 0, clip = 0;
  DevBuf dbg_clk, ws_uni, queue, ws_cl, ws_pre;
  DevBuf ws_kk, ws_gam, colcoef, t_mu, t_var, t_r, t_rv, t_am, t_cost, records, step_in;
  long long launches = 0;
  bool timing = false;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  bool ev_fwd = false, ev_bwd = false;
  size_t cl_cap_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // cache of uniform_max_clusters (key: shared-memory plan, cluster size)
  int cl_cap[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cl_cap_next = 0;
  // gpmpc_fit_eval: the factorisation + marginal-likelihood kernels of one objective evaluation as a CUDA graph
  cudaStream_t fit_stream = nullptr;
  cudaGraphExec_t fit_exec = nullptr;
  double* fit_pin = nullptr;               // pinned host staging: theta as {ls (E,D), s2 (E), noise (E)}, then out (E,3+D), then info (E ints)
  DevBuf fit_out;
  const double *fit_x = nullptr, *fit_y = nullptr;
  int fit_N = 0, fit_D = 0, fit_E = 0;
  size_t fit_bufs = 0;                     // fingerprint of the device buffers the graph was captured with
  long long fit_launches = 0;
};

static int fail(gpmpc_handle* h, int code, const char* msg, cudaError_t ce = cudaSuccess) {
  if (h) {
    h->err = msg;
    if (ce != cudaSuccess) {
      h->err += ": ";
      h->err += cudaGetErrorString(ce);
    }
  }
  return code;
}

#define CU(call)                                                            \
  do {                                                                      \
    cudaError_t ce_ = (call);                                               \
    if (ce_ != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, #call, ce_);     \
  } while (0)

__global__ void __launch_bounds__(512) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
      v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
}

extern "C" {

int gpmpc_version(void) { return 200; }   // 100: round 1; 200: round 2 (gpmpc_lbfgs_update added, tensor-core sweeps, general path rebuilt)

int gpmpc_fp64_peak(int device, double* flops_per_s) {
  if (!flops_per_s) return GPMPC_ERR_BAD_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return GPMPC_ERR_NO_DEVICE; }
  cudaSetDevice(device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GPMPC_ERR_CUDA;
  const int grid = prop.multiProcessorCount * 4, blk = 512, iters = 4096;
  double* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * grid * blk) != cudaSuccess) return GPMPC_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<grid, blk>>>(buf, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  if (cudaGetLastError() != cudaSuccess) return GPMPC_ERR_CUDA;
  *flops_per_s = 2.0 * 64.0 * (double)iters * grid * blk / (best * 1e-3);
  return GPMPC_OK;
}

const char* gpmpc_last_error(const gpmpc_handle* h) { return h ? h->err.c_str() : "null handle"; }

int gpmpc_create(gpmpc_handle** out, int device) {
  if (!out) return GPMPC_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return GPMPC_ERR_NO_DEVICE;
  }
  gpmpc_handle* h = new gpmpc_handle();
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete h; return GPMPC_ERR_NO_DEVICE; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return GPMPC_ERR_CUDA; }
  h->num_sms = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin - GPMPC_STATIC_SMEM;   // budget of the DYNAMIC allocation (static: exp table)
  for (int i = 0; i < 4; i++) cudaEventCreate(&h->ev[i]);
  {  // 2^(j/2048), rounded once from extended precision, stored pre-biased for exp2s: (j << 9) is subtracted from the
     // high word so that the kernels restore the entry and apply 2^(n >> 11) with one integer add (exp2s_entry)
    static double tab[EXP2S_N];
    for (int j = 0; j < EXP2S_N; j++) {
      const double v = (double)exp2l((long double)j / (long double)EXP2S_N);
      unsigned long long bits;
      memcpy(&bits, &v, sizeof(bits));
      bits -= (unsigned long long)j << (32 + 20 - EXP2S_LOG);
      memcpy(&tab[j], &bits, sizeof(bits));
    }
    if (h->exp2tab.ensure(sizeof(tab)) != cudaSuccess ||
        cudaMemcpy(h->exp2tab.ptr, tab, sizeof(tab), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError();
      h->exp2tab.release();
      delete h;
      return GPMPC_ERR_CUDA;
    }
  }
  *out = h;
  return GPMPC_OK;
}

int gpmpc_destroy(gpmpc_handle* h) {
  if (!h) return GPMPC_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  DevBuf* all[] = {&h->x, &h->il2, &h->s2, &h->ls, &h->noise, &h->beta, &h->betaT, &h->iK, &h->Kbuf, &h->Zbuf, &h->info,
                   &h->c_target, &h->c_W, &h->c_WT, &h->c_smin, &h->c_smax, &h->ws_kk, &h->ws_gam, &h->colcoef, &h->t_mu, &h->t_var,
                   &h->t_r, &h->t_rv, &h->t_am, &h->t_cost, &h->records, &h->step_in, &h->exp2tab, &h->dbg_clk, &h->ws_uni, &h->queue, &h->ws_cl, &h->ws_pre};
  for (DevBuf* b : all) b->release();
  h->fit_out.release();
  if (h->fit_exec) cudaGraphExecDestroy(h->fit_exec);
  if (h->fit_stream) cudaStreamDestroy(h->fit_stream);
  if (h->fit_pin) cudaFreeHost(h->fit_pin);
  for (int i = 0; i < 4; i++)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  delete h;
  return GPMPC_OK;
}

static int prepare_buffers(gpmpc_handle* h, int NP, int D, int E) {   // device buffers of the training block
  CU(h->x.ensure(sizeof(double) * NP * D));   // room for gpmpc_append up to the padded size
  CU(h->ls.ensure(sizeof(double) * E * D));
  CU(h->il2.ensure(sizeof(double) * E * D));
  CU(h->s2.ensure(sizeof(double) * E));
  CU(h->noise.ensure(sizeof(double) * E));
  CU(h->beta.ensure(sizeof(double) * E * NP));
  CU(h->betaT.ensure(sizeof(double) * E * NP));
  CU(h->iK.ensure(sizeof(double) * (size_t)E * NP * NP));
  CU(h->Kbuf.ensure(sizeof(double) * (size_t)E * NP * NP));
  CU(h->Zbuf.ensure(sizeof(double) * (size_t)E * NP * (NP + 64)));
  CU(h->info.ensure(sizeof(int) * GPMPC_MAX_STATE));
  return GPMPC_OK;
}

int gpmpc_prepare(gpmpc_handle* h, const double* x, const double* y, const double* lengthscale,
                  const double* outputscale, const double* noise, int N, int D, int E, void* stream) {
  if (!h) return GPMPC_ERR_BAD_ARG;
  if (!x || !y || !lengthscale || !outputscale || !noise || N < 1) return fail(h, GPMPC_ERR_BAD_ARG, "prepare: null pointer or N < 1");
  if (E < 1 || E > GPMPC_MAX_STATE || D < E || D > GPMPC_MAX_INPUT)
    return fail(h, GPMPC_ERR_UNSUPPORTED, "prepare: need 1 <= E <= 8 and E <= D <= 16");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(h->device));
  const int NP = (N + 63) / 64 * 64;
  h->prepared = false;
  { const int rc = prepare_buffers(h, NP, D, E); if (rc != GPMPC_OK) return rc; }
  CU(cudaMemcpyAsync(h->x.ptr, x, sizeof(double) * N * D, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(h->ls.ptr, lengthscale, sizeof(double) * E * D, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(h->s2.ptr, outputscale, sizeof(double) * E, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(h->noise.ptr, noise, sizeof(double) * E, cudaMemcpyDeviceToDevice, st));
  CU(launch_il2(h->ls.as<double>(), h->il2.as<double>(), E * D, st));
  h->launches += 1;
  CU(launch_prepare(h->x.as<double>(), y, h->ls.as<double>(), h->s2.as<double>(), h->noise.as<double>(), N, NP, D, E,
                    h->Kbuf.as<double>(), h->Zbuf.as<double>(), h->iK.as<double>(), h->beta.as<double>(),
                    h->betaT.as<double>(), h->info.as<int>(), st, &h->launches));
  int info[GPMPC_MAX_STATE];
  CU(cudaMemcpyAsync(info, h->info.ptr, sizeof(int) * E, cudaMemcpyDeviceToHost, st));
  double hyp[GPMPC_MAX_STATE * (GPMPC_MAX_INPUT + 2)];
  CU(cudaMemcpyAsync(hyp, h->ls.ptr, sizeof(double) * E * D, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hyp + E * D, h->s2.ptr, sizeof(double) * E, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hyp + E * D + E, h->noise.ptr, sizeof(double) * E, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  bool uni = true;   // identical hyper-parameters for every GP (bitwise)
  for (int a = 1; a < E && uni; a++) {
    for (int d = 0; d < D; d++) uni = uni && (hyp[a * D + d] == hyp[d]);
    uni = uni && (hyp[E * D + a] == hyp[E * D]) && (hyp[E * D + E + a] == hyp[E * D + E]);
  }
  h->uniform = uni;
  for (int a = 0; a < E; a++)
    if (info[a] != 0) {
      char msg[160];
      snprintf(msg, sizeof(msg), "prepare: K + noise*I of GP %d is not positive definite (pivot %d)", a, info[a] - 1);
      return fail(h, GPMPC_ERR_NOT_PD, msg);
    }
  h->N = N; h->NP = NP; h->D = D; h->DP = (D + 1) & ~1; h->E = E;
  h->prepared = true;
  return GPMPC_OK;
}

int gpmpc_append_room(const gpmpc_handle* h) { return (h && h->prepared) ? h->NP - h->N : 0; }

int gpmpc_append(gpmpc_handle* h, const double* x_new, const double* y_new, void* stream) {
  if (!h) return GPMPC_ERR_BAD_ARG;
  if (!h->prepared) return fail(h, GPMPC_ERR_NOT_PREPARED, "append: call gpmpc_prepare first");
  if (!x_new || !y_new) return fail(h, GPMPC_ERR_BAD_ARG, "append: null pointer");
  if (h->N >= h->NP) return fail(h, GPMPC_ERR_UNSUPPORTED, "append: padded size exhausted (N is a multiple of 64), call gpmpc_prepare");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(h->device));
  const int N = h->N, NP = h->NP, D = h->D, E = h->E;
  CU(launch_append(h->x.as<double>(), x_new, y_new, h->ls.as<double>(), h->s2.as<double>(), h->noise.as<double>(),
                   h->Kbuf.as<double>(), h->Zbuf.as<double>(), h->iK.as<double>(), h->beta.as<double>(),
                   h->betaT.as<double>(), h->info.as<int>(), N, NP, D, E, st, &h->launches));
  CU(cudaMemcpyAsync(h->x.as<double>() + (size_t)N * D, x_new, sizeof(double) * D, cudaMemcpyDeviceToDevice, st));
  int info[GPMPC_MAX_STATE];
  CU(cudaMemcpyAsync(info, h->info.ptr, sizeof(int) * E, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int a = 0; a < E; a++)
    if (info[a] != 0) {
      char msg[160];
      snprintf(msg, sizeof(msg), "append: Schur complement of GP %d is not positive (factorisation left unchanged)", a);
      return fail(h, GPMPC_ERR_NOT_PD, msg);
    }
  h->N = N + 1;
  return GPMPC_OK;
}

int gpmpc_get_factorization(gpmpc_handle* h, double* iK, double* beta, void* stream) {
  if (!h) return GPMPC_ERR_BAD_ARG;
  if (!h->prepared) return fail(h, GPMPC_ERR_NOT_PREPARED, "get_factorization: call gpmpc_prepare first");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(h->device));
  const int N = h->N, NP = h->NP, E = h->E;
  if (iK)
    for (int a = 0; a < E; a++)
      CU(cudaMemcpy2DAsync(iK + (size_t)a * N * N, sizeof(double) * N, h->iK.as<double>() + (size_t)a * NP * NP,
                           sizeof(double) * NP, sizeof(double) * N, N, cudaMemcpyDeviceToDevice, st));
  if (beta)
    CU(cudaMemcpy2DAsync(beta, sizeof(double) * N, h->beta.ptr, sizeof(double) * NP, sizeof(double) * N, E,
                         cudaMemcpyDeviceToDevice, st));
  return GPMPC_OK;
}

int gpmpc_set_cost(gpmpc_handle* h, const double* target, const double* W, const double* WT, double kappa,
                   int use_constraints, const double* state_min, const double* state_max,
                   int clip_lower_bound_cost_to_0, int Na, void* stream) {
  if (!h) return GPMPC_ERR_BAD_ARG;
  if (!h->prepared) return fail(h, GPMPC_ERR_NOT_PREPARED, "set_cost: call gpmpc_prepare first (defines E, D)");
  if (!target || !W || !WT) return fail(h, GPMPC_ERR_BAD_ARG, "set_cost: null pointer");
  if (use_constraints && (!state_min || !state_max)) return fail(h, GPMPC_ERR_BAD_ARG, "set_cost: constraints need state_min/state_max");
  const int E = h->E;
  if (Na < 1 || E + Na > h->D) return fail(h, GPMPC_ERR_BAD_ARG, "set_cost: need 1 <= Na and E + Na <= D");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(h->device));
  const int Dc = E + Na;
  CU(h->c_target.ensure(sizeof(double) * Dc));
  CU(h->c_W.ensure(sizeof(double) * Dc * Dc));
  CU(h->c_WT.ensure(sizeof(double) * E * E));
  CU(h->c_smin.ensure(sizeof(double) * E));
  CU(h->c_smax.ensure(sizeof(double) * E));
  CU(cudaMemcpyAsync(h->c_target.ptr, target, sizeof(double) * Dc, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(h->c_W.ptr, W, sizeof(double) * Dc * Dc, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(h->c_WT.ptr, WT, sizeof(double) * E * E, cudaMemcpyDeviceToDevice, st));
  if (use_constraints) {
    CU(cudaMemcpyAsync(h->c_smin.ptr, state_min, sizeof(double) * E, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(h->c_smax.ptr, state_max, sizeof(double) * E, cudaMemcpyDeviceToDevice, st));
  }
  h->kappa = kappa;
  h->use_constraints = use_constraints ? 1 : 0;
  h->clip = clip_lower_bound_cost_to_0 ? 1 : 0;
  h->Na = Na;
  h->cost_set = true;
  return GPMPC_OK;
}

static int fill_common(gpmpc_handle* h, RolloutParams& p, int EV, bool grad, int B, int H, int Na, size_t* smem,
                       int* grid, int* threads, bool uniform = false) {
  p.x = h->x.as<double>(); p.beta = h->beta.as<double>(); p.iK = h->iK.as<double>();
  p.il2 = h->il2.as<double>(); p.s2 = h->s2.as<double>(); p.exp2tab = h->exp2tab.as<double>();
  p.N = h->N; p.NP = h->NP; p.D = h->D; p.DP = h->DP; p.E = h->E; p.Na = Na;
  p.B = B; p.H = H;
  p.betaT = h->betaT.as<double>();
  p.seg = 256;
  p.seg_bwd = 32;
  if (uniform) {
    p.group = 1;
    p.seg = 32;   // forward sweep: columns per chunk of the static tile-triangle split (divides 64)
    *smem = 0;
    *grid = 0;
    *threads = 0;
    return GPMPC_OK;
  }
  // Launch plan of the general kernel: G = output pairs per N^2 phase (their column terms live in shared memory).  Two
  // CTAs per SM (half of the kernel's thread budget each) when that costs no extra phase -- one CTA's serial small-matrix
  // phases then overlap the other's sweep -- else one CTA with all the threads.
  const size_t half_sm = (size_t)(227 * 1024) / 2 - 1024 - GPMPC_STATIC_SMEM;
  const int P = h->E * (h->E + 1) / 2;
  const int maxt = rollout_max_threads(EV);
  int G1 = rollout_pick_group(EV, grad, h->NP, h->DP, h->D, h->E, H, Na, maxt / 32, h->smem_optin, false);
  p.lb_global = 0;
  if (G1 < 2 && P > 1) {   // tight plan (e.g. E = 8, N = 1000 with the gradient): move lb[E][NP] to the global scratch
    const int Gg = rollout_pick_group(EV, grad, h->NP, h->DP, h->D, h->E, H, Na, maxt / 32, h->smem_optin, true);
    if (Gg > G1) { G1 = Gg; p.lb_global = 1; }
  }
  const int G2 = p.lb_global ? 0 : rollout_pick_group(EV, grad, h->NP, h->DP, h->D, h->E, H, Na, maxt / 64, half_sm < h->smem_optin ? half_sm : h->smem_optin, false);
  if (G1 < 1) return fail(h, GPMPC_ERR_UNSUPPORTED, "rollout: training set too large for the shared-memory plan (N, D)");
  // measured at the headline shape (profiles/r02a_general_launch_plans.txt): 1 x 384 threads 237.6 k predictions/s,
  // 2 x 192 219.2 k, 2 x 128 223.3 k; the 128-register build: 1 x 512 236.5 k, 2 x 256 233.1 k -> one CTA per SM
  int ctas = 1;
  if (const char* e = getenv("GPMPC_GEN_CTAS")) { int v = atoi(e); if (v == 1 || (v == 2 && G2 >= 1 && maxt / 2 >= 128)) ctas = v; }
  // Small batches of rollouts: a thread-block cluster shares each candidate, its output pairs (a, b) dealt to the CTAs
  // (rollout_kernel).  Worth it once the N^2 sweeps dominate (NP >= 128).
  p.cluster = 1;
  if (H > 0) {
    // the clusters of a launch must all be resident at once (their CTAs walk the steps in lock step): capacity from the
    // driver -- a cluster lives inside one GPC, so e.g. 16 clusters of 8 one-CTA-per-SM kernels do NOT fit 148 SMs
    // (B = 16 candidates took 10.2 ms with c = 8 in two waves, 5.9 ms with 8 candidates); cached per (c, plan)
    auto capacity = [&](int c) -> int {
      const size_t sm_c = rollout_smem_bytes(EV, grad, h->NP, h->DP, h->D, h->E, G1, H, Na, maxt / 32, p.lb_global != 0);
      const size_t key = (sm_c * 16 + c) * 2 + (grad ? 1 : 0) + ((size_t)1 << 60);
      for (int k = 0; k < 8; k++)
        if (h->cl_cap_key[k] == key) return h->cl_cap[k];
      int n = 0;
      if (rollout_max_clusters(EV, grad, c, maxt, sm_c, &n) != cudaSuccess) { cudaGetLastError(); n = h->num_sms / (2 * c); }
      h->cl_cap_key[h->cl_cap_next] = key; h->cl_cap[h->cl_cap_next] = n; h->cl_cap_next = (h->cl_cap_next + 1) % 8;
      return n;
    };
    int c = 8;
    while (c > 1 && (B * c > h->num_sms || P < c || h->NP < 128 || B > capacity(c))) c /= 2;
    if (const char* e = getenv("GPMPC_GEN_CLUSTER")) { int v = atoi(e); if (v == 1 || ((v == 2 || v == 4 || v == 8) && B * v <= h->num_sms && P >= v)) c = v; }
    p.cluster = c;
    if (c > 1) ctas = 1;
  }
  if (B <= h->num_sms) ctas = 1;       // nothing to overlap with: give the candidate all the threads
  const int G = ctas == 2 ? G2 : G1;
  p.group = G;
  *threads = maxt / ctas;
  if (const char* e = getenv("GPMPC_GEN_THREADS")) { int v = atoi(e); if (v >= 128 && v <= maxt / ctas && v % 32 == 0) *threads = v; }
  p.seg = 32;                          // columns per chunk of the static split of the sweeps (divides 64)
  {  // small training sets / clusters: finer chunks so that every warp gets a run
    const int nrb = h->NP / 64, nw = *threads / 32;
    while (p.seg > 8 && (64 / p.seg) * nrb * (nrb + 1) / 2 * ((P + p.cluster - 1) / p.cluster) < 2 * nw) p.seg /= 2;
  }
  if (const char* e = getenv("GPMPC_GEN_SEG")) { int v = atoi(e); if (v == 8 || v == 16 || v == 32 || v == 64) p.seg = v; }
  *smem = rollout_smem_bytes(EV, grad, h->NP, h->DP, h->D, h->E, G, H, Na, *threads / 32, p.lb_global != 0);
  *grid = p.cluster > 1 ? B * p.cluster : (B < h->num_sms * ctas ? B : h->num_sms * ctas);
  cudaError_t ce = h->ws_kk.ensure(sizeof(double) * 2 * (size_t)(*grid) * h->E * h->NP);
  if (ce != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, "workspace allocation", ce);
  p.ws_kk = h->ws_kk.as<double>();
  ce = h->colcoef.ensure(sizeof(double) * 2 * (size_t)h->E * h->NP);
  if (ce != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, "workspace allocation", ce);
  p.colcoef = h->colcoef.as<double>();
  if (grad) {
    ce = h->ws_gam.ensure(sizeof(double) * (size_t)(*grid) * G * h->NP);
    if (ce != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, "workspace allocation", ce);
    p.ws_gam = h->ws_gam.as<double>();
  }
  return GPMPC_OK;
}

int gpmpc_predict_step(gpmpc_handle* h, const double* input_mu, const double* input_var, int B, int EV,
                       double* M, double* S, double* V, void* stream) {
  if (!h) return GPMPC_ERR_BAD_ARG;
  if (!h->prepared) return fail(h, GPMPC_ERR_NOT_PREPARED, "predict_step: call gpmpc_prepare first");
  if (!input_mu || !input_var || B < 1) return fail(h, GPMPC_ERR_BAD_ARG, "predict_step: null pointer or B < 1");
  if (EV < 1 || EV > GPMPC_MAX_EV || EV > h->D) return fail(h, GPMPC_ERR_UNSUPPORTED, "predict_step: need 1 <= EV <= min(8, D)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(h->device));
  RolloutParams p;
  memset(&p, 0, sizeof(p));
  size_t smem; int grid, threads;
  int rc = fill_common(h, p, EV, false, B, 0, 1, &smem, &grid, &threads);
  if (rc) return rc;
  p.mode = 1;
  p.obs_mu = input_mu; p.obs_var = input_var;
  p.stepM = M; p.stepS = S; p.stepV = V;
  CU(launch_colcoef(h->beta.as<double>(), h->colcoef.as<double>(), h->E * h->NP, st));
  CU(launch_rollout(EV, false, p, grid, threads, smem, st));
  h->launches += 2;
  return GPMPC_OK;
}

int gpmpc_rollout(gpmpc_handle* h, const double* actions_mpc, const double* obs_mu, const double* obs_var,
                  int per_candidate_init, int B, int H, int Na, int iter_ctrl, int limit_action_change,
                  const double* max_change, const double* action_prev, double* cost, double* grad,
                  double* states_mu, double* states_var, double* rewards, double* rewards_var,
                  double* actions_model, void* stream) {
  if (!h) return GPMPC_ERR_BAD_ARG;
  if (!h->prepared) return fail(h, GPMPC_ERR_NOT_PREPARED, "rollout: call gpmpc_prepare first");
  if (!h->cost_set) return fail(h, GPMPC_ERR_NOT_PREPARED, "rollout: call gpmpc_set_cost first");
  if (!actions_mpc || !obs_mu || !obs_var) return fail(h, GPMPC_ERR_BAD_ARG, "rollout: null input pointer");
  if (B < 1 || H < 1 || Na != h->Na) return fail(h, GPMPC_ERR_BAD_ARG, "rollout: need B >= 1, H >= 1 and Na as given to set_cost");
  const int E = h->E, D = h->D;
  const int include_time = (D == E + Na + 1) ? 1 : 0;
  if (D != E + Na + include_time) return fail(h, GPMPC_ERR_BAD_ARG, "rollout: D must be E + Na (+1 with a time input)");
  if (limit_action_change && (!max_change || !action_prev)) return fail(h, GPMPC_ERR_BAD_ARG, "rollout: derivative action mapper needs max_change and action_prev");
  if ((size_t)H * Na > 4096) return fail(h, GPMPC_ERR_UNSUPPORTED, "rollout: H * Na > 4096");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(h->device));
  const bool want_grad = grad != nullptr;
  RolloutParams p;
  memset(&p, 0, sizeof(p));
  size_t smem; int grid, threads;
  const bool use_uniform = h->uniform && h->path_mode == 0;
  int rc = fill_common(h, p, E, want_grad, B, H, Na, &smem, &grid, &threads, use_uniform);
  if (rc) return rc;
  p.mode = 0;
  p.include_time = include_time; p.iter_ctrl = iter_ctrl; p.per_cand_init = per_candidate_init ? 1 : 0;
  p.limit_change = limit_action_change ? 1 : 0;
  p.actions_mpc = actions_mpc; p.obs_mu = obs_mu; p.obs_var = obs_var;
  p.max_change = max_change; p.action_prev = action_prev;
  p.c_target = h->c_target.as<double>(); p.c_W = h->c_W.as<double>(); p.c_WT = h->c_WT.as<double>();
  p.c_smin = h->c_smin.as<double>(); p.c_smax = h->c_smax.as<double>();
  p.kappa = h->kappa; p.use_constraints = h->use_constraints; p.clip = h->clip;
  // outputs the caller did not ask for go to handle-owned scratch
  if (!cost) { CU(h->t_cost.ensure(sizeof(double) * B)); cost = h->t_cost.as<double>(); }
  if (!states_mu) { CU(h->t_mu.ensure(sizeof(double) * (size_t)B * (H + 1) * E)); states_mu = h->t_mu.as<double>(); }
  if (!states_var) { CU(h->t_var.ensure(sizeof(double) * (size_t)B * (H + 1) * E * E)); states_var = h->t_var.as<double>(); }
  if (!rewards) { CU(h->t_r.ensure(sizeof(double) * (size_t)B * (H + 1))); rewards = h->t_r.as<double>(); }
  if (!rewards_var) { CU(h->t_rv.ensure(sizeof(double) * (size_t)B * (H + 1))); rewards_var = h->t_rv.as<double>(); }
  if (!actions_model) { CU(h->t_am.ensure(sizeof(double) * (size_t)B * H * Na)); actions_model = h->t_am.as<double>(); }
  p.cost = cost; p.states_mu = states_mu; p.states_var = states_var; p.rewards = rewards; p.rewards_var = rewards_var;
  p.actions_model = actions_model;
  if (want_grad && !use_uniform) {   // general-path records (the uniform path sizes its own, 12x smaller, layout below)
    const RecLayout RL = rec_layout(E, D);
    CU(h->records.ensure(sizeof(double) * (size_t)B * H * RL.size));
    p.records = h->records.as<double>();
  }
  // candidate counters of the dynamic scheduling: [0] uniform forward, [1] uniform reverse sweep, [2] general kernel
  CU(h->queue.ensure(sizeof(int) * 4));
  CU(cudaMemsetAsync(h->queue.ptr, 0, sizeof(int) * 4, st));
  p.queue = h->queue.as<int>();
  if (use_uniform) {
    // uniform-kernel fast path: one exp per (i, j) for all output pairs; reverse-mode second sweep for the gradient
    const UniRecLayout UR = uni_rec_layout(E);
    if (want_grad) {
      CU(h->records.ensure(sizeof(double) * (size_t)B * H * UR.size));
      p.records = h->records.as<double>();
    } else {
      p.records = nullptr;
    }
    const size_t smf = uniform_smem_bytes(E, false, h->NP, h->DP, D, H, Na, false);
    size_t smb = uniform_smem_bytes(E, true, h->NP, h->DP, D, H, Na, false);
    // Reverse sweep as 3 CTAs x 128 threads (168-register build) instead of 2 x 256: more co-resident CTAs hide the serial
    // phases better (+9 % at N=200, +4 % at N=300, +1.8 % at N=500), but 444 slots of 4 warps make a longer last wave than
    // 296 slots of 8 (-3.4 % at 1024 candidates, N=500): taken from ~4 waves on, from 2 for small training sets
    // (profiles/r02x_reverse_sweep_plans.txt)
    const bool bwd3 = E <= 5 && (B >= 12 * h->num_sms || (h->NP <= 320 && B >= 6 * h->num_sms));
    {  // precomputed per-step matrices: only if they do not cost a resident CTA (or the launch itself)
      const size_t with = uniform_smem_bytes(E, true, h->NP, h->DP, D, H, Na, true);
      const size_t two = 112 * 1024 - GPMPC_STATIC_SMEM;   // two CTAs per SM
      const size_t lim = smb <= two ? two : h->smem_optin;
      p.premat = with <= lim ? 1 : 2;     // 2: the precomputed records go to a per-CTA global scratch (allocated below)
      // E <= 5: three 128-thread CTAs per SM (the 168-register build) beat two of 256 when they fit; at the headline shape
      // they do once the records are in the global scratch (reading them from L2 costs nothing measurable)
      auto fits3 = [&](size_t sm) { return (size_t)(227 * 1024) / (sm + GPMPC_STATIC_SMEM + 1024) >= 3; };
      if (bwd3 && !fits3(with) && fits3(smb)) p.premat = 2;
      if (const char* e = getenv("GPMPC_UNI_PREMAT")) { const int v = atoi(e); p.premat = v == 0 ? 0 : ((v == 1 && with <= lim) ? 1 : 2); }
      if (p.premat == 1) smb = with;
    }
    if (smf > h->smem_optin || (want_grad && smb > h->smem_optin))
      return fail(h, GPMPC_ERR_UNSUPPORTED, "rollout: training set too large for the shared-memory plan (N, D)");
    // several small CTAs per SM so that one CTA's serial small-matrix phases overlap another's N^2 sweep
    auto plan = [&](size_t sm, int* thr, int* grd, const char* env_thr, const char* env_ctas, bool fwd) {
      const size_t per_sm = 227 * 1024;
      int fit = (int)(per_sm / (sm + GPMPC_STATIC_SMEM + 1024));
      if (fit < 1) fit = 1;
      int ctas = fit > 2 ? 2 : fit;      // tuned on B200: 2 CTAs x 256 threads per SM (16 warps, <= 128 registers)
      *thr = 256;
      // state dimensions <= 5: 3 CTAs x 128 threads with <= 168 registers when they fit (forward: 57.9 vs 60.6 ms)
      if (E <= 5 && fit >= 3 && (fwd || bwd3)) { ctas = 3; *thr = 128; }
      if (const char* e = getenv(env_thr)) { int v = atoi(e); if (v == 128 || v == 256 || (!fwd && v == 192)) *thr = v; }   // tuning aid
      if (const char* e = getenv(env_ctas)) { int v = atoi(e); if (v >= 1 && v <= fit) ctas = v; }
      if (E > 5) { *thr = UNI_MAXT(E); ctas = 1; }   // large state dims: one CTA per SM (tensor-core sweeps: 384 threads)
      if (ctas * (*thr) > 512) ctas = 512 / (*thr);
      int g = h->num_sms * ctas;
      *grd = B < g ? B : g;
    };
    int thr_f, grid_f, thr_b, grid_b;
    plan(smf, &thr_f, &grid_f, "GPMPC_UNI_FWD_THREADS", "GPMPC_UNI_FWD_CTAS", true);
    plan(smb, &thr_b, &grid_b, "GPMPC_UNI_BWD_THREADS", "GPMPC_UNI_BWD_CTAS", false);
    // Small batches (fewer candidates than SMs, e.g. the single sequence scipy's L-BFGS-B evaluates): a thread-block
    // cluster of 2 / 4 / 8 CTAs shares each candidate, provided the tile triangle has >= 2 chunks of 8 columns per warp.
    p.cluster = 1;
    {
      const int nrb = h->NP / 64, chunks8 = 8 * nrb * (nrb + 1) / 2;
      // capacity: clusters of c CTAs (256 threads) the device holds at once -- two CTAs per SM are co-resident when the
      // shared memory allows it, so a mid-size batch (the batched optimiser's 64 .. 128 candidates) still gets a cluster
      // per candidate; asked from the driver (GPC boundaries), cached per (c, shared-memory plan)
      auto capacity = [&](int c) -> int {
        const size_t key = (want_grad ? smb : smf) * 16 + c;
        for (int k = 0; k < 8; k++)
          if (h->cl_cap_key[k] == key) return h->cl_cap[k];
        int nf = 0, nb = 1 << 30;
        if (uniform_max_clusters(E, false, c, 256, smf, &nf) != cudaSuccess) { cudaGetLastError(); nf = h->num_sms / c; }
        if (want_grad && uniform_max_clusters(E, true, c, 256, smb, &nb) != cudaSuccess) { cudaGetLastError(); nb = h->num_sms / c; }
        const int n = nf < nb ? nf : nb;
        h->cl_cap_key[h->cl_cap_next] = key; h->cl_cap[h->cl_cap_next] = n; h->cl_cap_next = (h->cl_cap_next + 1) % 8;
        return n;
      };
      int cap_mode = 2;
      if (const char* e = getenv("GPMPC_UNI_CLUSTER_CAP")) { int v = atoi(e); if (v == 1 || v == 2) cap_mode = v; }   // tuning aid: 1 = one CTA per SM
      int c = 8;
      while (c > 1 && (chunks8 < 2 * c * 8 || (cap_mode == 1 ? B * c > h->num_sms : B > capacity(c)))) c /= 2;   // cluster launches use 256 threads
      if (E <= 5) p.cluster = c;       // the 255-register kernels (E > 5) stay on the plain path
      if (const char* e = getenv("GPMPC_UNI_CLUSTER")) { int v = atoi(e); if (v == 1 || ((v == 2 || v == 4 || v == 8) && B <= capacity(v))) p.cluster = v; }
    }
    if (p.cluster > 1) {
      grid_f = grid_b = B * p.cluster;
      thr_f = thr_b = 256;
      CU(h->ws_cl.ensure(sizeof(double) * (size_t)B * 3 * 64));
      CU(cudaMemsetAsync(h->ws_cl.ptr, 0, sizeof(double) * (size_t)B * 3 * 64, st));
      p.ws_cl = h->ws_cl.as<double>();
    }
    {  // small training sets / clusters: finer chunks so that every warp of the sweeps gets a run
      const int nrb = h->NP / 64;
      while (p.seg > 8 && (64 / p.seg) * nrb * (nrb + 1) / 2 < 2 * p.cluster * (thr_f / 32)) p.seg /= 2;
      while (p.seg_bwd > 8 && (64 / p.seg_bwd) * nrb * (nrb + 1) / 2 < 2 * p.cluster * (thr_b / 32)) p.seg_bwd /= 2;
    }
    if (const char* e = getenv("GPMPC_UNI_SEG")) { int v = atoi(e); if (v == 8 || v == 16 || v == 32 || v == 64) p.seg = v; }
    if (const char* e = getenv("GPMPC_UNI_SEG_BWD")) { int v = atoi(e); if (v == 8 || v == 16 || v == 32 || v == 64) p.seg_bwd = v; }
    const bool dbg = getenv("GPMPC_DEBUG_CLOCKS") != nullptr;   // tuning aid: per-phase cycles of CTA 0 on stderr
    if (dbg) {
      CU(h->dbg_clk.ensure(sizeof(long long) * (16 + 4 * 2048)));
      CU(cudaMemsetAsync(h->dbg_clk.ptr, 0, sizeof(long long) * (16 + 4 * 2048), st));
      p.dbg_clk = h->dbg_clk.as<long long>();
    }
    if (h->timing) CU(cudaEventRecord(h->ev[0], st));
    CU(launch_uniform(E, false, p, nullptr, grid_f, thr_f, smf, st));
    h->launches += 1;
    if (h->timing) { CU(cudaEventRecord(h->ev[1], st)); h->ev_fwd = true; }
    h->ev_bwd = false;
    if (want_grad) {
      const size_t scr = sizeof(double) * (size_t)h->NP * (2 + E);
      if (p.cluster > 1) {             // three rotating buffers per cluster, zeroed here
        CU(h->ws_uni.ensure(scr * 3 * B));
        CU(cudaMemsetAsync(h->ws_uni.ptr, 0, scr * 3 * B, st));
      } else {
        CU(h->ws_uni.ensure(scr * grid_b));   // zeroed by the kernel itself
      }
      p.ws_uni = h->ws_uni.as<double>();
      p.ws_pre = nullptr;
      if (p.premat == 2) {
        CU(h->ws_pre.ensure(sizeof(double) * (size_t)grid_b * H * uniform_premat_len(E, h->NP, h->DP, D, H, Na)));
        p.ws_pre = h->ws_pre.as<double>();
      }
      if (h->timing) CU(cudaEventRecord(h->ev[2], st));
      CU(launch_uniform(E, true, p, grad, grid_b, thr_b, smb, st));
      h->launches += 1;
      if (h->timing) { CU(cudaEventRecord(h->ev[3], st)); h->ev_bwd = true; }
    }
    if (dbg) {
      static long long c[16 + 4 * 2048];
      CU(cudaMemcpyAsync(c, h->dbg_clk.ptr, sizeof(c), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      for (int k = 0; k < (want_grad ? 2 : 1); k++) {   // per-CTA life: cycles, wall time, start skew
        const long long* d = c + 16 + 4 * 1024 * k;
        const int g = k ? grid_b : grid_f;
        if (g > 1024) continue;
        long long cmin = d[0], cmax = d[0], nmin = d[1], nmax = d[1], t0min = d[3], endmax = d[3] + d[1];
        double csum = 0, nsum = 0;
        int imax = 0;
        for (int i = 0; i < g; i++) {
          const long long* q = d + 4 * i;
          if (q[0] < cmin) cmin = q[0];
          if (q[0] > cmax) cmax = q[0];
          if (q[1] < nmin) nmin = q[1];
          if (q[1] > nmax) { nmax = q[1]; imax = i; }
          if (q[3] < t0min) t0min = q[3];
          if (q[3] + q[1] > endmax) endmax = q[3] + q[1];
          csum += (double)q[0]; nsum += (double)q[1];
        }
        fprintf(stderr, "[gpmpc CTA life %s] grid %d: cycles min %.2fM mean %.2fM max %.2fM | ns min %.2fM mean %.2fM max %.2fM (CTA %d, SM %lld) | mean clock %.0f MHz | first start -> last end %.2f ms\n",
                k ? "bwd" : "fwd", g, cmin * 1e-6, csum / g * 1e-6, cmax * 1e-6, nmin * 1e-6, nsum / g * 1e-6, nmax * 1e-6, imax,
                d[4 * imax + 2], csum / nsum * 1e3, (endmax - t0min) * 1e-6);
      }
      const long long nf = (long long)((B + grid_f - 1) / grid_f) * H, nb = (long long)((B + grid_b - 1) / grid_b) * H;
      fprintf(stderr, "[gpmpc clocks/step, CTA 0] fwd: P0 %lld  P1a %lld  P1b %lld  P3 sweep %lld  P4 %lld | bwd: pre %lld  B0 %lld  B1a %lld  B1b %lld  B2 sweep %lld  B3a %lld  B3b %lld  B4 %lld\n",
              c[0] / nf, c[4] / nf, c[1] / nf, c[2] / nf, c[3] / nf, want_grad ? c[8] / nb : 0, want_grad ? c[9] / nb : 0,
              want_grad ? c[14] / nb : 0, want_grad ? c[10] / nb : 0, want_grad ? c[11] / nb : 0, want_grad ? c[15] / nb : 0,
              want_grad ? c[12] / nb : 0, want_grad ? c[13] / nb : 0);
    }
    return GPMPC_OK;
  }
  if (p.cluster > 1) {   // exchange buffers of the cluster path: (B, 2, 64) doubles, plain stores (no zeroing needed)
    CU(h->ws_cl.ensure(sizeof(double) * (size_t)B * 2 * 64));
    p.ws_cl = h->ws_cl.as<double>();
  }
  const bool dbg_gen = getenv("GPMPC_DEBUG_CLOCKS") != nullptr;   // tuning aid: per-phase cycles of CTA 0 on stderr
  if (dbg_gen) {
    CU(h->dbg_clk.ensure(sizeof(long long) * (16 + 4 * 2048)));
    CU(cudaMemsetAsync(h->dbg_clk.ptr, 0, sizeof(long long) * 64, st));
    p.dbg_clk = h->dbg_clk.as<long long>();
  }
  CU(launch_colcoef(h->beta.as<double>(), h->colcoef.as<double>(), h->E * h->NP, st));
  if (h->timing) CU(cudaEventRecord(h->ev[0], st));
  CU(launch_rollout(E, want_grad, p, grid, threads, smem, st));
  h->launches += 2;
  if (dbg_gen) {
    long long c[64];
    CU(cudaMemcpyAsync(c, h->dbg_clk.ptr, sizeof(c), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const long long ns = (long long)((B + grid / p.cluster - 1) / (grid / p.cluster)) * H;
    {  // per-warp duration of the sweep: the static split's balance
      long long mn = c[32], mx = c[32];
      for (int w = 0; w < threads / 32; w++) { if (c[32 + w] < mn) mn = c[32 + w]; if (c[32 + w] > mx) mx = c[32 + w]; }
      fprintf(stderr, "[gpmpc general sweep per warp, CTA 0, %d warps] fastest %lld  slowest %lld clocks/step:", threads / 32, mn / ns, mx / ns);
      for (int w = 0; w < threads / 32; w++) fprintf(stderr, " %lld", c[32 + w] / ns);
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "[gpmpc general clocks/step, CTA 0, cluster %d, group %d] P0a %lld  P0b %lld  P1 %lld  P2 %lld  P2b %lld | per step over its pairs: P3 setup %lld  sweep %lld  reduce %lld  finalize+next %lld | exchange %lld  P4 %lld\n",
            p.cluster, p.group, c[0] / ns, c[1] / ns, c[2] / ns, c[3] / ns, c[4] / ns, c[5] / ns, c[6] / ns, c[7] / ns, c[9] / ns, c[8] / ns, 0LL);
  }
  if (h->timing) { CU(cudaEventRecord(h->ev[1], st)); h->ev_fwd = true; }
  h->ev_bwd = false;
  if (want_grad) {
    BackwardParams b;
    memset(&b, 0, sizeof(b));
    b.il2 = p.il2; b.s2 = p.s2; b.D = D; b.E = E; b.Na = Na; b.B = B; b.H = H;
    b.limit_change = p.limit_change; b.include_time = include_time; b.max_change = max_change;
    b.c_target = p.c_target; b.c_W = p.c_W; b.c_WT = p.c_WT; b.c_smin = p.c_smin; b.c_smax = p.c_smax;
    b.kappa = p.kappa; b.use_constraints = p.use_constraints;
    b.states_mu = states_mu; b.states_var = states_var; b.rewards_var = rewards_var; b.actions_model = actions_model;
    b.records = p.records; b.grad = grad;
    if (h->timing) CU(cudaEventRecord(h->ev[2], st));
    CU(launch_backward(E, b, st));
    h->launches += 1;
    if (h->timing) { CU(cudaEventRecord(h->ev[3], st)); h->ev_bwd = true; }
  }
  return GPMPC_OK;
}

int gpmpc_mll(gpmpc_handle* h, const double* y, double* out, void* stream) {
  if (!h) return GPMPC_ERR_BAD_ARG;
  if (!h->prepared) return fail(h, GPMPC_ERR_NOT_PREPARED, "mll: call gpmpc_prepare first (same x, y, hyper-parameters)");
  if (!y || !out) return fail(h, GPMPC_ERR_BAD_ARG, "mll: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaSetDevice(h->device));
  CU(launch_mll(h->x.as<double>(), y, h->ls.as<double>(), h->s2.as<double>(), h->Kbuf.as<double>(), h->iK.as<double>(),
                h->beta.as<double>(), out, h->N, h->NP, h->D, h->E, 3 + h->D, st, &h->launches));
  return GPMPC_OK;
}

// The captured graph holds raw device pointers: any other call that re-allocates a buffer of the handle (a larger
// gpmpc_prepare in between) must invalidate it.
static size_t fit_buffer_fingerprint(const gpmpc_handle* h) {
  const void* ptrs[] = {h->x.ptr, h->ls.ptr, h->il2.ptr, h->s2.ptr, h->noise.ptr, h->beta.ptr, h->betaT.ptr, h->iK.ptr,
                        h->Kbuf.ptr, h->Zbuf.ptr, h->info.ptr, h->fit_out.ptr};
  size_t f = 1469598103934665603ull;
  for (const void* q : ptrs) f = (f ^ reinterpret_cast<size_t>(q)) * 1099511628211ull;
  return f;
}

// Everything one objective evaluation of the fit puts on the stream: trial hyper-parameters from the pinned staging
// buffer, Gram + Cholesky + inverse (launch_prepare), LML + gradient (launch_mll), results back to the staging buffer.
static int fit_enqueue(gpmpc_handle* h, const double* x, const double* y, int N, int NP, int D, int E, cudaStream_t st) {
  double* pin = h->fit_pin;
  double* pin_out = pin + E * (D + 2);
  int* pin_info = reinterpret_cast<int*>(pin_out + E * (3 + D));
  CU(cudaMemcpyAsync(h->x.ptr, x, sizeof(double) * N * D, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(h->ls.ptr, pin, sizeof(double) * E * D, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->s2.ptr, pin + E * D, sizeof(double) * E, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->noise.ptr, pin + E * D + E, sizeof(double) * E, cudaMemcpyHostToDevice, st));
  CU(launch_il2(h->ls.as<double>(), h->il2.as<double>(), E * D, st));
  h->launches += 1;
  CU(launch_prepare(h->x.as<double>(), y, h->ls.as<double>(), h->s2.as<double>(), h->noise.as<double>(), N, NP, D, E,
                    h->Kbuf.as<double>(), h->Zbuf.as<double>(), h->iK.as<double>(), h->beta.as<double>(),
                    h->betaT.as<double>(), h->info.as<int>(), st, &h->launches));
  CU(launch_mll(h->x.as<double>(), y, h->ls.as<double>(), h->s2.as<double>(), h->Kbuf.as<double>(), h->iK.as<double>(),
                h->beta.as<double>(), h->fit_out.as<double>(), N, NP, D, E, 3 + D, st, &h->launches));
  CU(cudaMemcpyAsync(pin_out, h->fit_out.ptr, sizeof(double) * E * (3 + D), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(pin_info, h->info.ptr, sizeof(int) * E, cudaMemcpyDeviceToHost, st));
  return GPMPC_OK;
}

int gpmpc_fit_eval(gpmpc_handle* h, const double* x, const double* y, const double* theta, int N, int D, int E,
                   double* out, int* info, void* stream) {
  if (!h) return GPMPC_ERR_BAD_ARG;
  if (!x || !y || !theta || !out || !info || N < 1) return fail(h, GPMPC_ERR_BAD_ARG, "fit_eval: null pointer or N < 1");
  if (E < 1 || E > GPMPC_MAX_STATE || D < E || D > GPMPC_MAX_INPUT)
    return fail(h, GPMPC_ERR_UNSUPPORTED, "fit_eval: need 1 <= E <= 8 and E <= D <= 16");
  CU(cudaSetDevice(h->device));
  const int NP = (N + 63) / 64 * 64;
  h->prepared = false;
  if (!h->fit_stream) CU(cudaStreamCreateWithFlags(&h->fit_stream, cudaStreamNonBlocking));
  if (!h->fit_pin) CU(cudaHostAlloc(reinterpret_cast<void**>(&h->fit_pin),
                                    sizeof(double) * GPMPC_MAX_STATE * (2 * GPMPC_MAX_INPUT + 6), cudaHostAllocDefault));
  double* pin = h->fit_pin;
  for (int a = 0; a < E; a++) {           // theta rows {ls[D], s2, noise} -> {ls (E,D), s2 (E), noise (E)}
    for (int d = 0; d < D; d++) pin[a * D + d] = theta[a * (D + 2) + d];
    pin[E * D + a] = theta[a * (D + 2) + D];
    pin[E * D + E + a] = theta[a * (D + 2) + D + 1];
  }
  if (!h->fit_exec || h->fit_x != x || h->fit_y != y || h->fit_N != N || h->fit_D != D || h->fit_E != E ||
      h->fit_bufs != fit_buffer_fingerprint(h)) {
    // (re)capture: buffers first (no allocation inside a capture), the caller's stream drained once (x, y uploads)
    if (h->fit_exec) { cudaGraphExecDestroy(h->fit_exec); h->fit_exec = nullptr; }
    { const int rc = prepare_buffers(h, NP, D, E); if (rc != GPMPC_OK) return rc; }
    CU(h->fit_out.ensure(sizeof(double) * GPMPC_MAX_STATE * (3 + GPMPC_MAX_INPUT)));
    CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    const long long l0 = h->launches;
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(h->fit_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = fit_enqueue(h, x, y, N, NP, D, E, h->fit_stream);
    const cudaError_t ce = cudaStreamEndCapture(h->fit_stream, &graph);
    if (rc != GPMPC_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, "fit_eval: stream capture", ce);
    const cudaError_t ci = cudaGraphInstantiate(&h->fit_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ci != cudaSuccess) { h->fit_exec = nullptr; return fail(h, GPMPC_ERR_CUDA, "fit_eval: graph instantiation", ci); }
    h->fit_launches = h->launches - l0;
    h->launches = l0;
    h->fit_x = x; h->fit_y = y; h->fit_N = N; h->fit_D = D; h->fit_E = E;
    h->fit_bufs = fit_buffer_fingerprint(h);
  }
  CU(cudaGraphLaunch(h->fit_exec, h->fit_stream));
  h->launches += h->fit_launches;
  CU(cudaStreamSynchronize(h->fit_stream));
  const double* pin_out = pin + E * (D + 2);
  const int* pin_info = reinterpret_cast<const int*>(pin_out + E * (3 + D));
  bool ok = true;
  for (int a = 0; a < E; a++) { info[a] = pin_info[a]; ok = ok && pin_info[a] == 0; }
  for (int k = 0; k < E * (3 + D); k++) out[k] = pin_out[k];
  bool uni = true;
  for (int a = 1; a < E && uni; a++)
    for (int k = 0; k < D + 2; k++) uni = uni && (theta[a * (D + 2) + k] == theta[k]);
  h->uniform = uni;
  h->N = N; h->NP = NP; h->D = D; h->DP = (D + 1) & ~1; h->E = E;
  h->prepared = ok;
  return GPMPC_OK;
}

int gpmpc_set_path(gpmpc_handle* h, int mode) {
  if (!h || mode < 0 || mode > 1) return GPMPC_ERR_BAD_ARG;
  h->path_mode = mode;
  return GPMPC_OK;
}

int gpmpc_uses_uniform_path(const gpmpc_handle* h) { return (h && h->uniform && h->path_mode == 0) ? 1 : 0; }

int gpmpc_enable_timing(gpmpc_handle* h, int on) {
  if (!h) return GPMPC_ERR_BAD_ARG;
  h->timing = on != 0;
  h->ev_fwd = h->ev_bwd = false;
  return GPMPC_OK;
}

long long gpmpc_launch_count(const gpmpc_handle* h) { return h ? h->launches : 0; }

float gpmpc_last_rollout_ms(gpmpc_handle* h) {
  if (!h || !h->ev_fwd) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) != cudaSuccess) { cudaGetLastError(); return -1.0f; }
  return ms;
}

float gpmpc_last_backward_ms(gpmpc_handle* h) {
  if (!h || !h->ev_bwd) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) != cudaSuccess) { cudaGetLastError(); return -1.0f; }
  return ms;
}

}  // extern "C"
