// Host-side dispatch of the rollout kernels (templates live in gpmpc_rollout_impl.cuh, one
// instantiation unit per EV in gpmpc_inst_evN.cu).
#include "gpmpc_rollout_layout.cuh"
#include "gpmpc_uniform_layout.cuh"

namespace gpmpc {

size_t rollout_smem_bytes(int EV, bool grad, int NP, int DP, int D, int E, int group, int H, int Na, int nwarps, bool lb_global) {
  SmemLayout L = make_layout(EV, grad, NP, DP, D, E, group, H, Na, nwarps, lb_global);
  return (size_t)L.total * sizeof(double);
}

int rollout_pick_group(int EV, bool grad, int NP, int DP, int D, int E, int H, int Na, int nwarps, size_t smem_limit, bool lb_global) {
  const int P = E * (E + 1) / 2;
  int best = 0;
  for (int g = 1; g <= P; g++) {
    if (rollout_smem_bytes(EV, grad, NP, DP, D, E, g, H, Na, nwarps, lb_global) <= smem_limit) best = g;
  }
  return best;
}


template <int EV> cudaError_t launch_uniform_inst(bool bwd, const RolloutParams& p, double* grad, int grid, int threads, size_t smem, cudaStream_t st);
template <int EV> cudaError_t launch_rollout_inst(bool grad, const RolloutParams& p, int grid, int threads, size_t smem, cudaStream_t st);
template <int E> cudaError_t launch_backward_inst(const BackwardParams& p, cudaStream_t st);
template <int EV> cudaError_t max_clusters_uniform_inst(bool bwd, int cluster, int threads, size_t smem, int* nclusters);
template <int EV> cudaError_t max_clusters_rollout_inst(bool grad, int cluster, int threads, size_t smem, int* nclusters);
#define GPMPC_DECL(n)                                                                                             \
  extern template cudaError_t launch_rollout_inst<n>(bool, const RolloutParams&, int, int, size_t, cudaStream_t);      \
  extern template cudaError_t launch_backward_inst<n>(const BackwardParams&, cudaStream_t);                       \
  extern template cudaError_t launch_uniform_inst<n>(bool, const RolloutParams&, double*, int, int, size_t, cudaStream_t); \
  extern template cudaError_t max_clusters_uniform_inst<n>(bool, int, int, size_t, int*);                         \
  extern template cudaError_t max_clusters_rollout_inst<n>(bool, int, int, size_t, int*);
GPMPC_DECL(1) GPMPC_DECL(2) GPMPC_DECL(3) GPMPC_DECL(4) GPMPC_DECL(5) GPMPC_DECL(6) GPMPC_DECL(7) GPMPC_DECL(8)
#undef GPMPC_DECL

cudaError_t launch_rollout(int EV, bool grad, const RolloutParams& p, int grid, int threads, size_t smem, cudaStream_t st) {
  switch (EV) {
    case 1: return launch_rollout_inst<1>(grad, p, grid, threads, smem, st);
    case 2: return launch_rollout_inst<2>(grad, p, grid, threads, smem, st);
    case 3: return launch_rollout_inst<3>(grad, p, grid, threads, smem, st);
    case 4: return launch_rollout_inst<4>(grad, p, grid, threads, smem, st);
    case 5: return launch_rollout_inst<5>(grad, p, grid, threads, smem, st);
    case 6: return launch_rollout_inst<6>(grad, p, grid, threads, smem, st);
    case 7: return launch_rollout_inst<7>(grad, p, grid, threads, smem, st);
    case 8: return launch_rollout_inst<8>(grad, p, grid, threads, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_uniform(int EV, bool bwd, const RolloutParams& p, double* grad, int grid, int threads, size_t smem, cudaStream_t st) {
  switch (EV) {
    case 1: return launch_uniform_inst<1>(bwd, p, grad, grid, threads, smem, st);
    case 2: return launch_uniform_inst<2>(bwd, p, grad, grid, threads, smem, st);
    case 3: return launch_uniform_inst<3>(bwd, p, grad, grid, threads, smem, st);
    case 4: return launch_uniform_inst<4>(bwd, p, grad, grid, threads, smem, st);
    case 5: return launch_uniform_inst<5>(bwd, p, grad, grid, threads, smem, st);
    case 6: return launch_uniform_inst<6>(bwd, p, grad, grid, threads, smem, st);
    case 7: return launch_uniform_inst<7>(bwd, p, grad, grid, threads, smem, st);
    case 8: return launch_uniform_inst<8>(bwd, p, grad, grid, threads, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t rollout_max_clusters(int EV, bool grad, int cluster, int threads, size_t smem, int* nclusters) {
  switch (EV) {
    case 1: return max_clusters_rollout_inst<1>(grad, cluster, threads, smem, nclusters);
    case 2: return max_clusters_rollout_inst<2>(grad, cluster, threads, smem, nclusters);
    case 3: return max_clusters_rollout_inst<3>(grad, cluster, threads, smem, nclusters);
    case 4: return max_clusters_rollout_inst<4>(grad, cluster, threads, smem, nclusters);
    case 5: return max_clusters_rollout_inst<5>(grad, cluster, threads, smem, nclusters);
    case 6: return max_clusters_rollout_inst<6>(grad, cluster, threads, smem, nclusters);
    case 7: return max_clusters_rollout_inst<7>(grad, cluster, threads, smem, nclusters);
    case 8: return max_clusters_rollout_inst<8>(grad, cluster, threads, smem, nclusters);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t uniform_max_clusters(int EV, bool bwd, int cluster, int threads, size_t smem, int* nclusters) {
  switch (EV) {
    case 1: return max_clusters_uniform_inst<1>(bwd, cluster, threads, smem, nclusters);
    case 2: return max_clusters_uniform_inst<2>(bwd, cluster, threads, smem, nclusters);
    case 3: return max_clusters_uniform_inst<3>(bwd, cluster, threads, smem, nclusters);
    case 4: return max_clusters_uniform_inst<4>(bwd, cluster, threads, smem, nclusters);
    case 5: return max_clusters_uniform_inst<5>(bwd, cluster, threads, smem, nclusters);
    case 6: return max_clusters_uniform_inst<6>(bwd, cluster, threads, smem, nclusters);
    case 7: return max_clusters_uniform_inst<7>(bwd, cluster, threads, smem, nclusters);
    case 8: return max_clusters_uniform_inst<8>(bwd, cluster, threads, smem, nclusters);
    default: return cudaErrorInvalidValue;
  }
}

int uniform_premat_len(int EV, int NP, int DP, int D, int H, int Na) {   // doubles per step of the precomputed records
  return make_uni_layout(EV, true, NP, DP, D, H, Na, false).prelen;
}
size_t uniform_smem_bytes(int EV, bool bwd, int NP, int DP, int D, int H, int Na, bool premat) {
  return (size_t)make_uni_layout(EV, bwd, NP, DP, D, H, Na, premat).total * sizeof(double);
}

cudaError_t launch_backward(int E, const BackwardParams& p, cudaStream_t st) {
  switch (E) {
    case 1: return launch_backward_inst<1>(p, st);
    case 2: return launch_backward_inst<2>(p, st);
    case 3: return launch_backward_inst<3>(p, st);
    case 4: return launch_backward_inst<4>(p, st);
    case 5: return launch_backward_inst<5>(p, st);
    case 6: return launch_backward_inst<6>(p, st);
    case 7: return launch_backward_inst<7>(p, st);
    case 8: return launch_backward_inst<8>(p, st);
    default: return cudaErrorInvalidValue;
  }
}

__global__ void il2_kernel(const double* __restrict__ ls, double* __restrict__ il2, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) il2[i] = 1.0 / (ls[i] * ls[i]);
}
// Column coefficients of the general kernel's off-diagonal pairs (gen_cols): the coefficient beta_b,j rides in the exponent
// as log|beta| (table units) and its sign in the magic constant of the exp's range reduction (exp2s_x4_signed).
__global__ void colcoef_kernel(const double* __restrict__ beta, double* __restrict__ cc, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double b = beta[i];
    cc[2 * i] = GPMPC_EXP2S_SCALE * log(fabs(b));
    cc[2 * i + 1] = __hiloint2double(0x43380000, b < 0.0 ? GPMPC_EXP2S_NEG_LO : 0);
  }
}
cudaError_t launch_colcoef(const double* beta, double* colcoef, int n, cudaStream_t st) {
  colcoef_kernel<<<(n + 127) / 128, 128, 0, st>>>(beta, colcoef, n);
  return cudaGetLastError();
}
cudaError_t launch_il2(const double* ls, double* il2, int n, cudaStream_t st) {
  il2_kernel<<<(n + 127) / 128, 128, 0, st>>>(ls, il2, n);
  return cudaGetLastError();
}

}  // namespace gpmpc
