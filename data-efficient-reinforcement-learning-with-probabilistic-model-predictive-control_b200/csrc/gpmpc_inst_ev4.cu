// Instantiation unit: rollout / reverse-sweep kernels for EV = 4.
#include "gpmpc_uniform_impl.cuh"
namespace gpmpc {
template cudaError_t launch_rollout_inst<4>(bool, const RolloutParams&, int, int, size_t, cudaStream_t);
template cudaError_t launch_backward_inst<4>(const BackwardParams&, cudaStream_t);
template cudaError_t launch_uniform_inst<4>(bool, const RolloutParams&, double*, int, int, size_t, cudaStream_t);
template cudaError_t max_clusters_uniform_inst<4>(bool, int, int, size_t, int*);
template cudaError_t max_clusters_rollout_inst<4>(bool, int, int, size_t, int*);
}  // namespace gpmpc
