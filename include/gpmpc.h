/*
 * gpmpc.h -- C ABI of the B200-native GP-MPC inner loop (libgpmpc.so).
 *
 * The reference (SimonRennotte/Data-Efficient-RL-with-Probabilistic-MPC) is pure Python and has
 * no FFI; this header is the boundary a maintainer binds with ctypes.  Every
 * entry point cites the reference interface it replaces (paths relative to rl_gp_mpc/).
 *
 * Conventions
 *   - all array arguments are DEVICE pointers to float64, row-major, contiguous, caller-owned
 *     (PyTorch CUDA allocations); the library owns only the handle and its workspace;
 *   - every call is enqueued on the caller's cudaStream_t (passed as void*); no call synchronises
 *     the stream except where the workspace has to grow (cudaMalloc) and gpmpc_destroy;
 *   - every function returns 0 on success or a negative gpmpc_status; gpmpc_last_error() gives
 *     the message.  Non-finite results (e.g. a negative predicted variance under the sqrt of the
 *     LCB, controllers/gp_mpc_controller.py:270) are NOT errors: NaN flows out like in the reference;
 *   - a handle is bound to one device and is not thread-safe (the reference's hot path is single
 *     threaded, controllers/gp_mpc_controller.py:114-153).
 */
#ifndef GPMPC_H_
#define GPMPC_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpmpc_handle gpmpc_handle;

typedef enum {
  GPMPC_OK = 0,
  GPMPC_ERR_BAD_ARG = -1,
  GPMPC_ERR_NOT_PREPARED = -2,
  GPMPC_ERR_NOT_PD = -3,      /* K + noise*I not positive definite */
  GPMPC_ERR_CUDA = -4,
  GPMPC_ERR_UNSUPPORTED = -5, /* shape outside the compiled limits */
  GPMPC_ERR_NO_DEVICE = -6
} gpmpc_status;

#define GPMPC_MAX_STATE 8   /* dim_state (one GP per state dimension) */
#define GPMPC_MAX_INPUT 16  /* dim_input = dim_state + dim_action (+1 with a time input) */

int gpmpc_version(void);
const char* gpmpc_last_error(const gpmpc_handle* h);

/* Fails with GPMPC_ERR_NO_DEVICE when no CUDA device is usable: there is no CPU fallback. */
int gpmpc_create(gpmpc_handle** out, int device);
int gpmpc_destroy(gpmpc_handle* h);

/*
 * Replaces GpStateTransitionModel.prepare_inference (control_objects/models/gp_model.py:182-191)
 * = calculate_factorizations (gp_model.py:400-431): RBF-ARD Gram matrices of the E GPs,
 * + noise*I, Cholesky, explicit inverse iK and beta = iK y, kept resident in the handle.
 *   x (N,D)  y (N,E)  lengthscale (E,D)  outputscale (E)  noise (E)
 * Returns GPMPC_ERR_NOT_PD (after a stream sync) when a pivot is not positive.
 */
int gpmpc_prepare(gpmpc_handle* h, const double* x, const double* y, const double* lengthscale,
                  const double* outputscale, const double* noise, int N, int D, int E, void* stream);

/*
 * Appends ONE training point to the factorisation of the last gpmpc_prepare in O(N^2) per GP (new Cholesky row, block
 * inverse update of iK, beta) instead of refactorising: the step the reference repeats from scratch every control
 * step after Memory.add (control_objects/memories/gp_memory.py:31-64 ->
 * controllers/gp_mpc_controller.py:114-118 -> gp_model.py:182-191).   x_new (D)  y_new (E), device pointers.
 * Works while N is below the padded size (next multiple of 64; gpmpc_append_room() = points left); beyond that it
 * returns GPMPC_ERR_UNSUPPORTED and the caller runs gpmpc_prepare, which also bounds the accumulated rounding of
 * successive rank-one updates.  GPMPC_ERR_NOT_PD (after a stream sync) leaves the factorisation unchanged.
 */
int gpmpc_append(gpmpc_handle* h, const double* x_new, const double* y_new, void* stream);
int gpmpc_append_room(const gpmpc_handle* h);

/* Copies the cached factorisation out: iK (E,N,N), beta (E,N)  (attributes read at gp_model.py:186). */
int gpmpc_get_factorization(gpmpc_handle* h, double* iK, double* beta, void* stream);

/*
 * Cost description = SetpointStateRewardMapper's config (config_classes/reward_config.py:4-64):
 *   target (E+Na)  = target_state_action_norm
 *   W ((E+Na)^2)   = weight_matrix_cost,   WT (E^2) = weight_matrix_cost_terminal
 *   kappa = exploration_factor; use_constraints/state_min/state_max (E) as in
 *   states_reward_mappers/setpoint_distance_reward_mapper.py:58-66 (variance-as-sigma quirk kept);
 *   clip_lower_bound_cost_to_0 as in controllers/gp_mpc_controller.py:272-274.
 */
int gpmpc_set_cost(gpmpc_handle* h, const double* target, const double* W, const double* WT,
                   double kappa, int use_constraints, const double* state_min,
                   const double* state_max, int clip_lower_bound_cost_to_0, int Na, void* stream);

/*
 * Replaces GpStateTransitionModel.predict_next_state_change (gp_model.py:112-180) for a batch of B
 * Gaussian inputs.  input_mu (B,D); input_var (B,EV,EV) = the leading EV x EV block of the input
 * covariance (all other entries zero, as built at gp_model.py:96-97; EV = E in the rollout).
 * Outputs (any may be NULL): M (B,E) = M.t() rows, S (B,E,E), V (B,D,E) = V.t().
 */
int gpmpc_predict_step(gpmpc_handle* h, const double* input_mu, const double* input_var, int B,
                       int EV, double* M, double* S, double* V, void* stream);

/*
 * Replaces GpMpcController.compute_mean_lcb_trajectory (controllers/gp_mpc_controller.py:229-285)
 * = action mapping (actions_mappers/*:transform_action_mpc_to_action_model)
 * + GpStateTransitionModel.predict_trajectory (gp_model.py:60-110)
 * + SetpointStateRewardMapper.get_rewards_trajectory (setpoint_distance_reward_mapper.py:144-149)
 * + LCB + gradient, for B candidate action sequences at once.
 *   actions_mpc (B, H*Na)        flat optimiser variables
 *   obs_mu (E) / obs_var (E,E)   shared by all candidates (per_candidate_init=0) or (B,E)/(B,E,E)
 *   iter_ctrl                    current_time_idx (used only when the model has a time input)
 *   limit_action_change          0: NormalizationActionMapper, 1: DerivativeActionMapper with
 *                                max_change (Na) and action_prev (Na)
 * Outputs (any may be NULL):
 *   cost (B) = mean_cost_traj_ucb,  grad (B, H*Na)  [NULL => forward only, nothing recorded]
 *   states_mu (B,H+1,E), states_var (B,H+1,E,E), rewards (B,H+1), rewards_var (B,H+1),
 *   actions_model (B,H,Na)
 */
int gpmpc_rollout(gpmpc_handle* h, const double* actions_mpc, const double* obs_mu,
                  const double* obs_var, int per_candidate_init, int B, int H, int Na, int iter_ctrl,
                  int limit_action_change, const double* max_change, const double* action_prev,
                  double* cost, double* grad, double* states_mu, double* states_var,
                  double* rewards, double* rewards_var, double* actions_model, void* stream);

/*
 * Exact-GP log marginal likelihood and its gradient w.r.t. the hyper-parameters, for the E GPs factorised by the
 * last gpmpc_prepare (same x, y, hyper-parameters) -- the objective GpStateTransitionModel.train minimises through
 * gpytorch's ExactMarginalLogLikelihood (control_objects/models/gp_model.py:193-306).  y (N,E) as given to prepare.
 * out (E, 3+D): { LML, dLML/d outputscale, dLML/d noise, dLML/d lengthscale[0..D) } (not divided by N).
 */
int gpmpc_mll(gpmpc_handle* h, const double* y, double* out, void* stream);

/*
 * One objective evaluation of the hyper-parameter fit -- what `loss = -mll(model(train_x), train_y); loss.backward()`
 * computes in control_objects/models/gp_model.py:262-277 -- for all E GPs at their trial hyper-parameters, as ONE CUDA
 * graph launch: gpmpc_prepare + gpmpc_mll are ~85 dependent kernels and launch-bound at N = 500 (2.1 ms as separate
 * launches), and a fit evaluates them several hundred times on the same (x, y).
 *   x (N,D), y (N,E): DEVICE pointers, fixed during a fit (the graph is re-captured when they, N, D or E change);
 *   theta (E, D+2): HOST, per GP { lengthscale[D], outputscale, noise };
 *   out (E, 3+D): HOST, as gpmpc_mll;  info (E): HOST, 0 or 1 + the pivot at which K + noise I stopped being positive
 *   definite (that GP's row of `out` is then meaningless; the other GPs are not affected).
 * Synchronous (returns after the results are on the host).  Leaves the handle prepared at theta iff every info is 0.
 */
int gpmpc_fit_eval(gpmpc_handle* h, const double* x, const double* y, const double* theta, int N, int D, int E,
                   double* out, int* info, void* stream);

/* Kernel-path selection.  When every GP has bitwise-identical hyper-parameters (the reference's state before
 * hyper-parameter training, examples/<env>/config_<env>.py:41-45) gpmpc_rollout uses the "uniform-kernel" path
 * (one exp per (i,j) for all output pairs).  mode 0: automatic (default); mode 1: always the general path.
 * gpmpc_uses_uniform_path reports which one the next rollout will take. */
int gpmpc_set_path(gpmpc_handle* h, int mode);
int gpmpc_uses_uniform_path(const gpmpc_handle* h);

/*
 * Batched projected L-BFGS on the box [0, 1]^n, one update per call and B candidates per launch -- the device-side
 * counterpart of the serial scipy L-BFGS-B restart loop of GpMpcController._get_optimal_actions
 * (control_objects/controllers/gp_mpc_controller.py:125-148; bounds: actions_mappers/normalization_action_mapper.py:13).
 * Stateless: all optimiser state lives in caller-owned device arrays (row-major, float64 unless noted):
 *   x, g (B, n), f (B): current point, its gradient and cost;  S, Y (history, B, n), rho (history, B): curvature pairs,
 *   ring buffer, `head` = slot the next pair goes to;  alpha (B), fails, first (B, int32): step scale, consecutive
 *   rejected trials, "no accepted step yet";  xt (B, n): trial point.
 * have_trial = 0 (first call): writes the first trial point xt from (x, g).  have_trial = 1: (ft (B), gt (B, n)) are the
 * cost and gradient evaluated at xt (e.g. by gpmpc_rollout); the trial is accepted per candidate by the Armijo test on the
 * projected step, the pair (s, y) enters slot `head` (the caller then advances head = (head + 1) % history), and the next
 * trial point is written to xt.  Non-finite ft rejects the trial, non-finite entries of gt count as 0.
 * rl_gp_mpc/control_objects/controllers/batched_optim.py::minimize_box_lbfgs is the executable specification.
 */
int gpmpc_lbfgs_update(int B, int n, int history, int head, int have_trial, double c1, double shrink,
                       double max_first_move, double* x, double* g, double* f, double* S, double* Y, double* rho,
                       double* alpha, int* fails, int* first, double* xt, const double* ft, const double* gt, void* stream);

/* Introspection for benchmarks/tests: number of kernels launched by this handle so far, and the
 * device time [ms] of the last rollout's forward kernel measured with CUDA events on `stream`
 * (valid after the stream has been synchronised; <0 if timing was not enabled). */
int gpmpc_enable_timing(gpmpc_handle* h, int on);
long long gpmpc_launch_count(const gpmpc_handle* h);
float gpmpc_last_rollout_ms(gpmpc_handle* h);
float gpmpc_last_backward_ms(gpmpc_handle* h);

/* Measurement aid for bench.py: times a register-resident DFMA loop (8 independent chains per thread,
 * 4 x 512 threads per SM) and returns the achieved float64 FLOP/s, the compute roofline of this path. */
int gpmpc_fp64_peak(int device, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* GPMPC_H_ */
