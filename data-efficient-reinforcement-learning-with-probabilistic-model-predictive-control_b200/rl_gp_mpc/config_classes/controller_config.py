"""ControllerConfig: same constructor and fields as the reference's config_classes/controller_config.py:1-26."""

_DEFAULT_OPTIMIZER_PARAMS = {
    "disp": None, "maxcor": 30, "ftol": 1e-99, "gtol": 1e-99, "eps": 1e-2, "maxfun": 30,
    "maxiter": 30, "iprint": -1, "maxls": 30, "finite_diff_rel_step": None,
}


class ControllerConfig:
    def __init__(self, len_horizon: int = 15, actions_optimizer_params: dict = None,
                 init_from_previous_actions: bool = True, restarts_optim: int = 1, optimize: bool = True,
                 num_repeat_actions: int = 1, batched_candidates: int = 0, batched_iters: int = 30,
                 batched_lr: float = 0.05, batched_method: str = "lbfgs", batched_seed: int = None,
                 batched_cuda_graph: bool = False, batched_fused: bool = True):
        """len_horizon: MPC steps; actions_optimizer_params: scipy L-BFGS-B options; init_from_previous_actions:
        warm start from the shifted previous solution; restarts_optim: optimiser restarts; optimize: False =
        random actions (debug); num_repeat_actions: each action is held this many env steps.
        NEW (additive, default off): batched_candidates > 0 replaces the serial scipy restarts by that many candidate
        sequences optimised simultaneously on the device, every objective/gradient evaluation being ONE batched rollout
        (SURVEY.md section 8(f) N1): batched_method "lbfgs" = projected L-BFGS (batched_iters + 1 evaluations),
        "adam" = projected Adam (batched_iters steps of size batched_lr, + 1 value-only evaluation); batched_seed: seed of
        the controller's own generator of the random restarts (None: drawn once; with a process group rank 0's is used);
        batched_fused: the L-BFGS update runs as ONE device kernel per iteration (gpmpc_lbfgs_update) instead of ~100
        tensor operations; batched_cuda_graph: replay the (unfused) iteration as a CUDA graph (capture costs ~100 ms)."""
        self.len_horizon = len_horizon
        self.actions_optimizer_params = dict(_DEFAULT_OPTIMIZER_PARAMS) if actions_optimizer_params is None \
            else actions_optimizer_params
        self.init_from_previous_actions = init_from_previous_actions
        self.restarts_optim = restarts_optim
        self.optimize = optimize
        self.num_repeat_actions = num_repeat_actions
        self.batched_candidates = batched_candidates
        self.batched_iters = batched_iters
        self.batched_lr = batched_lr
        self.batched_seed = batched_seed
        self.batched_cuda_graph = batched_cuda_graph
        self.batched_fused = batched_fused
        if batched_method not in ("lbfgs", "adam"):
            raise ValueError("batched_method must be 'lbfgs' or 'adam'")
        self.batched_method = batched_method
