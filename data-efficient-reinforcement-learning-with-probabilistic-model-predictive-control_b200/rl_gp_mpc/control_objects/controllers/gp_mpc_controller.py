"""GP-MPC controller with the reference's interface (control_objects/controllers/gp_mpc_controller.py),
its objective evaluated by the B200 CUDA engine.

  compute_mean_lcb_trajectory        gp_mpc_controller.py:229-285 -> one gpmpc_rollout call (value + gradient)
  compute_mean_lcb_trajectory_batch  NEW: B candidate sequences in one call -> costs (B,), grads (B, H*Na)
  get_action / _get_optimal_actions  gp_mpc_controller.py:52-153  same orchestration (scipy L-BFGS-B, restarts)
The five side-effect tensors (:279-283) are kept; for a batch they describe the best candidate.

Multi-GPU (one process per GPU, torch.distributed): GpMpcController(..., process_group=group) shards the candidate
batch of compute_mean_lcb_trajectory_batch and of the batched optimiser over the ranks of `group` (contiguous slices,
rl_gp_mpc/parallel.py): every rank scores / optimises its slice, ONE in-place all-gather brings every candidate's cost
to every rank, all ranks agree on the arg-min, and the winner's trajectory is broadcast from its owner.  Gradients stay
on the owning rank."""
import atexit
import multiprocessing
import sys
import threading
import weakref

import numpy as np
import torch
from scipy.optimize import minimize

from .batched_optim import minimize_box_adam, minimize_box_lbfgs

from rl_gp_mpc.config_classes.total_config import Config
from rl_gp_mpc.control_objects.actions_mappers.action_init_functions import (
    generate_mpc_action_init_frompreviousiter, generate_mpc_action_init_random)
from rl_gp_mpc.control_objects.actions_mappers.derivative_action_mapper import DerivativeActionMapper
from rl_gp_mpc.control_objects.actions_mappers.normalization_action_mapper import NormalizationActionMapper
from rl_gp_mpc.control_objects.memories.gp_memory import Memory
from rl_gp_mpc.control_objects.models.gp_model import GpStateTransitionModel
from rl_gp_mpc.control_objects.observations_states_mappers.normalization_observation_state_mapper import \
    NormalizationObservationStateMapper
from rl_gp_mpc.control_objects.states_reward_mappers.setpoint_distance_reward_mapper import SetpointStateRewardMapper
from rl_gp_mpc.control_objects.utils.pytorch_utils import Clamp

from .abstract_controller import BaseControllerObject
from .iteration_info_class import IterationInformation


class _TrainingThread(threading.Thread):
    """threading.Thread with the few multiprocessing.Process members the controller logic relies on
    (`_closed`, `close()`), so that check_and_close_processes reads like the reference's (:216-227).  The fit can be
    asked to stop (`stop_event`, polled once per objective evaluation); live threads are stopped and joined at
    interpreter exit, because a daemon thread killed inside a CUDA call aborts the process."""
    _live = weakref.WeakSet()

    def __init__(self, target, args):
        super().__init__(target=self._guarded, daemon=True)
        self._fn, self._fn_args, self._closed = target, args, False
        self.stop_event = threading.Event()
        _TrainingThread._live.add(self)

    def _guarded(self):
        queue, saved_state = self._fn_args[0], self._fn_args[1]
        try:
            self._fn(*self._fn_args, stop_event=self.stop_event)
        except Exception as exc:        # never leave the controller waiting on an empty queue
            print("training failed:", exc)
            saved_state.to_arrays()
            queue.put(saved_state.parameters)

    def close(self):
        self._closed = True

    def stop(self, timeout=60.0):
        self.stop_event.set()
        if self.is_alive():
            self.join(timeout)


@atexit.register
def _stop_training_threads():
    for t in list(_TrainingThread._live):
        t.stop()


class GpMpcController(BaseControllerObject):
    def __init__(self, observation_low, observation_high, action_low, action_high, config: Config, device=None,
                 process_group=None):
        """process_group (additive): None = single GPU; True / "world" = the default torch.distributed group; or a
        ProcessGroup.  Every rank must construct the controller with the same config and feed it the same memory."""
        self.config = config
        self.sharder = None
        if process_group is not None and process_group is not False:
            from rl_gp_mpc.parallel import CandidateSharder
            self.sharder = CandidateSharder(None if process_group is True or process_group == "world" else process_group)
        self.shard = None          # (lo, hi) of the last sharded batch: rows of the returned gradients
        self._cand_gen = None
        self.observation_state_mapper = NormalizationObservationStateMapper(
            config=config.observation, observation_low=observation_low, observation_high=observation_high)
        mapper_cls = DerivativeActionMapper if config.actions.limit_action_change else NormalizationActionMapper
        self.actions_mapper = mapper_cls(config=config.actions, action_low=action_low, action_high=action_high,
                                         len_horizon=config.controller.len_horizon)
        self.transition_model = GpStateTransitionModel(config=config.model,
                                                       dim_state=self.observation_state_mapper.dim_observation,
                                                       dim_action=self.actions_mapper.dim_action, device=device)
        self.state_reward_mapper = SetpointStateRewardMapper(config=config.reward)
        self.memory = Memory(config.memory, dim_input=self.transition_model.dim_input,
                             dim_state=self.transition_model.dim_state,
                             include_time_model=self.transition_model.config.include_time_model,
                             step_model=config.controller.num_repeat_actions)
        self.actions_mpc_previous_iter = None
        self.clamp_lcb_class = Clamp
        self.iter_ctrl = 0
        self.num_cores_main = multiprocessing.cpu_count()
        self.ctx = multiprocessing.get_context("spawn")
        self.queue_train = self.ctx.Queue()
        self.info_iters = {}
        self._cost_bound = False
        self._cost_key_values = None

    # ------------------------------------------------------------------ control step
    def get_action(self, obs_mu, obs_var=None, random: bool = False):
        """Optimal (or random) raw action for the current observation (reference :52-112)."""
        self.check_and_close_processes()
        ctl = self.config.controller
        if self.iter_ctrl % ctl.num_repeat_actions == 0:
            self.memory.prepare_for_model()
            state_mu, state_var = self.observation_state_mapper.get_state(obs=obs_mu, obs_var=obs_var,
                                                                          update_internals=True)
            actions_model = self._get_random_actions(state_mu, state_var) if random \
                else self._get_optimal_actions(state_mu, state_var)
            actions_raw = self.actions_mapper.transform_action_model_to_action_raw(actions_model, update_internals=True)
            next_action_raw = actions_raw[0]
            reward, reward_var = self.state_reward_mapper.get_reward(state_mu, state_var, actions_model[0])
            states_std_pred = torch.diagonal(self.states_var_pred, dim1=-2, dim2=-1).sqrt()
            idxs = np.arange(self.iter_ctrl, self.iter_ctrl + ctl.len_horizon * ctl.num_repeat_actions,
                             ctl.num_repeat_actions)
            self.iter_info = IterationInformation(
                iteration=self.iter_ctrl, state=self.states_mu_pred[0], cost=-reward.item(),
                cost_std=reward_var.sqrt().item(),
                mean_predicted_cost=np.min([-self.rewards_trajectory.mean().item(), 3]),
                mean_predicted_cost_std=self.rewards_traj_var.sqrt().mean().item(),
                lower_bound_mean_predicted_cost=self.cost_traj_mean_lcb.item(), predicted_idxs=idxs,
                predicted_states=self.states_mu_pred, predicted_states_std=states_std_pred,
                predicted_actions=actions_model, predicted_costs=-self.rewards_trajectory,
                predicted_costs_std=self.rewards_traj_var.sqrt())
            self.store_iter_info(self.iter_info)
            self.past_action = next_action_raw
        else:
            next_action_raw = self.past_action
        self.iter_ctrl += 1
        return np.array(torch.as_tensor(next_action_raw).detach().cpu().numpy())

    def _prepare(self):
        x_mem, y_mem = self.memory.get()
        self.transition_model.prepare_inference(x_mem, y_mem)

    def _candidate_inits(self, nb, n):
        """(nb, n) uniform random restarts from the controller's own generator (ControllerConfig.batched_seed; drawn once
        otherwise).  With a process group every rank draws the SAME matrix (rank 0's seed is broadcast once) and then
        works on its rows, so the sharded optimiser starts from the candidates a single GPU would start from."""
        if self._cand_gen is None:
            seed = getattr(self.config.controller, "batched_seed", None)
            if seed is None:
                seed = int(torch.seed() % (2 ** 31))
                if self.sharder is not None:
                    t = torch.tensor([seed], dtype=torch.int64, device=self.transition_model.engine.device)
                    self.sharder.broadcast(t, 0)
                    seed = int(t.item())
            self._cand_gen = torch.Generator(device="cpu").manual_seed(int(seed))
        return torch.rand((nb, n), dtype=torch.float64, generator=self._cand_gen)

    def _get_optimal_actions_batched(self, state_mu, state_var):
        """B candidate action sequences optimised at once on the device (batched_optim.py: projected L-BFGS, or
        projected Adam with ControllerConfig(batched_method="adam")) on the LCB objective.

        Replaces the serial restart loop of the reference (gp_mpc_controller.py:125-148): candidate 0 is the shifted
        previous solution (when init_from_previous_actions), the others are uniform random restarts; each iteration
        costs one batched rollout (value + gradient for all candidates).  Returns the model actions of the best one.
        With a process group the candidates are sharded over the ranks: each rank optimises its slice with no collective
        inside the iterations; the final costs are all-gathered once, the winner's sequence and trajectory broadcast."""
        ctl = self.config.controller
        h, na = ctl.len_horizon, self.actions_mapper.dim_action
        nb = int(ctl.batched_candidates)
        dev = self.transition_model.engine.device
        x = self._candidate_inits(nb, h * na).to(dev)
        if ctl.init_from_previous_actions and self.actions_mpc_previous_iter is not None:
            warm = generate_mpc_action_init_frompreviousiter(self.actions_mpc_previous_iter, dim_action=na)
            x[0] = torch.as_tensor(warm, dtype=torch.float64, device=dev)
        sh = self.sharder
        lo, hi = sh.bounds(nb) if sh is not None else (0, nb)
        x = x[lo:hi]

        # everything the objective needs is staged on the device ONCE, and its results land in the same buffers at every
        # call: an iteration then enqueues kernels only -- the batched rollout and ONE update kernel (gpmpc_lbfgs_update) --
        # with no host-to-device copies, allocations or synchronisations in between
        self._bind_cost()
        eng = self.transition_model.engine
        limit = bool(self.config.actions.limit_action_change)
        mu_d = torch.as_tensor(state_mu, dtype=torch.float64).to(dev)
        var_d = torch.as_tensor(state_var, dtype=torch.float64).to(dev)
        mc_d = torch.as_tensor(self.config.actions.max_change_action_norm, dtype=torch.float64).reshape(-1).to(dev) if limit else None
        ap_d = torch.as_tensor(self.actions_mapper.action_model_previous_iter, dtype=torch.float64).reshape(-1).to(dev) if limit else None
        bufs = {"cost": torch.empty((hi - lo,), dtype=torch.float64, device=dev),
                "grad": torch.empty((hi - lo, h * na), dtype=torch.float64, device=dev)}

        def fun(xb):
            out = eng.rollout(xb, mu_d, var_d, h, iter_ctrl=self.iter_ctrl, limit_action_change=limit, max_change=mc_d,
                              action_prev=ap_d, need_grad=True, need_traj=False, out=dict(bufs))
            return out["cost"], out["grad"]

        if hi > lo:
            if getattr(ctl, "batched_method", "lbfgs") == "adam":
                best_x, best_cost, x_last = minimize_box_adam(fun, x, ctl.batched_iters, lr=ctl.batched_lr)
                cost = self._rollout(x_last, state_mu, state_var, need_grad=False, need_traj=False)["cost"]
                better = torch.isfinite(cost) & (cost < best_cost)
                best_cost = torch.where(better, cost, best_cost)
                best_x = torch.where(better[:, None], x_last, best_x)
            else:
                best_x, best_cost = minimize_box_lbfgs(fun, x, ctl.batched_iters,
                                                       cuda_graph=bool(getattr(ctl, "batched_cuda_graph", False)),
                                                       fused=bool(getattr(ctl, "batched_fused", True)))
        else:
            best_x, best_cost = x, torch.empty((0,), dtype=torch.float64, device=dev)
        if sh is None:
            idx = int(torch.argmin(best_cost).item())
            final = self._rollout(best_x[idx:idx + 1], state_mu, state_var, need_grad=False)   # side effects of the winner
            self._store_side_effects(final, 0)
            self.batched_costs = best_cost
            self.last_optim_cost = float(best_cost[idx].item())
            self.actions_mpc_previous_iter = best_x[idx].cpu().numpy().copy()
            return self.actions_mapper.transform_action_mpc_to_action_model(best_x[idx].cpu())
        buf = sh.cost_buffer(nb, dev)
        sh.local_view(buf, nb).copy_(best_cost)
        costs = sh.gather(buf, nb)
        idx = sh.argmin(buf, nb)
        owner = sh.owner(idx, nb)
        winner = torch.empty((h * na,), dtype=torch.float64, device=dev)
        final = None
        if sh.rank == owner:
            winner.copy_(best_x[idx - lo])
            final = self._rollout(winner[None], state_mu, state_var, need_grad=False)
        sh.broadcast(winner, owner)
        self._share_side_effects(final, 0, owner, 1)
        self.batched_costs = costs.clone()
        self.last_optim_cost = float(costs[idx].item())
        self.actions_mpc_previous_iter = winner.cpu().numpy().copy()
        return self.actions_mapper.transform_action_mpc_to_action_model(winner.cpu())

    def _get_optimal_actions(self, state_mu, state_var):
        self._prepare()
        ctl = self.config.controller
        if getattr(ctl, "batched_candidates", 0) and ctl.optimize:
            return self._get_optimal_actions_batched(state_mu, state_var)
        h, na = ctl.len_horizon, self.actions_mapper.dim_action
        best_val, best_actions = np.inf, None
        for idx_restart in range(ctl.restarts_optim):
            warm = ctl.init_from_previous_actions and self.actions_mpc_previous_iter is not None and idx_restart == 0
            x0 = generate_mpc_action_init_frompreviousiter(self.actions_mpc_previous_iter, dim_action=na) if warm \
                else generate_mpc_action_init_random(len_horizon=h, dim_action=na)
            if ctl.optimize:
                res = minimize(fun=self.compute_mean_lcb_trajectory, x0=x0, jac=True, args=(state_mu, state_var),
                               method="L-BFGS-B", bounds=self.actions_mapper.bounds,
                               options=ctl.actions_optimizer_params)
                cand, val = res.x, res.fun
            else:
                cand = generate_mpc_action_init_random(len_horizon=h, dim_action=na)
                val, _ = self.compute_mean_lcb_trajectory(cand, state_mu, state_var)
            if val < best_val or (best_actions is None and np.isnan(val)):
                best_val, best_actions = val, cand
        self.actions_mpc_previous_iter = best_actions.copy()
        self.last_optim_cost = float(best_val)   # additive: objective value of the returned action sequence
        return self.actions_mapper.transform_action_mpc_to_action_model(torch.as_tensor(best_actions))

    def _get_random_actions(self, state_mu, state_var):
        h, na = self.config.controller.len_horizon, self.actions_mapper.dim_action
        actions_mpc = generate_mpc_action_init_random(len_horizon=h, dim_action=na)
        actions_model = self.actions_mapper.transform_action_mpc_to_action_model(torch.as_tensor(actions_mpc))
        self._prepare()
        self.compute_mean_lcb_trajectory(actions_mpc, state_mu, state_var)   # stores the trajectory info
        return actions_model

    # ------------------------------------------------------------------ memory / training hooks
    def add_memory(self, obs, action, obs_new, reward, predicted_state=None, predicted_state_std=None):
        state_mu, _ = self.observation_state_mapper.get_state(obs=obs, update_internals=False)
        state_mu_new, _ = self.observation_state_mapper.get_state(obs=obs_new, update_internals=False)
        action_model = self.actions_mapper.transform_action_raw_to_action_model(action)
        self.memory.add(state_mu, action_model, state_mu_new, reward, iter_ctrl=self.iter_ctrl - 1,
                        predicted_state=predicted_state, predicted_state_std=predicted_state_std)
        busy = "p_train" in self.__dict__ and not self.p_train._closed
        if self.iter_ctrl % self.config.training.training_frequency == 0 and not busy:
            self.start_training_process()

    def start_training_process(self):
        """Launch the hyper-parameter fit without blocking the control loop (reference :201-214 spawns a process;
        here a thread drives the device-side objective through its own engine handle).  The result is collected by
        check_and_close_processes exactly like in the reference."""
        saved_state = self.transition_model.save_state()
        saved_state.to_arrays()
        tr = self.config.training
        self.p_train = _TrainingThread(target=self.transition_model.train,
                                       args=(self.queue_train, saved_state, tr.lr_train, tr.iter_train,
                                             tr.clip_grad_value, tr.print_train, tr.step_print_train,
                                             self.transition_model._device))
        # the fit and the control loop both make many short blocking CUDA calls; with CPython's default 5 ms switch
        # interval every hand-over of the interpreter lock can stall that long (measured: 26 ms per objective evaluation
        # of the fit instead of ~1 ms), so the interval is shortened while both are active
        if getattr(self, "_switch_interval_saved", None) is None:
            self._switch_interval_saved = sys.getswitchinterval()
        sys.setswitchinterval(min(sys.getswitchinterval(), 2e-4))
        self.p_train.start()

    def _restore_switch_interval(self):
        """The shortened interpreter switch interval only serves the concurrent fit: put the process's own value back."""
        if getattr(self, "_switch_interval_saved", None) is not None:
            sys.setswitchinterval(self._switch_interval_saved)
            self._switch_interval_saved = None

    def check_and_close_processes(self):
        if "p_train" in self.__dict__ and not self.p_train._closed and not self.p_train.is_alive():
            params = self.queue_train.get()
            self.p_train.join()
            for model, p in zip(self.transition_model.models, params):
                model.initialize(**p)
            self.p_train.close()
            self._restore_switch_interval()
            self._prepare()

    def close(self):
        """Stops a hyper-parameter fit that is still running (its result is dropped) -- call at the end of a run."""
        if "p_train" in self.__dict__:
            self.p_train.stop()
        self._restore_switch_interval()

    # ------------------------------------------------------------------ objective
    def _bind_cost(self):
        """(Re)uploads the cost description when the reward config changed -- by identity OR in place: the reference
        re-reads config.reward on every objective evaluation (gp_mpc_controller.py:269), so an in-place edit of targets,
        weights, exploration factor or constraints must reach the fused kernels too (key = the values themselves)."""
        key = self._cost_fingerprint(self.config.reward)
        if not self._cost_bound or self._cost_key_values != key:
            self.transition_model.set_cost(self.config.reward)
            self._cost_bound, self._cost_key_values = True, key

    @staticmethod
    def _cost_fingerprint(r):
        """Cheap key of everything set_cost uploads: tensors by (identity, in-place version counter), the rest by value."""
        parts = [id(r)]
        for name in ("target_state_action_norm", "weight_matrix_cost", "weight_matrix_cost_terminal", "state_min",
                     "state_max", "target_state_norm", "target_action_norm", "weight_state", "weight_action",
                     "weight_state_terminal"):
            v = getattr(r, name, None)
            if torch.is_tensor(v):
                parts.append((id(v), v._version))
            elif v is not None:
                parts.append(tuple(np.asarray(v, dtype=np.float64).reshape(-1).tolist()))
        for name in ("exploration_factor", "use_constraints", "clip_lower_bound_cost_to_0", "area_multiplier"):
            parts.append(getattr(r, name, None))
        return tuple(parts)

    def _rollout(self, actions_mpc, obs_mu, obs_var, need_grad=True, need_traj=True, out=None):
        self._bind_cost()
        am = self.actions_mapper
        limit = bool(self.config.actions.limit_action_change)
        return self.transition_model.engine.rollout(
            actions_mpc, obs_mu, obs_var, self.config.controller.len_horizon, iter_ctrl=self.iter_ctrl,
            limit_action_change=limit,
            max_change=self.config.actions.max_change_action_norm if limit else None,
            action_prev=am.action_model_previous_iter if limit else None, need_grad=need_grad, need_traj=need_traj,
            out=out)

    def _store_side_effects(self, out, idx):
        self.cost_traj_mean_lcb = -out["cost"][idx].cpu()
        self.states_mu_pred = out["states_mu_pred"][idx].cpu()
        self.rewards_trajectory = out["rewards_trajectory"][idx].cpu()
        self.rewards_traj_var = out["rewards_traj_var"][idx].cpu()
        self.states_var_pred = out["states_var_pred"][idx].cpu()

    def _share_side_effects(self, out, idx, owner, _unused=None):
        """Sharded runs: the five side-effect tensors of candidate `idx` of `out` (held by rank `owner` only) are packed
        into one flat buffer, broadcast, and stored on every rank -- get_action then reads the same trajectory everywhere."""
        h, e = self.config.controller.len_horizon, self.transition_model.dim_state
        sizes = [1, (h + 1) * e, h + 1, h + 1, (h + 1) * e * e]
        flat = torch.empty((sum(sizes),), dtype=torch.float64, device=self.transition_model.engine.device)
        if self.sharder.rank == owner:
            parts = [out["cost"][idx].reshape(1), out["states_mu_pred"][idx].reshape(-1),
                     out["rewards_trajectory"][idx].reshape(-1), out["rewards_traj_var"][idx].reshape(-1),
                     out["states_var_pred"][idx].reshape(-1)]
            torch.cat(parts, out=flat)
        self.sharder.broadcast(flat, owner)
        host = flat.cpu()
        c, mu, r, rv, var = torch.split(host, sizes)
        self.cost_traj_mean_lcb = -c[0].clone()
        self.states_mu_pred = mu.reshape(h + 1, e).clone()
        self.rewards_trajectory = r.clone()
        self.rewards_traj_var = rv.clone()
        self.states_var_pred = var.reshape(h + 1, e, e).clone()

    def _packed_outputs(self, batch):
        """One flat device buffer holding every output of a rollout of `batch` sequences, plus views into it in the
        engine's naming -- so that the legacy single-sequence call brings everything back with ONE device-to-host copy
        (it is called once per L-BFGS-B iteration: six separate .cpu() syncs were 40 % of its latency)."""
        key = (batch, self.config.controller.len_horizon)
        if getattr(self, "_packed_key", None) != key:
            h, e, na = self.config.controller.len_horizon, self.transition_model.dim_state, self.actions_mapper.dim_action
            shapes = [("cost", (batch,)), ("grad", (batch, h * na)), ("states_mu_pred", (batch, h + 1, e)),
                      ("states_var_pred", (batch, h + 1, e, e)), ("rewards_trajectory", (batch, h + 1)),
                      ("rewards_traj_var", (batch, h + 1)), ("actions_model", (batch, h, na))]
            total = sum(int(np.prod(sh)) for _, sh in shapes)
            flat = torch.empty(total, dtype=torch.float64, device=self.transition_model.engine.device)
            host = torch.empty(total, dtype=torch.float64).pin_memory()
            views, hviews, off = {}, {}, 0
            for name, sh in shapes:
                n = int(np.prod(sh))
                views[name] = flat[off:off + n].view(sh)
                hviews[name] = host[off:off + n].view(sh)
                off += n
            self._packed_key, self._packed = key, (flat, host, views, hviews)
        return self._packed

    def compute_mean_lcb_trajectory(self, actions_mpc, obs_mu, obs_var):
        """(H*Na,) numpy -> (float, numpy (H*Na,)): LCB of the mean trajectory cost and its gradient."""
        flat, host, views, hviews = self._packed_outputs(1)
        # inputs: one pinned staging buffer, one host-to-device copy
        e = self.transition_model.dim_state
        na_h = self.config.controller.len_horizon * self.actions_mapper.dim_action
        if getattr(self, "_in_host", None) is None or self._in_host.numel() != na_h + e + e * e:
            self._in_host = torch.empty(na_h + e + e * e, dtype=torch.float64).pin_memory()
            self._in_dev = torch.empty_like(self._in_host, device=flat.device)
        self._in_host[:na_h] = torch.as_tensor(np.asarray(actions_mpc, dtype=np.float64)).reshape(-1)
        self._in_host[na_h:na_h + e] = torch.as_tensor(obs_mu, dtype=torch.float64).reshape(-1)
        self._in_host[na_h + e:] = torch.as_tensor(obs_var, dtype=torch.float64).reshape(-1)
        self._in_dev.copy_(self._in_host, non_blocking=True)
        a = self._in_dev[:na_h].view(1, na_h)
        obs_mu, obs_var = self._in_dev[na_h:na_h + e], self._in_dev[na_h + e:].view(e, e)
        self._bind_cost()
        am = self.actions_mapper
        limit = bool(self.config.actions.limit_action_change)
        self.transition_model.engine.rollout(
            a, obs_mu, obs_var, self.config.controller.len_horizon, iter_ctrl=self.iter_ctrl, limit_action_change=limit,
            max_change=self.config.actions.max_change_action_norm if limit else None,
            action_prev=am.action_model_previous_iter if limit else None, need_grad=True, out=dict(views))
        host.copy_(flat)                                   # one synchronous device-to-host copy
        self.cost_traj_mean_lcb = -hviews["cost"][0].clone()
        self.states_mu_pred = hviews["states_mu_pred"][0].clone()
        self.rewards_trajectory = hviews["rewards_trajectory"][0].clone()
        self.rewards_traj_var = hviews["rewards_traj_var"][0].clone()
        self.states_var_pred = hviews["states_var_pred"][0].clone()
        return float(hviews["cost"][0]), hviews["grad"][0].numpy().copy()

    def compute_mean_lcb_trajectory_batch(self, actions_mpc, obs_mu, obs_var, need_grad=True, need_traj=True):
        """(B, H*Na) -> costs (B,), grads (B, H*Na) as CUDA tensors; side effects = best candidate.

        With a process group: every rank passes the SAME (B, H*Na) batch (host or device tensor; only the rank's own rows
        are copied to its GPU), scores rows [lo, hi) = self.shard, and after ONE all-gather returns the costs of ALL B
        candidates; the gradients returned are those of the rank's own rows, (hi - lo, H*Na).  self.best_candidate is the
        global arg-min on every rank.  need_traj=False skips the trajectory outputs and the side-effect tensors."""
        a = torch.as_tensor(actions_mpc)
        a = a.reshape(a.shape[0], -1)
        sh = self.sharder
        if sh is None:
            out = self._rollout(a, obs_mu, obs_var, need_grad=need_grad, need_traj=need_traj)
            best = int(torch.argmin(torch.nan_to_num(out["cost"], nan=float("inf"))).item())
            if need_traj:
                self._store_side_effects(out, best)
            self.best_candidate = best
            return out["cost"], out.get("grad")
        nb = a.shape[0]
        lo, hi = sh.bounds(nb)
        dev = self.transition_model.engine.device
        buf = sh.cost_buffer(nb, dev)
        out = {"cost": sh.local_view(buf, nb)}        # the kernel writes its slice of the gather buffer directly
        if hi > lo:
            out = self._rollout(a[lo:hi], obs_mu, obs_var, need_grad=need_grad, need_traj=need_traj, out=out)
        costs = sh.gather(buf, nb)
        best = sh.argmin(buf, nb)
        self.best_candidate, self.shard = best, (lo, hi)
        if need_traj:
            self._share_side_effects(out, best - lo, sh.owner(best, nb))
        grads = out.get("grad") if hi > lo else torch.empty((0, a.shape[1]), dtype=torch.float64, device=dev)
        return costs, grads

    def compute_cost_unnormalized(self, obs, action, obs_var=None):
        state_mu, state_var = self.observation_state_mapper.get_state(obs=obs, obs_var=obs_var, update_internals=False)
        action_model = self.actions_mapper.transform_action_raw_to_action_model(action)
        reward_mu, reward_var = self.state_reward_mapper.get_reward(state_mu, state_var, action_model)
        return -reward_mu.item(), reward_var.item()

    def get_iter_info(self):
        return self.iter_info

    def store_iter_info(self, iter_info: IterationInformation):
        for key, value in iter_info.__dict__.items():
            self.info_iters.setdefault(key, []).append(value)
