"""GP state-transition model with the reference's interface (control_objects/models/gp_model.py), backed by
the B200 CUDA engine (rl_gp_mpc/_cabi.py -> libgpmpc.so).

  prepare_inference         gp_model.py:182-191   -> gpmpc_prepare  (Gram + Cholesky + iK + beta on device)
  calculate_factorizations  gp_model.py:400-431   -> gpmpc_prepare / gpmpc_get_factorization
  predict_next_state_change gp_model.py:112-180   -> gpmpc_predict_step   (also batched: (B,D), (B,D,D))
  predict_trajectory        gp_model.py:60-110    -> gpmpc_rollout        (also batched: (B,H,Na))
Legacy single-input calls take and return float64 tensors with the reference's shapes; the batched forms are
additive.  There is no CPU fallback: without the extension / a CUDA device these methods raise."""
import numpy as np
import torch

from rl_gp_mpc import _cabi
from rl_gp_mpc.config_classes.model_config import ModelConfig

from .abstract_model import AbstractStateTransitionModel
from .gp_hyperparameters import ExactGPModelMonoTask, Interval


class SavedState:
    """Picklable bundle of the model (reference gp_model.py:13-36)."""

    def __init__(self, inputs, states_change, parameters, constraints_hyperparams, models=None):
        self.inputs = inputs
        self.states_change = states_change
        self.parameters = parameters
        self.constraints_hyperparams = constraints_hyperparams
        self.state_dicts_models = None if models is None else [m.state_dict() for m in models]

    def to_arrays(self):
        self.inputs = np.asarray(self.inputs)
        self.states_change = np.asarray(self.states_change)
        self.parameters = [{k: np.asarray(v) for k, v in p.items()} for p in self.parameters]
        self.constraints_hyperparams = {k: (v.numpy() if isinstance(v, torch.Tensor) else v)
                                        for k, v in self.constraints_hyperparams.items()}

    def to_tensors(self):
        self.inputs = torch.as_tensor(self.inputs)
        self.states_change = torch.as_tensor(self.states_change)
        self.parameters = [{k: torch.as_tensor(v) for k, v in p.items()} for p in self.parameters]
        self.constraints_hyperparams = {k: (torch.as_tensor(v) if isinstance(v, np.ndarray) else v)
                                        for k, v in self.constraints_hyperparams.items()}


def create_models(gp_init_dict, constraints_gp, inputs=None, targets=None, num_models=None, num_inputs=None):
    """Builds the per-state hyper-parameter containers with Interval constraints and init values
    (reference gp_model.py:318-384)."""
    if inputs is not None and targets is not None:
        num_models = len(targets[0])
        models = [ExactGPModelMonoTask(inputs, targets[:, i], len(inputs[0])) for i in range(num_models)]
    else:
        if num_models is None or num_inputs is None:
            raise ValueError("If train_inputs or train_targets are None, num_models and num_inputs must be defined")
        models = [ExactGPModelMonoTask(None, None, num_inputs) for _ in range(num_models)]
    for i, model in enumerate(models):
        if constraints_gp is not None:
            if "min_std_noise" in constraints_gp and "max_std_noise" in constraints_gp:
                model.likelihood.noise_covar.register_constraint(
                    "raw_noise", Interval(torch.as_tensor(constraints_gp["min_std_noise"])[i] ** 2,
                                          torch.as_tensor(constraints_gp["max_std_noise"])[i] ** 2))
            if "min_outputscale" in constraints_gp:
                model.covar_module.register_constraint(
                    "raw_outputscale", Interval(constraints_gp["min_outputscale"][i], constraints_gp["max_outputscale"][i]))
            if "min_lengthscale" in constraints_gp:
                model.covar_module.base_kernel.register_constraint(
                    "raw_lengthscale", Interval(constraints_gp["min_lengthscale"][i], constraints_gp["max_lengthscale"][i]))
        if isinstance(gp_init_dict, list):
            model.load_state_dict(gp_init_dict[i])
        else:
            model.likelihood.initialize(**{"noise_covar.noise": gp_init_dict["noise_covar.noise"][i]})
            model.covar_module.initialize(**{"base_kernel.lengthscale": gp_init_dict["base_kernel.lengthscale"][i],
                                             "outputscale": gp_init_dict["outputscale"][i]})
    return models


def _hyperparameters(models):
    ls = torch.stack([m.covar_module.base_kernel.lengthscale[0] for m in models])
    s2 = torch.stack([m.covar_module.outputscale.reshape(()) for m in models])
    noise = torch.stack([m.likelihood.noise.reshape(()) for m in models])
    return ls, s2, noise


def calculate_factorizations(x, y, models, engine=None):
    """iK = (K + noise I)^-1 (E,N,N) and beta = iK y (E,N) (reference gp_model.py:400-431), on the device."""
    eng = engine or _cabi.Engine()
    ls, s2, noise = _hyperparameters(models)
    eng.prepare(x, y, ls, s2, noise)
    return eng.factorization()


class GpStateTransitionModel(AbstractStateTransitionModel):
    def __init__(self, config: ModelConfig, dim_state, dim_action, device=None):
        super().__init__(config, dim_state, dim_action)
        if self.config.include_time_model:
            self.dim_input += 1
        self.config.extend_dimensions_params(dim_state=self.dim_state, dim_input=self.dim_input)
        self.models = create_models(gp_init_dict=self.config.gp_init, constraints_gp=self.config.__dict__,
                                    inputs=None, targets=None, num_models=self.dim_state, num_inputs=self.dim_input)
        for m in self.models:
            m.eval()
        self._device = device
        self._engine = None
        self._cost_key = None
        self.x_mem = self.y_mem = None
        self.incremental_updates = True     # O(N^2) append instead of a full refactorisation when the memory grew
        self.last_prepare_mode = None
        self._prep_x = self._prep_y = self._prep_hyp = None

    # ------------------------------------------------------------------ engine plumbing
    @property
    def engine(self):
        if self._engine is None:
            self._engine = _cabi.Engine(self._device)   # raises without the extension / a CUDA device
        return self._engine

    def set_cost(self, reward_config):
        """Binds the cost description used by the fused rollout (SetpointStateRewardMapper's config)."""
        c = reward_config
        self.engine.set_cost(c.target_state_action_norm, c.weight_matrix_cost, c.weight_matrix_cost_terminal,
                             c.exploration_factor, c.use_constraints, c.state_min, c.state_max,
                             c.clip_lower_bound_cost_to_0)
        self._cost_key = id(reward_config)

    def _ensure_cost(self):
        if self._cost_key is None:   # predict_trajectory alone: a zero cost keeps the fused kernel happy
            n = self.dim_state + self.dim_action
            self.engine.set_cost(torch.zeros(n), torch.zeros((n, n)), torch.zeros((self.dim_state, self.dim_state)), 0.0)
            self._cost_key = "zero"

    # ------------------------------------------------------------------ reference API
    def prepare_inference(self, inputs, state_changes):
        """Caches the factorisation of the training set in the engine (reference gp_model.py:182-191).

        The reference refactorises from scratch at every control step.  Here, when the call brings the PREVIOUS
        training set plus a few new rows under unchanged hyper-parameters -- what GpMpcController.get_action does after
        Memory.add (gp_mpc_controller.py:114-118) -- the new rows are appended in O(N^2) each (gpmpc_append) while the
        padded size allows, i.e. at most 63 times in a row before a full refactorisation.
        `incremental_updates = False` restores the reference's behaviour; `last_prepare_mode` says which path ran."""
        inputs = torch.as_tensor(inputs)
        state_changes = torch.as_tensor(state_changes)
        ls, s2, noise = _hyperparameters(self.models)
        hyp = torch.cat([ls.detach().double().flatten(), s2.detach().double().flatten(), noise.detach().double().flatten()]).cpu()
        n_old = 0 if self._prep_x is None else len(self._prep_x)
        n_new = len(inputs) - n_old
        can_append = (self.incremental_updates and self._engine is not None and n_old > 0 and 0 < n_new <= self._engine.append_room()
                      and self._prep_hyp is not None and torch.equal(hyp, self._prep_hyp)
                      and inputs.shape[1:] == self._prep_x.shape[1:]
                      and torch.equal(torch.as_tensor(inputs[:n_old], dtype=torch.float64).cpu(), self._prep_x)
                      and torch.equal(torch.as_tensor(state_changes[:n_old], dtype=torch.float64).cpu(), self._prep_y))
        self.x_mem = inputs
        self.y_mem = state_changes
        self.lengthscales, self.variances = ls, s2
        if can_append:
            try:
                for i in range(n_old, n_old + n_new):
                    self.engine.append(inputs[i], state_changes[i])
                if self.engine.N != len(inputs):
                    raise RuntimeError("engine holds %d points after the appends, expected %d" % (self.engine.N, len(inputs)))
                self.last_prepare_mode = "append"
            except Exception:
                # a failed append (Schur complement not positive, CUDA error) leaves the engine with SOME of the new rows:
                # forget the cached training set (so that a retry cannot append the same rows twice) and refactorise
                self._prep_x = self._prep_y = self._prep_hyp = None
                can_append = False
        if not can_append:
            self.engine.prepare(inputs, state_changes, ls, s2, noise)
            self.last_prepare_mode = "full"
        self._prep_x = torch.as_tensor(inputs, dtype=torch.float64).cpu().clone()
        self._prep_y = torch.as_tensor(state_changes, dtype=torch.float64).cpu().clone()
        self._prep_hyp = hyp
        self.iL = torch.diag_embed(1.0 / self.lengthscales)
        self._fact = None

    def _factorization(self):
        if self._fact is None:
            self._fact = self.engine.factorization()
        return self._fact

    @property
    def iK(self):
        return self._factorization()[0].cpu()

    @property
    def beta(self):
        return self._factorization()[1].cpu()

    def predict_next_state_change(self, input_mu, input_var):
        """(D,),(D,D) -> (M^T (1,E), S (E,E), V^T (D,E)); or batched (B,D),(B,D,D) -> (B,E),(B,E,E),(B,D,E)."""
        mu = torch.as_tensor(input_mu)
        var = torch.as_tensor(input_var)
        single = mu.dim() == 1
        if single:
            mu, var = mu[None], var[None]
        ev = self._uncertain_block(var)
        M, S, V = self.engine.predict_step(mu, var[:, :ev, :ev])
        if single:
            return M.cpu(), S[0].cpu(), V[0].cpu()
        return M, S, V

    def _uncertain_block(self, var):
        """Smallest leading block of the input covariance holding all non-zeros (E for rollouts, gp_model.py:96-97)."""
        nz = (var != 0).any(0)
        idx = torch.nonzero(nz.any(0) | nz.any(1)).flatten()
        ev = int(idx.max().item()) + 1 if idx.numel() else 1
        ev = max(ev, min(self.dim_state, 8))
        if ev > 8:
            raise _cabi.GpmpcError("input covariance has variance on more than 8 leading input dimensions")
        return ev

    def predict_trajectory(self, actions, obs_mu, obs_var, len_horizon, current_time_idx):
        """actions (H,Na) -> (H+1,E),(H+1,E,E) like the reference; actions (B,H,Na) -> (B,H+1,E),(B,H+1,E,E)."""
        self._ensure_cost()
        a = torch.as_tensor(actions)
        single = a.dim() == 2
        if single:
            a = a[None]
        out = self.engine.rollout(a.reshape(a.shape[0], -1), obs_mu, obs_var, len_horizon, iter_ctrl=current_time_idx,
                                  need_grad=False)
        if single:
            return out["states_mu_pred"][0].cpu(), out["states_var_pred"][0].cpu()
        return out["states_mu_pred"], out["states_var_pred"]

    @staticmethod
    def train(queue, saved_state, lr_train, num_iter_train, clip_grad_value, print_train=False, step_print_train=25,
              device=None, stop_event=None, lockstep=True):
        """Hyper-parameter fitting with the reference's procedure (gp_model.py:193-306): for every GP, uniform random
        re-initialisation inside the Interval bounds, LBFGS with strong-Wolfe line search on the unconstrained
        parameters (lbfgs_lockstep.LbfgsStrongWolfe: torch.optim.LBFGS's algorithm as a generator), keep the best
        negative marginal log-likelihood (per data point, as gpytorch's ExactMarginalLogLikelihood reports it) and fall
        back to the previous hyper-parameters when nothing better is found.  The objective and its gradient come from
        the device (Gram, Cholesky, K^-1, 1/2 tr((aa^T-K^-1) dK)); only the O(D) optimiser state lives on the host.
        The result goes to `queue` as a list of
        {'covar_module.base_kernel.lengthscale', 'covar_module.outputscale', 'likelihood.noise'} dicts.

        lockstep (default): the E fits advance side by side and their objective evaluations are BATCHED -- every round,
        the trial hyper-parameters of all GPs still fitting go to the device in ONE CUDA-graph launch of the
        factorisation + likelihood kernels (gpmpc_fit_eval; a GP that is done rides along with its last values).  Each GP
        sees exactly the evaluations its own optimiser asks for, so the result equals the serial procedure's
        (lockstep=False: one GP after the other with gpmpc_prepare + gpmpc_mll, E times as many device calls)."""
        import time
        from .lbfgs_lockstep import LbfgsStrongWolfe, run_lockstep
        t0 = time.time()
        saved_state.to_tensors()
        x = torch.as_tensor(saved_state.inputs, dtype=torch.float64)
        y_all = torch.as_tensor(saved_state.states_change, dtype=torch.float64)
        cons = saved_state.constraints_hyperparams
        try:
            engine = _cabi.Engine(device)
        except Exception as exc:                       # no device / no library: keep the current hyper-parameters
            print("training skipped:", exc)
            queue.put([{k: np.asarray(v) for k, v in p.items()} for p in saved_state.parameters])
            return
        n, d = x.shape
        n_gp = len(saved_state.parameters)
        x_dev = x.to(engine.device).contiguous()
        y_dev = y_all.to(engine.device).contiguous()

        def bounds(idx):
            lo = torch.cat([torch.as_tensor(cons["min_lengthscale"], dtype=torch.float64)[idx].reshape(-1),
                            torch.as_tensor(cons["min_outputscale"], dtype=torch.float64)[idx].reshape(1),
                            torch.as_tensor(cons["min_std_noise"], dtype=torch.float64)[idx].reshape(1) ** 2])
            hi = torch.cat([torch.as_tensor(cons["max_lengthscale"], dtype=torch.float64)[idx].reshape(-1),
                            torch.as_tensor(cons["max_outputscale"], dtype=torch.float64)[idx].reshape(1),
                            torch.as_tensor(cons["max_std_noise"], dtype=torch.float64)[idx].reshape(1) ** 2])
            return lo.numpy().copy(), hi.numpy().copy()

        def unpack(row):
            """(E, 3+D) row of the device result -> (-LML / n, gradient w.r.t. theta = [lengthscale (D), outputscale, noise])."""
            row = np.asarray(row, dtype=np.float64)
            return float(-row[0] / n), -np.concatenate([row[3:3 + d], row[1:3]]) / n

        def eval_one(idx, theta):
            """Objective of GP idx alone (gpmpc_prepare + gpmpc_mll with E = 1)."""
            th = torch.as_tensor(theta, dtype=torch.float64)
            y_col = y_all[:, idx:idx + 1].contiguous()
            engine.prepare(x, y_col, th[:d].reshape(1, d), th[d:d + 1], th[d + 1:d + 2])
            return unpack(engine.mll(y_col)[0].cpu().numpy())

        prev_thetas = [np.concatenate([np.asarray(p["covar_module.base_kernel.lengthscale"], dtype=np.float64).reshape(-1),
                                       np.asarray(p["covar_module.outputscale"], dtype=np.float64).reshape(1),
                                       np.asarray(p["likelihood.noise"], dtype=np.float64).reshape(1)])
                       for p in saved_state.parameters]
        last = [t.copy() for t in prev_thetas]         # the values a GP rides along with when it has nothing to evaluate
        stats = [0, 0.0]                               # rounds of batched evaluations, seconds inside the device call

        def eval_all(points):
            """One device round for all GPs: ONE CUDA-graph launch (gpmpc_fit_eval); a GP whose trial point has a non
            positive definite kernel matrix gets its error, the others their values.  GP by GP if the joint call fails."""
            try:
                thetas = np.stack([points.get(i, last[i]) for i in range(n_gp)])
                tc = time.time()
                out, info = engine.fit_eval(x_dev, y_dev, torch.as_tensor(thetas))
                stats[0] += 1
                stats[1] += time.time() - tc
                out = out.numpy()
                res = {}
                for i in points:
                    if int(info[i]) != 0 or not np.isfinite(out[i]).all():
                        res[i] = RuntimeError("prepare: K + noise*I of GP %d is not positive definite" % i)
                    else:
                        res[i] = unpack(out[i])
                        last[i] = np.array(points[i], dtype=np.float64)
                return res
            except Exception:
                res = {}
                for i in points:
                    try:
                        res[i] = eval_one(i, points[i])
                    except Exception as exc:      # noqa: BLE001 -- handed to the GP's own fit
                        res[i] = exc
                return res

        def eval_serial(points):
            res = {}
            for i in points:
                try:
                    res[i] = eval_one(i, points[i])
                except Exception as exc:          # noqa: BLE001
                    res[i] = exc
            return res

        starts = [torch.rand(d + 2, dtype=torch.float64).numpy() for _ in range(n_gp)]       # random restarts (:229-247)

        def fit(idx):
            """Generator: yields trial hyper-parameters theta of GP idx, receives (loss, d loss / d theta)."""
            lo, hi = bounds(idx)
            prev_theta = prev_thetas[idx]
            best_theta = prev_theta.copy()
            stopped = False
            try:
                try:
                    best_loss, _ = yield prev_theta
                except InterruptedError:
                    raise
                except Exception:
                    best_loss = float("inf")
                prev_loss = best_loss
                start = lo + starts[idx] * (hi - lo)
                p0 = np.clip((start - lo) / (hi - lo), 1e-6, 1 - 1e-6)
                opt = LbfgsStrongWolfe(np.log(p0) - np.log1p(-p0), lr=lr_train)
                try:
                    for it in range(num_iter_train):
                        step = opt.step()
                        try:
                            raw = next(step)
                            while True:
                                sig = 1.0 / (1.0 + np.exp(-raw))
                                loss, g_theta = yield lo + (hi - lo) * sig
                                if print_train and it % step_print_train == 0:
                                    print("Iter %d/%d - Loss: %.5f" % (it + 1, num_iter_train, loss))
                                raw = step.send((loss, g_theta * (hi - lo) * sig * (1.0 - sig)))
                        except StopIteration as done:
                            loss = done.value            # as torch's optimizer.step: the loss at the START of the step
                        if loss < best_loss:
                            best_loss = loss
                            best_theta = lo + (hi - lo) / (1.0 + np.exp(-opt.x))
                except InterruptedError:
                    raise
                except Exception as exc:               # e.g. a trial point with a non positive definite kernel matrix
                    print(exc)
                print("training - model %d - time %.2f s - loss %.5f -> %.5f - outputscale %s - lengthscales %s - noise %s" % (
                    idx, time.time() - t0, prev_loss, best_loss, best_theta[d], best_theta[:d], best_theta[d + 1]))
            except InterruptedError:                   # controller shut down: this GP keeps what it has
                stopped = True
            if stopped and np.array_equal(best_theta, prev_theta):
                return {k: np.asarray(v) for k, v in saved_state.parameters[idx].items()}
            return {"covar_module.base_kernel.lengthscale": best_theta[:d].reshape(1, d).copy(),
                    "covar_module.outputscale": np.asarray(best_theta[d]).reshape(()),
                    "likelihood.noise": best_theta[d + 1].reshape(1).copy()}

        should_stop = (lambda: stop_event.is_set()) if stop_event is not None else None
        if lockstep and n_gp > 1:
            results = run_lockstep({i: fit(i) for i in range(n_gp)}, eval_all, should_stop)
        else:
            results = {}
            for i in range(n_gp):
                results.update(run_lockstep({i: fit(i)}, eval_serial, should_stop))
                if stop_event is not None and stop_event.is_set():
                    break
        if stats[0]:
            print("training - %d rounds of batched evaluations, %.2f s in the device calls of %.2f s" % (
                stats[0], stats[1], time.time() - t0))
        out = []
        for i in range(n_gp):
            r = results.get(i)
            out.append(r if r is not None else {k: np.asarray(v) for k, v in saved_state.parameters[i].items()})
        queue.put(out)

    def save_state(self):
        return SavedState(inputs=self.x_mem, states_change=self.y_mem,
                          parameters=[m.state_dict() for m in self.models],
                          constraints_hyperparams=self.config.__dict__)
