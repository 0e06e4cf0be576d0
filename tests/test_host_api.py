"""CPU: host-side mirror of the reference interface (config classes, mappers, reward mapper, memory)."""
import numpy as np
import pytest
import torch

from oracle import gpmpc_oracle as orc
from oracle.workloads import make_workload


def test_config_defaults_and_broadcast():
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.config_classes.reward_config import RewardConfig
    from rl_gp_mpc.config_classes.total_config import Config
    assert torch.get_default_dtype() == torch.float64            # total_config.py:11
    cfg = Config()
    assert cfg.controller.len_horizon == 15 and cfg.controller.actions_optimizer_params["maxfun"] == 30
    r = RewardConfig(target_state_norm=[0.5, 0.5], weight_state=[1, 2], weight_state_terminal=[3, 4],
                     target_action_norm=[0.1], weight_action=[0.2])
    assert torch.equal(r.weight_matrix_cost, torch.diag(torch.tensor([1.0, 2.0, 0.2])))
    assert torch.equal(r.target_state_action_norm, torch.tensor([0.5, 0.5, 0.1]))
    m = ModelConfig(gp_init={"noise_covar.noise": [1e-5, 1e-5], "base_kernel.lengthscale": [0.25, 0.25],
                             "outputscale": [5e-2, 5e-2]}, include_time_model=True, init_lengthscale_time=100)
    m.extend_dimensions_params(dim_state=2, dim_input=5)
    ls = m.gp_init["base_kernel.lengthscale"]
    assert ls.shape == (2, 5) and torch.all(ls[:, :-1] == 0.25) and torch.all(ls[:, -1] == 100)
    assert m.min_std_noise.shape == (2,)


@pytest.mark.needs_reference
def test_configs_match_reference_objects():
    from oracle.ref_loader import load_reference
    ref = load_reference()
    from rl_gp_mpc.config_classes import model_config, reward_config
    kw = dict(gp_init={"noise_covar.noise": [1e-5] * 3, "base_kernel.lengthscale": [0.5, 0.6, 0.7],
                       "outputscale": [5e-2] * 3}, min_lengthscale=4e-3, max_lengthscale=10.0)
    ours = model_config.ModelConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})
    theirs = ref["model_config"].ModelConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in kw.items()})
    ours.extend_dimensions_params(3, 4)
    theirs.extend_dimensions_params(3, 4)
    for key in ("min_std_noise", "max_outputscale", "min_lengthscale", "max_lengthscale"):
        assert torch.equal(getattr(ours, key), getattr(theirs, key)), key
    for key in ours.gp_init:
        assert torch.equal(ours.gp_init[key], theirs.gp_init[key]), key
    a = reward_config.RewardConfig()
    b = ref["reward_config"].RewardConfig()
    for key in ("weight_matrix_cost", "weight_matrix_cost_terminal", "target_state_action_norm", "state_min"):
        assert torch.equal(getattr(a, key), getattr(b, key)), key


def test_action_mappers_follow_oracle():
    from rl_gp_mpc.config_classes.actions_config import ActionsConfig
    from rl_gp_mpc.control_objects.actions_mappers.derivative_action_mapper import DerivativeActionMapper
    from rl_gp_mpc.control_objects.actions_mappers.normalization_action_mapper import NormalizationActionMapper
    h, na = 6, 2
    a = torch.rand(h * na, dtype=torch.float64)
    nm = NormalizationActionMapper(np.array([-1.0, 0.0]), np.array([1.0, 4.0]), h, ActionsConfig())
    assert torch.equal(nm.transform_action_mpc_to_action_model(a), a.reshape(h, na))
    assert nm.bounds == [(0, 1)] * (h * na)
    assert torch.allclose(nm.denorm_action(nm.norm_action(np.array([0.5, 1.0]))), torch.tensor([0.5, 1.0]))
    cfg = ActionsConfig(limit_action_change=True, max_change_action_norm=[0.1, 0.2])
    dm = DerivativeActionMapper(np.array([-1.0, 0.0]), np.array([1.0, 4.0]), h, cfg)
    prev = dm.action_model_previous_iter.clone()
    a.requires_grad_(True)
    got = dm.transform_action_mpc_to_action_model(a)
    want = orc.action_mpc_to_model(a.detach(), h, True, [0.1, 0.2], prev.numpy())
    assert torch.allclose(got.detach(), want)
    got.sum().backward()                                  # straight-through clamp: gradient of the cumsum
    expect = (torch.arange(h, 0, -1, dtype=torch.float64)[:, None] * torch.tensor([0.2, 0.4])).reshape(-1)
    assert torch.allclose(a.grad, expect)


def test_reward_mapper_matches_oracle():
    from rl_gp_mpc.config_classes.reward_config import RewardConfig
    from rl_gp_mpc.control_objects.states_reward_mappers.setpoint_distance_reward_mapper import SetpointStateRewardMapper
    cfg = make_workload(E=3, Na=2, N=5, H=4, B=1, use_constraints=True)
    r = cfg["reward"]
    rc = RewardConfig(target_state_norm=list(r["target_state"]), weight_state=list(r["weight_state"]),
                      weight_state_terminal=list(r["weight_state_terminal"]), target_action_norm=list(r["target_action"]),
                      weight_action=list(r["weight_action"]), use_constraints=True, state_min=list(r["state_min"]),
                      state_max=list(r["state_max"]))
    ours, theirs = SetpointStateRewardMapper(rc), orc.OracleReward(r)
    g = torch.Generator().manual_seed(0)
    mu = torch.rand((5, 3), generator=g)
    A = torch.randn((5, 3, 3), generator=g) * 0.1
    var = A @ A.transpose(-1, -2) + 1e-3 * torch.eye(3)
    act = torch.rand((4, 2), generator=g)
    r1, v1 = ours.get_rewards_trajectory(mu, var, act)
    r2, v2 = theirs.get_rewards_trajectory(mu, var, act)
    assert torch.allclose(r1, r2, atol=1e-13) and torch.allclose(v1, v2, atol=1e-13)
    s1 = ours.get_reward(mu[0], var[0], act[0])
    s2 = theirs.get_reward(mu[0], var[0], act[0])
    assert torch.allclose(s1[0], s2[0], atol=1e-13) and torch.allclose(s1[1], s2[1], atol=1e-13)


def test_memory_gate_and_empty_contract():
    from rl_gp_mpc.config_classes.memory_config import MemoryConfig
    from rl_gp_mpc.control_objects.memories.gp_memory import Memory
    mem = Memory(MemoryConfig(min_error_prediction_state_for_memory=[1e-2, 1e-2],
                              min_prediction_state_std_for_memory=[1e-2, 1e-2], points_batch_memory=4),
                 dim_input=3, dim_state=2)
    x, y = mem.get()
    assert x.shape == (1, 3) and y.shape == (1, 2) and not x.any() and not y.any()      # gp_memory.py:109-111
    s = torch.tensor([0.1, 0.2])
    for k in range(6):                                                                    # grows past the batch size
        good = k % 2 == 0
        mem.add(s, torch.tensor([0.5]), s + 0.1, 0.0, iter_ctrl=k,
                predicted_state=s + (0.0 if good else 0.1), predicted_state_std=torch.tensor([0.1, 0.1]))
    mem.prepare_for_model()
    x, y = mem.get()
    assert x.shape == (3, 3) and torch.allclose(y, torch.full((3, 2), 0.1))              # only large-error points kept


def test_controller_constructs_and_exposes_reference_attributes():
    from rl_gp_mpc import GpMpcController
    from rl_gp_mpc.config_classes.total_config import Config
    c = GpMpcController(-np.ones(3), np.ones(3), -np.ones(1), np.ones(1), Config())
    tm = c.transition_model
    assert tm.dim_input == 4 and tm.dim_state == 3 and len(tm.models) == 3
    m = tm.models[0]
    assert m.covar_module.base_kernel.lengthscale.shape == (1, 4) and m.likelihood.noise.shape == (1,)
    m.initialize(**{"covar_module.base_kernel.lengthscale": np.full((1, 4), 0.3), "covar_module.outputscale": 0.1,
                    "likelihood.noise": np.array([2e-5])})
    assert torch.allclose(m.covar_module.base_kernel.lengthscale, torch.full((1, 4), 0.3))
    assert abs(m.covar_module.outputscale.item() - 0.1) < 1e-15
    assert set(m.state_dict()) == {"covar_module.base_kernel.lengthscale", "covar_module.outputscale", "likelihood.noise"}
    assert c.compute_cost_unnormalized(np.zeros(3), np.zeros(1))[0] > 0


def test_prepare_inference_chooses_append_or_full_refactorisation():
    """Host logic of SURVEY 8(f) N3 with a recording stand-in for the engine (no device needed): the O(N^2) append is
    taken only for 'previous training set + new rows, same hyper-parameters, room left', else a full prepare."""
    from rl_gp_mpc.config_classes.model_config import ModelConfig
    from rl_gp_mpc.control_objects.models.gp_model import GpStateTransitionModel

    class FakeEngine:
        def __init__(self):
            self.calls, self.N, self.NP = [], 0, 0

        def prepare(self, x, y, ls, s2, noise):
            self.N = len(x)
            self.NP = (self.N + 63) // 64 * 64
            self.calls.append(("prepare", self.N))

        def append(self, x_new, y_new):
            assert self.N < self.NP
            self.N += 1
            self.calls.append(("append", self.N))

        def append_room(self):
            return self.NP - self.N

    model = GpStateTransitionModel(ModelConfig(gp_init={"noise_covar.noise": [1e-4] * 2,
                                                         "base_kernel.lengthscale": [[0.75] * 3] * 2,
                                                         "outputscale": [5e-2] * 2}), dim_state=2, dim_action=1)
    eng = model._engine = FakeEngine()
    g = torch.Generator().manual_seed(0)
    x, y = torch.rand(200, 3, generator=g), torch.rand(200, 2, generator=g)
    model.prepare_inference(x[:60], y[:60])
    assert model.last_prepare_mode == "full" and eng.calls[-1] == ("prepare", 60)
    model.prepare_inference(x[:62], y[:62])                           # grew by two rows
    assert model.last_prepare_mode == "append" and eng.calls[-2:] == [("append", 61), ("append", 62)]
    model.prepare_inference(x[:62], y[:62])                           # unchanged: the reference refactorises, so do we
    assert model.last_prepare_mode == "full"
    model.prepare_inference(x[:64], y[:64])
    assert model.last_prepare_mode == "append" and eng.append_room() == 0
    model.prepare_inference(x[:65], y[:65])                           # padded size used up
    assert model.last_prepare_mode == "full" and eng.calls[-1] == ("prepare", 65)
    x2 = x.clone(); x2[3, 1] += 0.5
    model.prepare_inference(x2[:66], y[:66])                          # an old row changed
    assert model.last_prepare_mode == "full"
    model.prepare_inference(x2[:67], y[:67])
    assert model.last_prepare_mode == "append"
    model.models[0].covar_module.outputscale = 0.07                  # hyper-parameters changed (training result)
    model.prepare_inference(x2[:68], y[:68])
    assert model.last_prepare_mode == "full"
    model.incremental_updates = False
    model.prepare_inference(x2[:69], y[:69])
    assert model.last_prepare_mode == "full"
    model.incremental_updates = True
    model.prepare_inference(x2[:30], y[:30])                          # shrank
    assert model.last_prepare_mode == "full" and eng.N == 30


def test_cost_binding_follows_in_place_changes_of_the_reward_config():
    """The reference re-reads config.reward at every objective evaluation (gp_mpc_controller.py:269); the fused kernels
    get the cost description uploaded once, so the controller must notice in-place edits (tensor contents, scalar
    fields) as well as a replaced object -- its cache key covers identities, tensor version counters and scalars."""
    from rl_gp_mpc import GpMpcController
    from rl_gp_mpc.config_classes.total_config import Config
    cfg = Config()
    r = cfg.reward
    key0 = GpMpcController._cost_fingerprint(r)
    assert GpMpcController._cost_fingerprint(r) == key0                 # stable while nothing changes
    r.weight_matrix_cost[0, 0] += 1.0                                    # in-place edit of a tensor
    key1 = GpMpcController._cost_fingerprint(r)
    assert key1 != key0
    r.exploration_factor = r.exploration_factor + 0.5                    # scalar field
    key2 = GpMpcController._cost_fingerprint(r)
    assert key2 != key1
    r.target_state_action_norm = r.target_state_action_norm.clone()     # replaced tensor object
    assert GpMpcController._cost_fingerprint(r) != key2
