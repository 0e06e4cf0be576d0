"""ProcessControl: the reference's custom environment (rl_gp_mpc/envs/process_control.py:6-154), a stirred tank.

State: liquid volume v and amount of product r; observation = (level v/s, concentration r/v) with Gaussian
measurement noise.  An uncontrolled inflow (rate fi, concentration ci) and a controlled inflow (rate action[1],
concentration cr) feed the tank, action[0] is the outflow.  The physical parameters and the (unused by the
controller) set points are redrawn every `period_change` steps when `change_params` is on, to test robustness.
Same constructor arguments, attribute names, observation / action spaces and step / reset / get_obs semantics."""
import numpy as np

from .spaces import Box, Env


def _log_uniform(lo, hi):
    return float(np.exp(np.random.uniform(np.log(lo), np.log(hi))))


class ProcessControl(Env):
    metadata = {"render.modes": []}

    def __init__(self, dt=1, s_range=(9, 11), fi_range=(0., 0.2), ci_range=(0, 0.2), cr_range=(0.5, 1),
                 noise_l_prop_range=(1e-5, 1e-3), noise_co_prop_range=(1e-5, 1e-3), sp_l_range=(0.2, 0.8),
                 sp_co_range=(0.2, 0.4), change_params=True, period_change=50):
        super().__init__()
        self.name = "processcontrol"
        self.observation_space = Box(low=np.array([0, 0]), high=np.array([10, 1]), shape=(2,), dtype=np.float32)
        self.action_space = Box(low=np.array([0, 0]), high=np.array([1, 1]), shape=(2,), dtype=np.float32)
        self.reward_range = (0, 1)
        self.dt = dt
        self.s_range, self.fi_range, self.ci_range, self.cr_range = s_range, fi_range, ci_range, cr_range
        self.noise_l_prop_range, self.noise_co_prop_range = noise_l_prop_range, noise_co_prop_range
        self.sp_l_range, self.sp_co_range = sp_l_range, sp_co_range
        self.change_params, self.period_change = change_params, period_change
        self.define_params()

    def define_params(self):
        """Draws the tank surface, disturbance flow, concentrations, noise levels and set points from their ranges."""
        self.s = np.random.uniform(*self.s_range)
        self.fi = np.random.uniform(*self.fi_range)
        self.ci = np.random.uniform(*self.ci_range)
        self.cr = np.random.uniform(*self.cr_range)
        self.noise_l_prop = _log_uniform(*self.noise_l_prop_range)
        self.noise_co_prop = _log_uniform(*self.noise_co_prop_range)
        self.sp_l = np.random.uniform(*self.sp_l_range)
        self.sp_co = np.random.uniform(*self.sp_co_range)
        if hasattr(self, "v"):
            self.clip_parameters()
        print("New params value: s: %.2f,  fi: %.2f,  ci: %.2f, cr: %.2f, noise_l: %.4f, noise_co: %.4f, sp_l: %.2f, "
              "sp_co: %.2f" % (self.s, self.fi, self.ci, self.cr, self.noise_l_prop, self.noise_co_prop, self.sp_l,
                               self.sp_co))

    def step(self, action):
        out_flow, in_flow = action[0], action[1]
        d_volume = self.fi + in_flow - out_flow
        d_product = self.fi * self.ci + in_flow * self.cr - out_flow * self.r / (self.v + 1e-3)
        self.v += d_volume * self.dt
        self.r += d_product * self.dt
        self.iter += 1
        lo, hi = self.observation_space.low, self.observation_space.high
        self.v = np.clip(self.v, lo[0] * self.s, hi[0] * self.s)
        self.r = np.clip(self.r, lo[1] * self.v, hi[1] * self.v)
        reward = -((self.v / self.s - self.sp_l) ** 2 + (self.r / (self.v + 1e-6) - self.sp_co) ** 2)
        if self.change_params and self.iter % self.period_change == 0:
            self.define_params()
        return self.get_obs(), reward, 0, {}

    def reset(self, min_prop=0.3, max_prop=0.7):
        self.iter = 0
        lo, hi = self.observation_space.low, self.observation_space.high
        start = np.clip(self.observation_space.sample(), lo + min_prop * (hi - lo), lo + max_prop * (hi - lo))
        self.v = start[0] * self.s
        self.r = start[1] * self.v
        return self.get_obs()

    def get_obs(self):
        """Noisy measurement (level, concentration) of the internal state, clipped to the observation space."""
        lo, hi = self.observation_space.low, self.observation_space.high
        level = self.v / self.s
        conc = self.r / (self.v + 1e-6)
        if self.noise_l_prop != 0:
            level += np.random.normal(0, self.noise_l_prop * hi[0])
        if self.noise_co_prop != 0:
            conc += np.random.normal(0, self.noise_co_prop * hi[1])
        return np.array([np.clip(level, lo[0], hi[0]), np.clip(conc, lo[1], hi[1])])

    def render(self, mode="human", close=False):
        pass

    def clip_parameters(self, prop_level_max_after_reset=0.9):
        """After a parameter change the volume must still fit the (new) tank; product scales with it."""
        v_before = self.v
        self.v = np.clip(self.v, a_min=0, a_max=prop_level_max_after_reset * self.s * self.observation_space.high[0])
        self.r = self.r * self.v / v_before
