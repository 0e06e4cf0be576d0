"""Stand-ins for the two gym environments the reference's examples drive (examples/pendulum/run_pendulum.py:20
'Pendulum-v0', examples/mountain_car/run_mountaincar.py:17 'MountainCarContinuous-v0') for installations without gym.

The reference takes both from gym 0.17.3 (environment.yml); these classes restate the published classic-control
dynamics of that release with the members the driver loop and the controller read (reset, step, observation_space,
action_space, close) -- no rendering.  `make(name)` returns gym's own environment when gym is importable."""
import numpy as np

from .spaces import Box, Env


class Pendulum(Env):
    """Torque-limited pendulum, angle 0 = upright.  Observation (cos th, sin th, th_dot), action = torque in [-2, 2];
    th_dot' = th_dot + (-3 g / (2 l) sin(th + pi) + 3 u / (m l^2)) dt, th' = th + th_dot' dt, |th_dot| <= 8;
    reward = -(wrap(th)^2 + 0.1 th_dot^2 + 0.001 u^2) evaluated before the step."""
    max_speed, max_torque, dt, g, m, l = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0

    def __init__(self, seed=None):
        super().__init__()
        self.name = "pendulum"
        high = np.array([1.0, 1.0, self.max_speed], dtype=np.float32)
        self.observation_space = Box(low=-high, high=high, shape=(3,), dtype=np.float32)
        self.action_space = Box(low=np.array([-self.max_torque], dtype=np.float32),
                                high=np.array([self.max_torque], dtype=np.float32), shape=(1,), dtype=np.float32)
        self.rng = np.random.default_rng(seed)
        self.state = np.zeros(2)

    def _obs(self):
        th, thdot = self.state
        return np.array([np.cos(th), np.sin(th), thdot])

    def reset(self):
        self.state = self.rng.uniform(low=[-np.pi, -1.0], high=[np.pi, 1.0])
        return self._obs()

    def step(self, action):
        th, thdot = self.state
        u = float(np.clip(np.asarray(action, dtype=np.float64).reshape(-1)[0], -self.max_torque, self.max_torque))
        wrapped = ((th + np.pi) % (2 * np.pi)) - np.pi
        cost = wrapped ** 2 + 0.1 * thdot ** 2 + 0.001 * u ** 2
        thdot = thdot + (-3 * self.g / (2 * self.l) * np.sin(th + np.pi) + 3.0 / (self.m * self.l ** 2) * u) * self.dt
        th = th + thdot * self.dt
        thdot = float(np.clip(thdot, -self.max_speed, self.max_speed))
        self.state = np.array([th, thdot])
        return self._obs(), -cost, False, {}

    def render(self, *args, **kwargs):
        return None


class MountainCarContinuous(Env):
    """Under-powered car in a valley.  Observation (position, velocity), action = force in [-1, 1];
    v' = v + 0.0015 f - 0.0025 cos(3 x), |v| <= 0.07, x' = x + v' in [-1.2, 0.6] (inelastic left wall);
    done at x >= 0.45; reward = 100 on arrival - 0.1 f^2."""
    min_position, max_position, max_speed, goal_position, goal_velocity, power = -1.2, 0.6, 0.07, 0.45, 0.0, 0.0015

    def __init__(self, seed=None):
        super().__init__()
        self.name = "mountaincar"
        self.observation_space = Box(low=np.array([self.min_position, -self.max_speed], dtype=np.float32),
                                     high=np.array([self.max_position, self.max_speed], dtype=np.float32),
                                     shape=(2,), dtype=np.float32)
        self.action_space = Box(low=np.array([-1.0], dtype=np.float32), high=np.array([1.0], dtype=np.float32),
                                shape=(1,), dtype=np.float32)
        self.rng = np.random.default_rng(seed)
        self.state = np.zeros(2)

    def reset(self):
        self.state = np.array([self.rng.uniform(-0.6, -0.4), 0.0])
        return self.state.copy()

    def step(self, action):
        x, v = self.state
        f = float(np.clip(np.asarray(action, dtype=np.float64).reshape(-1)[0], -1.0, 1.0))
        v = float(np.clip(v + f * self.power - 0.0025 * np.cos(3 * x), -self.max_speed, self.max_speed))
        x = float(np.clip(x + v, self.min_position, self.max_position))
        if x == self.min_position and v < 0:
            v = 0.0
        done = bool(x >= self.goal_position and v >= self.goal_velocity)
        reward = (100.0 if done else 0.0) - 0.1 * f ** 2
        self.state = np.array([x, v])
        return self.state.copy(), reward, done, {}

    def render(self, *args, **kwargs):
        return None


_STAND_INS = {"Pendulum-v0": Pendulum, "Pendulum-v1": Pendulum, "MountainCarContinuous-v0": MountainCarContinuous}


def make(name, seed=None):
    """gym.make(name) when gym is installed (what the reference's examples call), else the stand-in above."""
    try:                                # pragma: no cover - depends on the installation
        import gym
        return gym.make(name)
    except Exception:                   # noqa: BLE001
        return _STAND_INS[name](seed=seed)
