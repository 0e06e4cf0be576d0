"""Starting points for the action optimiser (reference actions_mappers/action_init_functions.py:4-18)."""
import numpy as np


def generate_mpc_action_init_random(len_horizon, dim_action):
    """Uniform random flat vector of H*Na optimiser variables in [0, 1]."""
    return np.random.uniform(0.0, 1.0, size=len_horizon * dim_action)


def generate_mpc_action_init_frompreviousiter(actions_mpc, dim_action):
    """Warm start: drop the first step of the previous solution and repeat its last step (in place)."""
    tail = actions_mpc[dim_action:].copy()
    actions_mpc[:tail.size] = tail
    return actions_mpc


def get_init_action(len_horizon, num_actions):
    return np.random.uniform(0.0, 1.0, size=(len_horizon, num_actions))


def get_init_action_change(len_horizon, max_change_action_norm):
    signs = np.random.uniform(-1.0, 1.0, size=(len_horizon, 1))
    return signs * np.asarray(max_change_action_norm).reshape(1, -1)
