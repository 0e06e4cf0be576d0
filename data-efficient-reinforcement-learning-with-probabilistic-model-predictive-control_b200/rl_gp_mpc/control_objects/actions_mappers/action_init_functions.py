"""Initial guesses for the action optimiser (reference actions_mappers/action_init_functions.py:4-18)."""
import numpy as np


def generate_mpc_action_init_random(len_horizon, dim_action):
    return np.random.uniform(0.0, 1.0, size=len_horizon * dim_action)


def generate_mpc_action_init_frompreviousiter(actions_mpc, dim_action):
    """Shift the previous solution by one step (in place, last step repeated)."""
    actions_mpc[:-dim_action] = actions_mpc[dim_action:]
    return actions_mpc


def get_init_action_change(len_horizon, max_change_action_norm):
    u = np.random.uniform(-1.0, 1.0, size=(len_horizon, 1))
    return u * np.asarray(max_change_action_norm)[None, :]


def get_init_action(len_horizon, num_actions):
    return np.random.uniform(0.0, 1.0, size=(len_horizon, num_actions))
