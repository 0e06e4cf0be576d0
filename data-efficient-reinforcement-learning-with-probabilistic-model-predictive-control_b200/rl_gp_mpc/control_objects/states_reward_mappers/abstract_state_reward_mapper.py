"""Reward mapper base (reference states_reward_mappers/abstract_state_reward_mapper.py)."""
from rl_gp_mpc.config_classes.reward_config import RewardConfig


class AbstractStateRewardMapper:
    def __init__(self, config: RewardConfig):
        self.config = config

    def get_reward(self, state_mu, state_var, action):
        raise NotImplementedError

    def get_reward_terminal(self, state_mu, state_var):
        raise NotImplementedError

    def get_rewards_trajectory(self, states_mu, states_var, actions):
        raise NotImplementedError
