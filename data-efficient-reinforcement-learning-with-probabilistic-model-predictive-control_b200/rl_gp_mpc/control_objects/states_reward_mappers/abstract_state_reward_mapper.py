"""Interface of a reward mapper: Gaussian state (and action) -> mean and variance of the reward
(reference states_reward_mappers/abstract_state_reward_mapper.py)."""


class AbstractStateRewardMapper:
    def __init__(self, config):
        self.config = config    # a RewardConfig

    def get_rewards_trajectory(self, states_mu, states_var, actions):
        """(H+1,E), (H+1,E,E), (H,Na) -> rewards (H+1,), reward variances (H+1,)."""
        raise NotImplementedError("trajectory reward not implemented by %s" % type(self).__name__)

    def get_reward_terminal(self, state_mu, state_var):
        raise NotImplementedError("terminal reward not implemented by %s" % type(self).__name__)

    def get_reward(self, state_mu, state_var, action):
        raise NotImplementedError("stage reward not implemented by %s" % type(self).__name__)
