"""Controller interface (reference control_objects/controllers/abstract_controller.py:4-20)."""


class BaseControllerObject:
    def __init__(self, config):
        raise NotImplementedError

    def add_memory(self, obs, action, obs_new, reward, **kwargs):
        raise NotImplementedError()

    def get_action(self, obs_mu, obs_var=None):
        raise NotImplementedError()

    def get_action_random(self, obs_mu, obs_var=None):
        raise NotImplementedError()

    def train(self):
        raise NotImplementedError()
