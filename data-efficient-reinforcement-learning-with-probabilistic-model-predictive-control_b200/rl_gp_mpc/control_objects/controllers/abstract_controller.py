"""What the environment loop expects from a controller (reference control_objects/controllers/
abstract_controller.py:4-20): act on an observation, remember a transition, optionally train."""


class BaseControllerObject:
    _REQUIRED = ("add_memory", "get_action", "get_action_random", "train")

    def __init__(self, config):
        raise NotImplementedError("BaseControllerObject is an interface; use GpMpcController")

    def get_action(self, obs_mu, obs_var=None):
        """Raw action for the (un-normalised) observation."""
        raise NotImplementedError

    def get_action_random(self, obs_mu, obs_var=None):
        """Exploratory action that ignores the model."""
        raise NotImplementedError

    def add_memory(self, obs, action, obs_new, reward, **kwargs):
        """Store one transition (obs, action) -> obs_new."""
        raise NotImplementedError

    def train(self):
        """Refit the model's hyper-parameters."""
        raise NotImplementedError
