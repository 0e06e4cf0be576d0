"""State-transition model interface (reference control_objects/models/abstract_model.py:5-28)."""
from rl_gp_mpc.config_classes.model_config import ModelConfig


class AbstractStateTransitionModel:
    def __init__(self, config: ModelConfig, dim_state, dim_action):
        self.config = config
        self.dim_state = dim_state
        self.dim_action = dim_action
        self.dim_input = dim_state + dim_action

    def predict_trajectory(self, input, input_var):
        raise NotImplementedError

    def predict_next_state(self, input, input_var):
        raise NotImplementedError

    def prepare_inference(self, x, y):
        raise NotImplementedError

    def train(self, x, y):
        raise NotImplementedError

    def save_state(self):
        raise NotImplementedError

    def load_state(self, saved_state):
        raise NotImplementedError
