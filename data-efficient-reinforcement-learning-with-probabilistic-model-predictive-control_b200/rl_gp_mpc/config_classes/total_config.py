"""Config bundle (reference config_classes/total_config.py:11,14-31).

Like the reference, importing this module makes float64 the default torch dtype (total_config.py:11):
the whole path computes in float64."""
import torch

from .actions_config import ActionsConfig
from .controller_config import ControllerConfig
from .memory_config import MemoryConfig
from .model_config import ModelConfig
from .observation_config import ObservationConfig
from .reward_config import RewardConfig
from .training_config import TrainingConfig

torch.set_default_dtype(torch.float64)


class Config:
    def __init__(self, observation_config: ObservationConfig = None, reward_config: RewardConfig = None,
                 actions_config: ActionsConfig = None, model_config: ModelConfig = None,
                 memory_config: MemoryConfig = None, training_config: TrainingConfig = None,
                 controller_config: ControllerConfig = None):
        self.observation = observation_config or ObservationConfig()
        self.reward = reward_config or RewardConfig()
        self.actions = actions_config or ActionsConfig()
        self.model = model_config or ModelConfig()
        self.memory = memory_config or MemoryConfig()
        self.training = training_config or TrainingConfig()
        self.controller = controller_config or ControllerConfig()
