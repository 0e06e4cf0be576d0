"""ActionsConfig (reference config_classes/actions_config.py:4-17)."""
from .utils.functions_process_config import convert_config_lists_to_tensor


class ActionsConfig:
    def __init__(self, limit_action_change: bool = False, max_change_action_norm: "list[float]" = None):
        """limit_action_change: bound the per-step change of the normalised action;
        max_change_action_norm: that bound, per action dimension."""
        self.limit_action_change = limit_action_change
        self.max_change_action_norm = [0.05] if max_change_action_norm is None else max_change_action_norm
        convert_config_lists_to_tensor(self)
