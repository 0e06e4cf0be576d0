"""MemoryConfig (reference config_classes/memory_config.py:4-23)."""
from .utils.functions_process_config import convert_config_lists_to_tensor


class MemoryConfig:
    def __init__(self, check_errors_for_storage: bool = True,
                 min_error_prediction_state_for_memory: "list[float]" = None,
                 min_prediction_state_std_for_memory: "list[float]" = None, points_batch_memory: int = 1500):
        self.check_errors_for_storage = check_errors_for_storage
        self.min_error_prediction_state_for_memory = [3e-4, 3e-4, 3e-4] \
            if min_error_prediction_state_for_memory is None else min_error_prediction_state_for_memory
        self.min_prediction_state_std_for_memory = [3e-3, 3e-3, 3e-3] \
            if min_prediction_state_std_for_memory is None else min_prediction_state_std_for_memory
        self.points_batch_memory = points_batch_memory
        convert_config_lists_to_tensor(self)
