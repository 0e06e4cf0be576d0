import torch

torch.set_default_dtype(torch.float64)  # reference: config_classes/total_config.py:11
