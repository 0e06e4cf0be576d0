"""Two small numerical helpers with the reference's semantics (control_objects/utils/pytorch_utils.py:4-17).

Host-side only: inside the fused CUDA rollout the same two rules are applied by the forward and reverse kernels."""
import math

import torch

_SQRT2 = math.sqrt(2.0)


def normal_cdf(x, mu, sigma):
    """Phi((x - mu) / sigma) written with erf, as the reference does."""
    return 0.5 * (1.0 + torch.erf((x - mu) / (sigma * _SQRT2)))


class Clamp(torch.autograd.Function):
    """Straight-through clamp: the value is clamped, the gradient passes unchanged, so that an optimiser variable
    sitting on a bound can still move (torch.clamp would zero its gradient)."""

    @staticmethod
    def forward(ctx, input, min, max):
        return torch.clamp(input, min=min, max=max)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone(), None, None
