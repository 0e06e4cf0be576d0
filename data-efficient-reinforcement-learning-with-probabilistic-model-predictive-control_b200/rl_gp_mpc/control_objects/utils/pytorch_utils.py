"""Straight-through clamp and normal CDF with the reference's semantics
(control_objects/utils/pytorch_utils.py:4-17).  Host-side helpers only: inside the fused CUDA rollout
the same two rules are applied in the forward and reverse kernels."""
import math

import torch


class Clamp(torch.autograd.Function):
    """clamp in the forward pass, identity in the backward pass (gradient flows at the bounds)."""

    @staticmethod
    def forward(ctx, input, min, max):
        return torch.clamp(input, min=min, max=max)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone(), None, None


def normal_cdf(x, mu, sigma):
    z = (x - mu) / (sigma * math.sqrt(2.0))
    return 0.5 * (1.0 + torch.erf(z))
