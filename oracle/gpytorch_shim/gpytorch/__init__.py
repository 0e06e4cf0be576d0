"""Stand-in for the `gpytorch` package -- TEST INFRASTRUCTURE ONLY.

gpytorch is a third-party dependency of the reference (environment.yml:17,
unpinned) that is not installed in this image and is not vendored under
/root/reference.  This module restates only the pieces the reference's hot
path touches so that the reference's *own* code can be imported verbatim and
run as the parity oracle (see oracle/ref_loader.py):

  * gp_model.py:387-397  ExactGP / ScaleKernel(RBFKernel(ard_num_dims)) /
                         ZeroMean / GaussianLikelihood containers
  * gp_model.py:358-374  register_constraint(..., Interval(lo, hi))
  * gp_model.py:376-383  .initialize(**{'base_kernel.lengthscale': ...,
                         'outputscale': ...}), likelihood.initialize(
                         **{'noise_covar.noise': ...}), load_state_dict
  * gp_model.py:425      covar_module(x).evaluate()
  * gp_model.py:427      likelihood.noise            shape (1,)
  * gp_model.py:189-190  base_kernel.lengthscale     shape (1, D)
                         covar_module.outputscale    shape ()
  * gp_mpc_controller.py:224  model.initialize(**{'covar_module.base_kernel.
                         lengthscale': ..., 'covar_module.outputscale': ...,
                         'likelihood.noise': ...})

Published semantics restated (gpytorch 1.x): ScaleKernel(RBFKernel) Gram
matrix K_ij = outputscale * exp(-0.5 * sum_d ((x_i-x_j)_d / l_d)^2), with an
exactly-zero distance on the diagonal (K_ii = outputscale); Interval constraint
value = lo + (hi-lo)*sigmoid(raw).  "parity unpinned": the reference has no
tests at this boundary, and the real gpytorch cannot be executed here.

Nothing under the product package imports this.
"""
import math
import sys
import types

import torch


class Interval(torch.nn.Module):
    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(lower_bound, dtype=torch.get_default_dtype()))
        self.register_buffer("upper_bound", torch.as_tensor(upper_bound, dtype=torch.get_default_dtype()))

    def transform(self, raw):
        return self.lower_bound + (self.upper_bound - self.lower_bound) * torch.sigmoid(raw)

    def inverse_transform(self, value):
        p = (value - self.lower_bound) / (self.upper_bound - self.lower_bound)
        return torch.log(p) - torch.log1p(-p)


class _Positive(torch.nn.Module):
    lower_bound = torch.tensor(0.0)
    upper_bound = torch.tensor(math.inf)

    def transform(self, raw):
        return torch.nn.functional.softplus(raw)

    def inverse_transform(self, value):
        return value + torch.log(-torch.expm1(-value))


class _Module(torch.nn.Module):
    """Minimal gpytorch.Module: raw parameters + constraints + initialize()."""

    def register_constraint(self, raw_name, constraint):
        # keep the constrained value fixed when the constraint is swapped
        old = getattr(self, raw_name + "_constraint", None)
        raw = getattr(self, raw_name)
        if old is not None:
            value = old.transform(raw.detach())
            lo, hi = constraint.lower_bound, constraint.upper_bound
            value = torch.minimum(torch.maximum(value, lo + 1e-12 * (hi - lo)), hi - 1e-12 * (hi - lo))
            with torch.no_grad():
                raw.copy_(constraint.inverse_transform(value).reshape(raw.shape))
        setattr(self, raw_name + "_constraint", constraint)

    def _set_constrained(self, raw_name, value):
        raw = getattr(self, raw_name)
        c = getattr(self, raw_name + "_constraint")
        value = torch.as_tensor(value, dtype=raw.dtype)
        with torch.no_grad():
            raw.copy_(c.inverse_transform(value).expand_as(raw) if value.numel() == 1
                      else c.inverse_transform(value).reshape(raw.shape))

    def _get_constrained(self, raw_name):
        return getattr(self, raw_name + "_constraint").transform(getattr(self, raw_name))

    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            obj = self
            parts = name.split(".")
            for p in parts[:-1]:
                obj = getattr(obj, p)
            setattr(obj, parts[-1], val)
        return self


class RBFKernel(_Module):
    def __init__(self, ard_num_dims=None):
        super().__init__()
        d = 1 if ard_num_dims is None else ard_num_dims
        self.raw_lengthscale = torch.nn.Parameter(torch.zeros(1, d))
        self.raw_lengthscale_constraint = _Positive()

    @property
    def lengthscale(self):
        return self._get_constrained("raw_lengthscale")

    @lengthscale.setter
    def lengthscale(self, value):
        self._set_constrained("raw_lengthscale", value)

    def gram(self, x):
        xs = x / self.lengthscale  # (N, D)
        diff = xs[:, None, :] - xs[None, :, :]
        sq = (diff * diff).sum(-1)
        return torch.exp(-0.5 * sq)


class _Lazy:
    def __init__(self, t):
        self._t = t

    def evaluate(self):
        return self._t

    def to_dense(self):
        return self._t


class ScaleKernel(_Module):
    def __init__(self, base_kernel):
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = torch.nn.Parameter(torch.zeros(()))
        self.raw_outputscale_constraint = _Positive()

    @property
    def outputscale(self):
        return self._get_constrained("raw_outputscale")

    @outputscale.setter
    def outputscale(self, value):
        self._set_constrained("raw_outputscale", value)

    def forward(self, x):
        return _Lazy(self.outputscale * self.base_kernel.gram(x))


class _HomoskedasticNoise(_Module):
    def __init__(self):
        super().__init__()
        self.raw_noise = torch.nn.Parameter(torch.zeros(1))
        self.raw_noise_constraint = _Positive()

    @property
    def noise(self):
        return self._get_constrained("raw_noise")

    @noise.setter
    def noise(self, value):
        self._set_constrained("raw_noise", value)


class GaussianLikelihood(_Module):
    def __init__(self):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise()

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value


class ZeroMean(_Module):
    def forward(self, x):
        return torch.zeros(x.shape[:-1], dtype=x.dtype)


class ExactGP(_Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        self.train_inputs = None if train_inputs is None else (train_inputs,)
        self.train_targets = train_targets
        self.likelihood = likelihood


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.mean = mean
        self.lazy_covariance_matrix = covariance_matrix


def _submodule(name, **attrs):
    m = types.ModuleType("gpytorch." + name)
    m.__dict__.update(attrs)
    sys.modules["gpytorch." + name] = m
    return m


models = _submodule("models", ExactGP=ExactGP)
kernels = _submodule("kernels", ScaleKernel=ScaleKernel, RBFKernel=RBFKernel)
likelihoods = _submodule("likelihoods", GaussianLikelihood=GaussianLikelihood)
means = _submodule("means", ZeroMean=ZeroMean)
constraints = _submodule("constraints", Interval=Interval)
distributions = _submodule("distributions", MultivariateNormal=MultivariateNormal)
mlls = _submodule("mlls")
