"""Hyper-parameter containers exposing the attribute paths the reference reads on its gpytorch models.

The reference holds one gpytorch ExactGP(ScaleKernel(RBFKernel(ard)), ZeroMean, GaussianLikelihood) per state
dimension (control_objects/models/gp_model.py:387-397) but, on the inference path, only touches
  model.covar_module.base_kernel.lengthscale   (1, D)      gp_model.py:189
  model.covar_module.outputscale               ()          gp_model.py:190
  model.likelihood.noise                       (1,)        gp_model.py:427
  model.initialize(**{'covar_module.base_kernel.lengthscale': .., 'covar_module.outputscale': ..,
                      'likelihood.noise': ..})             controllers/gp_mpc_controller.py:224
  model.covar_module.initialize(**{'base_kernel.lengthscale': .., 'outputscale': ..}),
  model.likelihood.initialize(**{'noise_covar.noise': ..})  gp_model.py:379-383
  register_constraint(name, Interval(lo, hi))              gp_model.py:358-374
  state_dict() / load_state_dict()                         gp_model.py:312, :376
gpytorch itself is not a dependency of this package: the kernel matrix is assembled by the CUDA prepare
kernels (csrc/gpmpc_prepare.cu).  Values are stored constrained (clamped into their Interval)."""
import torch


class Interval:
    def __init__(self, lower_bound, upper_bound):
        self.lower_bound = torch.as_tensor(lower_bound, dtype=torch.get_default_dtype())
        self.upper_bound = torch.as_tensor(upper_bound, dtype=torch.get_default_dtype())

    def clamp(self, value):
        return torch.minimum(torch.maximum(value, self.lower_bound), self.upper_bound)


class _Holder:
    """Object with named constrained parameters, `initialize(**{dotted.path: value})` and constraints."""
    _params = {}     # name -> shape

    def __init__(self):
        self._values = {}
        self._constraints = {}

    def register_constraint(self, raw_name, constraint):
        name = raw_name[4:] if raw_name.startswith("raw_") else raw_name
        self._constraints[name] = constraint
        setattr(self, "raw_%s_constraint" % name, constraint)
        if name in self._values:
            self._values[name] = constraint.clamp(self._values[name])

    def _set(self, name, value):
        v = torch.as_tensor(value, dtype=torch.get_default_dtype()).detach().clone()
        shape = self._params[name]
        v = v.reshape(shape) if v.numel() == int(torch.tensor(shape).prod()) or shape == () else v.expand(shape).clone()
        if name in self._constraints:
            v = self._constraints[name].clamp(v)
        self._values[name] = v

    def initialize(self, **kwargs):
        for path, value in kwargs.items():
            obj = self
            parts = path.split(".")
            for p in parts[:-1]:
                obj = getattr(obj, p)
            setattr(obj, parts[-1], value)
        return self


class RBFKernel(_Holder):
    def __init__(self, ard_num_dims):
        super().__init__()
        self._params = {"lengthscale": (1, ard_num_dims)}
        self.ard_num_dims = ard_num_dims
        self._set("lengthscale", torch.full((1, ard_num_dims), 0.6931471805599453))

    @property
    def lengthscale(self):
        return self._values["lengthscale"]

    @lengthscale.setter
    def lengthscale(self, value):
        self._set("lengthscale", value)


class ScaleKernel(_Holder):
    def __init__(self, base_kernel):
        super().__init__()
        self._params = {"outputscale": ()}
        self.base_kernel = base_kernel
        self._set("outputscale", 0.6931471805599453)

    @property
    def outputscale(self):
        return self._values["outputscale"]

    @outputscale.setter
    def outputscale(self, value):
        self._set("outputscale", value)


class _NoiseCovar(_Holder):
    def __init__(self):
        super().__init__()
        self._params = {"noise": (1,)}
        self._set("noise", [0.6931471805599453])

    @property
    def noise(self):
        return self._values["noise"]

    @noise.setter
    def noise(self, value):
        self._set("noise", value)


class GaussianLikelihood(_Holder):
    def __init__(self):
        super().__init__()
        self.noise_covar = _NoiseCovar()

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value

    def train(self):
        return self

    def eval(self):
        return self


class ExactGPModelMonoTask(_Holder):
    """One zero-mean GP with kernel s2 * exp(-1/2 sum_d ((x-x')_d / l_d)^2) and Gaussian noise
    (reference gp_model.py:387-397); a parameter container, inference runs in the CUDA engine."""

    def __init__(self, train_x, train_y, dim_input):
        super().__init__()
        self.train_inputs = None if train_x is None else (train_x,)
        self.train_targets = train_y
        self.likelihood = GaussianLikelihood()
        self.covar_module = ScaleKernel(RBFKernel(ard_num_dims=dim_input))

    def state_dict(self):
        return {
            "covar_module.base_kernel.lengthscale": self.covar_module.base_kernel.lengthscale.clone(),
            "covar_module.outputscale": self.covar_module.outputscale.clone(),
            "likelihood.noise": self.likelihood.noise.clone(),
        }

    def load_state_dict(self, state):
        self.initialize(**{k: v for k, v in state.items()})

    def train(self):
        return self

    def eval(self):
        return self
