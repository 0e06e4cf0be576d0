"""Interface of a state-transition model as the controller sees it.

Mirrors the method set of the reference's base class (control_objects/models/abstract_model.py:5-28) so that
alternative models can be swapped in; the only implementation here is the CUDA-backed GP model."""


def _abstract(name):
    def method(self, *args, **kwargs):
        raise NotImplementedError("%s.%s must be provided by a concrete model" % (type(self).__name__, name))
    method.__name__ = name
    return method


class AbstractStateTransitionModel:
    """dim_input = dim_state + dim_action (a time input, if any, is added by the subclass)."""

    def __init__(self, config, dim_state, dim_action):
        self.dim_state, self.dim_action = dim_state, dim_action
        self.dim_input = dim_state + dim_action
        self.config = config

    prepare_inference = _abstract("prepare_inference")      # (x, y): cache everything that depends on the memory only
    predict_next_state = _abstract("predict_next_state")    # (input, input_var): one moment-matched step
    predict_trajectory = _abstract("predict_trajectory")    # (input, input_var, ...): multi-step propagation
    train = _abstract("train")                              # hyper-parameter fitting
    save_state = _abstract("save_state")
    load_state = _abstract("load_state")
