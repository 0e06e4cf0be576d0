"""ctypes binding of libgpmpc.so (include/gpmpc.h) -- the drop-in boundary of the hot path.

There is NO CPU fallback: if the shared library is missing or no CUDA device is usable, every
entry point raises.  torch is used only for device memory and streams.
"""
import ctypes
import os

import torch

_LIB_PATH = os.environ.get("GPMPC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libgpmpc.so")
_lib = None

GPMPC_OK = 0
_STATUS = {-1: "bad argument", -2: "not prepared", -3: "matrix not positive definite", -4: "CUDA error",
           -5: "unsupported shape", -6: "no CUDA device"}

_SIGNATURES = {
    "gpmpc_version": (ctypes.c_int, []),
    "gpmpc_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "gpmpc_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "gpmpc_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "gpmpc_prepare": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 5 + [ctypes.c_int] * 3 + [ctypes.c_void_p]),
    "gpmpc_append": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "gpmpc_append_room": (ctypes.c_int, [ctypes.c_void_p]),
    "gpmpc_get_factorization": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "gpmpc_set_cost": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_double, ctypes.c_int, ctypes.c_void_p,
                                                              ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "gpmpc_predict_step": (ctypes.c_int, [ctypes.c_void_p] * 3 + [ctypes.c_int] * 2 + [ctypes.c_void_p] * 4),
    "gpmpc_rollout": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 6 + [ctypes.c_void_p] * 2
                      + [ctypes.c_void_p] * 7 + [ctypes.c_void_p]),
    "gpmpc_mll": (ctypes.c_int, [ctypes.c_void_p] * 4),
    "gpmpc_fit_eval": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3),
    "gpmpc_set_path": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "gpmpc_uses_uniform_path": (ctypes.c_int, [ctypes.c_void_p]),
    "gpmpc_enable_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "gpmpc_launch_count": (ctypes.c_longlong, [ctypes.c_void_p]),
    "gpmpc_last_rollout_ms": (ctypes.c_float, [ctypes.c_void_p]),
    "gpmpc_last_backward_ms": (ctypes.c_float, [ctypes.c_void_p]),
    "gpmpc_fp64_peak": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "gpmpc_lbfgs_update": (ctypes.c_int, [ctypes.c_int] * 5 + [ctypes.c_double] * 3 + [ctypes.c_void_p] * 13),
}


class GpmpcError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def exported_symbols():
    return sorted(_SIGNATURES)


def load_library():
    """Loads libgpmpc.so and declares every prototype of include/gpmpc.h; raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(_LIB_PATH):
        raise GpmpcError("CUDA extension not built: %s is missing (run `python -c 'import __graft_entry__ as g; "
                         "g.build()'` or `make -C csrc`). There is no CPU fallback." % _LIB_PATH)
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _f64(t, device, shape=None):
    t = torch.as_tensor(t, dtype=torch.float64, device=device).contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), tuple(t.shape)))
    return t


def lbfgs_update(x, g, f, S, Y, rho, alpha, fails, first, xt, ft, gt, head, c1=1e-4, shrink=0.25, max_first_move=0.1):
    """One batched projected L-BFGS update on the device (gpmpc_lbfgs_update, include/gpmpc.h); all tensors CUDA,
    contiguous; ft / gt = None on the first call (no trial point evaluated yet)."""
    lib = load_library()
    nb, n = x.shape
    with torch.cuda.device(x.device):
        rc = lib.gpmpc_lbfgs_update(int(nb), int(n), int(S.shape[0]), int(head), 0 if ft is None else 1, float(c1),
                                    float(shrink), float(max_first_move), _ptr(x), _ptr(g), _ptr(f), _ptr(S), _ptr(Y),
                                    _ptr(rho), _ptr(alpha), _ptr(fails), _ptr(first), _ptr(xt), _ptr(ft), _ptr(gt),
                                    ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    if rc != GPMPC_OK:
        raise GpmpcError("gpmpc_lbfgs_update failed: %s" % _STATUS.get(rc, rc))


def measure_fp64_peak(device_index=0):
    """Achieved float64 FLOP/s of a register-resident DFMA loop on this device (roofline denominator)."""
    out = ctypes.c_double(0.0)
    rc = load_library().gpmpc_fp64_peak(int(device_index), ctypes.byref(out))
    if rc != GPMPC_OK:
        raise GpmpcError("gpmpc_fp64_peak failed: %s" % _STATUS.get(rc, rc))
    return out.value


class Engine:
    """Thin owner of a gpmpc_handle; all tensors are float64 CUDA tensors on `device`."""

    def __init__(self, device=None):
        lib = load_library()
        if not torch.cuda.is_available():
            raise GpmpcError("no CUDA device available: the GP-MPC hot path has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = lib
        self._h = ctypes.c_void_p()
        rc = lib.gpmpc_create(ctypes.byref(self._h), self.device.index or 0)
        if rc != GPMPC_OK:
            raise GpmpcError("gpmpc_create failed: %s" % _STATUS.get(rc, rc))
        self.N = self.D = self.E = self.Na = 0

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.gpmpc_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    def _check(self, rc):
        if rc != GPMPC_OK:
            msg = self._lib.gpmpc_last_error(self._h)
            raise GpmpcError("%s (%s)" % (msg.decode() if msg else "", _STATUS.get(rc, rc)))

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- prepare_inference (gp_model.py:182-191)
    def prepare(self, x, y, lengthscale, outputscale, noise):
        dev = self.device
        x = _f64(x, dev); y = _f64(y, dev)
        N, D = x.shape
        E = y.shape[1]
        ls = _f64(lengthscale, dev, (E, D)); s2 = _f64(outputscale, dev, (E,)); nz = _f64(noise, dev, (E,))
        with torch.cuda.device(dev):
            self._check(self._lib.gpmpc_prepare(self._h, _ptr(x), _ptr(y), _ptr(ls), _ptr(s2), _ptr(nz), N, D, E,
                                                self._stream()))
        self.N, self.D, self.E = N, D, E

    def append_room(self):
        """Points that can still be appended before a full prepare() is needed (padded size)."""
        return int(self._lib.gpmpc_append_room(self._h))

    def append(self, x_new, y_new):
        """Adds one training point to the prepared factorisation in O(N^2) (new Cholesky row, block-inverse update)."""
        x_new = _f64(x_new, self.device, (self.D,)); y_new = _f64(y_new, self.device, (self.E,))
        with torch.cuda.device(self.device):
            self._check(self._lib.gpmpc_append(self._h, _ptr(x_new), _ptr(y_new), self._stream()))
        self.N += 1

    def factorization(self):
        iK = torch.empty((self.E, self.N, self.N), dtype=torch.float64, device=self.device)
        beta = torch.empty((self.E, self.N), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.gpmpc_get_factorization(self._h, _ptr(iK), _ptr(beta), self._stream()))
        return iK, beta

    def set_cost(self, target, W, WT, kappa, use_constraints=False, state_min=None, state_max=None, clip=False):
        dev = self.device
        target = _f64(target, dev)
        Na = target.numel() - self.E
        W = _f64(W, dev, (self.E + Na, self.E + Na)); WT = _f64(WT, dev, (self.E, self.E))
        smin = _f64(state_min, dev, (self.E,)) if use_constraints else None
        smax = _f64(state_max, dev, (self.E,)) if use_constraints else None
        with torch.cuda.device(dev):
            self._check(self._lib.gpmpc_set_cost(self._h, _ptr(target), _ptr(W), _ptr(WT), float(kappa),
                                                 int(bool(use_constraints)), _ptr(smin), _ptr(smax), int(bool(clip)),
                                                 Na, self._stream()))
        self.Na = Na

    # -- predict_next_state_change (gp_model.py:112-180), batched
    def predict_step(self, input_mu, input_var_block):
        dev = self.device
        mu = _f64(input_mu, dev)
        B = mu.shape[0]
        var = _f64(input_var_block, dev)
        EV = var.shape[-1]
        M = torch.empty((B, self.E), dtype=torch.float64, device=dev)
        S = torch.empty((B, self.E, self.E), dtype=torch.float64, device=dev)
        V = torch.empty((B, self.D, self.E), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            self._check(self._lib.gpmpc_predict_step(self._h, _ptr(mu), _ptr(var), B, EV, _ptr(M), _ptr(S), _ptr(V),
                                                     self._stream()))
        return M, S, V

    # -- compute_mean_lcb_trajectory (gp_mpc_controller.py:229-285), batched
    def rollout(self, actions_mpc, obs_mu, obs_var, H, iter_ctrl=0, limit_action_change=False, max_change=None,
                action_prev=None, need_grad=True, need_traj=True, out=None):
        dev = self.device
        a = _f64(actions_mpc, dev)
        B = a.shape[0]
        Na = self.Na
        a = a.reshape(B, H * Na)
        mu0 = _f64(obs_mu, dev); s0 = _f64(obs_var, dev)
        per = 1 if mu0.dim() == 2 else 0
        mc = _f64(max_change, dev, (Na,)) if limit_action_change else None
        ap = _f64(action_prev, dev, (Na,)) if limit_action_change else None
        E = self.E
        o = out if out is not None else {}
        def buf(name, shape, want=True):
            if not want:
                return None
            t = o.get(name)
            if t is None:
                t = torch.empty(shape, dtype=torch.float64, device=dev)
                o[name] = t
            return t
        cost = buf("cost", (B,))
        grad = buf("grad", (B, H * Na), need_grad)
        smu = buf("states_mu_pred", (B, H + 1, E), need_traj)
        svar = buf("states_var_pred", (B, H + 1, E, E), need_traj)
        rew = buf("rewards_trajectory", (B, H + 1), need_traj)
        rewv = buf("rewards_traj_var", (B, H + 1), need_traj)
        am = buf("actions_model", (B, H, Na), need_traj)
        with torch.cuda.device(dev):
            self._check(self._lib.gpmpc_rollout(self._h, _ptr(a), _ptr(mu0), _ptr(s0), per, B, H, Na, int(iter_ctrl),
                                                int(bool(limit_action_change)), _ptr(mc), _ptr(ap), _ptr(cost),
                                                _ptr(grad), _ptr(smu), _ptr(svar), _ptr(rew), _ptr(rewv), _ptr(am),
                                                self._stream()))
        return o

    def mll(self, y):
        """Log marginal likelihood and gradient of the GPs of the last prepare(); returns a (E, 3+D) CUDA tensor
        { LML, dLML/d outputscale, dLML/d noise, dLML/d lengthscale[0..D) }."""
        y = _f64(y, self.device, (self.N, self.E))
        out = torch.empty((self.E, 3 + self.D), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._lib.gpmpc_mll(self._h, _ptr(y), _ptr(out), self._stream()))
        return out

    def fit_eval(self, x, y, theta):
        """One objective evaluation of the hyper-parameter fit for all E GPs as ONE CUDA-graph launch (gpmpc_fit_eval):
        x (N,D), y (N,E) CUDA tensors that stay alive and unchanged during the fit, theta (E, D+2) host rows
        {lengthscale[D], outputscale, noise}.  Returns (out, info): out (E, 3+D) host tensor as mll(), info (E) int32 --
        non-zero where K + noise I of that GP is not positive definite (its row of out is meaningless)."""
        if not (x.is_cuda and y.is_cuda and x.dtype == torch.float64 and y.dtype == torch.float64
                and x.is_contiguous() and y.is_contiguous()):
            raise ValueError("fit_eval: x and y must be contiguous float64 CUDA tensors")
        N, D = x.shape
        E = y.shape[1]
        th = torch.as_tensor(theta, dtype=torch.float64, device="cpu").reshape(E, D + 2).contiguous()
        out = torch.empty((E, 3 + D), dtype=torch.float64)
        info = torch.zeros((E,), dtype=torch.int32)
        with torch.cuda.device(self.device):
            self._check(self._lib.gpmpc_fit_eval(self._h, _ptr(x), _ptr(y), th.data_ptr(), N, D, E, out.data_ptr(),
                                                 info.data_ptr(), self._stream()))
        self.N, self.D, self.E = N, D, E
        return out, info

    def set_path(self, mode):
        """0: automatic (uniform-kernel fast path when all GPs share their hyper-parameters), 1: general path."""
        self._check(self._lib.gpmpc_set_path(self._h, int(mode)))

    def uses_uniform_path(self):
        return bool(self._lib.gpmpc_uses_uniform_path(self._h))

    def enable_timing(self, on=True):
        self._lib.gpmpc_enable_timing(self._h, int(on))

    def launch_count(self):
        return int(self._lib.gpmpc_launch_count(self._h))

    def last_rollout_ms(self):
        return float(self._lib.gpmpc_last_rollout_ms(self._h))

    def last_backward_ms(self):
        return float(self._lib.gpmpc_last_backward_ms(self._h))
