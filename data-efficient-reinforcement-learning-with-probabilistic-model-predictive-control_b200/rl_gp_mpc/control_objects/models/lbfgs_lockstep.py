"""L-BFGS with strong-Wolfe line search as a GENERATOR, and a lock-step driver for several of them.

The reference fits every GP with `torch.optim.LBFGS(params, lr, line_search_fn="strong_wolfe")`
(control_objects/models/gp_model.py:262-277): an optimiser that calls its closure from deep inside its line search.
To batch the objective evaluations of the E fits into one device call per round the E optimisers
have to advance side by side, each one suspended wherever it needs a value.  A generator does exactly that:

    opt = LbfgsStrongWolfe(x0, lr)
    gen = opt.step()              # one optimiser.step(closure)
    x = next(gen)                 # the point it wants evaluated
    x = gen.send((f, g))          # ... until StopIteration, whose value is the loss at the start of the step

The arithmetic follows torch's implementation (torch/optim/lbfgs.py: two-loop recursion with unlimited-in-practice
history, cubic interpolation, bracketing and zoom phases, the same tolerances and evaluation budget per step) on float64
numpy vectors, so a fit takes the path torch's optimiser takes (tests/test_lbfgs_lockstep.py checks iterate by iterate
against torch.optim.LBFGS); there are no threads and no per-evaluation tensor bookkeeping -- the host side of a round of
E evaluations costs ~0.1 ms instead of the ~3 ms of E torch optimisers taking turns on the GIL."""
import math

import numpy as np


def _cubic_interpolate(x1, f1, g1, x2, f2, g2, bounds=None):
    """Minimiser of the cubic through two points with values and slopes, clipped to the bounds."""
    if bounds is not None:
        xmin_bound, xmax_bound = bounds
    else:
        xmin_bound, xmax_bound = (x1, x2) if x1 <= x2 else (x2, x1)
    d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
    d2_square = d1 ** 2 - g1 * g2
    if d2_square >= 0:
        d2 = math.sqrt(d2_square)
        if x1 <= x2:
            min_pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
        else:
            min_pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
        return min(max(min_pos, xmin_bound), xmax_bound)
    return (xmin_bound + xmax_bound) / 2.0


def _strong_wolfe(x, t, d, f, g, gtd, c1=1e-4, c2=0.9, tolerance_change=1e-9, max_ls=25):
    """Generator: yields trial points x + t d, receives (f, g); returns (f_new, g_new, t, evaluations)."""
    d_norm = float(np.abs(d).max())
    g = g.copy()
    f_new, g_new = yield x + t * d
    ls_func_evals = 1
    gtd_new = float(g_new.dot(d))
    t_prev, f_prev, g_prev, gtd_prev = 0.0, f, g, gtd
    done = False
    ls_iter = 0
    bracket = bracket_f = bracket_g = bracket_gtd = None
    while ls_iter < max_ls:        # bracketing phase
        if f_new > (f + c1 * t * gtd) or (ls_iter > 1 and f_new >= f_prev):
            bracket, bracket_f = [t_prev, t], [f_prev, f_new]
            bracket_g, bracket_gtd = [g_prev, g_new.copy()], [gtd_prev, gtd_new]
            break
        if abs(gtd_new) <= -c2 * gtd:
            bracket, bracket_f, bracket_g = [t], [f_new], [g_new]
            done = True
            break
        if gtd_new >= 0:
            bracket, bracket_f = [t_prev, t], [f_prev, f_new]
            bracket_g, bracket_gtd = [g_prev, g_new.copy()], [gtd_prev, gtd_new]
            break
        min_step = t + 0.01 * (t - t_prev)
        max_step = t * 10
        tmp = t
        t = _cubic_interpolate(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, bounds=(min_step, max_step))
        t_prev, f_prev, g_prev, gtd_prev = tmp, f_new, g_new.copy(), gtd_new
        f_new, g_new = yield x + t * d
        ls_func_evals += 1
        gtd_new = float(g_new.dot(d))
        ls_iter += 1
    if ls_iter == max_ls:          # budget spent while bracketing
        bracket, bracket_f, bracket_g = [0.0, t], [f, f_new], [g, g_new]
    insuf_progress = False
    low_pos, high_pos = (0, 1) if bracket_f[0] <= bracket_f[-1] else (1, 0)
    while not done and ls_iter < max_ls:   # zoom phase
        if abs(bracket[1] - bracket[0]) * d_norm < tolerance_change:
            break
        t = _cubic_interpolate(bracket[0], bracket_f[0], bracket_gtd[0], bracket[1], bracket_f[1], bracket_gtd[1])
        eps = 0.1 * (max(bracket) - min(bracket))
        if min(max(bracket) - t, t - min(bracket)) < eps:
            if insuf_progress or t >= max(bracket) or t <= min(bracket):
                t = max(bracket) - eps if abs(t - max(bracket)) < abs(t - min(bracket)) else min(bracket) + eps
                insuf_progress = False
            else:
                insuf_progress = True
        else:
            insuf_progress = False
        f_new, g_new = yield x + t * d
        ls_func_evals += 1
        gtd_new = float(g_new.dot(d))
        ls_iter += 1
        if f_new > (f + c1 * t * gtd) or f_new >= bracket_f[low_pos]:
            bracket[high_pos], bracket_f[high_pos] = t, f_new
            bracket_g[high_pos], bracket_gtd[high_pos] = g_new.copy(), gtd_new
            low_pos, high_pos = (0, 1) if bracket_f[0] <= bracket_f[1] else (1, 0)
        else:
            if abs(gtd_new) <= -c2 * gtd:
                done = True
            elif gtd_new * (bracket[high_pos] - bracket[low_pos]) >= 0:
                bracket[high_pos], bracket_f[high_pos] = bracket[low_pos], bracket_f[low_pos]
                bracket_g[high_pos], bracket_gtd[high_pos] = bracket_g[low_pos], bracket_gtd[low_pos]
            bracket[low_pos], bracket_f[low_pos] = t, f_new
            bracket_g[low_pos], bracket_gtd[low_pos] = g_new.copy(), gtd_new
    return bracket_f[low_pos], bracket_g[low_pos], bracket[low_pos], ls_func_evals


class LbfgsStrongWolfe:
    """State of one torch-style LBFGS optimiser over a float64 vector; `step()` is a generator (module docstring)."""

    def __init__(self, x0, lr=1.0, max_iter=20, max_eval=None, tolerance_grad=1e-7, tolerance_change=1e-9,
                 history_size=100):
        self.x = np.array(x0, dtype=np.float64)
        self.lr, self.max_iter = float(lr), int(max_iter)
        self.max_eval = int(max_eval) if max_eval is not None else self.max_iter * 5 // 4
        self.tolerance_grad, self.tolerance_change, self.history_size = tolerance_grad, tolerance_change, history_size
        self.func_evals = 0
        self.n_iter = 0
        self.d = self.t = self.prev_flat_grad = self.prev_loss = None
        self.old_dirs, self.old_stps, self.ro, self.H_diag = [], [], [], 1.0

    def step(self):
        f, g = yield self.x.copy()
        orig_loss = loss = float(f)
        flat_grad = np.asarray(g, dtype=np.float64)
        current_evals = 1
        self.func_evals += 1
        if np.abs(flat_grad).max() <= self.tolerance_grad:
            return orig_loss
        d, t, old_dirs, old_stps, ro, H_diag = self.d, self.t, self.old_dirs, self.old_stps, self.ro, self.H_diag
        prev_flat_grad, prev_loss = self.prev_flat_grad, self.prev_loss
        n_iter = 0
        while n_iter < self.max_iter:
            n_iter += 1
            self.n_iter += 1
            if self.n_iter == 1:
                d = -flat_grad
                old_dirs, old_stps, ro, H_diag = [], [], [], 1.0
            else:
                y = flat_grad - prev_flat_grad
                s = d * t
                ys = float(y.dot(s))
                if ys > 1e-10:
                    if len(old_dirs) == self.history_size:
                        old_dirs.pop(0)
                        old_stps.pop(0)
                        ro.pop(0)
                    old_dirs.append(y)
                    old_stps.append(s)
                    ro.append(1.0 / ys)
                    H_diag = ys / float(y.dot(y))
                num_old = len(old_dirs)
                al = [0.0] * num_old
                q = -flat_grad
                for i in range(num_old - 1, -1, -1):
                    al[i] = float(old_stps[i].dot(q)) * ro[i]
                    q = q - al[i] * old_dirs[i]
                d = q * H_diag
                for i in range(num_old):
                    be_i = float(old_dirs[i].dot(d)) * ro[i]
                    d = d + (al[i] - be_i) * old_stps[i]
            prev_flat_grad = flat_grad.copy()
            prev_loss = loss
            t = min(1.0, 1.0 / float(np.abs(flat_grad).sum())) * self.lr if self.n_iter == 1 else self.lr
            gtd = float(flat_grad.dot(d))
            if gtd > -self.tolerance_change:
                break
            loss, flat_grad, t, ls_func_evals = yield from _strong_wolfe(
                self.x, t, d, loss, flat_grad, gtd, tolerance_change=1e-9, max_ls=self.max_eval - current_evals)
            self.x = self.x + t * d
            opt_cond = np.abs(flat_grad).max() <= self.tolerance_grad
            current_evals += ls_func_evals
            self.func_evals += ls_func_evals
            if n_iter == self.max_iter or current_evals >= self.max_eval or opt_cond:
                break
            if np.abs(d * t).max() <= self.tolerance_change:
                break
            if abs(loss - prev_loss) < self.tolerance_change:
                break
        self.d, self.t, self.old_dirs, self.old_stps, self.ro, self.H_diag = d, t, old_dirs, old_stps, ro, H_diag
        self.prev_flat_grad, self.prev_loss = prev_flat_grad, prev_loss
        return orig_loss


def run_lockstep(gens, eval_batch, should_stop=None):
    """Drives generators that `yield` a point and expect `(f, g)` back, all of them side by side: every round the points
    of the generators still running go to `eval_batch({key: point})`, which returns `{key: (f, g) | Exception}`; an
    exception is thrown INTO its generator (which may handle it and go on, or end).  `gens`: {key: generator}.
    Returns {key: the generator's return value}."""
    results, pending = {}, {}
    for key, gen in gens.items():
        try:
            pending[key] = next(gen)
        except StopIteration as stop:
            results[key] = stop.value
    while pending:
        if should_stop is not None and should_stop():
            out = {key: InterruptedError("stopped") for key in pending}
        else:
            out = eval_batch(dict(pending))
        for key in list(pending):
            gen, res = gens[key], out[key]
            try:
                pending[key] = gen.throw(res) if isinstance(res, BaseException) else gen.send(res)
            except StopIteration as stop:
                results[key] = stop.value
                del pending[key]
    return results
