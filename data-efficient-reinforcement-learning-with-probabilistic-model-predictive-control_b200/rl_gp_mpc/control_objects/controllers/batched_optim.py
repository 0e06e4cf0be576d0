"""Box-constrained minimisers that advance B independent problems in lock step.

The reference optimises ONE action sequence at a time with scipy's L-BFGS-B inside a serial restart loop
(gp_mpc_controller.py:125-148, bounds [(0, 1)] * H * Na from normalization_action_mapper.py:13).  Here every
objective/gradient evaluation is one batched rollout of all B candidates, so the optimiser itself has to be batched:
all state is (B, n) tensors on the device of `x0`, no per-candidate Python.

`fun(x)` takes a (B, n) tensor inside the box [0, 1]^n and returns (cost (B,), grad (B, n)); non-finite costs are
treated as rejected trial points, non-finite gradient entries as zero.
"""
import torch


def _clean(cost, grad):
    return cost, torch.nan_to_num(grad, nan=0.0, posinf=0.0, neginf=0.0)


def minimize_box_adam(fun, x0, iters, lr=0.05, b1=0.9, b2=0.999, eps=1e-8):
    """Projected Adam: `iters` evaluations of `fun`; returns (best_x, best_cost) seen at the evaluated points.
    The last iterate is NOT evaluated here (the caller may do so with a cheaper value-only call)."""
    x = x0.clone()
    m = torch.zeros_like(x)
    v = torch.zeros_like(x)
    best_cost = torch.full((x.shape[0],), float("inf"), dtype=x.dtype, device=x.device)
    best_x = x.clone()
    for k in range(1, int(iters) + 1):
        cost, grad = _clean(*fun(x))
        better = torch.isfinite(cost) & (cost < best_cost)
        best_cost = torch.where(better, cost, best_cost)
        best_x = torch.where(better[:, None], x, best_x)
        m.mul_(b1).add_(grad, alpha=1 - b1)
        v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
        step = (m / (1 - b1 ** k)) / ((v / (1 - b2 ** k)).sqrt() + eps)
        x = (x - lr * step).clamp_(0.0, 1.0)
    return best_x, best_cost, x


def minimize_box_lbfgs(fun, x0, iters, history=8, c1=1e-4, shrink=0.25, max_first_move=0.1, cuda_graph=False, fused=False):
    """Projected L-BFGS with one evaluation per iteration.

    Per candidate and iteration: variables sitting on a bound with the gradient pushing outward are frozen, the
    quasi-Newton direction of the free variables comes from the two-loop recursion over that candidate's own (s, y)
    history, and the trial point clamp(x + alpha d) is evaluated (ONE batched call for all candidates).  Candidates whose
    trial passes the Armijo test move there and append (s, y) (pairs with s.y <= 1e-10 |s||y| are skipped); the others
    stay, and retry next iteration with alpha shrunk.  A candidate whose step keeps failing falls back to the
    projected steepest-descent direction.  Costs `iters` + 1 evaluations; the returned point is always the best
    evaluated one (accepted steps only ever decrease the cost).

    All state lives in tensors that are updated IN PLACE and the iteration has no data-dependent host control flow, so
    with cuda_graph=True (CUDA tensors; `fun` must then write its results into the same storage at every call and enqueue
    on the current stream only) the second iteration is captured into a CUDA graph and replayed: the ~50 small launches
    of the update cost one graph launch per iteration instead of ~1 ms of host time.  Falls back to eager execution when
    the capture fails.  (Measured: capture + instantiation cost ~100 ms per call, and the graph cannot be kept across
    control steps -- the training set grows every step -- so the controller does not use it.)

    fused=True (CUDA tensors): the whole update is ONE kernel per iteration (gpmpc_lbfgs_update, csrc/gpmpc_optim.cu), with
    this function as its specification; `fun` is called on the trial point only.
    """
    if fused and x0.is_cuda:
        return _minimize_box_lbfgs_fused(fun, x0, iters, history, c1, shrink, max_first_move)
    x = x0.clone()
    nb, n = x.shape
    dt, dev = x.dtype, x.device
    f, g = _clean(*fun(x))
    f, g = f.clone(), g.clone()
    f.copy_(torch.where(torch.isfinite(f), f, torch.full_like(f, float("inf"))))
    S = torch.zeros((history, nb, n), dtype=dt, device=dev)      # slot 0 = newest pair
    Y = torch.zeros_like(S)
    rho = torch.zeros((history, nb), dtype=dt, device=dev)        # 0 marks an empty / skipped slot
    alpha = torch.ones((nb,), dtype=dt, device=dev)
    fails = torch.zeros((nb,), dtype=torch.int64, device=dev)
    first = torch.ones((nb,), dtype=torch.bool, device=dev)       # no accepted step yet: scale the first move
    xt = torch.empty_like(x)

    def iteration():
        frozen = ((x <= 0.0) & (g > 0.0)) | ((x >= 1.0) & (g < 0.0))
        gf = torch.where(frozen, torch.zeros_like(g), g)
        # two-loop recursion, newest pair first; slots with rho = 0 drop out of both loops
        q = gf.clone()
        a = []
        for k in range(history):
            ak = rho[k] * (S[k] * q).sum(1)
            q = q - ak[:, None] * Y[k]
            a.append(ak)
        gamma = torch.ones((nb,), dtype=dt, device=dev)           # s.y / y.y of each candidate's newest pair
        found = torch.zeros((nb,), dtype=torch.bool, device=dev)
        for k in range(history):
            have = (rho[k] > 0) & ~found
            gamma = torch.where(have, 1.0 / (rho[k] * (Y[k] * Y[k]).sum(1)).clamp_min(1e-300), gamma)
            found = found | have
        r = gamma[:, None] * q
        for k in reversed(range(history)):
            bk = rho[k] * (Y[k] * r).sum(1)
            r = r + (a[k] - bk)[:, None] * S[k]
        d = torch.where(frozen, torch.zeros_like(r), -r)
        slope = (d * gf).sum(1)
        # not a descent direction, or the step keeps failing: projected steepest descent
        sd = (slope >= 0.0) | (fails >= 2) | ~torch.isfinite(slope)
        d = torch.where(sd[:, None], -gf, d)
        # moves without curvature information (first move, steepest descent): `max_first_move` in the largest component
        dmax = d.abs().amax(1).clamp_min(1e-300)
        a_eff = torch.where(first | sd, alpha * max_first_move / dmax, alpha)
        xt.copy_((x + a_eff[:, None] * d).clamp_(0.0, 1.0))
        ft, gt = _clean(*fun(xt))
        step = xt - x
        # Armijo test on the projected step; the directional term is capped at 0: when the clamp removes components of d,
        # gf . step can be positive, and a trial with ft > f must never be accepted (the returned point is the best seen)
        ok = torch.isfinite(ft) & (ft <= f + c1 * (gf * step).sum(1).clamp_max(0.0)) & (step.abs().amax(1) > 0)
        s_new = torch.where(ok[:, None], step, torch.zeros_like(step))
        y_new = torch.where(ok[:, None], gt - g, torch.zeros_like(step))
        sy = (s_new * y_new).sum(1)
        keep = ok & (sy > 1e-10 * s_new.norm(dim=1) * y_new.norm(dim=1))
        # the history shifts by one slot for everybody; candidates that did not produce a pair get an empty slot 0
        S[1:] = S[:-1].clone()
        Y[1:] = Y[:-1].clone()
        rho[1:] = rho[:-1].clone()
        S[0] = torch.where(keep[:, None], s_new, torch.zeros_like(s_new))
        Y[0] = torch.where(keep[:, None], y_new, torch.zeros_like(y_new))
        rho[0] = torch.where(keep, 1.0 / sy.clamp_min(1e-300), torch.zeros_like(sy))
        x.copy_(torch.where(ok[:, None], xt, x))
        g.copy_(torch.where(ok[:, None], gt, g))
        f.copy_(torch.where(ok, ft, f))
        first.copy_(first & ~ok)
        fails.copy_(torch.where(ok, torch.zeros_like(fails), fails + 1))
        alpha.copy_(torch.where(ok, torch.ones_like(alpha), alpha * shrink))

    iters = int(iters)
    done = 0
    if cuda_graph and x.is_cuda and iters >= 3:
        try:
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):        # an eager iteration off the default stream (allocator warm-up)
                iteration()
            cur.wait_stream(side)
            done = 1
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="relaxed"):
                iteration()
            for _ in range(iters - done):
                graph.replay()
            done = iters
        except Exception as exc:                 # noqa: BLE001 -- capture not possible here: finish eagerly
            import warnings
            warnings.warn("minimize_box_lbfgs: CUDA graph capture failed (%s), running eagerly" % (exc,))
            torch.cuda.synchronize(dev)
    for _ in range(iters - done):
        iteration()
    return x, f


def _minimize_box_lbfgs_fused(fun, x0, iters, history, c1, shrink, max_first_move):
    """minimize_box_lbfgs with the update done by gpmpc_lbfgs_update: per iteration one batched evaluation + one launch."""
    from rl_gp_mpc import _cabi
    x = x0.clone().contiguous()
    nb, n = x.shape
    dt, dev = x.dtype, x.device
    f, g = fun(x)
    f = torch.where(torch.isfinite(f), f, torch.full_like(f, float("inf"))).contiguous()
    g = torch.nan_to_num(g, nan=0.0, posinf=0.0, neginf=0.0).contiguous()
    S = torch.zeros((history, nb, n), dtype=dt, device=dev)
    Y = torch.zeros_like(S)
    rho = torch.zeros((history, nb), dtype=dt, device=dev)
    alpha = torch.ones((nb,), dtype=dt, device=dev)
    fails = torch.zeros((nb,), dtype=torch.int32, device=dev)
    first = torch.ones((nb,), dtype=torch.int32, device=dev)
    xt = torch.empty_like(x)
    head = 0
    _cabi.lbfgs_update(x, g, f, S, Y, rho, alpha, fails, first, xt, None, None, head, c1, shrink, max_first_move)
    for _ in range(int(iters)):
        ft, gt = fun(xt)
        _cabi.lbfgs_update(x, g, f, S, Y, rho, alpha, fails, first, xt, ft.contiguous(), gt.contiguous(), head, c1, shrink,
                           max_first_move)
        head = (head + 1) % history
    return x, f
