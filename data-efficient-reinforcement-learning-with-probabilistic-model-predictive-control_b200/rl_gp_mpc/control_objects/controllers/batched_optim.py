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


def minimize_box_lbfgs(fun, x0, iters, history=8, c1=1e-4, shrink=0.25, max_first_move=0.1):
    """Projected L-BFGS with one evaluation per iteration.

    Per candidate and iteration: variables sitting on a bound with the gradient pushing outward are frozen, the
    quasi-Newton direction of the free variables comes from the two-loop recursion over that candidate's own (s, y)
    history, and the trial point clamp(x + alpha d) is evaluated (ONE batched call for all candidates).  Candidates whose
    trial passes the Armijo test move there and append (s, y) (pairs with s.y <= 1e-10 |s||y| are skipped); the others
    stay, and retry next iteration with alpha shrunk.  A candidate whose step keeps failing falls back to the
    projected steepest-descent direction.  Costs `iters` + 1 evaluations; the returned point is always the best
    evaluated one (accepted steps only ever decrease the cost).
    """
    x = x0.clone()
    nb, n = x.shape
    dt, dev = x.dtype, x.device
    f, g = _clean(*fun(x))
    bad0 = ~torch.isfinite(f)
    f = torch.where(bad0, torch.full_like(f, float("inf")), f)
    S = torch.zeros((history, nb, n), dtype=dt, device=dev)
    Y = torch.zeros_like(S)
    rho = torch.zeros((history, nb), dtype=dt, device=dev)        # 0 marks an empty / skipped slot
    alpha = torch.ones((nb,), dtype=dt, device=dev)
    fails = torch.zeros((nb,), dtype=torch.int64, device=dev)
    first = torch.ones((nb,), dtype=torch.bool, device=dev)       # no accepted step yet: scale the first move
    head = 0                                                       # slot the next pair goes to (same for all)
    for _ in range(int(iters)):
        frozen = ((x <= 0.0) & (g > 0.0)) | ((x >= 1.0) & (g < 0.0))
        gf = torch.where(frozen, torch.zeros_like(g), g)
        # two-loop recursion, newest pair first; slots with rho = 0 drop out of both loops
        q = gf.clone()
        order = [(head - 1 - k) % history for k in range(history)]
        a = []
        for k in order:
            ak = rho[k] * (S[k] * q).sum(1)
            q = q - ak[:, None] * Y[k]
            a.append(ak)
        gamma = torch.ones((nb,), dtype=dt, device=dev)           # s.y / y.y of each candidate's newest pair
        found = torch.zeros((nb,), dtype=torch.bool, device=dev)
        for k in order:
            have = (rho[k] > 0) & ~found
            gamma = torch.where(have, 1.0 / (rho[k] * (Y[k] * Y[k]).sum(1)).clamp_min(1e-300), gamma)
            found = found | have
        r = gamma[:, None] * q
        for k, ak in zip(reversed(order), reversed(a)):
            bk = rho[k] * (Y[k] * r).sum(1)
            r = r + (ak - bk)[:, None] * S[k]
        d = torch.where(frozen, torch.zeros_like(r), -r)
        slope = (d * gf).sum(1)
        # not a descent direction, or the step keeps failing: projected steepest descent
        sd = (slope >= 0.0) | (fails >= 2) | ~torch.isfinite(slope)
        d = torch.where(sd[:, None], -gf, d)
        # moves without curvature information (first move, steepest descent): `max_first_move` in the largest component
        dmax = d.abs().amax(1).clamp_min(1e-300)
        a_eff = torch.where(first | sd, alpha * max_first_move / dmax, alpha)
        xt = (x + a_eff[:, None] * d).clamp_(0.0, 1.0)
        ft, gt = _clean(*fun(xt))
        step = xt - x
        # Armijo test on the projected step; the directional term is capped at 0: when the clamp removes components of d,
        # gf . step can be positive, and a trial with ft > f must never be accepted (the returned point is the best seen)
        ok = torch.isfinite(ft) & (ft <= f + c1 * (gf * step).sum(1).clamp_max(0.0)) & (step.abs().amax(1) > 0)
        s_new = torch.where(ok[:, None], step, torch.zeros_like(step))
        y_new = torch.where(ok[:, None], gt - g, torch.zeros_like(step))
        sy = (s_new * y_new).sum(1)
        keep = ok & (sy > 1e-10 * s_new.norm(dim=1) * y_new.norm(dim=1))
        # candidates that did not produce a pair keep their history aligned by writing an empty slot
        S[head] = torch.where(keep[:, None], s_new, torch.zeros_like(s_new))
        Y[head] = torch.where(keep[:, None], y_new, torch.zeros_like(y_new))
        rho[head] = torch.where(keep, 1.0 / sy.clamp_min(1e-300), torch.zeros_like(sy))
        head = (head + 1) % history
        x = torch.where(ok[:, None], xt, x)
        g = torch.where(ok[:, None], gt, g)
        f = torch.where(ok, ft, f)
        first = first & ~ok
        fails = torch.where(ok, torch.zeros_like(fails), fails + 1)
        alpha = torch.where(ok, torch.ones_like(alpha), alpha * shrink)
    return x, f
