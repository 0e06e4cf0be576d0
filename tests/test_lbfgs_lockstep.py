"""CPU: the generator form of L-BFGS / strong Wolfe (rl_gp_mpc/control_objects/models/lbfgs_lockstep.py) walks the path
torch.optim.LBFGS(line_search_fn="strong_wolfe") walks -- the optimiser the reference fits its GPs with
(gp_model.py:262-277) -- and the lock-step driver hands every generator exactly its own evaluations."""
import numpy as np
import pytest
import torch

from rl_gp_mpc.control_objects.models.lbfgs_lockstep import LbfgsStrongWolfe, run_lockstep


def rosenbrock(x):
    x = np.asarray(x, dtype=np.float64)
    f = np.sum(100.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1.0 - x[:-1]) ** 2)
    g = np.zeros_like(x)
    g[:-1] = -400.0 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2.0 * (1.0 - x[:-1])
    g[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
    return float(f), g


def bumpy(x):      # non-convex, bounded below: exercises the zoom phase and the evaluation budget
    x = np.asarray(x, dtype=np.float64)
    f = np.sum(np.sin(3.0 * x) + 0.1 * x ** 2) + 0.05 * np.sum(x[1:] * x[:-1])
    g = 3.0 * np.cos(3.0 * x) + 0.2 * x
    g[1:] += 0.05 * x[:-1]
    g[:-1] += 0.05 * x[1:]
    return float(f), g


class _Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fun):
        f, g = fun(x.detach().numpy())
        ctx.save_for_backward(torch.as_tensor(g))
        return torch.tensor(f, dtype=torch.float64)

    @staticmethod
    def backward(ctx, go):
        return go * ctx.saved_tensors[0], None


def torch_path(fun, x0, lr, steps):
    x = torch.tensor(x0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.LBFGS([x], lr=lr, line_search_fn="strong_wolfe")
    calls = [0]

    def closure():
        opt.zero_grad()
        loss = _Fn.apply(x, fun)
        loss.backward()
        calls[0] += 1
        return loss
    losses, xs = [], []
    for _ in range(steps):
        losses.append(float(opt.step(closure)))
        xs.append(x.detach().numpy().copy())
    return losses, xs, calls[0]


def generator_path(fun, x0, lr, steps):
    opt = LbfgsStrongWolfe(x0, lr=lr)
    calls = 0
    losses, xs = [], []
    for _ in range(steps):
        gen = opt.step()
        try:
            x = next(gen)
            while True:
                calls += 1
                x = gen.send(fun(x))
        except StopIteration as stop:
            losses.append(stop.value)
        xs.append(opt.x.copy())
    return losses, xs, calls


@pytest.mark.parametrize("fun,x0,lr,steps", [
    (rosenbrock, [-1.2, 1.0, 0.7, -0.3, 1.5], 1.0, 6),
    (rosenbrock, [2.0, -1.0, 0.5], 0.1, 8),
    (bumpy, [0.3, -2.0, 1.7, 0.9, -0.4, 2.5, -1.1, 0.0], 1.0, 5),
    (bumpy, [4.0, 3.0], 0.5, 4),
])
def test_generator_lbfgs_walks_the_path_of_torch_lbfgs(fun, x0, lr, steps):
    lt, xt, ct = torch_path(fun, x0, lr, steps)
    lg, xg, cg = generator_path(fun, x0, lr, steps)
    assert cg == ct                                          # the same number of objective evaluations
    np.testing.assert_allclose(lg, lt, rtol=1e-9, atol=1e-12)
    for a, b in zip(xg, xt):
        np.testing.assert_allclose(a, b, rtol=1e-8, atol=1e-10)


def test_lockstep_driver_gives_every_generator_its_own_evaluations_and_exceptions():
    funs = {0: rosenbrock, 1: bumpy, 2: bumpy}
    starts = {0: [-1.2, 1.0, 0.7], 1: [0.3, -2.0, 1.7], 2: [4.0, 3.0, 1.0]}
    alone = {k: generator_path(funs[k], starts[k], 1.0, 3) for k in funs}
    rounds = []

    def fit(key):
        opt = LbfgsStrongWolfe(starts[key], lr=1.0)
        losses = []
        for _ in range(3):
            try:
                losses.append((yield from opt.step()))
            except ArithmeticError:          # thrown into generator 2 once: it ends its fit early, like a failed factorisation
                break
        return losses, opt.x.copy()

    thrown = []

    def eval_batch(points):
        rounds.append(sorted(points))
        out = {k: funs[k](x) for k, x in points.items()}
        if 2 in points and len(rounds) == 4 and not thrown:
            out[2] = ArithmeticError("not positive definite")
            thrown.append(True)
        return out

    res = run_lockstep({k: fit(k) for k in funs}, eval_batch)
    for k in (0, 1):
        np.testing.assert_allclose(res[k][0], alone[k][0], rtol=0, atol=0)
        np.testing.assert_allclose(res[k][1], alone[k][1][-1], rtol=0, atol=0)
    assert len(res[2][0]) < 3                                 # stopped by the exception
    assert rounds[0] == [0, 1, 2] and rounds[-1] != [0, 1, 2]  # generators that are done drop out of the rounds
