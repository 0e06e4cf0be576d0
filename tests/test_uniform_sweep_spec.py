"""CPU spec of the uniform-kernel sweep (csrc/gpmpc_uniform_impl.cuh, csrc/gpmpc_common.cuh) in numpy, checked against
the plain formula of tests/algo_spec.py (which is itself pinned to the reference goldens):

  * exp2s: exponent in table units, magic-number rounding, 2048-entry table stored PRE-BIASED (entry j carries
    -(j << 9) in its high word) so that one integer add restores the entry and applies 2^(n >> 11); cubic;
    clamp of deep underflow on the high word;
  * row factor: Eh_ij = e_i exp(kap_j + u_i . nu_j + d_i), e_i = exp(max(kap_i, -600)), d_i the residual shift;
  * tile triangle: only 64 x 64 tiles on or above the diagonal, the diagonal tile in full with half weight.
"""
import numpy as np

from tests.algo_spec import Data, step_forward

LOG, NT = 11, 2048
SCALE = 2.95463944374059701659e+03            # 2048 / ln 2
C1, C2, C3 = 3.38450771757785784290e-04, 5.72744625649513507015e-08, 6.46152867293236580665e-12
SHIFT = 6755399441055744.0
HI_MIN = 0xC13FF000


def prebiased_table():
    # float64 exp2 of j/2048 (the host rounds from long double; a last-bit difference is irrelevant here)
    t = np.exp2(np.arange(NT, dtype=np.longdouble) / NT).astype(np.float64)
    bits = t.view(np.uint64) - (np.arange(NT, dtype=np.uint64) << np.uint64(32 + 20 - LOG))
    return bits


def exp2s(t2, table_bits):
    """numpy transcription of exp2s / exp2s_x4 (plain mul+add where the device uses FMA)."""
    t2 = np.asarray(t2, np.float64).copy()
    b = t2.view(np.uint64)
    hi = (b >> np.uint64(32)).astype(np.uint64)
    hi = np.minimum(hi, np.uint64(HI_MIN))                       # unsigned min on the high word = clamp from below
    b[:] = (hi << np.uint64(32)) | (b & np.uint64(0xFFFFFFFF))
    kd = t2 + SHIFT
    n = (kd.view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32).astype(np.int64)
    f = t2 - (kd - SHIFT)
    raw = table_bits[(n & (NT - 1)).astype(np.int64)]
    add = ((n << (20 - LOG)) & 0xFFFFFFFF).astype(np.uint64) << np.uint64(32)      # 32-bit add on the high word
    tv = ((raw + add) & np.uint64(0xFFFFFFFFFFFFFFFF)).view(np.float64)
    p = ((C3 * f + C2) * f + C1) * f
    return tv * p + tv


def test_exp2s_with_prebiased_table_is_accurate_and_safe():
    tab = prebiased_table()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-700.0, 3.0, 200000), rng.uniform(-1.0, 1.0, 50000), [0.0, -708.0, 2.9957]])
    t2 = x * SCALE                                   # the exponent as the kernels hold it (table units, float64)
    got = exp2s(t2, tab)
    want = np.exp2(t2.astype(np.longdouble) / NT)    # exact value of THAT exponent: 2^(t2 / 2048)
    rel = np.abs((got.astype(np.longdouble) - want) / want).astype(np.float64)
    assert rel[x > -700].max() < 4.5e-16            # ~2 ulp without FMA (device: <= ~1.1 ulp)
    deep = exp2s(np.array([-709.0, -1e3, -1e6, -1e300]) * SCALE, tab)
    assert np.all(np.isfinite(deep)) and np.all(deep >= 0.0) and np.all(deep < 1e-300)


def uniform_sweep(d, m, s, table_bits, kap_min=-600.0):
    """S_raw_ab / s2^2 of one step for GPs that share their hyper-parameters, organised like uniform_fwd_kernel."""
    E, N = d.E, d.N
    NP = (N + 63) // 64 * 64
    il2 = d.il2[0, :E]
    nu = np.zeros((NP, E)); nu[:N] = (d.x - m)[:, :E]
    tail = np.zeros(NP); tail[:N] = (((d.x - m)[:, E:]) ** 2 * d.il2[0, E:]).sum(1)
    beta = np.zeros((NP, E)); beta[:N] = d.beta.T
    iK = np.zeros((NP, NP)); iK[:N, :N] = d.iK[0]
    Wd = 2.0 * il2
    Q = 0.5 * np.linalg.solve(s * Wd[None, :] + np.eye(E), s)
    z = nu * il2
    kap = SCALE * (-0.5 * ((nu ** 2 * il2).sum(1) + tail) + np.einsum("ie,ef,if->i", z, Q, z))   # table units
    kap[N:] = 0.0
    u = 2.0 * SCALE * (z @ Q) * il2
    c = np.maximum(kap, kap_min * SCALE)                        # row factor e_i and residual shift d_i
    e = exp2s(c, table_bits)
    resid = kap - c
    acc = np.zeros((E, E)); tr = 0.0
    nrb = NP // 64
    for I in range(nrb):
        rows = slice(64 * I, 64 * I + 64)
        r = np.zeros((64, E)); trow = np.zeros(64)
        for J in range(I, nrb):
            cols = slice(64 * J, 64 * J + 64)
            t = kap[cols][None, :] + u[rows] @ nu[cols].T + resid[rows][:, None]
            Eh = exp2s(t, table_bits)
            wgt = 0.5 if J == I else 1.0                          # diagonal tile in full with half weight
            r += wgt * (Eh @ beta[cols])
            trow += (2.0 * wgt) * (Eh * iK[rows, cols]).sum(1)
        bi = beta[rows] * e[rows][:, None]                        # the row factor, applied to the finished sums
        X = bi.T @ r                                              # X_ab = sum_i beta_a,i e_i r_b,i
        acc += X + X.T
        tr += (e[rows] * trow).sum()
    return acc - tr * np.eye(E), bool((resid != 0.0).any())


def _case(ls, obs_var, seed, N=150, E=2, Na=1):
    rng = np.random.default_rng(seed)
    D = E + Na
    x = rng.uniform(0, 1, (N, D))
    y = 0.05 * np.sin(3.0 * x @ rng.standard_normal((D, E))) + 1e-3 * rng.standard_normal((N, E))
    d = Data(x, y, np.full((E, D), ls), np.full(E, 5e-2), np.full(E, 1e-4))
    m = np.concatenate([rng.uniform(0.3, 0.7, E), rng.uniform(0, 1, Na)])
    s = obs_var * (np.eye(E) + 0.1)
    return d, m, s


import pytest


@pytest.mark.parametrize("E,Na,N", [(2, 1, 150), (1, 1, 64), (3, 2, 70), (4, 2, 129)])
def test_uniform_sweep_matches_the_pairwise_formula(E, Na, N):
    tab = prebiased_table()
    d, m, s = _case(ls=0.3, obs_var=1e-3, seed=1, N=N, E=E, Na=Na)
    ref = step_forward(d, m, s)["Sraw"]
    got, far = uniform_sweep(d, m, s, tab)
    assert not far
    got = got * d.s2[0] ** 2
    iu = np.triu_indices(d.E)
    np.testing.assert_allclose(got[iu], ref[iu], rtol=0, atol=2e-9 * max(1.0, np.abs(ref).max()))


def test_far_rows_keep_their_contribution_through_the_residual_shift():
    tab = prebiased_table()
    d, m, s = _case(ls=0.02, obs_var=4.0, seed=2)              # kap_i < -600 for some rows, yet Eh_ij ~ 1 for neighbours
    ref = step_forward(d, m, s)["Sraw"]
    got, far = uniform_sweep(d, m, s, tab)
    assert far
    got = got * d.s2[0] ** 2
    iu = np.triu_indices(d.E)
    np.testing.assert_allclose(got[iu], ref[iu], rtol=0, atol=1e-9 * max(1.0, np.abs(ref).max()))


def test_reverse_sweep_sums_over_the_tile_triangle_equal_the_full_symmetric_sums():
    """uniform_bwd_kernel sweeps w_ij = (beta_i^T Om beta_j - wbar iK_ij) Eh_ij over the tiles on or above the diagonal
    (diagonal tile in full, half weight), with the row factor e_i folded into the coefficient row vector and the trace
    weight.  Its consumers only use  g_i = rho_i + gam_i  and the (i <-> j)-symmetric contraction of xi with z, which
    must equal the full row sums and HALF the full symmetric contraction (B4 doubles it)."""
    tab = prebiased_table()
    d, m, s = _case(ls=0.3, obs_var=2e-3, seed=7, N=150, E=3, Na=1)
    E, N = d.E, d.N
    NP = (N + 63) // 64 * 64
    rng = np.random.default_rng(3)
    Om = rng.standard_normal((E, E)); Om = Om + Om.T                    # adjoint of S_raw (symmetric)
    wbar = float(np.trace(Om))
    il2 = d.il2[0, :E]
    nu = np.zeros((NP, E)); nu[:N] = (d.x - m)[:, :E]
    tail = np.zeros(NP); tail[:N] = (((d.x - m)[:, E:]) ** 2 * d.il2[0, E:]).sum(1)
    beta = np.zeros((NP, E)); beta[:N] = d.beta.T
    iK = np.zeros((NP, NP)); iK[:N, :N] = d.iK[0]
    Q = 0.5 * np.linalg.solve(s * (2.0 * il2)[None, :] + np.eye(E), s)
    z = nu * il2
    kap = SCALE * (-0.5 * ((nu ** 2 * il2).sum(1) + tail) + np.einsum("ie,ef,if->i", z, Q, z)); kap[N:] = 0.0
    u = 2.0 * SCALE * (z @ Q) * il2
    # full symmetric reference
    Eh = np.exp((kap[:, None] + kap[None, :] + u @ nu.T) / SCALE)
    w = (beta @ Om @ beta.T - wbar * iK) * Eh
    assert np.abs(w - w.T).max() <= 1e-9 * np.abs(w).max()
    g_full = w.sum(1)
    sym_full = np.einsum("ij,ik,jl->kl", w, z, z); sym_full = sym_full + sym_full.T
    # tile-triangle sweep with the row factor folded in
    e = exp2s(np.maximum(kap, -600.0 * SCALE), tab)
    resid = kap - np.maximum(kap, -600.0 * SCALE)
    rho = np.zeros(NP); gam = np.zeros(NP); xi = np.zeros((NP, E))
    nrb = NP // 64
    for I in range(nrb):
        rows = slice(64 * I, 64 * I + 64)
        p = (beta[rows] @ Om) * e[rows][:, None]                        # coefficient row vectors, row factor folded in
        wb = wbar * e[rows]
        for J in range(I, nrb):
            cols = slice(64 * J, 64 * J + 64)
            half = 0.5 if J == I else 1.0
            c = half * (p @ beta[cols].T - wb[:, None] * iK[rows, cols])
            wt = c * exp2s(kap[cols][None, :] + u[rows] @ nu[cols].T + resid[rows][:, None], tab)
            rho[rows] += wt.sum(1); gam[cols] += wt.sum(0); xi[rows] += wt @ nu[cols]
    scale = np.abs(g_full).max()
    np.testing.assert_allclose(rho + gam, g_full, rtol=0, atol=1e-10 * scale)
    X = z.T @ (xi * il2)
    np.testing.assert_allclose(2.0 * (X + X.T), sym_full, rtol=0, atol=1e-10 * np.abs(sym_full).max())
