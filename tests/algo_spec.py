"""numpy executable SPEC of the factorised algorithm the CUDA kernels implement (tests only).

Same mathematics as the reference (gp_model.py:112-180, :60-110; reward mapper; controller
:229-285) but organised the way the device code is:

  * only the E x E state block of the input covariance is non-zero (gp_model.py:96-97), so all
    "D x D" solves collapse to E x E ones: A_a = (s + Lambda_a)^-1, Q_ab = 1/2 (s W_ab + I)^-1 s;
  * exp(k_i + k_j + maha_ij) is evaluated as ONE exponent
        t_ij = kap_i + kap_j + u_i . nu_j ,   u_i = 2 W_b Q z_a,i   (state dims only)
    so the N^2 loop is an E-term dot product + one exp + two FMAs;
  * the gradient is obtained in "forward mode per scalar": every N- or N^2-sum also emits its
    partial derivatives w.r.t. its small local parameters (m, A_a, Q_ab); the reverse sweep over
    the horizon is then pure small-matrix algebra (no N^2 recomputation, no autograd).

The test-suite checks this spec against the reference golden vectors on CPU; the CUDA kernels are
a transcription of it and are checked against the oracle and the goldens on the GPU.
"""
import math

import numpy as np
from scipy.special import erf


class Data:
    def __init__(self, x, y, lengthscale, outputscale, noise, iK=None, beta=None):
        self.x = np.asarray(x, np.float64)
        self.N, self.D = self.x.shape
        self.ls = np.asarray(lengthscale, np.float64)
        self.E = self.ls.shape[0]
        self.s2 = np.asarray(outputscale, np.float64)
        self.noise = np.asarray(noise, np.float64)
        if iK is None:
            xs = self.x[None] / self.ls[:, None, :]
            d2 = ((xs[:, :, None, :] - xs[:, None, :, :]) ** 2).sum(-1)
            K = self.s2[:, None, None] * np.exp(-0.5 * d2) + self.noise[:, None, None] * np.eye(self.N)
            iK = np.linalg.inv(K)
            beta = np.einsum("aij,ja->ai", iK, np.asarray(y, np.float64))
        self.iK, self.beta = iK, beta
        self.il2 = 1.0 / self.ls ** 2            # (E, D)


def step_forward(d, m, s):
    """One moment-matching step. Returns outputs + the local Jacobian structures."""
    E, D, N = d.E, d.D, d.N
    nu = d.x - m                                  # (N, D)
    out = dict(m=m.copy(), s=s.copy())
    h = np.zeros(E); g = np.zeros((E, D)); c = np.zeros(E); A = np.zeros((E, E, E))
    dh_dm = np.zeros((E, D)); dh_dA = np.zeros((E, E, E))
    dg_dm = np.zeros((E, E, D)); dg_dA = np.zeros((E, E, E, E))
    kk = np.zeros((E, N))
    for a in range(E):
        lam = d.ls[a, :E] ** 2
        Ca = s + np.diag(lam)
        A[a] = np.linalg.inv(Ca)
        detB = np.linalg.det(Ca) / np.prod(lam)
        c[a] = d.s2[a] / math.sqrt(detB)
        an = np.concatenate([nu[:, :E] @ A[a], nu[:, E:] * d.il2[a, E:]], 1)     # (N, D) = A_full nu
        q = (an * nu).sum(1)
        lb = d.beta[a] * np.exp(-0.5 * q)
        h[a] = lb.sum()
        g[a] = lb @ nu
        Gam = np.einsum("i,ie,id->ed", lb, nu[:, :E], nu)                        # (E, D)
        T = np.einsum("i,ie,ik,il->ekl", lb, nu[:, :E], nu[:, :E], nu[:, :E])
        dh_dm[a] = lb @ an
        dh_dA[a] = -0.5 * Gam[:, :E]
        dg_dm[a] = np.concatenate([Gam[:, :E] @ A[a], Gam[:, E:] * d.il2[a, E:]], 1)
        dg_dm[a][:, :E] -= h[a] * np.eye(E)
        dg_dA[a] = -0.5 * T
        kk[a] = math.log(d.s2[a]) - 0.5 * (nu ** 2 * d.il2[a]).sum(1)
    M = c * h
    V = np.zeros((E, D))                          # V[a, :] (reference returns V.t(): (D, E))
    for a in range(E):
        V[a, :E] = c[a] * (A[a] @ g[a, :E])
        V[a, E:] = c[a] * g[a, E:] * d.il2[a, E:]
    Sraw = np.zeros((E, E)); detR = np.ones((E, E))
    dS_dm = np.zeros((E, E, D)); dS_dQ = np.zeros((E, E, E, E))
    for a in range(E):
        for b in range(a, E):
            Wd = d.il2[a, :E] + d.il2[b, :E]
            R = s * Wd[None, :] + np.eye(E)
            Q = 0.5 * np.linalg.solve(R, s)
            detR[a, b] = np.linalg.det(R)
            za = nu * d.il2[a]; zb = nu * d.il2[b]
            kap_a = kk[a] + np.einsum("ie,ef,if->i", za[:, :E], Q, za[:, :E])
            kap_b = kk[b] + np.einsum("ie,ef,if->i", zb[:, :E], Q, zb[:, :E])
            u = 2.0 * (za[:, :E] @ Q) * d.il2[b, :E]                              # (N, E)
            t = kap_a[:, None] + kap_b[None, :] + u @ nu[:, :E].T
            cc = d.beta[a][:, None] * d.beta[b][None, :]
            if a == b:
                cc = cc - d.iK[a]
            w = cc * np.exp(t)
            Sraw[a, b] = w.sum()
            rho = w.sum(1); gam = w.sum(0)
            xi = w @ zb[:, :E]                                                     # (N, E)
            ybar = rho @ za[:, :E] + gam @ zb[:, :E]
            gm = rho @ za + gam @ zb
            gm[:E] -= 2.0 * Wd * (Q @ ybar)
            X = za[:, :E].T @ xi
            gQ = np.einsum("i,ie,if->ef", rho, za[:, :E], za[:, :E]) \
                + np.einsum("j,je,jf->ef", gam, zb[:, :E], zb[:, :E]) + X + X.T
            dS_dm[a, b] = gm; dS_dQ[a, b] = gQ
    S = np.zeros((E, E))
    for a in range(E):
        for b in range(a, E):
            S[a, b] = Sraw[a, b] / math.sqrt(detR[a, b]) - M[a] * M[b] + (d.s2[a] if a == b else 0.0)
            S[b, a] = S[a, b]
    out.update(h=h, g=g, c=c, A=A, M=M, V=V, S=S, Sraw=Sraw, detR=detR, dh_dm=dh_dm, dh_dA=dh_dA,
               dg_dm=dg_dm, dg_dA=dg_dA, dS_dm=dS_dm, dS_dQ=dS_dQ)
    return out


def step_backward(d, st, M_bar, S_bar, V_bar):
    """Adjoint of step_forward: given dL/dM (E), dL/dS (E,E), dL/dV^E (E[a], E[e]) returns (m_bar (D), s_bar (E,E))."""
    E, D = d.E, d.D
    s = st["s"]
    m_bar = np.zeros(D); s_bar = np.zeros((E, E))
    M_bar = M_bar.copy()
    # S_ab = Sraw/sqrt(detR) + .. - M_a M_b   (S symmetric: fold both triangles onto a<=b)
    for a in range(E):
        for b in range(a, E):
            sb = S_bar[a, b] + (S_bar[b, a] if a != b else 0.0)
            M_bar[a] -= sb * st["M"][b]
            M_bar[b] -= sb * st["M"][a]
            Wd = d.il2[a, :E] + d.il2[b, :E]
            R = s * Wd[None, :] + np.eye(E)
            Rinv = np.linalg.inv(R)
            Q = 0.5 * Rinv @ s
            rs = 1.0 / math.sqrt(st["detR"][a, b])
            Sraw_bar = sb * rs
            detR_bar = -0.5 * sb * st["Sraw"][a, b] * rs / st["detR"][a, b]
            m_bar += Sraw_bar * st["dS_dm"][a, b]
            Q_bar = Sraw_bar * st["dS_dQ"][a, b]
            RitQb = Rinv.T @ Q_bar
            s_bar += 0.5 * RitQb - (RitQb @ Q.T) * Wd[None, :]
            s_bar += detR_bar * st["detR"][a, b] * Rinv.T * Wd[None, :]
    for a in range(E):
        A = st["A"][a]; c = st["c"][a]
        Ag = A @ st["g"][a, :E]
        c_bar = M_bar[a] * st["h"][a] + V_bar[a] @ Ag
        h_bar = M_bar[a] * c
        g_bar = c * (A.T @ V_bar[a])                      # (E,)
        A_bar = c * np.outer(V_bar[a], st["g"][a, :E])
        m_bar += h_bar * st["dh_dm"][a] + g_bar @ st["dg_dm"][a]
        A_bar += h_bar * st["dh_dA"][a] + np.einsum("e,ekl->kl", g_bar, st["dg_dA"][a])
        s_bar += -0.5 * c_bar * c * A.T                   # c = s2 det(C)^-1/2 prod(l)
        s_bar += -A.T @ A_bar @ A.T
    return m_bar, 0.5 * (s_bar + s_bar.T)


class Cost:
    """Stage / terminal cost with adjoints (setpoint_distance_reward_mapper.py:12-68,124-149)."""

    def __init__(self, r, E, Na):
        self.E, self.Na = E, Na
        self.tgt = np.concatenate([np.asarray(r["target_state"], float), np.asarray(r["target_action"], float)])
        self.W = np.diag(np.concatenate([np.asarray(r["weight_state"], float), np.asarray(r["weight_action"], float)]))
        self.WT = np.diag(np.asarray(r["weight_state_terminal"], float))
        self.kappa = float(r["exploration_factor"])
        self.use_c = bool(r["use_constraints"]); self.clip = bool(r["clip_lower_bound_cost_to_0"])
        self.smin = np.asarray(r["state_min"], float); self.smax = np.asarray(r["state_max"], float)

    def stage(self, mu, s, a):
        """returns cost_mu, cost_var and their partials w.r.t. (mu, s, a)."""
        E = self.E
        e = np.concatenate([mu, a]) - self.tgt
        Wss = self.W[:E, :E]
        We = self.W @ e
        cmu = np.trace(s @ Wss) + e @ We
        TS = Wss @ s
        Wse = self.W[:E, :] @ e                       # (E,)  rows of W for the state block
        cvar = 2.0 * np.trace(TS @ TS) + 4.0 * Wse @ s @ Wse
        d = dict(cmu_mu=2.0 * We[:E], cmu_a=2.0 * We[E:], cmu_s=Wss.T.copy(),
                 cvar_s=4.0 * Wss.T @ s.T @ Wss.T + 4.0 * np.outer(Wse, Wse),
                 cvar_e=8.0 * self.W[:E, :].T @ (s @ Wse))
        if self.use_c:   # variance passed as sigma (quirk, :60-64)
            sig = np.diag(s)
            rt2 = math.sqrt(2.0)
            zmin = (self.smin - mu) / (sig * rt2); zmax = (self.smax - mu) / (sig * rt2)
            cmu += (0.5 * (1 + erf(zmin))).sum() + (1 - 0.5 * (1 + erf(zmax))).sum()
            pdf_min = np.exp(-zmin ** 2) / math.sqrt(math.pi); pdf_max = np.exp(-zmax ** 2) / math.sqrt(math.pi)
            d["cmu_mu"] = d["cmu_mu"] + (pdf_min - pdf_max) * (-1.0 / (sig * rt2))
            d["cmu_s"] = d["cmu_s"] + np.diag((pdf_min * (-zmin / sig)) - (pdf_max * (-zmax / sig)))
        return cmu, cvar, d

    def terminal(self, mu, s):
        e = mu - self.tgt[:self.E]
        We = self.WT @ e
        cmu = np.trace(s @ self.WT) + e @ We
        TS = self.WT @ s
        cvar = 2.0 * np.trace(TS @ TS) + 4.0 * We @ s @ We
        d = dict(cmu_mu=2.0 * We, cmu_s=self.WT.T.copy(),
                 cvar_s=4.0 * self.WT.T @ s.T @ self.WT.T + 4.0 * np.outer(We, We),
                 cvar_mu=8.0 * self.WT.T @ (s @ We))
        return cmu, cvar, d


def rollout(d, cost, actions_mpc, mu0, s0, H, iter_ctrl=0, include_time=False, limit_change=False,
            max_change=None, action_prev=None, need_grad=True):
    """LCB objective + gradient for ONE candidate (controller :229-285)."""
    E, D = d.E, d.D
    Na = cost.Na
    a2 = np.asarray(actions_mpc, float).reshape(H, Na)
    if limit_change:
        mc = np.asarray(max_change, float)
        raw = a2 * 2 * mc - mc
        raw[0] += np.asarray(action_prev, float)
        am = np.clip(np.cumsum(raw, 0), 0.0, 1.0)
    else:
        am = a2.copy()
    mus = [np.asarray(mu0, float)]; ss = [np.asarray(s0, float)]; steps = []
    for t in range(1, H + 1):
        m = np.concatenate([mus[-1], am[t - 1]] + ([[float(iter_ctrl + t - 1)]] if include_time else []))
        st = step_forward(d, m, ss[-1])
        steps.append(st)
        mus.append(mus[-1] + st["M"])
        sv = ss[-1] @ st["V"][:, :E].T          # (s @ V^E) with V^E[e, a] = V[a, e]
        ss.append(st["S"] + ss[-1] + sv + sv.T)
    r = np.zeros(H + 1); rv = np.zeros(H + 1); dstage = []
    for t in range(H):
        cm, cv, dd = cost.stage(mus[t], ss[t], am[t]); r[t] = -cm; rv[t] = cv; dstage.append(dd)
    cm, cv, dT = cost.terminal(mus[H], ss[H]); r[H] = -cm; rv[H] = cv
    ucb = r + cost.kappa * np.sqrt(rv)
    ucb_c = np.minimum(ucb, 0.0) if cost.clip else ucb
    J = -ucb_c.mean()
    res = dict(cost=J, states_mu_pred=np.stack(mus), states_var_pred=np.stack(ss), rewards_trajectory=r,
               rewards_traj_var=rv, actions_model=am)
    if not need_grad:
        return res
    # ---- reverse sweep.  J = -(1/(H+1)) sum_t (-cmu_t + kappa sqrt(cvar_t)); clamp is straight-through
    wmu = 1.0 / (H + 1)                                    # dJ/dcmu_t
    wvar = -cost.kappa / (H + 1) * 0.5 / np.sqrt(rv)       # dJ/dcvar_t
    mu_bar = wmu * dT["cmu_mu"] + wvar[H] * dT["cvar_mu"]
    s_bar = wmu * dT["cmu_s"] + wvar[H] * dT["cvar_s"]
    am_bar = np.zeros((H, Na))
    for t in range(H, 0, -1):
        st = steps[t - 1]; sp = ss[t - 1]
        # mu_t = mu_{t-1} + M ; s_t = S + s_{t-1} + s_{t-1} V^E + (..)^T
        U_bar = s_bar + s_bar.T
        V_bar_mat = sp.T @ U_bar                    # dL/dV^E[e, a]
        m_bar, s_prev_bar = step_backward(d, st, mu_bar, s_bar, V_bar_mat.T)
        s_prev_bar = s_prev_bar + s_bar + 0.5 * (U_bar @ st["V"][:, :E] + (U_bar @ st["V"][:, :E]).T)
        mu_prev_bar = mu_bar + m_bar[:E]
        am_bar[t - 1] += m_bar[E:E + Na]
        dd = dstage[t - 1]
        mu_prev_bar = mu_prev_bar + wmu * dd["cmu_mu"] + wvar[t - 1] * dd["cvar_e"][:E]
        am_bar[t - 1] += wmu * dd["cmu_a"] + wvar[t - 1] * dd["cvar_e"][E:]
        s_prev_bar = s_prev_bar + wmu * 0.5 * (dd["cmu_s"] + dd["cmu_s"].T) + wvar[t - 1] * 0.5 * (dd["cvar_s"] + dd["cvar_s"].T)
        mu_bar, s_bar = mu_prev_bar, s_prev_bar
    if limit_change:
        g = np.cumsum(am_bar[::-1], 0)[::-1] * 2 * np.asarray(max_change, float)   # clamp is straight-through
    else:
        g = am_bar
    res["grad"] = g.reshape(-1)
    return res
