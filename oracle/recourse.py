"""CPU restatement of the control (recourse) optimisation (oracle; test infrastructure only).

PARITY STATUS: **unpinned** against the reference.  The reference solves, per design point, the NLP

    max (1/N) sum_s (1 - y_s)   s.t.  g_k(d, theta_s, F) <= M_k y_s,  0 <= y_s <= 1,  0 <= F <= 2.5

(reference pyds/utils.py:7-51, pyds/solver.py:50-55) with GAMS/CONOPT from the warm start of
run_twostage.py:25-26 and keeps whatever local optimum CONOPT returns; neither solver nor stored
solutions are available here.  Eliminating y_s gives the equivalent reduced problem

    min_{F in box}  Phi(F) = sum_s w_s max(0, max_k g_ks(F)/M_k)          (a hinge loss)

which is what this module minimises -- with the SAME algorithm the CUDA path uses (log-sum-exp
smoothing with parameter mu -> mu_min, projected Levenberg-Marquardt / Gauss-Newton steps with
forward control sensitivities), so that iterates can be compared, and with an independent scipy
SLSQP solve of the epigraph NLP as a cross-check of the optimum value.

Two groupings: ``shared`` -- one control vector per design point, shared by all theta (the
reference's two-stage formulation) -- and ``per_scenario`` -- one control vector per (d, theta).
"""
from __future__ import annotations

import numpy as np


def smoothed_hinge(c, mu):
    """phi_mu(c) = mu*log(1 + sum_k exp(c_k/mu)), its gradient p (softmax weights) -- c: (..., K)."""
    m = np.maximum(0.0, c.max(axis=-1))
    e0 = np.exp(-m / mu)
    ek = np.exp((c - m[..., None]) / mu)
    Z = e0 + ek.sum(axis=-1)
    phi = m + mu * np.log(Z)
    p = ek / Z[..., None]
    return phi, p


def group_terms(c, dc, w, mu):
    """Per group: Phi_mu, raw hinge sum, gradient (nz) and Gauss-Newton Hessian (nz, nz).
    c: (G, S, K) scaled constraints, dc: (G, S, K, nz), w: (S,)."""
    phi, p = smoothed_hinge(c, mu)
    raw = np.maximum(0.0, c.max(axis=-1))
    v = np.einsum("gsk,gskz->gsz", p, dc)                      # sum_k p_k grad c_k
    grad = np.einsum("s,gsz->gz", w, v)
    H = (np.einsum("s,gsk,gska,gskb->gab", w, p, dc, dc) - np.einsum("s,gsa,gsb->gab", w, v, v)) / mu
    return (phi * w).sum(axis=1), (raw * w).sum(axis=1), grad, H


def lm_step(gr, H, lam, z, lo, hi, mu):
    """Projected LM step for one group."""
    nz = len(z)
    free = ~(((z <= lo) & (gr > 0)) | ((z >= hi) & (gr < 0)))
    step = np.zeros(nz)
    if not free.any():
        return z.copy()
    idx = np.nonzero(free)[0]
    Hf = H[np.ix_(idx, idx)].copy()
    # H is positive semi-definite in exact arithmetic; where one constraint dominates every scenario it is zero up to
    # round-off of either sign (~1e-16 gr_i^2 / mu): a diagonal entry below 1e-10 gr_i^2 / mu counts as "no curvature";
    # those directions are dropped (the damped matrix stays positive definite and the step along them goes to the bound)
    dg = np.where(np.diag(Hf) > 1e-10 * gr[idx] ** 2 / mu, np.diag(Hf), 0.0)
    flat = dg <= 0.0
    Hf[flat, :] = 0.0; Hf[:, flat] = 0.0
    np.fill_diagonal(Hf, dg)
    floor = 1e-8 * dg.max() + 1e-300
    l = lam
    for _ in range(12):
        A = Hf + np.diag(l * (dg + floor) + 1e-30)
        try:
            L = np.linalg.cholesky(A)
            step[idx] = -np.linalg.solve(L.T, np.linalg.solve(L, gr[idx]))
            break
        except np.linalg.LinAlgError:
            l *= 10.0
    return np.minimum(np.maximum(z + step, lo), hi)


def minimize_hinge(eval_c, z0, lo, hi, w, mu0=1e-4, mu_min=1e-9, max_iter=200, lam0=1e-3, step_tol=1e-8,
                   trace=None):
    """Minimise sum_s w_s hinge_s(z) for G independent groups.

    ``eval_c(z, active)`` -> (c (G,S,K), dc (G,S,K,nz)) at the group control vectors z (G, nz); only the rows
    flagged in ``active`` need to be valid.  Returns dict(z, phi_raw, status, iters, evals).
    status: 1 = every scenario of the group feasible (Phi = 0 exactly), 2 = converged at mu_min,
    3 = iteration cap."""
    z_cur = np.array(z0, dtype=np.float64, copy=True)
    G, nz = z_cur.shape
    z_try = z_cur.copy()
    F_cur = np.full(G, np.inf)
    raw_cur = np.full(G, np.inf)
    gr = np.zeros((G, nz)); H = np.zeros((G, nz, nz))
    lam = np.full(G, lam0); mu = np.full(G, mu0)
    status = np.zeros(G, dtype=np.int32)
    first = np.ones(G, dtype=bool)
    iters = np.zeros(G, dtype=np.int32)
    span = np.where(np.isfinite(hi - lo), hi - lo, 1.0)
    evals = 0
    for _ in range(max_iter):
        act = status == 0
        if not act.any():
            break
        c, dc = eval_c(z_try, act)
        evals += 1
        for g in np.nonzero(act)[0]:
            F_t, raw_t, g_t, H_t = group_terms(c[g:g + 1], dc[g:g + 1], w, mu[g])
            F_t, raw_t, g_t, H_t = F_t[0], raw_t[0], g_t[0], H_t[0]
            iters[g] += 1
            converged_here = False
            if raw_t == 0.0:
                z_cur[g] = z_try[g]; raw_cur[g] = 0.0; status[g] = 1
                continue
            if first[g] or F_t < F_cur[g]:
                moved = np.abs(z_try[g] - z_cur[g]) / span
                small = (not first[g]) and (moved.max() <= step_tol or F_cur[g] - F_t <= 1e-13 * max(F_t, 1e-300))
                z_cur[g] = z_try[g]; F_cur[g] = F_t; raw_cur[g] = raw_t; gr[g] = g_t; H[g] = H_t
                if not first[g]:
                    lam[g] = max(lam[g] / 3.0, 1e-12)
                first[g] = False
                converged_here = small
            else:
                lam[g] = lam[g] * 4.0
                if lam[g] >= 1e12:
                    converged_here = True
            if not converged_here:
                z_new = lm_step(gr[g], H[g], lam[g], z_cur[g], lo, hi, mu[g])
                if np.all(z_new == z_cur[g]):
                    converged_here = True
                else:
                    z_try[g] = z_new
            if converged_here:
                if mu[g] <= mu_min * 1.0001:
                    status[g] = 2
                else:
                    mu[g] = max(mu[g] / 10.0, mu_min); first[g] = True; z_try[g] = z_cur[g]; lam[g] = lam0
            if trace is not None:
                trace.append((int(g), int(iters[g]), float(mu[g]), float(F_t), float(raw_t), float(lam[g])))
        status[(status == 0) & (iters >= max_iter)] = 3
    status[status == 0] = 3
    return dict(z=z_cur, phi_raw=raw_cur, status=status, iters=iters, evals=evals)


def shared_control_problem(vm, d, p, w, z0=None, **kw):
    """The reference's two-stage formulation: one control vector per design point (shared over theta).
    ``vm`` is an oracle TapeVM built on the blob's program 1."""
    d = np.atleast_2d(d); p = np.atleast_2d(p)
    n_d = len(d)
    bigm = vm.big_m
    lo, hi = vm.sec["ZLOW"][:vm.nz], vm.sec["ZHIG"][:vm.nz]
    if z0 is None:
        z0 = np.zeros((n_d, vm.nz))
        r0 = vm.evaluate(d[:1], p[:1])            # model defaults are parameter independent for real controls
    z0 = np.broadcast_to(np.asarray(z0, dtype=np.float64), (n_d, vm.nz)).copy()

    def eval_c(z, active):
        r = vm.evaluate(d, p, z=z, tol=1e-13, want_sens=True)
        return r["g"] / bigm, r["dg"] / bigm[:, None]

    out = minimize_hinge(eval_c, z0, lo, hi, np.asarray(w, dtype=np.float64), **kw)
    r = vm.evaluate(d, p, z=out["z"], tol=1e-13)
    out["g"] = r["g"]
    return out


def global_control_problem(vm, d, p, w, z0=None, **kw):
    """The reference's three-stage formulation (pyds/run.py:76-91): ONE control vector for the whole
    (d, theta) tree; objective = average over design points of the weighted hinge sum."""
    d = np.atleast_2d(d); p = np.atleast_2d(p)
    n_d, n_p = len(d), len(p)
    bigm = vm.big_m
    lo, hi = vm.sec["ZLOW"][:vm.nz], vm.sec["ZHIG"][:vm.nz]
    z0 = np.zeros((1, vm.nz)) if z0 is None else np.asarray(z0, float).reshape(1, vm.nz).copy()
    ww = np.tile(np.asarray(w, float), n_d) / n_d

    def eval_c(z, active):
        r = vm.evaluate(d, p, z=z[0], tol=1e-13, want_sens=True)
        return (r["g"] / bigm).reshape(1, n_d * n_p, vm.ng), (r["dg"] / bigm[:, None]).reshape(1, n_d * n_p, vm.ng, vm.nz)

    out = minimize_hinge(eval_c, z0, lo, hi, ww, **kw)
    out["z"] = out["z"][0]
    out["g"] = vm.evaluate(d, p, z=out["z"], tol=1e-13)["g"]
    return out


def per_scenario_problem(vm, d, p, z0=None, **kw):
    """Per-scenario recourse: every (d, theta) pair optimises its own control vector."""
    d = np.atleast_2d(d); p = np.atleast_2d(p)
    n_d, n_p = len(d), len(p)
    bigm = vm.big_m
    lo, hi = vm.sec["ZLOW"][:vm.nz], vm.sec["ZHIG"][:vm.nz]
    G = n_d * n_p
    z0 = np.zeros((G, vm.nz)) if z0 is None else np.broadcast_to(np.asarray(z0, float), (n_d, n_p, vm.nz)).reshape(G, vm.nz).copy()

    def eval_c(z, active):
        r = vm.evaluate(d, p, z=z.reshape(n_d, n_p, vm.nz), tol=1e-13, want_sens=True)
        return (r["g"] / bigm).reshape(G, 1, vm.ng), (r["dg"] / bigm[:, None]).reshape(G, 1, vm.ng, vm.nz)

    out = minimize_hinge(eval_c, z0, lo, hi, np.ones(1), **kw)
    out["z"] = out["z"].reshape(n_d, n_p, vm.nz)
    out["phi_raw"] = out["phi_raw"].reshape(n_d, n_p)
    out["status"] = out["status"].reshape(n_d, n_p)
    return out


def slsqp_shared(vm, d_row, p, w, z0, maxiter=300):
    """Independent cross-check: scipy SLSQP on the epigraph NLP  min sum w_s y_s  s.t. g_ks/M_k <= y_s, y >= 0."""
    from scipy.optimize import minimize
    p = np.atleast_2d(p)
    n_p, nz, ng = len(p), vm.nz, vm.ng
    lo, hi = vm.sec["ZLOW"][:nz], vm.sec["ZHIG"][:nz]
    bigm = vm.big_m
    w = np.asarray(w, float)
    cache = {}

    def ev(z):
        key = z.tobytes()
        if key not in cache:
            cache.clear()
            r = vm.evaluate(d_row[None, :], p, z=z, tol=1e-13, want_sens=True)
            cache[key] = (r["g"][0] / bigm, r["dg"][0] / bigm[:, None])
        return cache[key]

    r0 = ev(np.asarray(z0, float))
    y0 = np.maximum(0.0, r0[0].max(axis=1)) + 1e-6
    x0 = np.concatenate([z0, y0])
    fun = lambda x: float(w @ x[nz:])
    jac = lambda x: np.concatenate([np.zeros(nz), w])

    def con(x):
        c, _ = ev(x[:nz].copy())
        return (x[nz:, None] - c).reshape(-1)

    def con_jac(x):
        _, dc = ev(x[:nz].copy())
        Jm = np.zeros((n_p * ng, nz + n_p))
        Jm[:, :nz] = -dc.reshape(n_p * ng, nz)
        for s in range(n_p):
            Jm[s * ng:(s + 1) * ng, nz + s] = 1.0
        return Jm

    bounds = [(lo[i], hi[i]) for i in range(nz)] + [(0.0, None)] * n_p
    res = minimize(fun, x0, jac=jac, bounds=bounds, constraints=[{"type": "ineq", "fun": con, "jac": con_jac}],
                   method="SLSQP", options={"maxiter": maxiter, "ftol": 1e-14})
    z = res.x[:nz]
    c, _ = ev(z.copy())
    return dict(z=z, phi_raw=float(w @ np.maximum(0.0, c.max(axis=1))), success=res.success, nit=res.nit)


def slsqp_group(vm, d_rows, p, w, z0, maxiter=300):
    """``slsqp_shared`` for a group of scenarios spread over SEVERAL design points that share one control vector (the
    three-stage formulation, reference pyds/run.py:76-91; with one row it is the two-stage problem, with one row and one
    theta the per-scenario problem).  ``w``: weights of the theta samples; the objective averages over the rows."""
    from scipy.optimize import minimize
    d_rows, p = np.atleast_2d(d_rows), np.atleast_2d(p)
    n_d, n_p, nz, ng = len(d_rows), len(p), vm.nz, vm.ng
    S = n_d * n_p
    lo, hi = vm.sec["ZLOW"][:nz], vm.sec["ZHIG"][:nz]
    bigm = vm.big_m
    ws = np.tile(np.asarray(w, float), n_d) / n_d
    cache = {}

    def ev(z):
        key = z.tobytes()
        if key not in cache:
            cache.clear()
            r = vm.evaluate(d_rows, p, z=z, tol=1e-13, want_sens=True)
            cache[key] = ((r["g"] / bigm).reshape(S, ng), (r["dg"] / bigm[:, None]).reshape(S, ng, nz))
        return cache[key]

    z0 = np.asarray(z0, float)
    y0 = np.maximum(0.0, ev(z0)[0].max(axis=1)) + 1e-6
    x0 = np.concatenate([z0, y0])
    fun = lambda x: float(ws @ x[nz:])
    jac = lambda x: np.concatenate([np.zeros(nz), ws])

    def con(x):
        return (x[nz:, None] - ev(x[:nz].copy())[0]).reshape(-1)

    def con_jac(x):
        dc = ev(x[:nz].copy())[1]
        Jm = np.zeros((S * ng, nz + S))
        Jm[:, :nz] = -dc.reshape(S * ng, nz)
        for s in range(S):
            Jm[s * ng:(s + 1) * ng, nz + s] = 1.0
        return Jm

    bounds = [(lo[i], hi[i]) for i in range(nz)] + [(0.0, None)] * S
    res = minimize(fun, x0, jac=jac, bounds=bounds, constraints=[{"type": "ineq", "fun": con, "jac": con_jac}],
                   method="SLSQP", options={"maxiter": maxiter, "ftol": 1e-14})
    z = np.clip(res.x[:nz], lo, hi)
    c = ev(z.copy())[0]
    return dict(z=z, phi_raw=float(ws @ np.maximum(0.0, c.max(axis=1))), success=res.success, nit=res.nit)


def hinge_value(vm, d_rows, p, w, z):
    """The reduced objective  (1/n_d) sum_i sum_s w_s max(0, max_k g_k / M_k)  at the controls ``z`` (independent of any
    optimiser: one state solve per scenario with the tape interpreter)."""
    d_rows, p = np.atleast_2d(d_rows), np.atleast_2d(p)
    c = vm.evaluate(d_rows, p, z=np.asarray(z, float), tol=1e-13)["g"] / vm.big_m
    return float((np.maximum(0.0, c.max(axis=-1)) * np.asarray(w, float)[None, :]).sum() / len(d_rows))
