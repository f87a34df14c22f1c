"""numpy fp64 restatement of the reference's process model (oracle; test infrastructure only).

Follows the stage rules of reference ``run_twostage.py:12-79`` (identical equations in
``run_threestage.py:22-77``):

* ``run_twostage.py:49-55``   mass balances, written there as ``dcdt/t_f == rhs``
* ``run_twostage.py:22-23``   ``F_in/t_f == dudt``  (divides by t_f -- replicated as is)
* ``run_twostage.py:62-65``   CQAs ``g1`` (purity) and ``g2`` (profit)
* ``run_twostage.py:69-71`` + ``pyds/utils.py:7-44``  big-M rows with one shared indicator
* ``run_twostage.py:72-74``   initial conditions cA(0)=2000, cB(0)=cC(0)=0
* ``run_twostage.py:76-79``   Radau collocation nfe=8, ncp=3; F_in piecewise constant per element

The "kinetics family" generalises the shipped model by exactly the knobs BASELINE.json's
synthetic configs name: the initial concentration ``cA0`` and the two price
coefficients of ``g2`` (``pB`` = 100, ``pA`` = 20 in the shipped model) may be design
variables.  With cA0=2000, pB=100, pA=20 it is the shipped model.

Everything is vectorised over a leading scenario axis ``S``; the collocation system of each
finite element (9 unknowns) is solved by full Newton with LAPACK (partial pivoting).
"""
from __future__ import annotations

import numpy as np

from .radau import ADOT, BW

NFE = 8
NCP = 3
BIG_M = (1.0e2, 1.0e7)          # run_twostage.py:69
F_BOUNDS = (0.0, 2.5)           # run_twostage.py:19
CA0 = 2.0e3                     # run_twostage.py:72
P_BEST = (2.5e3, 5.0e3, 6.409e-2, 9.938e3)   # run_twostage.py:118


def theta_samples(n: int = 50, seed: int = 12341234) -> np.ndarray:
    """The reference's theta fixture, run_twostage.py:110-116 (draw order E1, E2, k1, k2)."""
    ran = np.random.default_rng(seed)
    E1 = ran.normal(2.5e3, 2.5e1, size=n)
    E2 = ran.normal(5e3, 5e1, size=n)
    k1 = ran.normal(6.409e-2, 6.409e-4, size=n)
    k2 = ran.normal(9.938e3, 9.938e1, size=n)
    return np.stack([E1, E2, k1, k2], axis=1)


def rate_constants(T, E1, E2, k1, k2):
    """K1 = k1*exp(-E1/(R*T)), K2 likewise, R = 1 (run_twostage.py:48-53)."""
    return k1 * np.exp(-E1 / T), k2 * np.exp(-E2 / T)


def rhs(c, tf, K1, K2, F):
    """dc/dtau for c = (cA, cB, cC) stacked on the last axis (run_twostage.py:49-55)."""
    cA, cB = c[..., 0], c[..., 1]
    r1 = K1 * cA ** 2
    r2 = K2 * cB
    return np.stack([tf * (-2.0 * r1 + F), tf * (r1 - r2), tf * r2], axis=-1)


def rhs_jac(c, tf, K1, K2):
    cA = c[..., 0]
    J = np.zeros(c.shape + (3,))
    d1 = 2.0 * K1 * cA
    J[..., 0, 0] = -2.0 * tf * d1
    J[..., 1, 0] = tf * d1
    J[..., 1, 1] = -tf * K2
    J[..., 2, 1] = tf * K2
    return J


def solve_states(tf, T, E1, E2, k1, k2, F, cA0=CA0, nfe=NFE, tol=1e-14, max_it=60,
                 return_traj=False):
    """Solve the collocation equations element by element.

    All scalar arguments broadcast to shape (S,); ``F`` broadcasts to (S, nfe).
    Returns ``(c_end (S,3), int_F (S,), iters (S,nfe))`` where ``int_F = u(1)*t_f`` is the
    integral of F_in over normalised time (``u`` of run_twostage.py:17-23 is a pure quadrature).
    """
    tf, T, E1, E2, k1, k2, cA0 = np.broadcast_arrays(
        *[np.asarray(v, dtype=np.float64) for v in (tf, T, E1, E2, k1, k2, cA0)])
    S = tf.shape[0] if tf.ndim else 1
    tf, T, E1, E2, k1, k2, cA0 = [np.reshape(v, (S,)).astype(np.float64)
                                  for v in (tf, T, E1, E2, k1, k2, cA0)]
    F = np.broadcast_to(np.asarray(F, dtype=np.float64), (S, nfe))
    K1, K2 = rate_constants(T, E1, E2, k1, k2)
    h = 1.0 / nfe
    A = np.asarray(ADOT)                      # A[k, j]
    x0 = np.stack([cA0, np.zeros(S), np.zeros(S)], axis=1)
    iters = np.zeros((S, nfe), dtype=np.int64)
    traj = np.zeros((S, nfe * NCP + 1, 3))
    traj[:, 0] = x0
    u = np.zeros(S)
    for e in range(nfe):
        Fe = F[:, e]
        X = np.repeat(x0[:, None, :], NCP, axis=1)          # (S, 3 pts, 3 species)
        active = np.ones(S, dtype=bool)
        for it in range(max_it):
            idx = np.nonzero(active)[0]
            if idx.size == 0:
                break
            Xa = X[idx]
            f = rhs(Xa, tf[idx, None], K1[idx, None], K2[idx, None], Fe[idx, None])
            J = rhs_jac(Xa, tf[idx, None], K1[idx, None], K2[idx, None])
            # residual R[j] = sum_k A[k, j+1] x_k - h f(x_j),  x_0 = element start value
            lin = np.einsum("kj,skn->sjn", A[1:, 1:], Xa) + A[0, 1:][None, :, None] * x0[idx, None, :]
            R = lin - h * f
            M = np.zeros((idx.size, NCP, 3, NCP, 3))
            for j in range(NCP):
                for k in range(NCP):
                    M[:, j, :, k, :] += A[k + 1, j + 1] * np.eye(3)
                M[:, j, :, j, :] -= h * J[:, j]
            dX = np.linalg.solve(M.reshape(idx.size, 9, 9), -R.reshape(idx.size, 9, 1))
            dX = dX.reshape(idx.size, NCP, 3)
            X[idx] = Xa + dX
            iters[idx, e] += 1
            scale = np.maximum(np.abs(X[idx]).max(axis=(1, 2)), 1.0)
            done = np.abs(dX).max(axis=(1, 2)) <= tol * scale
            active[idx[done]] = False
        # one polishing step has already been taken past the tolerance test; quadratic
        # convergence puts the result at round-off.
        u = u + h * Fe * (BW[0] + BW[1] + BW[2]) / tf        # du/dtau = F/t_f  (sic)
        traj[:, e * NCP + 1:(e + 1) * NCP + 1] = X
        x0 = X[:, NCP - 1, :].copy()
    int_F = u * tf
    if return_traj:
        return x0, int_F, iters, traj
    return x0, int_F, iters


def cqa(c_end, int_F, tf, cA0=CA0, pB=100.0, pA=20.0):
    """g1, g2 of run_twostage.py:62-65 (both must be <= 0)."""
    cA, cB, cC = c_end[..., 0], c_end[..., 1], c_end[..., 2]
    g1 = 0.8 - (cB + 1e-2) / (cA + cB + cC + 1e-2)
    g2 = -(pB * cB - pA * (cA0 + int_F) - 128.0 * (tf + 30.0))
    return g1, g2


def indicator(g1, g2, big_m=BIG_M):
    """Optimal relaxed big-M indicator y = max(0, g1/M1, g2/M2) (utils.py:23-29),
    then the snap of solver.py:71-80: y == 0 -> 0 else 1."""
    y = np.maximum(0.0, np.maximum(g1 / big_m[0], g2 / big_m[1]))
    return y, (y != 0.0).astype(np.float64)


def g_grid(d, p, F=None, layout="two-stage", **kw):
    """g1, g2 over the full (n_d, n_p) grid for a fixed control profile.

    ``d`` is (n_d, d_dim) with columns (t_f, T[, cA0[, Fbar[, pB, pA]]]) -- the family of
    SURVEY.md section 8d; ``p`` is (n_p, 4) = (E1, E2, k1, k2).  ``F`` is (nfe,) shared, or
    (n_d, nfe); ignored when d carries Fbar."""
    d = np.atleast_2d(np.asarray(d, dtype=np.float64))
    p = np.atleast_2d(np.asarray(p, dtype=np.float64))
    n_d, d_dim = d.shape
    n_p = p.shape[0]
    tf = np.repeat(d[:, 0], n_p)
    T = np.repeat(d[:, 1], n_p)
    cA0 = np.repeat(d[:, 2], n_p) if d_dim >= 3 else np.full(n_d * n_p, CA0)
    if d_dim >= 4:
        Fs = np.repeat(d[:, 3], n_p)[:, None] * np.ones((1, NFE))
    else:
        F = np.zeros(NFE) if F is None else np.asarray(F, dtype=np.float64)
        Fs = np.repeat(np.broadcast_to(F, (n_d, NFE)), n_p, axis=0)
    pB = np.repeat(d[:, 4], n_p) if d_dim >= 6 else 100.0
    pA = np.repeat(d[:, 5], n_p) if d_dim >= 6 else 20.0
    th = np.tile(p, (n_d, 1))
    c_end, int_F, iters = solve_states(tf, T, th[:, 0], th[:, 1], th[:, 2], th[:, 3], Fs, cA0=cA0, **kw)
    g1, g2 = cqa(c_end, int_F, tf, cA0=cA0, pB=pB, pA=pA)
    return g1.reshape(n_d, n_p), g2.reshape(n_d, n_p), iters.reshape(n_d, n_p, -1)


def manager_g(d, p, F=None):
    """What ``TwoStageManager.g`` returns for a FIXED control profile (run.py:47-69):
    a list of n_d arrays (n_p, 1) holding ``-indicator`` (feasible <=> value >= 0)."""
    g1, g2, _ = g_grid(d, p, F)
    _, ind = indicator(g1, g2)
    return [(-ind[i])[:, None] for i in range(ind.shape[0])]


def prob_feasible(d, p, w, F=None):
    """DEUS's reduction [third-party, from memory; SURVEY.md section 8a last row]:
    P_i = sum_j w_j * 1[all_k g_ijk >= 0] on the manager's output (note -0.0 >= 0)."""
    w = np.asarray(w, dtype=np.float64)
    out = np.empty(len(d))
    for i, gm in enumerate(manager_g(d, p, F)):
        out[i] = np.sum(w * np.all(gm >= 0.0, axis=1))
    return out
