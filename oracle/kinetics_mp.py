"""50-digit mpmath restatement of the same model (oracle cross-check; test infrastructure only).

Independent of ``oracle/kinetics.py``: own Newton loop, own linear solve (mpmath LU), own
collocation constants at working precision.  Equations: reference ``run_twostage.py:49-65``,
discretisation ``run_twostage.py:76-79``.  One scenario per call -- use for a handful of points.
"""
from __future__ import annotations

import mpmath as mp

from .radau import radau_adot, radau_weights


def g_values_mp(tf, T, E1, E2, k1, k2, F, cA0=2000, pB=100, pA=20, nfe=8, dps=50, return_traj=False):
    """``return_traj``: also the state profiles at tau = 0 and the collocation points, rows [cA, cB, cC, u]."""
    with mp.workdps(dps):
        A = radau_adot(3, dps)
        b = radau_weights(3, dps)
        tf, T, E1, E2, k1, k2, cA0, pB, pA = [mp.mpf(float(v)) for v in (tf, T, E1, E2, k1, k2, cA0, pB, pA)]
        F = [mp.mpf(float(v)) for v in F]
        K1 = k1 * mp.exp(-E1 / T)
        K2 = k2 * mp.exp(-E2 / T)
        h = mp.mpf(1) / nfe
        x0 = [cA0, mp.mpf(0), mp.mpf(0)]
        u = mp.mpf(0)
        traj = [[cA0, mp.mpf(0), mp.mpf(0), mp.mpf(0)]]
        tau = [(4 - mp.sqrt(6)) / 10, (4 + mp.sqrt(6)) / 10, mp.mpf(1)]

        def f(c, Fe):
            r1 = K1 * c[0] ** 2
            r2 = K2 * c[1]
            return [tf * (-2 * r1 + Fe), tf * (r1 - r2), tf * r2]

        for e in range(nfe):
            X = [list(x0) for _ in range(3)]
            for _ in range(100):
                R = mp.matrix(9, 1)
                M = mp.matrix(9, 9)
                for j in range(3):
                    fj = f(X[j], F[e])
                    for n in range(3):
                        R[3 * j + n] = A[0][j + 1] * x0[n] + sum(A[k + 1][j + 1] * X[k][n] for k in range(3)) - h * fj[n]
                        for k in range(3):
                            M[3 * j + n, 3 * k + n] += A[k + 1][j + 1]
                    d1 = 2 * K1 * X[j][0]
                    M[3 * j + 0, 3 * j + 0] -= h * (-2 * tf * d1)
                    M[3 * j + 1, 3 * j + 0] -= h * (tf * d1)
                    M[3 * j + 1, 3 * j + 1] -= h * (-tf * K2)
                    M[3 * j + 2, 3 * j + 1] -= h * (tf * K2)
                dX = mp.lu_solve(M, -R)
                for j in range(3):
                    for n in range(3):
                        X[j][n] += dX[3 * j + n]
                if max(abs(v) for v in dX) < mp.mpf(10) ** (-(dps - 10)):
                    break
            for j in range(3):                       # u' = F_in / t_f with a piecewise-constant feed: linear in the element
                traj.append([X[j][0], X[j][1], X[j][2], u + h * tau[j] * F[e] / tf])
            u += h * sum(b) * F[e] / tf
            x0 = list(X[2])
        cA, cB, cC = x0
        g1 = mp.mpf("0.8") - (cB + mp.mpf("1e-2")) / (cA + cB + cC + mp.mpf("1e-2"))
        g2 = -(pB * cB - pA * (cA0 + u * tf) - 128 * (tf + 30))
        if return_traj:
            return g1, g2, x0, traj
        return g1, g2, x0
