"""Lagrange-Radau collocation constants (oracle side; test infrastructure only).

Restates what ``TransformationFactory('dae.collocation').apply_to(m, nfe=8, ncp=3)``
(reference ``run_twostage.py:76-78``) produces with Pyomo's default LAGRANGE-RADAU
scheme [third-party, from memory; Pyomo is absent from this image]: on every finite
element of width ``h`` the state is the cubic through ``tau = {0, (4-sqrt6)/10,
(4+sqrt6)/10, 1}`` and, at the three collocation points ``j = 1..3``,

    (1/h) * sum_k adot[k][j] * x_k  =  dx/dt (x_j)        with adot[k][j] = l_k'(tau_j).

All constants are evaluated with 50-digit mpmath and rounded once to fp64.
"""
from __future__ import annotations

import mpmath as mp


def radau_tau(ncp: int = 3, dps: int = 50):
    """Collocation nodes on [0,1] (tau_0 = 0 prepended) as mpmath numbers."""
    if ncp != 3:
        # Right-Radau nodes are the roots of P_{n-1}(2t-1) - P_n(2t-1) shifted to [0,1].
        with mp.workdps(dps):
            n = ncp
            f = lambda t: mp.legendre(n - 1, 2 * t - 1) - mp.legendre(n, 2 * t - 1)
            roots = []
            # bracket roots on a fine grid, polish with findroot
            grid = [mp.mpf(i) / 2000 for i in range(1, 2001)]
            prev_t, prev_f = grid[0], f(grid[0])
            for t in grid[1:]:
                ft = f(t)
                if prev_f == 0:
                    roots.append(prev_t)
                elif prev_f * ft < 0:
                    roots.append(mp.findroot(f, (prev_t, t), solver="anderson"))
                prev_t, prev_f = t, ft
            roots.append(mp.mpf(1))
            roots = sorted(set(roots))[-n:]
            return [mp.mpf(0)] + roots
    with mp.workdps(dps):
        s6 = mp.sqrt(6)
        return [mp.mpf(0), (4 - s6) / 10, (4 + s6) / 10, mp.mpf(1)]


def radau_adot(ncp: int = 3, dps: int = 50):
    """adot[k][j] = d/dtau of the k-th Lagrange basis polynomial at tau_j (mpmath)."""
    with mp.workdps(dps):
        tau = radau_tau(ncp, dps)
        n = len(tau)
        out = [[mp.mpf(0)] * n for _ in range(n)]
        for k in range(n):
            for j in range(n):
                s = mp.mpf(0)
                for m in range(n):
                    if m == k:
                        continue
                    p = 1 / (tau[k] - tau[m])
                    for l in range(n):
                        if l in (k, m):
                            continue
                        p *= (tau[j] - tau[l]) / (tau[k] - tau[l])
                    s += p
                out[k][j] = s
        return out


def radau_weights(ncp: int = 3, dps: int = 50):
    """Quadrature weights b_j (j = 1..ncp): integral over [0,1] of the Lagrange basis on
    the ncp collocation nodes.  For a state whose derivative does not depend on any state
    the collocation end value is exactly x_0 + h * sum_j b_j f_j."""
    with mp.workdps(dps):
        tau = radau_tau(ncp, dps)[1:]
        b = []
        for j in range(ncp):
            def basis(t, j=j):
                p = mp.mpf(1)
                for l in range(ncp):
                    if l != j:
                        p *= (t - tau[l]) / (tau[j] - tau[l])
                return p
            b.append(mp.quad(basis, [0, 1]))
        return b


def as_float(x):
    if isinstance(x, list):
        return [as_float(v) for v in x]
    return float(x)


TAU = as_float(radau_tau())
ADOT = as_float(radau_adot())
BW = as_float(radau_weights())
