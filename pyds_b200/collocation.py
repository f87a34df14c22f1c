"""Lagrange-Radau collocation constants for the kernel (product side).

``dae.collocation`` with ``nfe=8, ncp=3`` is what the reference applies to every model
(``run_twostage.py:76-78``); Pyomo's default scheme is LAGRANGE-RADAU.  The tables are built in
50-digit ``decimal`` arithmetic and rounded once to fp64:

* ``tau``  -- nodes on [0,1] with tau_0 = 0 prepended
* ``adot`` -- adot[k][j] = l_k'(tau_j), so that  sum_k adot[k][j] x_k = h * f(x_j)  at j = 1..ncp
* ``bw``   -- Radau quadrature weights (end value of a pure-quadrature state)
"""
from __future__ import annotations

from decimal import Decimal, getcontext
from functools import lru_cache

import numpy as np


@lru_cache(maxsize=None)
def radau_tables(ncp: int = 3):
    if ncp != 3:
        raise NotImplementedError("only the reference's ncp = 3 Radau scheme is tabulated")
    getcontext().prec = 60
    s6 = Decimal(6).sqrt()
    tau = [Decimal(0), (4 - s6) / 10, (4 + s6) / 10, Decimal(1)]
    n = len(tau)
    adot = [[Decimal(0)] * n for _ in range(n)]
    for k in range(n):
        for j in range(n):
            s = Decimal(0)
            for m in range(n):
                if m == k:
                    continue
                p = 1 / (tau[k] - tau[m])
                for l in range(n):
                    if l in (k, m):
                        continue
                    p *= (tau[j] - tau[l]) / (tau[k] - tau[l])
                s += p
            adot[k][j] = s
    bw = [(16 - s6) / 36, (16 + s6) / 36, Decimal(1) / 9]
    f = lambda v: float(v)
    return (np.array([f(t) for t in tau]),
            np.array([[f(v) for v in row] for row in adot]),
            np.array([f(v) for v in bw]))
