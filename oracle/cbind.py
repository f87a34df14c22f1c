"""ctypes loader for the plain-C oracle (test infrastructure / CPU baseline only)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .radau import ADOT, BW

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_kinetics.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kinetics_c.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_kinetics_grid.restype = ctypes.c_longlong
        L.oracle_kinetics_grid.argtypes = [ctypes.c_longlong, ctypes.c_int, dp, ctypes.c_longlong, dp, dp, dp,
                                           ctypes.c_longlong, ctypes.c_double, ctypes.c_int, dp, dp]
        L.oracle_set_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_set_formulation.restype = ctypes.c_int
        L.oracle_set_formulation.argtypes = [ctypes.c_int]
        L.oracle_set_tables.restype = None
        L.oracle_set_tables.argtypes = [dp, dp]
        a = np.ascontiguousarray(np.asarray(ADOT, dtype=np.float64).reshape(-1))
        b = np.ascontiguousarray(np.asarray(BW, dtype=np.float64))
        L.oracle_set_tables(a.ctypes.data_as(dp), b.ctypes.data_as(dp))
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


FORMULATIONS = {"dense9": 0, "blocktri": 1}
# algorithmic flops of one evaluation (SURVEY.md 8d convention, FMA = 2, per Newton iteration and element):
#   dense9:   3 points x (f, J) ~ 45, residual 2 n (ncp + 1) = 72, LU 2/3 n^3 = 486, solves 2 n^2 = 162   (n = 9)
#   blocktri: cA iteration 18 + 24 + 18 + 18 = 78; per element cB solve 78 and the quadratures ~ 20
FLOPS_PER_ITERATION = {"dense9": 45 + 72 + 486 + 162, "blocktri": 78}
FLOPS_PER_ELEMENT = {"dense9": 6, "blocktri": 78 + 20}


def flops_per_eval(formulation, iters_per_eval, nfe=8):
    return FLOPS_PER_ITERATION[formulation] * iters_per_eval + nfe * FLOPS_PER_ELEMENT[formulation] + 40


def kinetics_grid(d, p, w=None, F=None, tol=1e-13, max_it=60, want_g=True, want_P=True, threads=None, formulation="dense9"):
    """Returns (g (n_d,n_p,2) or None, P (n_d,) or None, newton_iterations).  ``formulation``: "dense9" (one 9-unknown
    Newton system per element) or "blocktri" (cA Newton, cB linear, cC quadrature -- what the GPU kernels solve)."""
    L = lib()
    L.oracle_set_formulation(FORMULATIONS[formulation])
    d = np.ascontiguousarray(np.atleast_2d(d), dtype=np.float64)
    p = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
    n_d, d_dim = d.shape
    n_p = p.shape[0]
    w = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
    f_rows = 0
    if F is not None:
        F = np.ascontiguousarray(np.atleast_2d(F), dtype=np.float64)
        f_rows = F.shape[0]
    g = np.empty((n_d, n_p, 2)) if want_g else None
    P = np.empty(n_d) if want_P else None
    if threads is not None:
        L.oracle_set_threads(int(threads))
    it = L.oracle_kinetics_grid(n_d, d_dim, _ptr(d), n_p, _ptr(p), _ptr(w), _ptr(F), f_rows, tol, max_it, _ptr(g), _ptr(P))
    return g, P, int(it)
