"""Semi-batch A -> B -> C reactor: the process model of the reference's two example scripts, written
once against the Pyomo-shaped layer (``pyds_b200.environ`` / ``pyds_b200.dae``).

Equations and constants are those of reference ``run_twostage.py:12-79`` (identical in
``run_threestage.py:22-77``): second-order A -> B, first-order B -> C with Arrhenius rates,
feed ``F_in`` of A, normalised time, ``du/dt = F_in/t_f`` (sic), CQAs g1 (purity) and g2 (profit) in
big-M form with M = (1e2, 1e7), cA(0) = 2000.
"""
import pyds_b200.environ as pyo
from pyds_b200 import dae
from pyds_b200.utils import add_BigMConstraint

R_GAS = 1


def declare_feed(m, t_f, warm_start):
    """Normalised time, feed control F_in in [0, 2.5] and its integral u (reference run_twostage.py:13-26)."""
    m.t = dae.ContinuousSet(bounds=(0, 1))
    m.u = pyo.Var(m.t, initialize=0)
    m.dudt = dae.DerivativeVar(m.u, initialize=0, wrt=m.t)
    m.F_in = pyo.Var(m.t, initialize=0, bounds=(0, 2.5))
    m.u[0].fix(0)
    m.c_F_in = pyo.Constraint(m.t, rule=lambda _m, t: _m.F_in[t] / t_f() == _m.dudt[t])
    m.F_in_default = pyo.Suffix(direction=pyo.Suffix.LOCAL)
    m.F_in_default[m.F_in] = dict(warm_start)


def declare_kinetics(m, t, t_f, T, u, F_in):
    """Uncertain kinetic parameters, mass balances and CQAs (reference run_twostage.py:29-74)."""
    m.I = pyo.Set(initialize=["A", "B", "C"])
    m.E1 = pyo.Param(mutable=True, initialize=0)
    m.E2 = pyo.Param(mutable=True, initialize=0)
    m.k1 = pyo.Param(mutable=True, initialize=0)
    m.k2 = pyo.Param(mutable=True, initialize=0)
    m.c = pyo.Var(m.I, t, bounds=(0, 1e8), initialize=0)
    m.dcdt = dae.DerivativeVar(m.c, wrt=t)

    def rate1(_m, s):
        return _m.k1 * pyo.exp(-_m.E1 / (R_GAS * T())) * _m.c["A", s] ** 2

    def rate2(_m, s):
        return _m.k2 * pyo.exp(-_m.E2 / (R_GAS * T())) * _m.c["B", s]

    m.mb1 = pyo.Constraint(t, rule=lambda _m, s: _m.dcdt["A", s] / t_f() == -2 * rate1(_m, s) + F_in[s])
    m.mb2 = pyo.Constraint(t, rule=lambda _m, s: _m.dcdt["B", s] / t_f() == rate1(_m, s) - rate2(_m, s))
    m.mb3 = pyo.Constraint(t, rule=lambda _m, s: _m.dcdt["C", s] / t_f() == rate2(_m, s))

    m.g1 = pyo.Var(initialize=0)
    m.g2 = pyo.Var(initialize=0)
    m.c_g1 = pyo.Constraint(rule=lambda _m: _m.g1 == (0.8 - (_m.c["B", 1] + 1e-2) /
                                                      (_m.c["A", 1] + _m.c["B", 1] + _m.c["C", 1] + 1e-2)))
    m.c_g2 = pyo.Constraint(rule=lambda _m: _m.g2 == -(100 * _m.c["B", 1] - 20 * (_m.c["A", 0] + u[1] * t_f())
                                                       - 128 * (t_f() + 30)))
    m.bigM_constant = pyo.Param([1, 2], initialize={1: 1e2, 2: 1e7})
    add_BigMConstraint(m, "cqa1", m.bigM_constant[1], expr=m.g1 <= 0)
    add_BigMConstraint(m, "cqa2", m.bigM_constant[2], expr=m.g2 <= 0)
    m.c["A", 0].fix(2e3)
    m.c["B", 0].fix(0)
    m.c["C", 0].fix(0)


def apply_collocation(m):
    """Radau collocation 8 x 3, feed piecewise constant per element (reference run_twostage.py:76-79)."""
    discretizer = pyo.TransformationFactory("dae.collocation")
    discretizer.apply_to(m, nfe=8, ncp=3)
    discretizer.reduce_collocation_points(m, var=m.F_in, ncp=1, contset=m.t)


def theta_samples(n_samples=50, seed=12341234):
    """The uncertain-parameter scenarios of reference run_twostage.py:110-120."""
    from numpy import random
    ran = random.default_rng(seed)
    E1 = ran.normal(2.5e3, 2.5e1, size=n_samples)
    E2 = ran.normal(5e3, 5e1, size=n_samples)
    k1 = ran.normal(6.409e-2, 6.409e-4, size=n_samples)
    k2 = ran.normal(9.938e3, 9.938e1, size=n_samples)
    p_best = [2.5e3, 5e3, 6.409e-2, 9.938e3]
    p_samples = [{"c": [E1[i], E2[i], k1[i], k2[i]], "w": 1.0 / n_samples} for i in range(n_samples)]
    return p_best, p_samples


def solver_and_simulator_config(**io_options):
    opts = {"solver": "conopt", "warmstart": True}
    opts.update(io_options)
    return ({"name": "gams", "solve trajectories": False, "tee": False, "io options": opts, "save output": False,
             "save solution state": False, "warn infeasible": True},
            {"enabled": True, "package": "scipy", "suffix name": "F_in_default", "save output": False, "kwargs": {}})
