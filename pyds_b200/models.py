"""The process models BASELINE.json's configs name, written against :class:`DaeModel`.

``kinetics(variant)`` is the semi-batch A -> B -> C reactor of the reference's example scripts
(``run_twostage.py:12-79``; same equations in ``run_threestage.py:22-77``), equation for equation,
including the reference's ``du/dtau = F_in/t_f`` (sic) and the ``dcdt/t_f == rhs`` form.

variants:

* ``"twostage"``  d = (t_f, T); control F_in piecewise constant on the 8 elements, bounds [0, 2.5],
  warm-start profile F_in = 0                                   (configs C1, C2, C4)
* ``"design4"``   d = (t_f, T, cA0, Fbar); F_in = Fbar on every element, no recourse  (config C3)
* ``"design6"``   d = (t_f, T, cA0, Fbar, pB, pA); the two price coefficients of g2 become
  design variables (100 and 20 in the shipped model)            (config C5)
"""
from __future__ import annotations

from .model import DaeModel

THETA_NAMES = ["E1", "E2", "k1", "k2"]


def kinetics(variant: str = "twostage", nfe: int = 8, f_init=0.0, dense9: bool = False, **kw) -> DaeModel:
    """``dense9=True`` keeps cC inside the Newton system (3 dynamic states, 9 unknowns per element -- the
    formulation SURVEY.md A.3/A.5 counts flops for) instead of treating it as a quadrature."""
    if dense9:
        kw["quadratures"] = ["u"]
        kw["block_triangular"] = False
    d_names = {"twostage": ["t_f", "T"], "design4": ["t_f", "T", "cA0", "Fbar"],
               "design6": ["t_f", "T", "cA0", "Fbar", "pB", "pA"]}[variant]
    m = DaeModel(d_names, THETA_NAMES, nfe=nfe, ncp=3, **kw)
    g = m.graph
    t_f, T = m.d["t_f"], m.d["T"]
    E1, E2, k1, k2 = (m.p[n] for n in THETA_NAMES)
    cA0 = m.d["cA0"] if "cA0" in m.d else g.const(2.0e3)          # run_twostage.py:72
    pB = m.d["pB"] if "pB" in m.d else g.const(100.0)
    pA = m.d["pA"] if "pA" in m.d else g.const(20.0)

    u = m.state("u", 0.0)                                          # run_twostage.py:17, 21
    cA = m.state("cA", cA0)
    cB = m.state("cB", 0.0)
    cC = m.state("cC", 0.0)
    if variant == "twostage":
        F_in = m.control("F_in", 0.0, 2.5, init=f_init)            # run_twostage.py:19, 25-26, 79
    else:
        F_in = m.fixed_control("F_in", m.d["Fbar"])

    R = 1.0
    r1 = k1 * g.exp(-E1 / (R * T)) * cA ** 2
    r2 = k2 * g.exp(-E2 / (R * T)) * cB
    m.ode(u, F_in / t_f)                                           # :22-23  F_in/t_f == dudt
    m.ode(cA, t_f * (-2 * r1 + F_in))                              # :49-50  dcdt/t_f == ...
    m.ode(cB, t_f * (r1 - r2))                                     # :51-53
    m.ode(cC, t_f * r2)                                            # :54-55

    cA1, cB1, cC1, u1 = m.end(cA), m.end(cB), m.end(cC), m.end(u)
    g1 = 0.8 - (cB1 + 1e-2) / (cA1 + cB1 + cC1 + 1e-2)             # :62-63
    g2 = -(pB * cB1 - pA * (m.initial(cA) + u1 * t_f) - 128 * (t_f + 30))   # :64-65
    m.constraint("cqa1", g1, 1e2)                                  # :69-70
    m.constraint("cqa2", g2, 1e7)                                  # :69, 71
    return m
