"""Random chain-shaped models (test infrastructure): 1-3 one-state blocks with right-hand sides
c0 + c1 x + c2 x^2 whose coefficients are single terms  +-narrow [* control] [* (earlier state)^p], 0-3 quadratures with
such integrands, one piecewise-constant control -- the family `Assembler::try_chain` accepts (csrc/assembler.h), drawn with
decaying dynamics so that every solve converges.  Used by the CPU test of the unit compaction and the GPU parity test."""
import numpy as np

from pyds_b200.model import DaeModel


def random_chain_model(seed, **kw):
    rng = np.random.default_rng(seed)
    nfe = int(rng.choice([3, 4, 6]))
    m = DaeModel(["a", "b"], ["k1", "k2", "k3"], nfe=nfe, **kw)
    a, b = m.d["a"], m.d["b"]
    k = [m.p["k1"], m.p["k2"], m.p["k3"]]
    nb = int(rng.integers(1, 4))
    free = bool(rng.integers(0, 2))
    init = [float(v) for v in rng.uniform(0.0, 1.0, nfe)]
    F = m.control("F", 0.0, 1.0, init=init) if free else m.fixed_control("F", 0.3 + b)

    def narrow():
        c = int(rng.integers(0, 5))
        if c == 0: return k[int(rng.integers(0, 3))]
        if c == 1: return float(rng.uniform(0.2, 1.5))
        if c == 2: return k[int(rng.integers(0, 3))] * (0.5 + a)
        if c == 3: return float(rng.uniform(0.2, 0.9)) + b
        return k[int(rng.integers(0, 3))] * float(rng.uniform(0.3, 1.2))

    x = []
    for i in range(nb):
        x0 = float(rng.uniform(0.1, 1.0)) + (a if rng.integers(0, 2) else 0.0)
        x.append(m.state(f"x{i}", x0))
    for i in range(nb):
        rhs = None

        def add(term, sign=1.0):
            nonlocal rhs
            rhs = (term if sign > 0 else -term) if rhs is None else (rhs + term if sign > 0 else rhs - term)

        # source c0: nothing, narrow, narrow * control, or narrow * (earlier state)^p
        c = int(rng.integers(0, 4)) if i else int(rng.integers(0, 3))
        if c == 1: add(narrow())
        elif c == 2: add(narrow() * F)
        elif c == 3:
            src = x[int(rng.integers(0, i))]
            add(narrow() * (src ** 2 if rng.integers(0, 2) else src))
        # decay c1 x: constant rate, or rate * earlier state
        c = int(rng.integers(0, 3)) if i else int(rng.integers(0, 2))
        if c == 0: add(narrow() * x[i], -1.0)
        elif c == 1: add(float(rng.uniform(0.3, 2.0)) * x[i], -1.0)
        else: add(narrow() * x[int(rng.integers(0, i))] * x[i], -1.0)
        # second order c2 x^2
        if rng.integers(0, 2): add(narrow() * x[i] ** 2, -1.0)
        m.ode(x[i], rhs)
    ends = [m.end(v) for v in x]
    for q in range(int(rng.integers(0, 3))):
        qs = m.state(f"q{q}", float(rng.uniform(0.0, 0.5)))
        c = int(rng.integers(0, 3))
        src = x[int(rng.integers(0, nb))]
        m.ode(qs, F if c == 0 else (narrow() * src if c == 1 else src * src))
        ends.append(m.end(qs))
    m.constraint("c1", ends[0] - float(rng.uniform(0.2, 0.6)), 10.0)
    m.constraint("c2", float(rng.uniform(0.5, 1.5)) - sum(ends[1:], ends[-1]), 10.0)
    return m, nb


def random_inputs(seed, n_d, n_t):
    rng = np.random.default_rng(1000 + seed)
    d = np.stack([rng.uniform(0.0, 1.0, n_d), rng.uniform(0.0, 0.5, n_d)], axis=1)
    p = np.stack([rng.uniform(0.3, 2.5, n_t) for _ in range(3)], axis=1)
    return d, p
