"""Host-side lowering (expression graph -> tapes -> blob) checked on the CPU: the blob executed
by the oracle's numpy VM must reproduce the hand-written restatement and the mpmath fixtures."""
import json
import math
import os

import numpy as np
import pytest

from oracle import kinetics as K, tape_vm
from pyds_b200 import models, tape as T
from pyds_b200.model import DaeModel, parse_blob


def test_graph_cse_and_folding():
    g = T.Graph()
    a, b = g.param(0), g.param(1)
    assert (a * b) is (b * a)
    assert (a + 0.0) is a and (a * 1.0) is a
    assert (g.const(2.0) * g.const(3.0)).value == 6.0
    assert (a * a).op == "sqr"
    assert (-(-a)) is a
    assert (a - a).is_const(0.0)
    e = g.exp(a) * b
    assert e.level == T.LEVEL_PARAM


def test_levels_and_hoisting():
    g = T.Graph()
    p, x, u = g.param(0), g.x(0), g.ctrl(0)
    e = p * (g.exp(p) * x * x + u)             # distributes p over the sum
    assert e.level == T.LEVEL_STATE
    low = T.lower(g, n_param=1, n_ctrl_elem=1, ns=1, nq=0, roots={"x0": [g.const(1.0)], "z": [], "f": [e], "J": [g.diff(e, x)], "fq": [], "g": [g.xe(0)]})
    # point tape: sqr, fma for f; one mul for J  -- everything parameter-only was hoisted
    assert low.stats["pt"] <= 3
    names = [T.decode(w)[0] for w in low.code["pt"]]
    assert "EXP" not in names and "EXP" in [T.decode(w)[0] for w in low.code["pro"]]


@pytest.mark.parametrize("fn", ["exp", "log", "sqrt", "sin", "cos", "tanh"])
def test_symbolic_derivatives_unary(fn):
    g = T.Graph()
    x = g.x(0)
    e = getattr(g, fn)(x * x + 1.5) / (x + 2.0)
    de = g.diff(e, x)
    f = lambda v: getattr(math, fn)(v * v + 1.5) / (v + 2.0)
    for v in (0.3, 1.1, 2.7):
        num = (f(v + 1e-6) - f(v - 1e-6)) / 2e-6
        assert abs(_eval(de, {("x", 0): v}) - num) < 1e-6 * max(1.0, abs(num))


def test_symbolic_derivative_pow_and_div():
    g = T.Graph()
    x, y = g.x(0), g.x(1)
    e = x ** 2.5 / y + y ** x
    for wrt, idx in ((x, 0), (y, 1)):
        de = g.diff(e, wrt)
        base = {("x", 0): 1.7, ("x", 1): 2.2}
        f = lambda vals: vals[0] ** 2.5 / vals[1] + vals[1] ** vals[0]
        v0 = [1.7, 2.2]; vp = list(v0); vm = list(v0)
        vp[idx] += 1e-6; vm[idx] -= 1e-6
        num = (f(vp) - f(vm)) / 2e-6
        assert abs(_eval(de, base) - num) < 1e-6 * max(1.0, abs(num))


def _eval(e, leaves):
    memo = {}
    for n in T.topo_order([e]):
        if n.op == "const": v = n.value
        elif not n.args: v = leaves[(n.op, n.value)]
        else:
            a = [memo[c.id] for c in n.args]
            v = {"add": lambda: a[0] + a[1], "sub": lambda: a[0] - a[1], "mul": lambda: a[0] * a[1], "div": lambda: a[0] / a[1],
                 "neg": lambda: -a[0], "sqr": lambda: a[0] * a[0], "rcp": lambda: 1.0 / a[0], "exp": lambda: math.exp(a[0]),
                 "log": lambda: math.log(a[0]), "sqrt": lambda: math.sqrt(a[0]), "sin": lambda: math.sin(a[0]),
                 "cos": lambda: math.cos(a[0]), "tanh": lambda: math.tanh(a[0]), "pow": lambda: a[0] ** a[1]}[n.op]()
        memo[n.id] = v
    return memo[e.id]


def test_encode_decode_roundtrip():
    w = T.encode("FMA", T.REF_WIDE | 77, 5, T.REF_CONST | 9, T.REF_WIDE | 4000)
    assert T.decode(w) == ("FMA", T.REF_WIDE | 77, 5, T.REF_CONST | 9, T.REF_WIDE | 4000)


def test_kinetics_structure():
    lm = models.kinetics("twostage").lower()
    # u and cC appear in no right-hand side -> quadratures; the Newton system is 6x6
    assert lm.dyn_states == ["cA", "cB"] and lm.quad_states == ["u", "cC"]
    assert (lm.ns, lm.nq, lm.ng, lm.nz, lm.d_dim, lm.p_dim) == (2, 2, 2, 8, 2, 4)
    np.testing.assert_array_equal(lm.z_lo, 0.0)
    np.testing.assert_array_equal(lm.z_hi, 2.5)
    np.testing.assert_array_equal(lm.big_m, [1e2, 1e7])
    sec = parse_blob(lm.blob)
    # block-triangular solve: cA alone (nonlinear), then cB (linear given cA), then the quadrature integrands
    assert lm.blocks == [["cA"], ["cB"]] and lm.block_linear == [0, 1]
    st = lm.tapes.stats
    assert sec["TPNT"].size == st["pt0"] + st["pt1"] + st["ptq"] <= 8
    assert list(sec["BLKT"]) == [st["pt0"], st["pt1"], st["ptq"]]
    # both blocks are polynomials in their own state (kind bit 1): f_A = c0 + c2 cA^2, f_B = c0(cA) + c1 cB
    assert list(sec["BLKS"]) == [1, 2, 0, 0, 0, 1, 3, 1, 0, 0] and lm.block_poly == [True, True]
    assert [int(r) == T.REF_NONE for r in (sec["RF__"][0], sec["RJ__"][0], sec["RP2_"][0])] == [False, True, False]
    assert [int(r) == T.REF_NONE for r in (sec["RF__"][1], sec["RJ__"][1], sec["RP2_"][1])] == [False, False, True]
    assert st["pt0"] == 0                                   # cA's coefficients are parameter / control values
    lmg = models.kinetics("twostage", polynomial_blocks=False).lower()      # generic point tapes + f / J tables
    assert list(parse_blob(lmg.blob)["BLKS"]) == [1, 0, 0, 0, 0, 1, 1, 1, 0, 0] and lmg.block_poly == [False, False]
    # the coupled single-block form (what the sensitivity program and dense9 use)
    lmc = models.kinetics("twostage", block_triangular=False).lower()
    assert lmc.blocks == [["cA", "cB"]]
    assert [int(r) == T.REF_NONE for r in parse_blob(lmc.blob)["RJ__"]] == [False, True, False, False]   # dfA/dcB == 0
    # every opcode the kernel may meet is known on both sides
    for name in ("TPRO", "TELM", "TPNT", "TGEE"):
        for w in sec[name]:
            assert T.decode(w)[0] in T.OPS
    lm4 = models.kinetics("design4").lower()
    assert lm4.nz == 1 and bool(lm4.z_fixed[0])


@pytest.mark.parametrize("variant,ncol", [("twostage", 2), ("design6", 6)])
def test_blob_vm_vs_mpmath_golden(golden_dir, variant, ncol):
    cases = json.load(open(os.path.join(golden_dir, "kinetics_g_mp.json")))["cases"]
    vm = tape_vm.TapeVM(models.kinetics(variant).lower().blob)
    n = 0
    for c in cases:
        if len(c["d"]) != ncol:
            continue
        z = np.array(c["F"]) if variant == "twostage" else None
        r = vm.evaluate(np.array(c["d"])[None, :], np.array(c["theta"])[None, :], z=z, tol=1e-13)
        assert abs(r["g"][0, 0, 0] - c["g1"]) <= 1e-13
        assert abs(r["g"][0, 0, 1] - c["g2"]) <= 1e-13 * 1e5
        n += 1
    assert n >= 12


def test_blob_vm_vs_numpy_oracle_grid_and_prob():
    rng = np.random.default_rng(5)
    d = np.stack([rng.uniform(250, 350, 25), rng.uniform(250, 300, 25), rng.uniform(1500, 2500, 25), rng.uniform(0, 2.5, 25)], axis=1)
    p = K.theta_samples()
    w = rng.uniform(0.5, 1.5, 50); w /= w.sum()
    vm = tape_vm.TapeVM(models.kinetics("design4").lower().blob)
    r = vm.evaluate(d, p)                      # model's own Newton tolerance
    g1, g2, _ = K.g_grid(d, p)
    np.testing.assert_allclose(r["g"][..., 0], g1, atol=1e-12)
    np.testing.assert_allclose(r["g"][..., 1], g2, atol=1e-7)
    np.testing.assert_allclose(vm.prob_feasible(d, p, w), K.prob_feasible(d, p, w), atol=1e-15)
    assert r["status"].sum() == 0


def test_generic_model_without_quadrature_and_with_transcendentals():
    """A second model family: 1 dynamic state, exercises log/sqrt/tanh/pow through blob + VM and
    compares with scipy on a fine grid (collocation error is tiny for this smooth problem)."""
    from scipy.integrate import solve_ivp
    m = DaeModel(["a"], ["b"], nfe=16)
    g = m.graph
    a, b = m.d["a"], m.p["b"]
    x = m.state("x", 1.0)
    m.ode(x, -a * g.tanh(x) * g.sqrt(1.0 + x * x) + b * g.log(2.0 + x ** 2))
    m.constraint("c", m.end(x) - 1.2, 10.0)
    lm = m.lower()
    assert lm.ns == 1 and lm.nq == 0
    vm = tape_vm.TapeVM(lm.blob)
    r = vm.evaluate(np.array([[0.7], [1.9]]), np.array([[0.4], [1.1]]), tol=1e-13)
    for i, av in enumerate((0.7, 1.9)):
        for j, bv in enumerate((0.4, 1.1)):
            s = solve_ivp(lambda t, y: [-av * math.tanh(y[0]) * math.sqrt(1 + y[0] ** 2) + bv * math.log(2 + y[0] ** 2)], (0, 1), [1.0], rtol=1e-12, atol=1e-13)
            assert abs(r["g"][i, j, 0] - (s.y[0, -1] - 1.2)) < 1e-8


def test_lowering_rejects_bad_models():
    m = DaeModel(["a"], ["b"])
    x = m.state("x", 1.0)
    with pytest.raises(ValueError):
        m.lower()                               # no ODE
    m.ode(x, -x)
    with pytest.raises(ValueError):
        m.lower()                               # no constraint


def test_poly_coeffs():
    g = T.Graph()
    x, y, p = g.x(0), g.x(1), g.param(0)
    vals = {("x", 0): 1.7, ("x", 1): -0.6, ("param", 0): 2.5}
    e = p * (3.0 * x ** 2 - y * x) / (1.0 + p) + g.exp(y) - 4.0
    co = g.poly_coeffs(e, x, 2)
    assert len(co) == 3
    for c in co:                                            # free of x
        assert all(not (n.op == "x" and n.value == 0) for n in T.topo_order([c]))
    xv = vals[("x", 0)]
    assert abs(sum(_eval(c, vals) * xv ** k for k, c in enumerate(co)) - _eval(e, vals)) < 1e-12
    assert len(g.poly_coeffs(p * x - y, x, 2)) == 2         # linear in x
    assert len(g.poly_coeffs(p * y, x, 2)) == 1             # free of x
    assert g.poly_coeffs(g.exp(x) + y, x, 2) is None        # transcendental in x
    assert g.poly_coeffs(x ** 3 + y, x, 2) is None          # degree 3
    assert g.poly_coeffs(y / (1.0 + x), x, 2) is None       # rational in x


def test_polynomial_blocks_agree_with_generic_blocks():
    rng = np.random.default_rng(11)
    d = np.stack([rng.uniform(250, 350, 12), rng.uniform(250, 300, 12), rng.uniform(1500, 2500, 12), rng.uniform(0, 2.5, 12)], axis=1)
    p = K.theta_samples()[:20]
    a = tape_vm.TapeVM(models.kinetics("design4").lower().blob).evaluate(d, p)
    b = tape_vm.TapeVM(models.kinetics("design4", polynomial_blocks=False).lower().blob).evaluate(d, p)
    np.testing.assert_allclose(a["g"][..., 0], b["g"][..., 0], rtol=0, atol=1e-12)
    np.testing.assert_allclose(a["g"][..., 1], b["g"][..., 1], rtol=0, atol=1e-12 * 1e5)
    assert (a["feasible"] == b["feasible"]).all() and not a["status"].any()


def test_second_derivative_tapes_opt_in():
    """second_derivatives=True adds d2g/de2 (end values) and d2f/dx2 to the sensitivity program; checked against
    central differences of the first-derivative tables the same blob carries."""
    lm0 = models.kinetics("twostage").lower()
    lm2 = models.kinetics("twostage", second_derivatives=True).lower()
    assert "RGEE" not in parse_blob(lm0.blob)["sens"] and "RGEE" in parse_blob(lm2.blob)["sens"]
    assert lm2.blob != lm0.blob and parse_blob(lm2.blob)["TPNT"].tolist() == parse_blob(lm0.blob)["TPNT"].tolist()   # program 0 untouched
    lm = lm2
    ns, nq, ng = lm.ns, lm.nq, lm.ng
    vm = tape_vm.TapeVM(lm.blob, program=1)
    r = vm.evaluate(np.array([[300.0, 280.0]]), np.array([[2500.0, 5000.0, 0.064, 9900.0]]), z=np.full(8, 0.5), want_sens=True)
    gee = r["gee"][0, 0]
    assert gee.shape == (ng, ns + nq, ns + nq)
    np.testing.assert_allclose(gee, gee.transpose(0, 2, 1), rtol=1e-12, atol=0)           # symmetric
    assert np.abs(gee[0]).max() > 0 and np.abs(gee[1]).max() == 0                          # g1 rational, g2 linear in e
    # d2g1/de2 against central differences of g1 = 0.8 - (cB + 1e-2)/(cA + cB + cC + 1e-2) at the VM's end values
    xe, qe = r["x_end"][0, 0], r["q_end"][0, 0]
    e0 = np.array([xe[0], xe[1], qe[0], qe[1]])                                            # (cA, cB, u, cC)
    g1 = lambda e: 0.8 - (e[1] + 1e-2) / (e[0] + e[1] + e[3] + 1e-2)
    h = 1e-3 * np.maximum(np.abs(e0), 1.0)
    H = np.zeros((4, 4))
    for a in range(4):
        for b in range(4):
            ea, eb = np.eye(4)[a] * h[a], np.eye(4)[b] * h[b]
            H[a, b] = (g1(e0 + ea + eb) - g1(e0 + ea - eb) - g1(e0 - ea + eb) + g1(e0 - ea - eb)) / (4 * h[a] * h[b])
    np.testing.assert_allclose(gee[0], H, rtol=1e-5, atol=1e-12)
    # d2f/dx2 at the end point: f_A = t_f(-2 K1 cA^2 + F) -> d2f_A/dcA2 = -4 t_f K1, everything else zero but d2f_B/dcA2
    fxx = r["fxx_end"][0, 0]
    K1 = 0.064 * np.exp(-2500.0 / 280.0)
    np.testing.assert_allclose(fxx[0, 0, 0], -4 * 300.0 * K1, rtol=1e-12)
    np.testing.assert_allclose(fxx[1, 0, 0], 2 * 300.0 * K1, rtol=1e-12)
    assert np.count_nonzero(fxx) == 2
