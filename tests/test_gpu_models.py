"""Other kernel instantiations and the generic interpreter paths on the GPU: NS = 3 (the 9-unknown
form of the shipped model), NS = 1 with transcendental tape ops (out-of-line cold path), parameter-free
design (d_dim = 0), state-dependent nonlinear quadrature integrand."""
import numpy as np
import pytest

from oracle import cbind, kinetics as K, tape_vm
from pyds_b200 import capi, models
from pyds_b200.model import DaeModel

pytestmark = pytest.mark.gpu
G_SCALE = np.array([1.0, 1e5])


def test_dense9_variant_ns3():
    lm = models.kinetics("design4", dense9=True).lower()
    assert (lm.ns, lm.nq) == (3, 1)
    eng = capi.Engine(lm.blob, 0)
    assert eng.info["ns"] == 3
    rng = np.random.default_rng(12)
    n_d, n_t = 41, 333
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d), rng.uniform(1500, 2500, n_d), rng.uniform(0, 2.5, n_d)], axis=1)
    p = K.theta_samples(n_t)
    g, ind, P = eng.eval_host(d, p, None, want_g=True, want_ind=True)
    go, Po, _ = cbind.kinetics_grid(d, p, np.full(n_t, 1.0 / n_t), tol=1e-13)
    assert np.all(np.abs(g - go) <= 1e-10 * np.maximum(np.abs(go), G_SCALE))
    np.testing.assert_allclose(P, Po, atol=1e-12)
    assert eng.counters()["failed"] == 0


def test_generic_block_triangular_path_matches_polynomial_blocks():
    """polynomial_blocks=False runs the block-triangular kinetics through the generic ops (NB, point ops, NS / NSL)
    instead of the in-register POLY1 solve; both against the C oracle."""
    rng = np.random.default_rng(19)
    n_d, n_t = 29, 301
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d), rng.uniform(1500, 2500, n_d), rng.uniform(0, 2.5, n_d)], axis=1)
    p = K.theta_samples(n_t)
    go, Po, _ = cbind.kinetics_grid(d, p, np.full(n_t, 1.0 / n_t), tol=1e-13)
    for poly in (True, False):
        lm = models.kinetics("design4", polynomial_blocks=poly).lower()
        assert lm.block_poly == [poly, poly]
        eng = capi.Engine(lm.blob, 0)
        g, ind, P = eng.eval_host(d, p, None, want_g=True, want_ind=True)
        assert np.all(np.abs(g - go) <= 1e-10 * np.maximum(np.abs(go), G_SCALE))
        np.testing.assert_allclose(P, Po, atol=1e-12)
        c = eng.counters()
        assert c["failed"] == 0 and c["scenarios"] == n_d * n_t


def test_scenarios_per_thread_variants_agree(monkeypatch):
    """PYDSB_SPT selects feas_kernel<1,1> / <1,2> for the all-polynomial model.  A scenario keeps iterating until
    its whole warp has converged and the two variants group scenarios into warps differently, so g agrees to
    round-off, not bit for bit."""
    rng = np.random.default_rng(23)
    n_d, n_t = 13, 700                       # ragged against both tile sizes (128 and 256)
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d), rng.uniform(1500, 2500, n_d), rng.uniform(0, 2.5, n_d)], axis=1)
    p = K.theta_samples(n_t)
    w = rng.uniform(0.5, 1.5, n_t); w /= w.sum()
    blob = models.kinetics("design4").lower().blob
    out = {}
    for spt in ("1", "2"):
        monkeypatch.setenv("PYDSB_SPT", spt)
        eng = capi.Engine(blob, 0)
        assert eng.info["scenarios_per_thread"] == int(spt)
        out[spt] = eng.eval_host(d, p, w, want_g=True, want_ind=True)
    assert np.all(np.abs(out["1"][0] - out["2"][0]) <= 1e-12 * np.maximum(np.abs(out["2"][0]), G_SCALE))
    near = np.any(np.abs(out["2"][0]) < 1e-8 * G_SCALE, axis=-1)
    assert np.all((out["1"][1] == out["2"][1]) | near)
    np.testing.assert_allclose(out["1"][2], out["2"][2], rtol=0, atol=1e-12 + float(w.max()) * near.sum())
    go, Po, _ = cbind.kinetics_grid(d, p, w, tol=1e-13)
    assert np.all(np.abs(out["2"][0] - go) <= 1e-10 * np.maximum(np.abs(go), G_SCALE))


def _transcendental_model():
    m = DaeModel(["a"], ["b", "c"], nfe=6)
    g = m.graph
    a, b, c = m.d["a"], m.p["b"], m.p["c"]
    x = m.state("x", 1.0 + 0.1 * b)
    q = m.state("q", 0.0)
    m.ode(x, -a * g.tanh(x) * g.sqrt(1.0 + x * x) + b * g.log(2.0 + x ** 2) - g.sin(c * x) / (3.0 + g.cos(x)) + g.abs(x) ** 1.5 * 0.01)
    m.ode(q, g.exp(-x) * x * x + g.minimum(x, 0.5) + g.maximum(c, x))          # nonlinear, state-dependent integrand
    m.constraint("c1", m.end(x) - 1.2, 10.0)
    m.constraint("c2", m.end(q) - 1.5 * c, 5.0)
    return m


def test_ns1_transcendental_ops_and_nonlinear_quadrature():
    lm = _transcendental_model().lower()
    assert (lm.ns, lm.nq, lm.ng) == (1, 1, 2)
    eng = capi.Engine(lm.blob, 0)
    vm = tape_vm.TapeVM(lm.blob)
    rng = np.random.default_rng(5)
    d = rng.uniform(0.3, 2.0, (23, 1))
    p = np.stack([rng.uniform(0.2, 1.5, 150), rng.uniform(0.5, 1.2, 150)], axis=1)
    g, ind, P = eng.eval_host(d, p, None, want_g=True, want_ind=True)
    r = vm.evaluate(d, p, tol=1e-13)
    # the kernel corrects the quadrature integrand to first order in the last Newton update instead of
    # re-evaluating it: second-order term is far below the 1e-10 tolerance
    assert np.all(np.abs(g - r["g"]) <= 1e-10 * np.maximum(np.abs(r["g"]), 1.0))
    near = np.any(np.abs(r["g"]) < 1e-8, axis=-1)
    assert np.all(((ind == 0.0) == r["feasible"]) | near)
    assert 0.0 < P.mean() < 1.0


def test_model_without_design_variables():
    m = DaeModel([], ["k"], nfe=4)
    x = m.state("x", 1.0)
    m.ode(x, -m.p["k"] * x)
    m.constraint("c", m.end(x) - 0.5, 1.0)
    eng = capi.Engine(m.lower().blob, 0)
    p = np.linspace(0.2, 1.5, 77)[:, None]
    g, ind, P = eng.eval_host(np.zeros((2, 0)), p, None, want_g=True, want_ind=True)
    r = tape_vm.TapeVM(m.lower().blob).evaluate(np.zeros((2, 0)), p, tol=1e-13)
    np.testing.assert_allclose(g, r["g"], atol=1e-12)
    np.testing.assert_allclose(g[0, :, 0], np.exp(-p[:, 0]) - 0.5, atol=1e-6)      # Radau 4x3 discretisation error
    np.testing.assert_array_equal(g[0], g[1])
    np.testing.assert_allclose(P[0], np.mean(g[0, :, 0] <= 0.0), rtol=0, atol=1e-15)      # 77 weights of 1/77, summation order


def test_unsupported_shapes_fail_loudly():
    m = DaeModel(["a"], ["b"])
    xs = [m.state(f"x{i}", 1.0) for i in range(4)]
    for i in range(4):
        m.ode(xs[i], -xs[(i + 1) % 4] * m.p["b"])          # 4 coupled dynamic states
    m.constraint("c", m.end(xs[0]) - m.d["a"], 1.0)
    with pytest.raises(ValueError, match="limit of 3"):
        m.lower()
    # five states in blocks of (2, 1, 2) are fine: the limit is per coupled block
    m2 = DaeModel(["a"], ["b"], nfe=4)
    ys = [m2.state(f"y{i}", 1.0 + 0.1 * i) for i in range(5)]
    b_ = m2.p["b"]
    m2.ode(ys[0], -b_ * ys[0] + 0.3 * ys[1])
    m2.ode(ys[1], 0.2 * ys[0] - ys[1] * ys[1])
    m2.ode(ys[2], ys[0] - 0.5 * ys[2])
    m2.ode(ys[3], ys[2] * ys[4] - ys[3])
    m2.ode(ys[4], -ys[3] * 0.1 - 0.7 * ys[4] + ys[1])
    m2.constraint("c", m2.end(ys[3]) + m2.end(ys[4]) - m2.d["a"], 2.0)
    lm2 = m2.lower()
    assert sorted(len(b) for b in lm2.blocks) == [1, 2, 2]
    eng = capi.Engine(lm2.blob, 0)
    d = np.linspace(0.2, 1.5, 9)[:, None]
    p = np.linspace(0.3, 1.2, 40)[:, None]
    g, ind, P = eng.eval_host(d, p, None, want_g=True, want_ind=True)
    r = tape_vm.TapeVM(lm2.blob).evaluate(d, p, tol=1e-13)
    np.testing.assert_allclose(g, r["g"], rtol=1e-10, atol=1e-12)
    assert eng.counters()["failed"] == 0


def _mixed_chain_model():
    """A -> B -> C -> D chain, three dynamic states + one quadrature: block A is a quadratic polynomial (POLY1),
    block B has a cubic rate (not a polynomial of degree <= 2 in its own state: generic NB1 / point ops / NS1),
    block C is linear in its own state with a state-dependent coefficient (POLY1, linear)."""
    m = DaeModel(["tf"], ["k1", "k2", "k3"], nfe=5)
    tf = m.d["tf"]
    k1, k2, k3 = m.p["k1"], m.p["k2"], m.p["k3"]
    a, b, c = m.state("a", 1.0), m.state("b", 0.2), m.state("c", 0.0)
    dq = m.state("dq", 0.0)
    m.ode(a, tf * (-k1 * a * a))
    m.ode(b, tf * (k1 * a * a - k2 * b ** 3))
    m.ode(c, tf * (k2 * b ** 3 - k3 * (1.0 + a) * c))
    m.ode(dq, tf * k3 * (1.0 + a) * c)
    m.constraint("yield", 0.25 - m.end(dq), 1.0)
    m.constraint("residual", m.end(b) - 0.6, 1.0)
    return m


def test_mixed_polynomial_and_generic_blocks_in_one_program():
    lm = _mixed_chain_model().lower()
    assert lm.blocks == [["a"], ["b"], ["c"]] and lm.block_poly == [True, False, True] and lm.block_linear == [0, 0, 1]
    ops, info = capi.assemble(lm.blob)
    names = [o[1] for o in ops]
    assert names.count("POLY1") == 2 and names.count("NB1") == 1 and names.count("NS1") == 1
    assert info["scenarios_per_thread"] == 1                 # not every block is polynomial
    eng = capi.Engine(lm.blob, 0)
    vm = tape_vm.TapeVM(lm.blob)
    rng = np.random.default_rng(31)
    d = rng.uniform(0.5, 3.0, (19, 1))
    p = np.stack([rng.uniform(0.5, 2.0, 211), rng.uniform(0.5, 3.0, 211), rng.uniform(0.2, 1.5, 211)], axis=1)
    w = np.full(211, 1 / 211)
    g, ind, P = eng.eval_host(d, p, w, want_g=True, want_ind=True)
    r = vm.evaluate(d, p, tol=1e-13)
    assert not r["status"].any() and eng.counters()["failed"] == 0
    assert np.all(np.abs(g - r["g"]) <= 1e-10 * np.maximum(np.abs(r["g"]), 1.0))
    near = np.any(np.abs(r["g"]) < 1e-8, axis=-1)
    assert np.all(((ind == 0.0) == r["feasible"]) | near)
    np.testing.assert_allclose(P, (r["feasible"] * w[None, :]).sum(1), atol=1e-12 + near.sum() / 211)
    assert 0.02 < P.mean() < 0.98


def test_state_solve_trajectories_against_two_cpu_restatements():
    """pydsb_state_solve: dynamic states against the tape interpreter's trajectories, and the quadrature state (cC is
    integrated, not solved for, in the default lowering) against the independent 9-unknown kinetics solve in which cC
    is an ordinary collocation unknown -- that pins the integration-matrix rows used for the inner points."""
    from oracle import kinetics as K, tape_vm
    lm = models.kinetics("twostage").lower()
    eng = capi.Engine(lm.blob, 0)
    rng = np.random.default_rng(3)
    d = np.stack([rng.uniform(250, 350, 5), rng.uniform(250, 300, 5)], axis=1)
    p = K.theta_samples(37)
    z = rng.uniform(0.0, 2.5, size=(5, lm.nz))
    x, failed = eng.state_solve(d, p, z=z)
    ns, nq = lm.ns, lm.nq
    assert x.shape == (5, 37, 25, ns + nq) and failed.shape == (5, 37) and not failed.any()
    tg = eng.time_grid()
    assert tg.shape == (25,) and tg[0] == 0.0 and tg[-1] == 1.0 and np.all(np.diff(tg) > 0)
    vm = tape_vm.TapeVM(lm.blob)
    ref = vm.evaluate(d, p, z=z, return_traj=True)["traj"]
    np.testing.assert_allclose(x[..., :ns], ref, rtol=1e-10, atol=1e-9)
    # all three concentrations from the dense restatement (columns A, B, C)
    names = lm.dyn_states + lm.quad_states
    col = {n: names.index(n) for n in names}
    for i in range(5):
        traj = K.solve_states(d[i, 0], d[i, 1], p[:, 0], p[:, 1], p[:, 2], p[:, 3], F=z[i], return_traj=True)[3]
        for k, nm in enumerate(("cA", "cB", "cC")):
            np.testing.assert_allclose(x[i, :, :, col[nm]], traj[:, :, k], rtol=1e-9, atol=1e-8)
        # u' = F_in / t_f with a piecewise-constant feed: exact piecewise-linear integral
        e_of = np.minimum((tg * 8 - 1e-12).astype(int).clip(0), 7)
        u_ref = (np.concatenate([[0.0], np.cumsum(z[i])])[e_of] / 8 + z[i][e_of] * (tg - e_of / 8)) / d[i, 0]
        np.testing.assert_allclose(x[i, :, :, col["u"]], np.broadcast_to(u_ref, (37, 25)), rtol=1e-12, atol=1e-14)
    # the end point of the trajectory is what g is computed from: same solve as the feasibility kernel
    g = eng.eval_host(d, p, None, z=z, want_g=True, want_P=False)[0]
    g_ref = vm.evaluate(d, p, z=z)["g"]
    np.testing.assert_allclose(g, g_ref, rtol=1e-10, atol=1e-9)
    # no-control model, two scenarios per thread, ragged tile
    lm3 = models.kinetics("design4").lower()
    eng3 = capi.Engine(lm3.blob, 0)
    assert eng3.info["scenarios_per_thread"] == 2
    d3, p3, _ = __import__("bench").make_inputs(3, 301)
    x3, f3 = eng3.state_solve(d3, p3)
    ref3 = tape_vm.TapeVM(lm3.blob).evaluate(d3, p3, return_traj=True)["traj"]
    np.testing.assert_allclose(x3[..., :lm3.ns], ref3, rtol=1e-10, atol=1e-9)
    assert not f3.any()
