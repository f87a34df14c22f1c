"""The fused polynomial-chain element loop (csrc/poly_chain.cuh, X_PCHAIN) on the GPU: against the interpreted loop
(POLY1 per block and element -- the path round 1 validated), against the numpy tape interpreter (oracle) on a chain
whose shape is NOT registered (generic instantiation, three blocks, bimolecular term, implicit-one narrow factor), with
per-scenario controls, with the small-theta PACK tiles, and the registered instantiation against the generic one."""
import numpy as np
import pytest

from oracle import cbind, kinetics as K, tape_vm
from pyds_b200 import capi, models
from pyds_b200.model import DaeModel

pytestmark = pytest.mark.gpu
G_SCALE = np.array([1.0, 1e5])


def _d4(rng, n):
    return np.stack([rng.uniform(250, 350, n), rng.uniform(250, 300, n), rng.uniform(1500, 2500, n), rng.uniform(0, 2.5, n)], axis=1)


def test_fused_chain_matches_interpreted_loop_and_oracle():
    rng = np.random.default_rng(101)
    n_d, n_t = 37, 611                                   # ragged against the 256-row tile
    d, p = _d4(rng, n_d), K.theta_samples(n_t)
    w = rng.uniform(0.5, 1.5, n_t); w /= w.sum()
    out = {}
    for fused in (True, False):
        lm = models.kinetics("design4", fused_chain=fused).lower()
        ops, info = capi.assemble(lm.blob)
        assert (info["chain_blocks"] == 2) == fused and ("PCHAIN" in [o[1] for o in ops]) == fused
        eng = capi.Engine(lm.blob, 0)
        out[fused] = eng.eval_host(d, p, w, want_g=True, want_ind=True)
        c = eng.counters()
        assert c["failed"] == 0 and c["scenarios"] == n_d * n_t
    g1, g0 = out[True][0], out[False][0]
    assert np.all(np.abs(g1 - g0) <= 1e-12 * np.maximum(np.abs(g0), G_SCALE))
    near = np.any(np.abs(g0) < 1e-8 * G_SCALE, axis=-1)
    assert np.all((out[True][1] == out[False][1]) | near)
    go, Po, _ = cbind.kinetics_grid(d, p, w, tol=1e-13)
    assert np.all(np.abs(g1 - go) <= 1e-10 * np.maximum(np.abs(go), G_SCALE))
    np.testing.assert_allclose(out[True][2], Po, rtol=0, atol=1e-12 + float(w.max()) * near.sum())


def test_fused_chain_with_per_scenario_and_per_design_controls():
    """The twostage model: F_in is a per-element control, so the chain's c0 of cA and the u integrand are
    (narrow) x (control of the element); the controls come from z in every z_mode."""
    rng = np.random.default_rng(103)
    n_d, n_t = 9, 300
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    p = K.theta_samples(n_t)
    engs = {f: capi.Engine(models.kinetics("twostage", fused_chain=f).lower().blob, 0) for f in (True, False)}
    for z in (rng.uniform(0, 2.5, 8), rng.uniform(0, 2.5, (n_d, 8)), rng.uniform(0, 2.5, (n_d, n_t, 8))):
        g1 = engs[True].eval_host(d, p, None, z=z, want_g=True)[0]
        g0 = engs[False].eval_host(d, p, None, z=z, want_g=True)[0]
        assert np.all(np.abs(g1 - g0) <= 1e-12 * np.maximum(np.abs(g0), G_SCALE))
    # and the model's own warm-start profile (z_mode 0) against the C oracle
    g1 = engs[True].eval_host(d, p, None, want_g=True)[0]
    go, _, _ = cbind.kinetics_grid(d, p, np.full(n_t, 1.0 / n_t), tol=1e-13)
    assert np.all(np.abs(g1 - go) <= 1e-10 * np.maximum(np.abs(go), G_SCALE))


def test_fused_chain_in_packed_tiles():
    """50 theta samples (the shipped scripts): several design points share a tile (PACK), per-design-point controls."""
    rng = np.random.default_rng(107)
    n_d, n_t = 203, 50
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    p = K.theta_samples(n_t)
    z = rng.uniform(0, 2.5, (n_d, 8))
    out = {}
    for fused in (True, False):
        eng = capi.Engine(models.kinetics("twostage", fused_chain=fused).lower().blob, 0)
        out[fused] = eng.eval_host(d, p, None, z=z, want_g=True, want_ind=True)
    assert np.all(np.abs(out[True][0] - out[False][0]) <= 1e-12 * np.maximum(np.abs(out[False][0]), G_SCALE))
    near = np.any(np.abs(out[False][0]) < 1e-8 * G_SCALE, axis=-1)
    np.testing.assert_allclose(out[True][2], out[False][2], rtol=0, atol=1e-12 + near.sum() / n_t)


def _three_block_model(**kw):
    """A -> B -> C with a bimolecular step: cA' = F - k1 cA^2 ; cB' = k1 cA^2 - k2 cA cB ; cC' = k3 cB - k4 cC ;
    quadratures q1' = cC^2 (no narrow factor), q2' = F (control only)."""
    m = DaeModel(["a"], ["k1", "k2"], nfe=6, **kw)
    g = m.graph
    a, k1, k2 = m.d["a"], m.p["k1"], m.p["k2"]
    cA, cB, cC = m.state("cA", 1.0 + a), m.state("cB", 0.1), m.state("cC", 0.0)
    q1, q2 = m.state("q1", 0.0), m.state("q2", 0.5)
    F = m.control("F", 0.0, 1.0, init=[0.1, 0.3, 0.0, 0.2, 0.5, 0.4])
    m.ode(cA, F - k1 * cA ** 2)
    m.ode(cB, k1 * cA ** 2 - k2 * cA * cB)
    m.ode(cC, 0.7 * cB - (0.2 + a) * cC)
    m.ode(q1, cC * cC)
    m.ode(q2, F)
    m.constraint("c1", m.end(cC) - 0.4, 10.0)
    m.constraint("c2", m.end(q1) + m.end(q2) - m.end(cB) - 0.55, 10.0)
    return m


def test_unregistered_three_block_chain_against_the_tape_interpreter():
    lm = _three_block_model().lower()
    ops, info = capi.assemble(lm.blob)
    assert info["chain_blocks"] == 3 and info["chain_shape_registered"] == 0
    pc = [o for o in ops if o[1] == "PCHAIN"][0][2]
    kinds = [int(pc[8 + 12 * b + 2]) for b in range(3)]
    assert kinds == [0, 1, 2]                            # Newton; linear, c1 varies with cA (one LU step); constant decay
    eng = capi.Engine(lm.blob, 0)
    vm = tape_vm.TapeVM(lm.blob)
    rng = np.random.default_rng(109)
    d = rng.uniform(0.0, 1.0, (19, 1))
    p = np.stack([rng.uniform(0.5, 3.0, 333), rng.uniform(0.2, 2.0, 333)], axis=1)
    for z in (None, rng.uniform(0, 1, (19, 6))):
        g, ind, P = eng.eval_host(d, p, None, z=z, want_g=True, want_ind=True)
        r = vm.evaluate(d, p, z=z, tol=1e-13)
        assert np.all(np.abs(g - r["g"]) <= 1e-10 * np.maximum(np.abs(r["g"]), 1.0))
        near = np.any(np.abs(r["g"]) < 1e-8, axis=-1)
        assert np.all(((ind == 0.0) == r["feasible"]) | near)
    assert eng.counters()["failed"] == 0
    # the interpreted loop gives the same values
    g0 = capi.Engine(_three_block_model(fused_chain=False).lower().blob, 0).eval_host(d, p, None, z=z, want_g=True)[0]
    assert np.all(np.abs(g - g0) <= 1e-12 * np.maximum(np.abs(g0), 1.0))


@pytest.mark.parametrize("seed", range(24))
def test_random_chains_against_the_tape_interpreter(seed):
    """Randomly drawn chain-shaped models (tests/random_chains.py: every combination of term kinds, free and fixed
    controls, 1-3 blocks, 0-2 quadratures; generic instantiation, renumbered units, inexact Newton steps) against the numpy
    interpreter of the same blob, and against the interpreted element loop."""
    from random_chains import random_chain_model, random_inputs
    m, nb = random_chain_model(seed)
    lm = m.lower()
    info = capi.assemble(lm.blob)[1]
    eng = capi.Engine(lm.blob, 0)
    vm = tape_vm.TapeVM(lm.blob)
    d, p = random_inputs(seed, 9, 301)                   # ragged against the 256-row tile
    rng = np.random.default_rng(seed)
    zs = [None] + ([rng.uniform(0, 1, (9, lm.nz))] if lm.nz and not all(lm.z_fixed) else [])
    for z in zs:
        g, ind, P = eng.eval_host(d, p, None, z=z, want_g=True, want_ind=True)
        r = vm.evaluate(d, p, z=z, tol=1e-13)
        assert np.all(np.abs(g - r["g"]) <= 1e-10 * np.maximum(np.abs(r["g"]), 1.0))
        near = np.any(np.abs(r["g"]) < 1e-8, axis=-1)
        assert np.all(((ind == 0.0) == r["feasible"]) | near)
    assert eng.counters()["failed"] == 0
    if info["chain_blocks"]:
        g0 = capi.Engine(random_chain_model(seed, fused_chain=False)[0].lower().blob, 0).eval_host(d, p, None, z=z, want_g=True)[0]
        assert np.all(np.abs(g - g0) <= 1e-11 * np.maximum(np.abs(g0), 1.0))


def test_registered_shape_instantiation_equals_the_generic_one(monkeypatch):
    rng = np.random.default_rng(113)
    d, p = _d4(rng, 11), K.theta_samples(520)
    blob = models.kinetics("design4").lower().blob
    assert capi.assemble(blob)[1]["chain_shape_registered"] == 1
    eng = capi.Engine(blob, 0)
    g_reg = eng.eval_host(d, p, None, want_g=True)[0]
    monkeypatch.setenv("PYDSB_GENERIC_CHAIN", "1")       # looked up at every launch
    g_gen = eng.eval_host(d, p, None, want_g=True)[0]
    assert np.all(np.abs(g_reg - g_gen) <= 1e-13 * np.maximum(np.abs(g_gen), G_SCALE))


def test_chain_reports_failed_newton_solves():
    """A scenario whose Newton iteration cannot converge is reported failed (indicator 1, g = NaN) by the fused loop as
    by the interpreted one: the iteration cap of 2 is too small for the nonlinear block."""
    rng = np.random.default_rng(127)
    d, p = _d4(rng, 5), K.theta_samples(130)
    res = {}
    for fused in (True, False):
        eng = capi.Engine(models.kinetics("design4", fused_chain=fused, newton_max_it=2).lower().blob, 0)
        g, ind, P = eng.eval_host(d, p, None, want_g=True, want_ind=True)
        res[fused] = (np.isnan(g).any(-1), ind, eng.counters()["failed"])
    assert res[True][2] == res[False][2] > 0
    np.testing.assert_array_equal(res[True][0], res[False][0])
    np.testing.assert_array_equal(res[True][1], res[False][1])


def test_handles_of_different_models_alternate():
    """All handles of a process share the kernels' constant-bank windows (program, chain descriptor): every launch
    re-loads its own model's program, stream ordered.  Two models -- a registered chain and a generic three-block one --
    evaluated alternately (feasibility, optimiser, trajectories) must each give what they give alone."""
    rng = np.random.default_rng(131)
    d4, p4 = _d4(rng, 5), K.theta_samples(300)
    d3 = rng.uniform(0.0, 1.0, (5, 1))
    p3 = np.stack([rng.uniform(0.5, 3.0, 300), rng.uniform(0.2, 2.0, 300)], axis=1)
    ea = capi.Engine(models.kinetics("design4").lower().blob, 0)
    eb = capi.Engine(_three_block_model().lower().blob, 0)
    ec = capi.Engine(models.kinetics("twostage").lower().blob, 0)
    d2 = np.stack([rng.uniform(250, 350, 4), rng.uniform(250, 300, 4)], axis=1)
    ga = ea.eval_host(d4, p4, None, want_g=True)[0]
    gb = eb.eval_host(d3, p3, None, want_g=True)[0]
    rc = ec.recourse_solve(d2, p4[:50], np.full(50, 0.02))
    xa = ea.state_solve(d4[:2], p4[:40])[0]
    for _ in range(3):
        gb2 = eb.eval_host(d3, p3, None, want_g=True)[0]
        rc2 = ec.recourse_solve(d2, p4[:50], np.full(50, 0.02))
        ga2 = ea.eval_host(d4, p4, None, want_g=True)[0]
        xa2 = ea.state_solve(d4[:2], p4[:40])[0]
        assert np.array_equal(ga, ga2) and np.array_equal(gb, gb2) and np.array_equal(xa, xa2)
        assert np.array_equal(rc["z"], rc2["z"]) and np.array_equal(rc["phi"], rc2["phi"])

