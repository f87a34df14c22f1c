"""Parity of the CUDA path (through the C-ABI in libpyds_b200.so) with the oracle.

Tolerances (BASELINE.json north_star): g within 1e-10 relative -- applied as
|dg_k| <= 1e-10 * max(|g_k|, s_k) with s = (1, 1e5): g2 is a difference of O(1e5) terms, so its
round-off floor is absolute; classification identical except where |g_k| < 1e-8 * s_k.
"""
import json
import os

import numpy as np
import pytest

from oracle import cbind, kinetics as K, tape_vm
from pyds_b200 import capi, models

pytestmark = pytest.mark.gpu

G_SCALE = np.array([1.0, 1e5])


def _engine(variant):
    return capi.Engine(models.kinetics(variant).lower().blob, 0)


@pytest.fixture(scope="module")
def eng2():
    return _engine("twostage")


@pytest.fixture(scope="module")
def eng4():
    return _engine("design4")


@pytest.fixture(scope="module")
def eng6():
    return _engine("design6")


def _assert_g_close(g, go):
    tol = 1e-10 * np.maximum(np.abs(go), G_SCALE)
    assert np.all(np.abs(g - go) <= tol), float(np.max(np.abs(g - go) / np.maximum(np.abs(go), G_SCALE)))


def _assert_class_equal(ind, go):
    feas_o = np.all(go <= 0.0, axis=-1)
    feas_g = (ind == 0.0)
    near = np.any(np.abs(go) < 1e-8 * G_SCALE, axis=-1)
    assert np.all((feas_o == feas_g) | near)


def test_loaded_library_is_the_in_tree_cuda_build(eng2):
    assert os.path.samefile(capi.library_path(), os.path.join(os.path.dirname(capi.__file__), "libpyds_b200.so"))
    assert eng2.info["num_sms"] >= 100 and eng2.info["ns"] == 2


def test_golden_mpmath_cases(golden_dir, eng2, eng6):
    cases = json.load(open(os.path.join(golden_dir, "kinetics_g_mp.json")))["cases"]
    n = 0
    for c in cases:
        d = np.array(c["d"])[None, :]
        th = np.array(c["theta"])[None, :]
        if d.shape[1] == 2:
            g, ind, P = eng2.eval_host(d, th, z=np.array(c["F"]), want_g=True, want_ind=True)
        else:
            g, ind, P = eng6.eval_host(d, th, want_g=True, want_ind=True)
        _assert_g_close(g[0, 0], np.array([c["g1"], c["g2"]]))
        assert (ind[0, 0] == 0.0) == (c["g1"] <= 0 and c["g2"] <= 0)
        assert P[0] == (1.0 if ind[0, 0] == 0.0 else 0.0)
        n += 1
    assert n == len(cases)


@pytest.mark.parametrize("n_d,n_t", [(1, 1), (3, 50), (17, 127), (5, 128), (9, 129), (2, 1000), (33, 301)])
def test_grid_vs_c_oracle_ragged_shapes(eng2, n_d, n_t):
    rng = np.random.default_rng(100 + n_d * 1000 + n_t)
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    p = K.theta_samples(n_t, seed=n_t)
    w = rng.uniform(0.5, 1.5, n_t); w /= w.sum()
    F = rng.uniform(0, 2.5, 8)
    g, ind, P = eng2.eval_host(d, p, w, z=F, want_g=True, want_ind=True)
    go, Po, it = cbind.kinetics_grid(d, p, w, F=F, tol=1e-13)
    assert it > 0
    _assert_g_close(g, go)
    _assert_class_equal(ind, go)
    np.testing.assert_allclose(P, Po, atol=1e-12)
    assert set(np.unique(ind)).issubset({0.0, -1.0})


def test_reference_shipped_shape_c1(eng2):
    """Config C1: n_theta = 50 (run_twostage.py:110-120), 150 live points, warm-start profile F = 0."""
    rng = np.random.default_rng(1989)
    d = np.stack([rng.uniform(250, 350, 150), rng.uniform(250, 300, 150)], axis=1)
    p = K.theta_samples()
    w = np.full(50, 1 / 50)
    g, ind, P = eng2.eval_host(d, p, w, want_g=True, want_ind=True)
    go, Po, _ = cbind.kinetics_grid(d, p, w, tol=1e-13)
    _assert_g_close(g, go)
    _assert_class_equal(ind, go)
    np.testing.assert_allclose(P, Po, atol=1e-13)
    # best-estimate call with n_p = 1 (DEUS score evaluation)
    g1, ind1, P1 = eng2.eval_host(d, np.array([K.P_BEST]), None, want_g=True, want_ind=True)
    go1, Po1, _ = cbind.kinetics_grid(d, np.array([K.P_BEST]), np.ones(1), tol=1e-13)
    _assert_g_close(g1, go1)
    np.testing.assert_array_equal(P1, Po1)


def test_design4_and_design6_vs_oracle(eng4, eng6):
    rng = np.random.default_rng(11)
    n_d, n_t = 60, 257
    cols = [rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d), rng.uniform(1500, 2500, n_d), rng.uniform(0, 2.5, n_d),
            rng.uniform(90, 110, n_d), rng.uniform(15, 25, n_d)]
    p = K.theta_samples(n_t)
    for eng, k in ((eng4, 4), (eng6, 6)):
        d = np.stack(cols[:k], axis=1)
        g, ind, P = eng.eval_host(d, p, None, want_g=True, want_ind=True)
        go, Po, _ = cbind.kinetics_grid(d, p, np.full(n_t, 1.0 / n_t), tol=1e-13)
        _assert_g_close(g, go)
        _assert_class_equal(ind, go)
        np.testing.assert_allclose(P, Po, atol=1e-12)


def test_control_modes(eng2):
    rng = np.random.default_rng(21)
    n_d, n_t = 6, 70
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    p = K.theta_samples(n_t)
    vm = tape_vm.TapeVM(models.kinetics("twostage").lower().blob)
    z_d = rng.uniform(0, 2.5, (n_d, 8))
    z_s = rng.uniform(0, 2.5, (n_d, n_t, 8))
    for z in (None, z_d[0], z_d, z_s):
        g, ind, P = eng2.eval_host(d, p, None, z=z, want_g=True, want_ind=True)
        r = vm.evaluate(d, p, z=z, tol=1e-13)
        _assert_g_close(g, r["g"])
        _assert_class_equal(ind, r["g"])


def test_empty_inputs(eng2):
    p = K.theta_samples(5)
    g, ind, P = eng2.eval_host(np.zeros((0, 2)), p, None, want_g=True, want_ind=True)
    assert g.shape == (0, 5, 2) and P.shape == (0,)
    g, ind, P = eng2.eval_host(np.array([[300.0, 275.0]]), np.zeros((0, 4)), None, want_P=True)
    assert P.shape == (1,) and P[0] == 0.0


def test_unaligned_theta_pointer_uses_fallback_path(eng2):
    import torch
    n_d, n_t = 4, 200
    rng = np.random.default_rng(2)
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    p = K.theta_samples(n_t)
    buf = torch.zeros(n_t * 4 + 1, dtype=torch.float64, device="cuda")
    th = buf[1:].view(n_t, 4)                       # 8-byte aligned only -> no bulk TMA
    th.copy_(torch.from_numpy(p))
    assert th.data_ptr() % 16 == 8
    P = torch.empty(n_d, dtype=torch.float64, device="cuda")
    eng2.eval_device(torch.from_numpy(d).cuda(), th, P_out=P)
    _, Po, _ = cbind.kinetics_grid(d, p, np.full(n_t, 1.0 / n_t), tol=1e-13)
    np.testing.assert_allclose(P.cpu().numpy(), Po, atol=1e-12)


def test_failed_state_solve_is_reported_infeasible(eng2):
    # a NaN parameter makes Newton fail: scenario must come back infeasible with NaN g and be counted
    d = np.array([[300.0, 275.0]])
    p = K.theta_samples(3)
    p[1, 0] = np.nan
    eng2.counters(reset=True)
    g, ind, P = eng2.eval_host(d, p, None, want_g=True, want_ind=True)
    assert np.isnan(g[0, 1]).all() and ind[0, 1] == -1.0
    assert np.isfinite(g[0, 0]).all() and np.isfinite(g[0, 2]).all()
    assert eng2.counters()["failed"] == 1


def test_full_size_properties_c3(eng4):
    """BASELINE config C3 shape (1e4 x 1e4 is run by bench.py; here 2000 x 10000 = 2e7 evaluations):
    size-independent properties -- determinism, linearity of P in w, count consistency, bounds,
    permutation invariance -- plus a random sample of rows against the oracle."""
    import torch
    rng = np.random.default_rng(1989)
    n_d, n_t = 2000, 10000
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d), rng.uniform(1500, 2500, n_d), rng.uniform(0, 2.5, n_d)], axis=1)
    p = K.theta_samples(n_t)
    dd, pp = torch.from_numpy(d).cuda(), torch.from_numpy(p).cuda()
    w1 = rng.uniform(0, 1, n_t); w2 = rng.uniform(0, 1, n_t)
    outs = []
    for w in (None, None, w1, w2, w1 + w2):
        P = torch.empty(n_d, dtype=torch.float64, device="cuda")
        eng4.eval_device(dd, pp, w=None if w is None else torch.from_numpy(w).cuda(), P_out=P)
        outs.append(P.cpu().numpy())
    np.testing.assert_array_equal(outs[0], outs[1])                       # deterministic
    assert outs[0].min() >= 0.0 and outs[0].max() <= 1.0 + 1e-12
    np.testing.assert_allclose(outs[2] + outs[3], outs[4], rtol=1e-12, atol=1e-9)   # linear in w
    counts = np.rint(outs[0] * n_t)
    np.testing.assert_allclose(outs[0], counts / n_t, atol=1e-12)         # uniform weights -> k / n_t
    perm = rng.permutation(n_t)
    P = torch.empty(n_d, dtype=torch.float64, device="cuda")
    eng4.eval_device(dd, torch.from_numpy(p[perm]).cuda(), P_out=P)
    np.testing.assert_allclose(P.cpu().numpy(), outs[0], atol=1e-12)      # order of theta is irrelevant
    rows = rng.choice(n_d, 40, replace=False)
    _, Po, _ = cbind.kinetics_grid(d[rows], p, np.full(n_t, 1.0 / n_t), tol=1e-13, want_g=False)
    np.testing.assert_allclose(outs[0][rows], Po, atol=1e-12)
    assert 0.01 < outs[0].mean() < 0.9                                    # the grid straddles the feasible band


@pytest.mark.parametrize("kw", [{}, {"polynomial_blocks": False}, {"block_triangular": False}])
def test_far_outside_the_design_box_same_failures_same_values(kw):
    """Long horizons, hot reactor, large feeds: the first finite elements are stiff and Newton from x_0 does not
    converge for a third of the scenarios.  The kernel must fail exactly where the C oracle fails (iteration cap,
    reported as NaN g / infeasible) and agree to round-off everywhere else -- for all three formulations."""
    rng = np.random.default_rng(77)
    n_d, n_t = 60, 256
    d = np.stack([rng.uniform(50, 2000, n_d), rng.uniform(230, 400, n_d), rng.uniform(100, 5000, n_d), rng.uniform(0, 10, n_d)], axis=1)
    p = K.theta_samples(n_t) * rng.uniform(0.7, 1.3, (n_t, 4))
    w = np.full(n_t, 1.0 / n_t)
    go, Po, _ = cbind.kinetics_grid(d, p, w, tol=1e-13)
    eng = capi.Engine(models.kinetics("design4", **kw).lower().blob, 0)
    g, ind, P = eng.eval_host(d, p, w, want_g=True, want_ind=True)
    bad_o, bad_g = ~np.isfinite(go).all(-1), ~np.isfinite(g).all(-1)
    assert 0.1 < bad_o.mean() < 0.9
    assert np.array_equal(bad_o, bad_g)
    assert eng.counters()["failed"] == int(bad_g.sum())
    ok = ~bad_o
    assert np.all(np.abs(g - go)[ok] <= 1e-10 * np.maximum(np.abs(go)[ok], G_SCALE))
    assert np.all(ind[bad_g] == -1.0)                          # a failed state solve is reported infeasible
    np.testing.assert_allclose(P, Po, atol=1e-12)


@pytest.mark.parametrize("n_d,n_t", [(301, 1), (13, 50), (257, 2), (40, 64), (7, 128)])
def test_small_theta_sets_share_tiles(eng2, eng4, n_d, n_t):
    """n_t <= half a tile: several design points per tile (the PACK instantiation), every output and control mode."""
    rng = np.random.default_rng(7 * n_d + n_t)
    p = K.theta_samples(n_t, seed=3 + n_t)
    w = rng.uniform(0.5, 1.5, n_t); w /= w.sum()
    d2 = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    vm = tape_vm.TapeVM(models.kinetics("twostage").lower().blob)
    for z in (rng.uniform(0, 2.5, (n_d, 8)), rng.uniform(0, 2.5, (n_d, n_t, 8))):
        g, ind, P = eng2.eval_host(d2, p, w, z=z, want_g=True, want_ind=True)
        r = vm.evaluate(d2, p, z=z, tol=1e-13)
        _assert_g_close(g, r["g"])
        _assert_class_equal(ind, r["g"])
        np.testing.assert_allclose(P, ((r["g"] <= 0).all(-1) * w).sum(1), atol=1e-12)
    d4 = np.concatenate([d2, np.stack([rng.uniform(1500, 2500, n_d), rng.uniform(0, 2.5, n_d)], axis=1)], axis=1)
    g, ind, P = eng4.eval_host(d4, p, w, want_g=True, want_ind=True)
    go, Po, _ = cbind.kinetics_grid(d4, p, w, tol=1e-13)
    _assert_g_close(g, go)
    _assert_class_equal(ind, go)
    np.testing.assert_allclose(P, Po, atol=1e-12)
    # the fused P-only call (what the managers and the nested-sampling driver use)
    np.testing.assert_array_equal(eng4.eval_host(d4, p, w)[2], P)


def test_state_solve_against_the_mpmath_golden_profiles(golden_dir, eng2):
    """pydsb_state_solve -- dynamic states, the cC quadrature at the inner points (integration-matrix rows) and u --
    against 50-digit arithmetic at tau = 0 and the 24 collocation points."""
    lm = models.kinetics("twostage").lower()
    names = lm.dyn_states + lm.quad_states
    for c in json.load(open(os.path.join(golden_dir, "kinetics_traj_mp.json")))["cases"]:
        ref = np.array(c["traj"])
        x, failed = eng2.state_solve(np.array(c["d"])[None, :], np.array(c["theta"])[None, :], z=np.array(c["F"]))
        assert not failed.any() and x.shape == (1, 1, 25, 4)
        for k, nm in enumerate(c["columns"]):
            np.testing.assert_allclose(x[0, 0, :, names.index(nm)], ref[:, k], rtol=1e-10, atol=1e-9, err_msg=nm)
