"""Control (recourse) optimisation on the GPU against the CPU restatement of the same algorithm
(oracle/recourse.py) and an independent scipy SLSQP solve.  Parity vs the reference's CONOPT optimum
is UNPINNED (no solver, no stored solutions); tolerance from BASELINE.json: optima within 1e-6."""
import numpy as np
import pytest

from oracle import kinetics as K, recourse as RC, tape_vm
from pyds_b200 import capi, models

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    lm = models.kinetics("twostage").lower()
    return lm, capi.Engine(lm.blob, 0), tape_vm.TapeVM(lm.blob, program=1)


D6 = np.array([[300., 275.], [340., 285.], [350., 282.5], [260., 290.], [330., 270.], [345., 295.]])


def test_shared_control_matches_cpu_restatement(setup):
    lm, eng, vm = setup
    p = K.theta_samples(50)
    w = np.full(50, 1 / 50)
    ref = RC.shared_control_problem(vm, D6, p, w)
    out = eng.recourse_solve(D6, p, w, z0=np.zeros((1, 8)))
    assert set(out["status"]) <= {1, 2}
    np.testing.assert_allclose(out["phi"], ref["phi_raw"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(out["z"], ref["z"], atol=1e-5)
    # the optimised controls must beat (or equal) the warm start F = 0 on every design point
    g0 = vm.evaluate(D6, p)["g"] / vm.big_m
    phi0 = (np.maximum(0.0, g0.max(-1)) * w).sum(1)
    assert np.all(out["phi"] <= phi0 + 1e-15)
    # classification at the optimised controls, through the plain feasibility kernel
    g, ind, P = eng.eval_host(D6, p, w, z=out["z"], want_g=True, want_ind=True)
    c = g / vm.big_m
    feas_ref = (ref["g"] / vm.big_m).max(-1) <= 1e-7
    feas_gpu = c.max(-1) <= 1e-7
    near = np.abs((ref["g"] / vm.big_m).max(-1)) < 1e-6
    assert np.all((feas_ref == feas_gpu) | near)


def test_shared_control_vs_slsqp_optimum(setup):
    lm, eng, vm = setup
    p = K.theta_samples(50)
    w = np.full(50, 1 / 50)
    out = eng.recourse_solve(D6[:2], p, w)
    for i in range(2):
        s = RC.slsqp_shared(vm, D6[i], p, w, np.zeros(8))
        assert out["phi"][i] <= s["phi_raw"] + 1e-6 * max(s["phi_raw"], 1e-6)


def test_per_scenario_recourse(setup):
    lm, eng, vm = setup
    p = K.theta_samples(7)
    d = D6[:3]
    ref = RC.per_scenario_problem(vm, d, p)
    out = eng.recourse_solve(d, p, per_scenario=True)
    assert out["z"].shape == (3, 7, 8)
    np.testing.assert_allclose(out["phi"], ref["phi_raw"], rtol=1e-6, atol=1e-9)
    # feasible scenarios stop at the first exactly-feasible iterate (status 1, phi == 0)
    assert np.array_equal(out["status"] == 1, out["phi"] == 0.0)
    # a scenario that was feasible at F = 0 needs exactly one evaluation
    g0 = vm.evaluate(d, p)["g"]
    feas0 = (g0 <= 0).all(-1)
    assert np.all(out["iters"][feas0] == 1)
    # verify through the plain kernel with per-scenario controls
    g, ind, P = eng.eval_host(d, p, None, z=out["z"], want_g=True, want_ind=True)
    assert np.array_equal(ind == 0.0, out["phi"] == 0.0)


def test_recourse_larger_theta_set_multi_tile(setup):
    """n_theta > 128: the shared-mode reduction spans several tiles."""
    lm, eng, vm = setup
    p = K.theta_samples(300)
    w = np.full(300, 1 / 300)
    d = D6[:2]
    ref = RC.shared_control_problem(vm, d, p, w)
    out = eng.recourse_solve(d, p, w)
    np.testing.assert_allclose(out["phi"], ref["phi_raw"], rtol=1e-6, atol=1e-9)
