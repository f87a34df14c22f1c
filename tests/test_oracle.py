"""The oracle against every pinned artefact the reference offers, and against itself (three
independent restatements).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import cbind, kinetics as K, radau, tree


def _load(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def test_theta_fixture_matches_reference_sampling(golden_dir):
    g = _load(golden_dir, "theta_fixture.json")
    p = K.theta_samples()
    # sha256 quoted in SURVEY.md section 4 for run_twostage.py:110-116
    assert g["sha256"] == "f8c4fe33c6fbeeddb0ebc23645dbb20c5b3674a3871d2e04dfd28305d1402342"
    assert hashlib.sha256(p.tobytes()).hexdigest() == g["sha256"]
    assert p[0].tolist() == g["row0"]
    np.testing.assert_allclose(p.sum(axis=0), g["col_sums"], rtol=1e-15)


def test_stages_to_update_known_answer(golden_dir):
    g = _load(golden_dir, "stages_to_update.json")
    assert [list(t) for t in tree.stages_to_update(g["input"])] == g["output"]


def test_radau_tables(golden_dir):
    g = _load(golden_dir, "radau3.json")
    A = np.array(radau.ADOT)
    np.testing.assert_array_equal(A, np.array(g["adot"]))
    # SURVEY.md A.3 quotes these to six digits
    ref6 = np.array([[-9.0, -4.139388, 1.739388, -3.0], [10.048809, 3.224745, -3.567840, 5.531973],
                     [-1.382143, 1.167840, 0.775255, -7.531973], [0.333333, -0.253197, 1.053197, 5.0]])
    np.testing.assert_allclose(A, ref6, atol=5e-7)
    # differentiation matrix annihilates constants and differentiates tau exactly
    np.testing.assert_allclose(A.sum(axis=0), 0.0, atol=1e-14)
    np.testing.assert_allclose(np.array(radau.TAU) @ A, 1.0, atol=1e-14)
    np.testing.assert_allclose(sum(radau.BW), 1.0, atol=1e-15)
    # Radau IIA is stiffly accurate: last row of the RK matrix equals the weights, i.e. the
    # collocation polynomial integrates its own derivative values with BW
    tau = np.array(radau.TAU)
    for deg in range(3):
        np.testing.assert_allclose(np.dot(radau.BW, tau[1:] ** deg), 1.0 / (deg + 1), atol=1e-15)


def test_product_side_tables_agree_with_oracle():
    from pyds_b200.collocation import radau_tables
    tau, adot, bw = radau_tables(3)
    np.testing.assert_array_equal(adot, np.array(radau.ADOT))
    np.testing.assert_array_equal(bw, np.array(radau.BW))
    np.testing.assert_array_equal(tau, np.array(radau.TAU))


def _cases(golden_dir):
    return _load(golden_dir, "kinetics_g_mp.json")["cases"]


def test_numpy_oracle_vs_mpmath_golden(golden_dir):
    for c in _cases(golden_dir):
        d = np.array(c["d"])
        kw = dict(cA0=d[2]) if len(d) >= 3 else {}
        xe, iF, _ = K.solve_states(d[0], d[1], *c["theta"], np.array(c["F"]), **kw)
        pB, pA = (d[4], d[5]) if len(d) >= 6 else (100.0, 20.0)
        g1, g2 = K.cqa(xe, iF, d[0], cA0=d[2] if len(d) >= 3 else K.CA0, pB=pB, pA=pA)
        assert abs(g1[0] - c["g1"]) <= 1e-13
        assert abs(g2[0] - c["g2"]) <= 1e-13 * 1e5          # g2 is a difference of O(1e5) terms
        np.testing.assert_allclose(xe[0], c["c_end"], rtol=1e-13, atol=1e-11)


def test_c_oracle_vs_mpmath_golden(golden_dir):
    for c in _cases(golden_dir):
        d = np.array(c["d"])[None, :]
        F = None if d.shape[1] >= 4 else np.array(c["F"])
        g, P, it = cbind.kinetics_grid(d, np.array(c["theta"])[None, :], None, F=F, tol=1e-13)
        assert it > 0
        assert abs(g[0, 0, 0] - c["g1"]) <= 1e-13
        assert abs(g[0, 0, 1] - c["g2"]) <= 1e-13 * 1e5


def test_c_and_numpy_oracles_agree_on_grid():
    rng = np.random.default_rng(3)
    d = np.stack([rng.uniform(250, 350, 40), rng.uniform(250, 300, 40)], axis=1)
    p = K.theta_samples()
    w = np.full(50, 1 / 50)
    g, P, _ = cbind.kinetics_grid(d, p, w, tol=1e-13)
    g1, g2, _ = K.g_grid(d, p)
    np.testing.assert_allclose(g[..., 0], g1, atol=1e-13)
    np.testing.assert_allclose(g[..., 1], g2, atol=1e-8)
    np.testing.assert_allclose(P, K.prob_feasible(d, p, w), atol=1e-15)


def test_collocation_converges_to_ode_solution():
    """Independent sanity bound: the Radau 8x3 solution differs from an
    accurate ODE integration by the discretisation error, and that error shrinks at high order
    when the number of finite elements grows."""
    from scipy.integrate import solve_ivp
    tf, T = 300.0, 275.0
    E1, E2, k1, k2 = K.P_BEST
    K1, K2 = K.rate_constants(T, E1, E2, k1, k2)
    rhs = lambda t, c: [tf * (-2 * K1 * c[0] ** 2), tf * (K1 * c[0] ** 2 - K2 * c[1]), tf * K2 * c[1]]
    sol = solve_ivp(rhs, (0, 1), [2000.0, 0, 0], method="LSODA", rtol=1e-12, atol=1e-10)
    ref = sol.y[:, -1]
    errs = []
    for nfe in (4, 8, 16):
        xe, _, _ = K.solve_states(tf, T, E1, E2, k1, k2, np.zeros(nfe), nfe=nfe)
        errs.append(np.abs(xe[0] - ref).max())
    assert errs[1] < 1e-2 and errs[2] < errs[1] / 8 and errs[1] < errs[0] / 8
    g1_col = K.cqa(K.solve_states(tf, T, E1, E2, k1, k2, np.zeros(8))[0], 0.0, tf)[0][0]
    assert abs(g1_col - 0.01241505) < 5e-9          # SURVEY.md Appendix D probe value


def test_tree_ordering_and_slicing():
    assert tree.get_all_idx([1, 2, 3]) == [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 1, 0), (0, 1, 1), (0, 1, 2)]
    assert tree.scenario_indices_at_stage([1, 2, 3], 1) == [(0, 0), (0, 1)]
    rows = tree.three_stage_rows(3, 4)
    assert rows[:5] == [(0, 0), (0, 1), (0, 2), (0, 3), (1, 0)]
    plan = tree.load_input_plan([1, 3], {0: ["t_f", "T"], 1: ["E1"]}, {0: np.array([[300.0, 275.0]]), 1: np.array([[1.0], [2.0], [3.0]])})
    assert plan[((0,), "t_f")] == 300.0 and plan[((0, 2), "E1")] == 3.0
    with pytest.raises(ValueError):
        tree.load_input_plan([1, 3], {0: ["t_f"]}, {1: np.zeros((3, 1))})


def test_indicator_rule_and_negative_zero():
    y = np.array([0.0, 1e-300, 0.3, -0.0])
    np.testing.assert_array_equal(tree.snap_indicator(y), [0.0, 1.0, 1.0, 0.0])
    g = -tree.snap_indicator(y)
    assert np.all((g >= 0) == np.array([True, False, False, True]))      # -0.0 >= 0 is feasible
    yy, ind = K.indicator(np.array([-0.1, 0.05, -0.1]), np.array([-5.0, -5.0, 3.0]))
    np.testing.assert_array_equal(ind, [0.0, 1.0, 1.0])
    np.testing.assert_allclose(yy, [0.0, 0.05 / 1e2, 3.0 / 1e7])


def test_state_profiles_against_the_mpmath_golden(golden_dir):
    """Both fp64 restatements (the dense 9-unknown solve and the tape interpreter with cC / u as quadratures) against
    the 50-digit profiles at tau = 0 and the 24 collocation points."""
    from oracle import tape_vm
    from pyds_b200 import models
    lm = models.kinetics("twostage").lower()
    vm = tape_vm.TapeVM(lm.blob)
    names = lm.dyn_states + lm.quad_states
    for c in _load(golden_dir, "kinetics_traj_mp.json")["cases"]:
        ref = np.array(c["traj"])                                   # (25, 4): cA, cB, cC, u
        assert ref.shape == (25, 4) and c["columns"] == ["cA", "cB", "cC", "u"]
        d, th, F = np.array(c["d"]), np.array(c["theta"]), np.array(c["F"])
        traj = K.solve_states(d[0], d[1], *th, F=F, return_traj=True)[3][0]
        np.testing.assert_allclose(traj, ref[:, :3], rtol=1e-11, atol=1e-9)
        tv = vm.evaluate(d[None, :], th[None, :], z=F, return_traj=True)["traj"][0, 0]
        for k, nm in enumerate(lm.dyn_states):
            np.testing.assert_allclose(tv[:, k], ref[:, ["cA", "cB", "cC", "u"].index(nm)], rtol=1e-11, atol=1e-9)
        assert set(names) == {"cA", "cB", "cC", "u"}
