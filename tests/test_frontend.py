"""Host-side front end on the CPU: the Pyomo-shaped modelling layer, its lowering, the reference's
helper semantics (utils / output manager) and the nested-sampling driver.  No GPU calls."""
import os
import pickle

import numpy as np
import pytest

import pyds_b200.environ as pyo
from oracle import kinetics as K, tape_vm
from pyds_b200 import dae, models, ns, utils
from pyds_b200.output_manager import OutputManager
from pyds_b200.simulator import LoweringError, lower_stage_model
from pyds_b200.user_utils import read_last_file, read_output_file


def _lower_example(mod, tmp_path):
    cfg = mod.build_config(str(tmp_path))
    m = utils.create_flattened_model(cfg["problem"]["stage rules"])
    cfg["problem"]["model transformation"](m)
    return lower_stage_model(m, cfg["problem"]["input map"], m.component("F_in_default")).lower(), cfg


def test_twostage_script_lowers_to_the_reference_equations(tmp_path):
    from examples import twostage_reactor as ex
    lm, cfg = _lower_example(ex, tmp_path)
    assert (lm.ns, lm.nq, lm.ng, lm.nz, lm.d_dim, lm.p_dim, lm.nfe) == (2, 2, 2, 8, 2, 4, 8)
    np.testing.assert_array_equal(lm.big_m, [1e2, 1e7])
    np.testing.assert_array_equal(lm.z_lo, 0.0)
    np.testing.assert_array_equal(lm.z_hi, 2.5)
    assert [e.value for e in lm.z_init] == [0.0] * 8                       # F_in_default = {0: 0}
    vm = tape_vm.TapeVM(lm.blob)
    rng = np.random.default_rng(8)
    d = np.stack([rng.uniform(250, 350, 9), rng.uniform(250, 300, 9)], axis=1)
    p = K.theta_samples()
    F = rng.uniform(0, 2.5, 8)
    r = vm.evaluate(d, p, z=F, tol=1e-13)
    g1, g2, _ = K.g_grid(d, p, F=F)
    np.testing.assert_allclose(r["g"][..., 0], g1, atol=1e-13)
    np.testing.assert_allclose(r["g"][..., 1], g2, atol=1e-8)


def test_threestage_repaired_script_lowers_and_keeps_warm_start(tmp_path):
    from examples import threestage_reactor as ex
    lm, cfg = _lower_example(ex, tmp_path)
    assert cfg["problem"]["input map"] == {1: ["t_f", "T"], 2: ["E1", "E2", "k1", "k2"]}
    # warm start {0: 2.5, 0.2: 0} sampled at the element mid-points of 8 elements
    assert [e.value for e in lm.z_init] == [2.5, 2.5, 0, 0, 0, 0, 0, 0]
    vm = tape_vm.TapeVM(lm.blob)
    d = np.array([[320.0, 280.0]])
    p = K.theta_samples(5)
    F = np.array([e.value for e in lm.z_init])
    r = vm.evaluate(d, p, z=F, tol=1e-13)
    g1, g2, _ = K.g_grid(d, p, F=F)
    np.testing.assert_allclose(r["g"][..., 0], g1, atol=1e-13)


def test_modelling_layer_rejects_what_the_kernel_cannot_express():
    m = pyo.ConcreteModel()
    m.t = dae.ContinuousSet(bounds=(0, 1))
    m.k = pyo.Param(initialize=0, mutable=True)
    m.x = pyo.Var(m.t)
    m.dx = dae.DerivativeVar(m.x, wrt=m.t)
    m.q = pyo.Var(m.t, bounds=(0, 1))
    m.ode = pyo.Constraint(m.t, rule=lambda _m, t: _m.dx[t] == -_m.k * _m.x[t] + _m.q[t])
    m.y = pyo.Var()
    m.cy = pyo.Constraint(rule=lambda _m: _m.y == _m.x[1] - 0.5)
    utils.add_BigMConstraint(m, "c1", 10.0, expr=m.y <= 0)
    imap = {0: [], 1: ["k"]}
    with pytest.raises(LoweringError, match="dae.collocation"):
        lower_stage_model(m, imap)
    tf = pyo.TransformationFactory("dae.collocation")
    tf.apply_to(m, nfe=4, ncp=3)
    with pytest.raises(LoweringError, match="initial value"):
        lower_stage_model(m, imap)
    m.x[0].fix(1.0)
    with pytest.raises(LoweringError, match="piecewise constant"):
        lower_stage_model(m, imap)
    tf.reduce_collocation_points(m, var=m.q, ncp=1, contset=m.t)
    lm = lower_stage_model(m, imap).lower()
    assert (lm.ns, lm.nz, lm.ng, lm.nfe) == (1, 4, 1, 4)
    # x' = -k x + q with q = 0: x(1) = exp(-k); collocation error of Radau-3 on 4 elements is tiny
    r = tape_vm.TapeVM(lm.blob).evaluate(np.zeros((1, 0)), np.array([[0.7]]), tol=1e-13)
    assert abs(r["g"][0, 0, 0] - (np.exp(-0.7) - 0.5)) < 1e-7


def test_bigm_forms():
    m = pyo.ConcreteModel()
    m.a = pyo.Var()
    utils.add_BigMConstraint(m, "up", 5.0, expr=m.a <= 2)
    utils.add_BigMConstraint(m, "lo", 7.0, expr=m.a >= -1)
    utils.add_BigMConstraint(m, "eq", 9.0, expr=m.a == 3)
    names = [n for n, _, _ in m._bigm]
    assert names == ["up", "lo", "eq[1]", "eq[2]"]
    assert [M for _, _, M in m._bigm] == [5.0, 7.0, 9.0, 9.0]
    assert hasattr(m, "_indicator_var") and m._indicator_var.bounds == (0, 1)      # one shared indicator


def test_utils_tree_helpers_and_load_input():
    m = utils.create_EF([lambda mm, st: None, lambda mm, st: None], [1, 3])
    assert m.BFs == [1, 3] and m.n_stages == 2
    assert list(utils.get_all_idx([1, 2, 2])) == [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1)]
    assert utils.get_final_scenarios(m) == [(0, 0), (0, 1), (0, 2)]
    out = utils.load_input(m, {0: ["a"], 1: ["b", "c"]}, {0: np.array([[1.0]]), 1: np.arange(6.0).reshape(3, 2)})
    assert out[1].shape == (3, 2)
    with pytest.raises(ValueError, match="Undefined input binding"):
        utils.load_input(m, {0: ["a"]}, {1: np.zeros((3, 1))})
    with pytest.raises(ValueError):
        utils.create_EF([lambda mm, st: None], [1])


def test_output_manager_stream(tmp_path):
    path = tmp_path / "pyds_output.pkl"
    path.write_bytes(b"stale")
    om = OutputManager(str(tmp_path))
    assert not path.exists()                                   # previous log is removed (reference output_manager.py:9-11)
    for i in range(3):
        om.clear_buffer()
        om.add_input({0: np.array([[float(i)]])})
        om.add_solver_solution({"relaxation": {"objective": i}})
        om.add_simulator_solution([{"user time": 0.0}])
        om.write_data_to_disk()
    recs = read_output_file(str(path))
    assert len(recs) == 3 and recs[2]["solver"]["relaxation"]["objective"] == 2
    assert read_last_file(str(path))["input"][0][0, 0] == 2.0
    assert set(recs[0].keys()) == {"input", "solver", "simulator"}


def test_manager_config_contract(tmp_path):
    """Missing config keys raise KeyError as in the reference (no defaults except 'use relaxation' and
    'io options'.warmstart); construction must not need a GPU (the engine is created on first use)."""
    from examples import twostage_reactor as ex
    from pyds_b200.run import TwoStageManager
    cfg = ex.build_config(str(tmp_path))
    m = TwoStageManager(cfg)
    assert m.n_stages == 2 and m.model is None and m.solver.use_relaxation is True
    assert cfg["solver"]["io options"]["warmstart"] is True
    assert m.simulator.lowered.nz == 8 and m.simulator.input_names == ["t_f", "T", "E1", "E2", "k1", "k2"]
    bad = ex.build_config(str(tmp_path))
    del bad["problem"]["output map"]
    with pytest.raises(KeyError):
        TwoStageManager(bad)
    with pytest.raises(ValueError):
        m._check(np.zeros((3, 3)), np.zeros((5, 4)))


def test_ns_driver_with_a_cpu_constraint_function():
    """The driver against a cheap analytic g (disc of radius 0.3 feasible with probability 1)."""
    calls = []

    def g(d, p):
        calls.append(len(d))
        out = []
        for row in d:
            r = np.hypot(row[0] - 0.5, row[1] - 0.5)
            out.append(np.where(p[:, 0] * 0 + r <= 0.3, 0.0, -1.0)[:, None] * np.ones((len(p), 1)))
        return out

    form = {"activity_type": "ds", "problem": {"parameters_best_estimate": [0.0], "parameters_samples": [{"c": [0.0], "w": 0.5}, {"c": [1.0], "w": 0.5}],
                                              "target_reliability": 0.9, "design_variables": [{"a": [0, 1]}, {"b": [0, 1]}]},
            "solver": {"name": "ds-ns", "settings": {"efp_evaluation": {"constraints_func_ptr": g}, "points_schedule": [(0.0, 40, 8), (0.5, 60, 12)],
                                                     "stop_criteria": [{"inside_fraction": 1.0}]},
                       "algorithms": {"sampling": {"settings": {"prng_seed": 1989, "f0": 0.1, "alpha": 0.3, "stop_criteria": [{"max_iterations": 500}]}}}}}
    res = ns.DEUS(form).solve()
    assert res["inside_fraction"] == 1.0 and len(res["live_points"]) >= 40
    assert np.all(np.hypot(res["live_points"][:, 0] - 0.5, res["live_points"][:, 1] - 0.5) <= 0.3)
    assert res["evaluations"] == sum(calls)
    # efp reduction: -0.0 >= 0 counts as feasible
    assert ns.efp_from_g([np.array([[-0.0], [-1.0]])], np.array([0.25, 0.75]))[0] == 0.25


def test_native_live_set_update_matches_the_python_rule():
    """pydsb_ns_replace_worst (host-only C++) against the driver's Python loop, on probabilities with many ties (P is a
    sum of a few equal weights, as in the shipped scripts) on both sides of the target."""
    from pyds_b200 import capi
    rng = np.random.default_rng(5)
    drv = ns.DEUS.__new__(ns.DEUS)
    for n_live, n_cand, levels, target in ((50, 40, 5, 0.5), (1000, 700, 11, 0.65), (7, 30, 3, 0.0), (64, 0, 4, 0.5)):
        drv.target = target
        live = rng.uniform(size=(n_live, 3))
        live_p = rng.integers(0, levels, n_live) / (levels - 1)
        cand = rng.uniform(size=(n_cand, 3))
        cand_p = rng.integers(0, levels, n_cand) / (levels - 1)
        a_live, a_p = live.copy(), live_p.copy()
        a_live, a_p, _ = drv._replace_worst_py(a_live, a_p, cand, cand_p, n_live)
        b_live, b_p = live.copy(), live_p.copy()
        slots = capi.ns_replace_worst(b_p, cand_p, target)
        taken = slots >= 0
        b_live[slots[taken]] = cand[taken]
        np.testing.assert_array_equal(a_p, b_p)
        np.testing.assert_array_equal(a_live, b_live)
        assert slots.shape == (n_cand,) and np.all(slots < n_live)
    with pytest.raises(ValueError):
        capi.ns_replace_worst(np.zeros(4, dtype=np.float32), np.zeros(2), 0.5)


def test_native_bounding_ellipsoid_matches_numpy():
    from pyds_b200 import capi
    rng = np.random.default_rng(8)
    lo, hi = np.array([250.0, 250.0, 1500.0, 0.0, -1.0]), np.array([350.0, 300.0, 2500.0, 2.5, 1.0])
    A = rng.normal(size=(5, 5)) * 0.1
    live = np.clip(lo + (0.5 + rng.normal(size=(4000, 5)) @ A) * (hi - lo), lo, hi)
    mu, L, r = capi.ns_ellipsoid(live, lo, hi)
    u = (live - lo) / (hi - lo)
    np.testing.assert_allclose(mu, u.mean(axis=0), rtol=1e-12)
    cov = np.cov(u.T) + 1e-12 * np.eye(5)
    np.testing.assert_allclose(L @ L.T, cov, rtol=1e-10, atol=1e-15)
    assert np.all(np.triu(L, 1) == 0.0)
    y = np.linalg.solve(L, (u - mu).T)
    np.testing.assert_allclose(r, np.sqrt((y * y).sum(axis=0).max()), rtol=1e-10)
    # fewer points than dimensions: the fixed ball
    mu2, L2, _ = capi.ns_ellipsoid(live[:3], lo, hi)
    np.testing.assert_array_equal(L2, 0.5 * np.eye(5))
    # both proposal paths draw the same points from the same generator state (up to round-off of the statistics)
    drv = ns.DEUS.__new__(ns.DEUS)
    drv.lo, drv.hi = lo, hi
    out = []
    for native in (True, False):
        drv.native_update = native
        out.append(drv._propose(np.random.default_rng(3), live, 500, 0.1))
    assert out[0].shape == (500, 5)
    np.testing.assert_allclose(out[0], out[1], rtol=1e-9, atol=1e-9)
    with pytest.raises(capi.PydsbError):
        capi.ns_ellipsoid(live, hi, lo)


def test_proposal_samplers_draw_from_the_same_distribution():
    """Rejection from the ellipsoid and rejection from its clipped bounding box both give uniform samples of
    (ellipsoid intersected with the box): same first and second moments within Monte Carlo error, all inside both."""
    from pyds_b200 import capi
    rng = np.random.default_rng(2)
    lo, hi = np.zeros(3), np.ones(3)
    live = np.clip(np.array([0.15, 0.8, 0.5]) + rng.normal(size=(400, 3)) @ np.diag([0.12, 0.1, 0.3]), 0.0, 1.0)
    drv = ns.DEUS.__new__(ns.DEUS)
    drv.lo, drv.hi, drv.native_update = lo, hi, True
    a = drv._propose(np.random.default_rng(10), live, 150000, 0.2, method="ellipsoid")
    b = drv._propose(np.random.default_rng(11), live, 150000, 0.2, method="box")
    c = drv._propose(np.random.default_rng(12), live, 20000, 0.2)
    mu, L, r = capi.ns_ellipsoid(live, lo, hi)
    for pts in (a, b, c):
        assert pts.min() >= 0.0 and pts.max() <= 1.0
        y = np.linalg.solve(L, (pts - mu).T)
        assert np.sqrt((y * y).sum(axis=0)).max() <= r * 1.2 * (1 + 1e-12)
    assert (a.min(axis=0) < 0.01).any() or (a.max(axis=0) > 0.99).any()        # the box does cut the ellipsoid
    np.testing.assert_allclose(a.mean(axis=0), b.mean(axis=0), atol=4e-3)
    np.testing.assert_allclose(np.cov(a.T), np.cov(b.T), atol=2e-3)
    np.testing.assert_allclose(c.mean(axis=0), a.mean(axis=0), atol=1e-2)


def test_ns_driver_native_and_python_updates_follow_the_same_trajectory():
    """The whole driver with the native host passes against the pure-numpy / Python-loop version: same seeds, same
    constraint function -> the same live set after every schedule level (growth phases included), up to round-off of
    the ellipsoid statistics."""
    def g(d, p):
        r = np.hypot(d[:, 0] - 0.4, (d[:, 1] - 0.6) * 1.5)
        lvl = np.clip(np.round((0.6 - r) * 10) / 4, 0.0, 1.0)            # tie-heavy probabilities 0, .25, .5, .75, 1
        return [np.where(np.arange(len(p))[:, None] < lv * len(p), 0.0, -1.0) for lv in lvl]

    def run(native):
        form = {"activity_type": "ds", "problem": {"parameters_best_estimate": [0.0],
                                                  "parameters_samples": [{"c": [float(k)], "w": 0.25} for k in range(4)],
                                                  "target_reliability": 0.75, "design_variables": [{"a": [0, 1]}, {"b": [0, 1]}]},
                "solver": {"name": "ds-ns", "settings": {"efp_evaluation": {"constraints_func_ptr": g},
                                                         "points_schedule": [(0.0, 30, 6), (0.25, 45, 9), (0.5, 60, 12)],
                                                         "stop_criteria": [{"inside_fraction": 1.0}]},
                           "algorithms": {"sampling": {"settings": {"prng_seed": 7, "stop_criteria": [{"max_iterations": 400}]}}}}}
        drv = ns.DEUS(form)
        drv.native_update = native
        return drv.solve()

    a, b = run(True), run(False)
    assert a["iterations"] == b["iterations"] and a["evaluations"] == b["evaluations"]
    assert a["inside_fraction"] == 1.0 and len(a["live_points"]) == len(b["live_points"]) >= 45
    np.testing.assert_array_equal(a["live_efp"], b["live_efp"])
    np.testing.assert_allclose(a["live_points"], b["live_points"], rtol=0, atol=1e-9)


def test_output_record_value_dicts_have_the_reference_shape():
    """``Solver._scenario_values``: what ``utils.parse_value(scen, 'c')`` returns in the reference for an indexed Var over
    (species, t) -- {('A', t): value} with t the discretised ContinuousSet points rounded to 6 decimals -- built here
    from a trajectory array; no device needed."""
    import types
    from pyds_b200.collocation import radau_tables
    from pyds_b200.solver import Solver
    tau = radau_tables(3)[0][1:]
    tgrid = np.concatenate([[0.0], ((np.arange(8)[:, None] + tau[None, :]) / 8).reshape(-1)])
    lowered = types.SimpleNamespace(dyn_states=["c['A']", "c['B']"], quad_states=["u", "c['C']"])
    dae_model = types.SimpleNamespace(state_keys={"u": ("u", (), (0, 1)), "c['A']": ("c", ("A",), (0, 1)),
                                                  "c['B']": ("c", ("B",), (0, 1)), "c['C']": ("c", ("C",), (0, 2))})
    parent = types.SimpleNamespace(simulator=types.SimpleNamespace(lowered=lowered, dae_model=dae_model),
                                   output_map={1: ["c", "u", "nope"]})
    s = Solver.__new__(Solver)
    s.parent = parent
    s._time = tgrid
    s._traj = np.arange(2 * 3 * 25 * 4, dtype=np.float64).reshape(2, 3, 25, 4)
    v = s._scenario_values(1, 2)
    assert set(v) == {"c", "u", "nope"} and v["nope"] is None
    assert len(v["c"]) == 75 and len(v["u"]) == 25
    assert v["c"][("A", 0.0)] == s._traj[1, 2, 0, 0] and v["c"][("B", 1.0)] == s._traj[1, 2, 24, 1]
    assert v["c"][("A", round(0.155051 / 8, 6))] == s._traj[1, 2, 1, 0]
    assert v["c"][("C", 2.0)] == s._traj[1, 2, 24, 3]                # a state on a [0, 2] time set keeps its own bounds
    assert v["u"][0.125] == s._traj[1, 2, 3, 2]                      # scalar-indexed Var: keyed by t alone


def test_device_resident_sampler_matches_the_host_sampler_statistically():
    """ns.DEUS with settings 'device' (torch tensors; the CPU here): same loop, live set / proposals / probabilities resident.
    Both variants must converge to the same region: live set inside the target, similar centre and spread."""
    import torch

    class Owner:
        def g(self, d, p):
            raise NotImplementedError

        def efp(self, d, p, w):
            d = np.asarray(d)
            return np.clip(1.2 - ((d - 0.6) ** 2).sum(1) * 4.0, 0.0, 1.0)

        def efp_device(self, d, p, w):
            return torch.clamp(1.2 - ((d - 0.6) ** 2).sum(dim=1) * 4.0, 0.0, 1.0)

    o = Owner()

    def form(device):
        alg = {"prng_seed": 11, "stop_criteria": [{"max_iterations": 200}]}
        if device:
            alg["device"] = device
        return {"activity_type": "ds",
                "problem": {"design_variables": [{"a": [0, 1]}, {"b": [0, 1]}, {"c": [0, 1]}], "parameters_best_estimate": [1.0],
                            "parameters_samples": [{"c": [1.0], "w": 1.0}], "target_reliability": 0.9},
                "solver": {"name": "ds-ns", "settings": {"efp_evaluation": {"constraints_func_ptr": o.g},
                                                         "points_schedule": [(0.0, 300, 30), (0.5, 400, 40)],
                                                         "stop_criteria": [{"inside_fraction": 1.0}]},
                           "algorithms": {"sampling": {"settings": alg}}}}

    host, dev = ns.DEUS(form(None)).solve(), ns.DEUS(form("cpu")).solve()
    for r in (host, dev):
        assert r["inside_fraction"] == 1.0 and r["live_efp"].min() >= 0.9 and len(r["live_points"]) == 400
        assert r["initial_live_set_seconds"] >= 0.0
    # the feasible region is the ball |d - 0.6|^2 <= 0.075: both live sets fill it
    np.testing.assert_allclose(host["live_points"].mean(0), dev["live_points"].mean(0), atol=0.03)
    np.testing.assert_allclose(host["live_points"].std(0), dev["live_points"].std(0), rtol=0.25)
    assert abs(host["iterations"] - dev["iterations"]) <= 0.5 * host["iterations"]
    # an owner without efp_device is refused, not silently run on the host
    class HostOnly:
        def g(self, d, p):
            raise NotImplementedError
    f = form("cpu")
    f["solver"]["settings"]["efp_evaluation"]["constraints_func_ptr"] = HostOnly().g
    with pytest.raises(ValueError):
        ns.DEUS(f).solve()
