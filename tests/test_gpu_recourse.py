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


def _two_sided(vm, eng_phi, eng_z, d_rows, p, w, rng, z_start=None, n_random=2):
    """Independent check of an optimum the GPU reports, from both sides (BASELINE.json: optima within 1e-6):
    (i) the objective at the reported controls, re-evaluated by the tape interpreter, IS the reported value;
    (ii) SLSQP on the epigraph NLP started AT the reported controls finds nothing lower (local optimality);
    (iii) SLSQP started from the SAME warm start finds nothing lower either, i.e. an independent solver on the reference's
    full-space formulation, from the reference's start, does not beat the reported optimum.
    The hinge objective is not convex -- its minimisers are vertices of the control box (bang-bang profiles) and another
    start can reach another vertex -- so random starts are informational: returns how many of them found a lower value."""
    tol = 1e-6 * max(eng_phi, 1e-3)
    lo, hi = vm.sec["ZLOW"][:vm.nz], vm.sec["ZHIG"][:vm.nz]
    assert abs(RC.hinge_value(vm, d_rows, p, w, eng_z) - eng_phi) <= 1e-9 + 1e-9 * eng_phi
    polish = RC.slsqp_group(vm, d_rows, p, w, eng_z)
    assert polish["phi_raw"] >= eng_phi - tol, (polish["phi_raw"], eng_phi)
    same = RC.slsqp_group(vm, d_rows, p, w, np.zeros(vm.nz) if z_start is None else z_start)
    assert eng_phi <= same["phi_raw"] + tol, (eng_phi, same["phi_raw"])
    better = 0
    for _ in range(n_random):
        r = RC.slsqp_group(vm, d_rows, p, w, rng.uniform(lo, hi))
        better += r["phi_raw"] < eng_phi - tol
    return better, abs(same["phi_raw"] - eng_phi) <= tol


def test_two_sided_independent_optimum_check_two_stage(setup):
    """Every design point of D6, shared control across the 50 shipped theta samples."""
    lm, eng, vm = setup
    p = K.theta_samples(50)
    w = np.full(50, 1 / 50)
    out = eng.recourse_solve(D6, p, w, z0=np.zeros((1, 8)))
    rng = np.random.default_rng(41)
    res = [_two_sided(vm, float(out["phi"][i]), out["z"][i], D6[i:i + 1], p, w, rng) for i in range(len(D6))]
    # from the reference's warm start the independent solver lands on the same optimum (to 1e-6) on at least 5 of the 6
    # points -- where it does not, the GPU's is the lower one (asserted above); a random start finds a better vertex on
    # at most one of them (the problem is multi-modal: which local optimum CONOPT reports is unpinned)
    assert sum(same for _, same in res) >= 5 and sum(b > 0 for b, _ in res) <= 1


def test_two_sided_independent_optimum_check_three_stage(setup):
    """One control vector for the whole (d, theta) tree (reference pyds/run.py:76-91)."""
    lm, eng, vm = setup
    p = K.theta_samples(20)
    w = np.full(20, 1 / 20)
    d = D6[[0, 2, 4]]
    z0 = np.array([2.5, 0, 0, 0, 0, 0, 0, 0.0])
    out = eng.recourse_solve(d, p, w, z0=z0[None, :], global_group=True)
    _two_sided(vm, float(out["phi"][0]), out["z"], d, p, w, np.random.default_rng(43), z_start=z0, n_random=1)


def test_two_sided_independent_optimum_check_per_scenario_c4_shape(setup):
    """Config C4's shape on a sample: 24 (d, theta) scenarios, 8 recourse controls each."""
    lm, eng, vm = setup
    rng = np.random.default_rng(47)
    d = np.stack([rng.uniform(250, 350, 6), rng.uniform(250, 300, 6)], axis=1)
    p = K.theta_samples(4)
    out = eng.recourse_solve(d, p, per_scenario=True, max_iter=50)
    n_infeasible = 0
    for i in range(6):
        for j in range(4):
            phi = float(out["phi"][i, j])
            n_infeasible += phi > 0
            _two_sided(vm, phi, out["z"][i, j], d[i:i + 1], p[j:j + 1], np.ones(1), rng, n_random=1)
    assert 0 < n_infeasible < 24                                   # the sample exercises both outcomes


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


def _coupled_model():
    """Two mutually coupled states (no triangular structure) with a piecewise-constant control: exercises the
    dense 6x6 Newton + sensitivity path of sens_kernel (the kinetics models take the triangular path)."""
    from pyds_b200.model import DaeModel
    m = DaeModel(["a"], ["b", "c"], nfe=4)
    a, b, c = m.d["a"], m.p["b"], m.p["c"]
    x, y = m.state("x", 1.0), m.state("y", 0.5)
    q = m.state("q", 0.0)
    u = m.control("u", -1.0, 1.0, init=0.0)
    m.ode(x, a * x - b * x * y + u)
    m.ode(y, -c * y + 0.5 * b * x * y)
    m.ode(q, u * u + 0.1 * x)                      # quadrature with a control-dependent integrand
    m.constraint("c1", m.end(x) - 1.6, 10.0)
    m.constraint("c2", 0.9 - m.end(y) + 0.05 * m.end(q), 10.0)
    return m


def test_dense_coupled_block_sensitivities_and_optimum():
    lm = _coupled_model().lower()
    assert lm.blocks == [["x", "y"]] and lm.nz == 4
    eng = capi.Engine(lm.blob, 0)
    vm = tape_vm.TapeVM(lm.blob, program=1)
    rng = np.random.default_rng(3)
    d = rng.uniform(0.4, 0.9, (5, 1))
    p = np.stack([rng.uniform(0.8, 1.2, 40), rng.uniform(0.3, 0.6, 40)], axis=1)
    w = np.full(40, 1 / 40)
    # interior optimum on a kink of the hinge: the trust region collapses stage after stage (~160 evaluations)
    ref = RC.shared_control_problem(vm, d, p, w, max_iter=500)
    out = eng.recourse_solve(d, p, w, max_iter=500)
    assert set(out["status"]) <= {1, 2} and set(ref["status"]) <= {1, 2}
    np.testing.assert_allclose(out["phi"], ref["phi_raw"], rtol=1e-6, atol=1e-9)
    g0 = vm.evaluate(d, p)["g"] / vm.big_m          # never worse than the warm start u = 0, better where it can be
    phi0 = (np.maximum(0.0, g0.max(-1)) * w).sum(1)
    assert np.all(out["phi"] <= phi0 + 1e-15) and out["phi"][0] < 0.8 * phi0[0]
    # the state solve itself (feasibility program, generic NB2 / NS2 path) at the optimised controls
    g, _, _ = eng.eval_host(d, p, w, z=out["z"], want_g=True)
    r = tape_vm.TapeVM(lm.blob).evaluate(d, p, z=out["z"], tol=1e-13)
    assert np.all(np.abs(g - r["g"]) <= 1e-10 * np.maximum(np.abs(r["g"]), 1.0))
    # per-scenario mode on the same model
    ref_ps = RC.per_scenario_problem(vm, d[:2], p[:5], max_iter=500)
    out_ps = eng.recourse_solve(d[:2], p[:5], per_scenario=True, max_iter=500)
    np.testing.assert_allclose(out_ps["phi"], ref_ps["phi_raw"], rtol=1e-6, atol=1e-9)


def test_blob_with_second_derivative_tapes_runs_and_agrees(setup):
    """The opt-in second-derivative tables ride in the sensitivity program (longer tapes, more slots); the
    optimiser ignores them and must return the same optimum."""
    lm, eng, vm = setup
    lm2 = models.kinetics("twostage", second_derivatives=True).lower()
    eng2 = capi.Engine(lm2.blob, 0)
    p = K.theta_samples(50)
    w = np.full(50, 1 / 50)
    a = eng.recourse_solve(D6, p, w)
    b = eng2.recourse_solve(D6, p, w)
    np.testing.assert_allclose(a["phi"], b["phi"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a["z"], b["z"], atol=1e-7)


def test_global_control_sharded_entry_point_against_the_single_process_solve(setup):
    """pydsb_recourse_solve_global_dist on one GPU with a stand-in for the all-reduce: this "rank" holds the batch once,
    a second identical rank is imitated by doubling the statistics, n_d_total = 2 n_d.  The average over the doubled
    batch is the same objective, so the optimum must agree with the plain global solve; a rank without rows is imitated
    by a callback that supplies the other rank's statistics (captured from the first run)."""
    import torch
    lm, eng, vm = setup
    dev = torch.device("cuda", 0)
    p = K.theta_samples(40)
    w = np.full(40, 1 / 40)
    one = eng.recourse_solve(D6, p, w, z0=np.zeros((1, lm.nz)), global_group=True)
    lib = capi.load()
    nz = lm.nz
    nv = 2 + nz + nz * (nz + 1) // 2
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    td, tp, tw, tz0 = up(D6), up(p), up(w), up(np.zeros((1, nz)))
    buf = torch.zeros(6 * nv, dtype=torch.float64, device=dev)
    trace = []

    def run(d_t, n_total, fn):
        z = torch.empty((1, nz), dtype=torch.float64, device=dev)
        phi = torch.empty(1, dtype=torch.float64, device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        it = torch.zeros(1, dtype=torch.int32, device=dev)
        cb = capi.ALLREDUCE_FN(fn)
        torch.cuda.synchronize()
        rc = lib.pydsb_recourse_solve_global_dist(eng._h, None if d_t is None else d_t.data_ptr(), 0 if d_t is None else len(d_t),
                                                  n_total, tp.data_ptr(), tw.data_ptr(), 40, tz0.data_ptr(), None, z.data_ptr(),
                                                  phi.data_ptr(), st.data_ptr(), it.data_ptr(), buf.data_ptr(), cb, None, None)
        assert rc == 0, lib.pydsb_last_error()
        return z.cpu().numpy()[0], float(phi.cpu()[0]), int(st.cpu()[0]), int(it.cpu()[0])

    def doubled(ptr, n, user):
        assert ptr == buf.data_ptr() and n == buf.numel()
        trace.append(buf.cpu().numpy().copy())               # this rank's own statistics, per iteration
        buf.mul_(2.0)
        torch.cuda.synchronize()
        return 0

    z2, phi2, st2, it2 = run(td, 12, doubled)
    np.testing.assert_allclose(phi2, one["phi"][0], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(z2, one["z"], rtol=1e-7, atol=1e-9)
    assert st2 == one["status"][0] and it2 == one["iters"][0]

    k = [0]

    def other_rank(ptr, n, user):                            # a rank with no rows: everything comes from the peer
        buf.copy_(torch.from_numpy(2.0 * trace[k[0]]).to(dev))
        k[0] += 1
        torch.cuda.synchronize()
        return 0

    z0r, phi0r, st0r, it0r = run(None, 12, other_rank)
    np.testing.assert_array_equal(z0r, z2)                   # same statistics -> the same decisions, bit for bit
    assert (phi0r, st0r, it0r) == (phi2, st2, it2)

    def failing(ptr, n, user):
        return 1

    z = torch.empty((1, nz), dtype=torch.float64, device=dev)
    rc = lib.pydsb_recourse_solve_global_dist(eng._h, td.data_ptr(), 6, 6, tp.data_ptr(), tw.data_ptr(), 40, None, None,
                                              z.data_ptr(), None, None, None, buf.data_ptr(), capi.ALLREDUCE_FN(failing), None, None)
    assert rc != 0 and b"all-reduce" in lib.pydsb_last_error()


def test_threestage_manager_on_two_gpus():
    """Design points sharded over two ranks, one NCCL all-reduce per optimiser iteration (tools/run_threestage_dist.py
    compares with the single-GPU solve on every rank)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29733", os.path.join(root, "tools", "run_threestage_dist.py"),
                        "--n-d", "21", "--n-p", "20"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_evaluators_and_fused_loop_agree(setup, monkeypatch):
    """Three ways through the same optimisation of a shared-control group with 50 theta samples: the chain evaluator
    inside the single-launch loop (rfused_kernel), the chain evaluator under the host loop (rsens_kernel +
    lm_update_kernel), the coupled-block interpreter under the host loop (sens_kernel).  The first two run the same
    code in a different harness: identical bits.  The third is another evaluation of the same derivatives."""
    lm, eng, vm = setup
    p = K.theta_samples(50)
    w = np.full(50, 1 / 50)
    rng = np.random.default_rng(53)
    d = np.stack([rng.uniform(250, 350, 40), rng.uniform(250, 300, 40)], axis=1)
    fused = eng.recourse_solve(d, p, w, z0=np.zeros((1, 8)))
    monkeypatch.setenv("PYDSB_NO_FUSED_LM", "1")
    host = eng.recourse_solve(d, p, w, z0=np.zeros((1, 8)))
    for k in ("z", "phi", "status", "iters"):
        np.testing.assert_array_equal(fused[k], host[k])
    monkeypatch.setenv("PYDSB_NO_RCHAIN", "1")
    interp = eng.recourse_solve(d, p, w, z0=np.zeros((1, 8)))
    np.testing.assert_allclose(interp["phi"], fused["phi"], rtol=1e-6, atol=1e-9)
    np.testing.assert_array_equal(interp["status"], fused["status"])
    # per-scenario mode: chain evaluator against the interpreter
    monkeypatch.delenv("PYDSB_NO_RCHAIN")
    a = eng.recourse_solve(d[:5], p[:9], per_scenario=True)
    monkeypatch.setenv("PYDSB_NO_RCHAIN", "1")
    b = eng.recourse_solve(d[:5], p[:9], per_scenario=True)
    np.testing.assert_allclose(a["phi"], b["phi"], rtol=1e-6, atol=1e-9)
    np.testing.assert_array_equal(a["status"], b["status"])


def test_chain_sensitivities_against_the_tape_interpreter(setup):
    """The derivatives themselves: one LM iteration from a generic interior point moves along -H^-1 grad; comparing the
    objective decrease is indirect, so take the gradient of the smoothed objective by central differences of the GPU's own
    objective values (per-scenario mode, mu large enough to be smooth) against the interpreter's analytic dg."""
    lm, eng, vm = setup
    p = K.theta_samples(3)
    d = D6[:2]
    z = np.random.default_rng(59).uniform(0.3, 2.2, 8)
    r = vm.evaluate(d, p, z=z, tol=1e-13, want_sens=True)            # dg (n_d, n_p, ng, nz) by the numpy interpreter
    eps = 1e-5
    g0 = eng.eval_host(d, p, None, z=z, want_g=True)[0]
    np.testing.assert_allclose(g0, r["g"], rtol=1e-10, atol=1e-8)
    for c in range(8):
        zp, zm = z.copy(), z.copy()
        zp[c] += eps; zm[c] -= eps
        fd = (eng.eval_host(d, p, None, z=zp, want_g=True)[0] - eng.eval_host(d, p, None, z=zm, want_g=True)[0]) / (2 * eps)
        np.testing.assert_allclose(fd, r["dg"][..., c], rtol=2e-5, atol=1e-6 * np.abs(r["dg"]).max())


@pytest.mark.parametrize("seed", [0, 1, 4, 6, 7, 9, 11, 13, 14, 15, 16, 22, 24, 26, 35, 37])      # the draws with a free control
def test_random_chain_models_optimise_like_the_oracle(seed):
    """Randomly drawn chain-shaped models with a free piecewise-constant control (tests/random_chains.py): the chain
    evaluator with sensitivities in registers (rsens_kernel, generic instantiation) and the fused single-launch loop against
    the numpy restatement of the optimiser, shared controls and per-scenario controls."""
    from random_chains import random_chain_model, random_inputs
    m, nb = random_chain_model(seed)
    lm = m.lower()
    if all(lm.z_fixed):
        pytest.skip("the drawn control is fixed: nothing to optimise")
    eng = capi.Engine(lm.blob, 0)
    vm = tape_vm.TapeVM(lm.blob, program=1)
    d, p = random_inputs(seed, 4, 24)
    w = np.full(len(p), 1.0 / len(p))
    ref = RC.shared_control_problem(vm, d, p, w)
    out = eng.recourse_solve(d, p, w)
    np.testing.assert_allclose(out["phi"], ref["phi_raw"], rtol=1e-6, atol=1e-9)
    ref_ps = RC.per_scenario_problem(vm, d[:2], p[:6])
    out_ps = eng.recourse_solve(d[:2], p[:6], per_scenario=True)
    np.testing.assert_allclose(out_ps["phi"], ref_ps["phi_raw"], rtol=1e-6, atol=1e-9)
    # the evaluator in use is the chain one whenever the model has that shape
    assert eng.info["recourse_chain_blocks"] in (0, nb)

