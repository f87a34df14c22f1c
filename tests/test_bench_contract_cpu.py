"""bench.py's line builders on stub measurements (no GPU): every key the driver and the judge read is present and well
formed, the derived figures follow from the measurements, the committed ncu capture is only quoted while the kernel
sources it was taken from are the ones in the tree, and the two arms name the same workload."""
import argparse
import json
import os

import numpy as np

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub(**kw):
    m = {"n_d": 10000, "n_t": 10000, "world": 1, "steps": 5, "warmup": 3, "ms_per_step": 17.5, "kernel_ms": 17.49,
         "f_alg": 2630.76, "newton_iters_per_element": 4.6, "peak_tflops": 34.2, "kernel_name": "feas_kernel<1,2> fused chain, shape registered",
         "traffic": 815360, "ncu_pipes": {"fp64_pipe_active_pct": 60.5, "issue_active_pct": 53.9}, "ncu_note": "x",
         "d_dim": 4, "p_dim": 4, "cpu": {"value": 1.2e7, "unit": bench.UNIT, "cores": 16, "kind": "port", "sample": "..."},
         "e2e_s": 0.0177, "e2e_matches": True, "configs": [{"name": "C4"}], "ns_iteration": {"seconds": 0.2},
         "counters": {"launches": 10, "scenarios": 500000000}, "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []},
         "formulation": "blocktri", "state_blocks": []}
    m.update(kw)
    return m


def test_own_arm_line_has_every_contract_key_and_consistent_figures():
    b = bench.build_line(_stub())
    json.dumps(b)                                              # serialisable as it stands
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "configs",
              "ns_iteration"):
        assert k in b, k
    assert b["n_gpus"] == 1 and b["warmup"] >= 3 and b["higher_is_better"] is True and b["scaling"] == "weak"
    assert b["vs_baseline"] is None and b["dtype"] == "f64" and b["data"] == "synthetic" and "workload" in b["config"]
    assert "model" not in b["config"] and "flush" in b["config"]["l2"]
    # value = units of all ranks / max-over-ranks time; roofline = algorithmic flops of one launch / its own time
    assert abs(b["value"] - 1e8 / 17.5e-3) / b["value"] < 1e-12
    r = b["roofline"]
    assert abs(r["achieved"] - 1e8 * 2630.76 / 17.49e-3 / 1e12) < 1e-9 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-15
    assert r["frac_fp64_pipe_ncu"] == 0.605 and r["traffic"] == 815360 and r["unit"] == "TFLOP/s" and r["bound"] == "fp64"
    assert r["algorithmic_bytes"] == 8 * (10000 * 4 + 10000 * 5 + 10000)
    e = b["e2e"]
    assert e["h2d_bytes_per_step"] == 8 * (10000 * 4 + 10000 * 4 + 10000) and e["d2h_bytes_per_step"] == 80000
    assert abs(e["value"] - 1e8 / 0.0177) / e["value"] < 1e-12 and e["value"] <= b["value"]
    assert b["gpu_launches"] == 10
    # N ranks: the whole job's units over the same time
    b4 = bench.build_line(_stub(world=4))
    assert abs(b4["value"] - 4 * b["value"]) / b["value"] < 1e-12 and b4["n_gpus"] == 4
    assert b4["e2e"]["h2d_bytes_per_step"] == 4 * e["h2d_bytes_per_step"]
    # without an ncu capture for the running sources nothing is quoted
    bn = bench.build_line(_stub(traffic=None, ncu_pipes=None))
    assert bn["roofline"]["traffic"] is None and bn["roofline"]["frac_fp64_pipe_ncu"] is None


def test_reference_arm_line_names_the_same_workload():
    args = argparse.Namespace(ref_rows=1000, n_theta=10000, n_d=10000, gpus=1, steps=5, warmup=3, ref_formulation="blocktri")
    r = bench.reference_line(args, 1.25e7, 800.0, 16, 29.1)
    assert r["impl"] == "reference" and r["metric"] == bench.METRIC and r["unit"] == bench.UNIT
    assert r["config"]["workload"] == bench.build_line(_stub())["config"]["workload"]
    assert r["cpu_baseline"]["value"] == r["value"] and r["cpu_baseline"]["kind"] == "port" and r["cpu_baseline"]["cores"] == 16
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["gpu_launches"] == 0 and r["higher_is_better"] is True
    # the two formulations of the CPU arm differ by the flops they need, and say so
    f_tri, f_d9 = r["config"]["flops_per_eval"], bench.reference_line(
        argparse.Namespace(**{**vars(args), "ref_formulation": "dense9"}), 1.8e6, 5500.0, 16, 29.1)["config"]["flops_per_eval"]
    assert 2500 < f_tri < 4000 and 6 < f_d9 / f_tri < 9


def test_committed_ncu_figures_are_stamped_with_the_kernel_sources(tmp_path):
    h = bench.source_hash()
    assert len(h) == 64 and h == bench.source_hash()
    rec = {"kernel": "k", "n_d": 10, "n_theta": 20, "dram_bytes_read": 7, "dram_bytes_write": 1, "fp64_pipe_active_pct": 60.0,
           "source": "ncu ...", "source_sha256": h}
    f = tmp_path / "t.json"
    f.write_text(json.dumps(rec))
    assert bench.committed_ncu("k", 10, 20, str(f))[:2] == (8, {"fp64_pipe_active_pct": 60.0})
    assert bench.committed_ncu("k", 11, 20, str(f))[:2] == (None, None)                 # another shape
    assert bench.committed_ncu("other", 10, 20, str(f))[:2] == (None, None)             # another kernel
    f.write_text(json.dumps({**rec, "source_sha256": "0" * 64}))
    t, pipes, why = bench.committed_ncu("k", 10, 20, str(f))
    assert t is None and pipes is None and "stale" in why                                # the kernels changed since
    assert bench.committed_ncu("k", 10, 20, str(tmp_path / "absent.json"))[0] is None


def test_sens_flops_model():
    from pyds_b200 import models
    lm = models.kinetics("twostage").lower()
    f0, f1 = bench.sens_flops_per_solve(lm, 29.0), bench.sens_flops_per_solve(lm, 30.0)
    fl = lm.tapes_sens.stats["flops"]
    assert f1 - f0 == 3 * fl["pt0"] + 2 * 6 * 4 + 36 + 42                               # one more Newton iteration
    assert 5000 < f0 < 20000


def test_inputs_are_the_frozen_config():
    d, th, w = bench.make_inputs(100, 50)
    assert d.shape == (100, 4) and th.shape == (50, 4) and np.isclose(w.sum(), 1.0)
    assert (d.min(0) >= [250, 250, 1500, 0]).all() and (d.max(0) <= [350, 300, 2500, 2.5]).all()
    np.testing.assert_array_equal(bench.make_inputs(100, 50, rows=10)[0], d[:10])        # bounded samples are prefixes
    assert not np.array_equal(bench.make_inputs(100, 50, rank=1)[0], d)                  # every rank its own design points
