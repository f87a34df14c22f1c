"""Run BASELINE.json's five configs once on the GPU box and write gpurun_out/configs.json
(development / evidence tool; timings are wall clock around synchronous calls unless stated)."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from bench import make_inputs  # noqa: E402
from examples import threestage_reactor as ex3, twostage_reactor as ex2  # noqa: E402
from examples.reactor_model import theta_samples  # noqa: E402
from pyds_b200 import capi, models, ns  # noqa: E402
from pyds_b200.run import ThreeStageManager, TwoStageManager  # noqa: E402

out = {"device": torch.cuda.get_device_name(0)}

# ---- C1: run_twostage.py as shipped (nested sampling through the DEUS-shaped form) -------------------------------
m = TwoStageManager(ex2.build_config(tempfile.mkdtemp()))
p_best, p_samples = theta_samples()
rng = np.random.default_rng(1989)
d150 = np.stack([rng.uniform(250, 350, 150), rng.uniform(250, 300, 150)], axis=1)
p50 = np.array([s["c"] for s in p_samples])
m.g(d150[:10], p50)                                                     # warm-up (engine creation)
t0 = time.time(); g_list = m.g(d150, p50); t_g = time.time() - t0
t0 = time.time(); P = m.efp(d150, p50, np.full(50, 0.02)); t_efp = time.time() - t0
res = ns.DEUS(ex2.activity_form(m, p_best, p_samples)).solve()
out["C1"] = {"g_call_150x50_s": t_g, "efp_call_150x50_s": t_efp, "evals_per_g_call": 7500,
             "lm_iterations_mean": float(np.mean(m.solver.result["iters"])), "P_mean": float(P.mean()),
             "ns": {k: (float(v) if np.isscalar(v) else None) for k, v in res.items() if k in
                    ("iterations", "evaluations", "inside_fraction", "seconds", "seconds_per_iteration")},
             "live_box": [res["live_points"].min(0).tolist(), res["live_points"].max(0).tolist()]}
print("C1", out["C1"], flush=True)

# ---- C2: run_threestage.py (repaired), one control vector for the whole tree ---------------------------------------
m3 = ThreeStageManager(ex3.build_config(tempfile.mkdtemp()))
d100 = np.stack([rng.uniform(250, 400, 100), rng.uniform(250, 300, 100)], axis=1)
m3.g(d100[:5], p50)
t0 = time.time(); g3 = m3.g(d100, p50); t_g3 = time.time() - t0
form3 = ex3.activity_form(m3, p_best, p_samples)
form3["solver"]["algorithms"]["sampling"]["settings"]["stop_criteria"] = [{"max_iterations": 300}]
res3 = ns.DEUS(form3).solve()
out["C2"] = {"g_call_100x50_s": t_g3, "controls": np.asarray(m3.solver.controls).tolist(),
             "ns": {k: float(res3[k]) for k in ("iterations", "evaluations", "inside_fraction", "seconds", "seconds_per_iteration")}}
print("C2", out["C2"], flush=True)

# ---- C4: synthetic two-stage, 8 recourse controls, 1e3 x 1e4, per-scenario batched optimisation ----------------------
lm2 = models.kinetics("twostage").lower()
eng2 = capi.Engine(lm2.blob, 0)
d1k = np.stack([rng.uniform(250, 350, 1000), rng.uniform(250, 300, 1000)], axis=1)
_, p1e4, _ = make_inputs(10, 10000)
eng2.recourse_solve(d1k[:4], p1e4[:256], per_scenario=True)
torch.cuda.synchronize()
eng2.counters(reset=True)
t0 = time.time(); r4 = eng2.recourse_solve(d1k, p1e4, per_scenario=True, max_iter=50); t4 = time.time() - t0
c4 = eng2.counters()
t0 = time.time(); r4s = eng2.recourse_solve(d1k, p1e4, np.full(10000, 1e-4), z0=np.zeros((1, 8))); t4s = time.time() - t0
out["C4"] = {"per_scenario": {"seconds": t4, "scenarios": 10_000_000, "state_solves": c4["scenarios"], "feasible_frac": float((r4["phi"] == 0).mean()),
                              "status_counts": {int(k): int(v) for k, v in zip(*np.unique(r4["status"], return_counts=True))},
                              "iters_mean": float(r4["iters"].mean()), "iters_max": int(r4["iters"].max()),
                              "state_solves_per_s": c4["scenarios"] / t4},
             "shared": {"seconds": t4s, "iters_mean": float(r4s["iters"].mean()), "phi_mean": float(r4s["phi"].mean())}}
print("C4", out["C4"], flush=True)

# ---- C5: 6 design variables, 1e5 x 1e5, fused P(feasible) only (one GPU here; 8 GPUs shard the rows) -------------------
lm6 = models.kinetics("design6").lower()
eng6 = capi.Engine(lm6.blob, 0)
n = 100_000
d6 = np.stack([rng.uniform(250, 350, n), rng.uniform(250, 300, n), rng.uniform(1500, 2500, n), rng.uniform(0, 2.5, n),
               rng.uniform(90, 110, n), rng.uniform(15, 25, n)], axis=1)
_, p1e5, w1e5 = make_inputs(10, n)
dd, pp, ww = (torch.from_numpy(a).cuda() for a in (d6, p1e5, w1e5))
Pd = torch.empty(n, dtype=torch.float64, device="cuda")
eng6.eval_device(dd[:1000], pp, ww, P_out=Pd[:1000]); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng6.eval_device(dd, pp, ww, P_out=Pd); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
out["C5"] = {"n_d": n, "n_theta": n, "ms": ms, "evals_per_s": n * n / ms * 1e3, "P_mean": float(Pd.mean()),
             "note": "whole 1e10 sweep on ONE B200; on 8 GPUs each rank owns 12500 rows"}
print("C5", out["C5"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
