"""Config C4 alone (synthetic two-stage, 8 recourse controls, 1e3 x 1e4, per-scenario batched optimisation) --
development tool for profiling the recourse kernels.  usage: run_c4.py [n_d] [n_theta]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_inputs  # noqa: E402
from pyds_b200 import capi, models  # noqa: E402

n_d = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_t = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
rng = np.random.default_rng(1989)
eng = capi.Engine(models.kinetics("twostage").lower().blob, 0)
d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
_, p, _ = make_inputs(10, n_t)
if not os.environ.get("PYDSB_NO_WARM"):
    eng.recourse_solve(d[:4], p[:256], per_scenario=True)
torch.cuda.synchronize()
eng.counters(reset=True)
td, tp = torch.from_numpy(d).cuda(), torch.from_numpy(p).cuda()
torch.cuda.synchronize()
t0 = time.time(); rd = eng.recourse_solve_device(td, tp, per_scenario=True, max_iter=50); t_dev = time.time() - t0
c = eng.counters()
del rd
if os.environ.get("PYDSB_C4_PROFILE"):               # CUDA events around every kernel of a second, warm solve
    eng.set_profiling(True); eng.profile(reset=True)
    t0 = time.time(); rd = eng.recourse_solve_device(td, tp, per_scenario=True, max_iter=50); t_w = time.time() - t0
    print({"warm_seconds": t_w, **eng.profile()}, flush=True)
    eng.set_profiling(False)
    del rd
t0 = time.time(); r = eng.recourse_solve(d, p, per_scenario=True, max_iter=50); t = time.time() - t0
print({"seconds_device_resident": t_dev, "seconds": t, "scenarios": n_d * n_t, "state_solves": c["scenarios"], "launches": c["launches"],
       "feasible_frac": float((r["phi"] == 0).mean()), "iters_mean": float(r["iters"].mean()), "iters_max": int(r["iters"].max()),
       "state_solves_per_s": c["scenarios"] / t_dev}, flush=True)
t0 = time.time(); rs = eng.recourse_solve(d, p, np.full(n_t, 1.0 / n_t), z0=np.zeros((1, 8))); ts = time.time() - t0
print({"shared_seconds": ts, "iters_mean": float(rs["iters"].mean()), "phi_mean": float(rs["phi"].mean())}, flush=True)
