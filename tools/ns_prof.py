"""Where a device-resident nested-sampling iteration of config C5 spends its time (development tool; PYDSB_NS_TRACE=1)."""
import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
from bench import make_inputs, C5_BOX, _EfpOwner
from pyds_b200 import capi, models, ns
from pyds_b200.parallel import ShardedEfp
eng6 = capi.Engine(models.kinetics("design6").lower().blob, 0)
sh6 = ShardedEfp(eng6, 1, 0)
_, th5, w5 = make_inputs(10, 100000)
owner = _EfpOwner(sh6)
def form(n_live, th, w, its, dev):
    return {"activity_type": "ds", "problem": {"design_variables": C5_BOX, "parameters_best_estimate": [2.5e3, 5e3, 6.409e-2, 9.938e3],
            "parameters_samples": [{"c": row, "w": float(wi)} for row, wi in zip(th, w)], "target_reliability": 0.65},
            "solver": {"name": "ds-ns", "settings": {"efp_evaluation": {"constraints_func_ptr": owner.g}, "points_schedule": [(0.0, n_live, n_live // 10)],
                       "stop_criteria": [{"inside_fraction": 1.0}]},
                       "algorithms": {"sampling": {"settings": {"prng_seed": 1989, "device": dev, "stop_criteria": [{"max_iterations": its}]}}}}}
ns.DEUS(form(2000, th5[:2000], np.full(2000, 1 / 2000), 2, "cuda:0")).solve()
for n_theta in (100000, 1000):
    r = ns.DEUS(form(100000, th5[:n_theta], np.full(n_theta, 1 / n_theta), 4, "cuda:0")).solve()
    print(n_theta, {k: r[k] for k in ("seconds_per_iteration", "seconds_per_iteration_evaluation", "initial_live_set_seconds")}, flush=True)
