"""Throughput with the reference's own theta-set size (50 samples, run_twostage.py:110-120) and many design points:
tiles shared by several design points (feas_kernel<..., PACK>).  usage: run_small_theta.py [n_d] [n_theta]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs  # noqa: E402
from pyds_b200 import capi, models  # noqa: E402

n_d = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
out = {}
eng = capi.Engine(models.kinetics("design4").lower().blob, 0)
for n_t in ([int(sys.argv[2])] if len(sys.argv) > 2 else [50, 100, 128, 129, 256]):
    d, p, w = make_inputs(n_d, n_t)
    td, tp, tw = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (d, p, w))
    P = torch.empty(n_d, dtype=torch.float64, device="cuda")
    for _ in range(2):
        eng.eval_device(td, tp, tw, P_out=P)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.eval_device(td, tp, tw, P_out=P)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[n_t] = {"n_d": n_d, "n_theta": n_t, "ms": ms, "evals_per_s": n_d * n_t / (ms * 1e-3), "P_mean": float(P.mean())}
    print(out[n_t], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/small_theta.json", "w") as f:
    json.dump(out, f, indent=1)
