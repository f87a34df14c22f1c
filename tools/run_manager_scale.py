"""The drop-in managers at grid sizes the reference cannot reach: TwoStageManager.efp on 1e3 x 1e4 (config C4's
shape, one control vector per design point) -- control optimisation, classification and the weighted reduction all on
the device, n_d doubles back."""
import json
import sys
import tempfile
import time
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs  # noqa: E402
from examples import twostage_reactor as ex2  # noqa: E402
from pyds_b200.run import TwoStageManager  # noqa: E402

m = TwoStageManager(ex2.build_config(tempfile.mkdtemp()))
rng = np.random.default_rng(1989)
out = {"device": torch.cuda.get_device_name(0)}
for n_d, n_p in ((150, 50), (1000, 10000), (10000, 1000)):
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    _, p, _ = make_inputs(4, n_p)
    w = np.full(n_p, 1.0 / n_p)
    m.efp(d[:8], p[:64], w[:64] * (n_p / 64))
    torch.cuda.synchronize()
    t0 = time.time()
    P = m.efp(d, p, w)
    t = time.time() - t0
    out[f"{n_d}x{n_p}"] = {"seconds": t, "scenarios": n_d * n_p, "P_mean": float(P.mean()),
                           "lm_iterations_mean": float(np.mean(m.solver.result["iters"]))}
    print(n_d, n_p, out[f"{n_d}x{n_p}"], flush=True)
with open("gpurun_out/manager_scale.json", "w") as f:
    json.dump(out, f, indent=1)
