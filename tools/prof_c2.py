"""Latency breakdown of the three-stage (global control) efp call on one GPU."""
import os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from examples import threestage_reactor as ex3
from oracle import kinetics as K
from pyds_b200.run import ThreeStageManager
m = ThreeStageManager(ex3.build_config(tempfile.mkdtemp()))
rng = np.random.default_rng(7)
for n_d, n_p in ((100, 50), (2000, 50), (1000, 50)):
    d = np.stack([rng.uniform(250, 400, n_d), rng.uniform(250, 300, n_d)], axis=1)
    p = K.theta_samples(n_p); w = np.full(n_p, 1.0 / n_p)
    m.solver.efp(d[:4], p, w, mode="global")
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        P = m.solver.efp(d, p, w, mode="global")
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(n_d, n_p, "efp ms", (t1 - t0) * 1e3, "iters", m.solver.result["iters"], flush=True)
    eng = m.simulator.engine
    dev = torch.device("cuda", 0)
    td, tp, tw = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (d, p, w))
    tz0 = torch.from_numpy(m.simulator.default_controls()[None, :]).to(dev)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = eng.recourse_solve_device(td, tp, tw, tz0, global_group=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        Pt = torch.empty(n_d, dtype=torch.float64, device=dev)
        eng.eval_device(td, tp, tw, z=r["z"].contiguous(), z_mode=1, P_out=Pt, feas_tol=1e-7)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print("   solve ms", (t1 - t0) * 1e3, "eval ms", (t2 - t1) * 1e3, flush=True)
