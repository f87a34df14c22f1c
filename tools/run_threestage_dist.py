"""Three-stage (one global control vector) on several GPUs: design points sharded over the ranks, one all-reduce of
the optimiser statistics per iteration (pydsb_recourse_solve_global_dist).  Launch with torchrun; every rank also
solves the whole batch alone and the two answers are compared.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tools/run_threestage_dist.py [--n-d 101] [--n-p 50] [--out file.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from examples import threestage_reactor as ex3  # noqa: E402
from oracle import kinetics as K  # noqa: E402  (theta sample generator only)
from pyds_b200.run import ThreeStageManager  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-d", type=int, default=101)
ap.add_argument("--n-p", type=int, default=50)
ap.add_argument("--out", default=None)
a = ap.parse_args()

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ["PYDSB_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

cfg = ex3.build_config(tempfile.mkdtemp())
cfg["simulator"]["device"] = local
m = ThreeStageManager(cfg)
rng = np.random.default_rng(7)
d = np.stack([rng.uniform(250, 400, a.n_d), rng.uniform(250, 300, a.n_d)], axis=1)
p = K.theta_samples(a.n_p)
w = np.full(a.n_p, 1.0 / a.n_p)

m.efp(d[:4], p, w)                                      # warm-up: engine creation, NCCL communicator
torch.cuda.synchronize()
dist.barrier()
t0 = time.time()
P_dist = m.efp(d, p, w)                                 # sharded rows + all-reduce per iteration
torch.cuda.synchronize()
t_dist = time.time() - t0
z_dist = np.asarray(m.solver.controls).copy()
it_dist = int(m.solver.result["iters"][0])
phi_dist = float(m.solver.result["phi"][0])

m.solver.efp(d[:4], p, w, mode="global")
torch.cuda.synchronize()
t0 = time.time()
P_single = m.solver.efp(d, p, w, mode="global")         # the whole batch on this GPU alone, same device-side path
torch.cuda.synchronize()
t_single = time.time() - t0
z_single = np.asarray(m.solver.controls).copy()
phi_single = float(m.solver.result["phi"][0])

res = {"rank": rank, "world": world, "n_d": a.n_d, "n_p": a.n_p, "seconds_sharded": t_dist, "seconds_single": t_single,
       "iters_sharded": it_dist, "iters_single": int(m.solver.result["iters"][0]), "phi_sharded": phi_dist,
       "phi_single": phi_single, "max_abs_dz": float(np.abs(z_dist - z_single).max()),
       "P_mismatch": int((np.abs(P_dist - P_single) > 1e-12).sum()), "P_mean": float(P_dist.mean())}
allres = [None] * world
dist.all_gather_object(allres, res)
zs = [None] * world
dist.all_gather_object(zs, z_dist.tolist())
if rank == 0:
    same = all(z == zs[0] for z in zs)
    out = {"ranks": allres, "controls_identical_on_all_ranks": same}
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
    ok = same and all(abs(r["phi_sharded"] - r["phi_single"]) <= 1e-6 * max(1.0, abs(r["phi_single"])) + 1e-9 and
                      r["P_mismatch"] == 0 for r in allres)
    if not ok:
        dist.destroy_process_group()
        sys.exit(1)
dist.barrier()
dist.destroy_process_group()
