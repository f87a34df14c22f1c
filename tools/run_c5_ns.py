"""Config C5 (BASELINE.json configs[4]): nested sampling with 6 design variables, 1e5 live points x 1e5 theta, the
design points sharded over the ranks of one box.  Times a few nested-sampling iterations of pyds_b200.ns (ellipsoid
proposals on the host, fused P(feasible) of the proposals on the GPUs, live-set update).

    python tools/run_c5_ns.py [--n-live 100000] [--n-theta 100000] [--iterations 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_c5_ns.py ...
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import _EfpOwner, make_inputs  # noqa: E402
from pyds_b200 import capi, models, ns  # noqa: E402
from pyds_b200.parallel import ShardedEfp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n-live", type=int, default=100000)
ap.add_argument("--n-theta", type=int, default=100000)
ap.add_argument("--iterations", type=int, default=3)
args = ap.parse_args()

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = capi.Engine(models.kinetics("design6").lower().blob, local)
sharded = ShardedEfp(eng, world, rank)
_, theta, w = make_inputs(10, args.n_theta)
owner = _EfpOwner(sharded)
n_rep = max(args.n_live // 10, 1)
form = {
    "activity_type": "ds",
    "problem": {"design_variables": [{"t_f": [250.0, 350.0]}, {"T": [250.0, 300.0]}, {"cA0": [1500.0, 2500.0]}, {"Fbar": [0.0, 2.5]},
                                     {"pB": [90.0, 110.0]}, {"pA": [15.0, 25.0]}],
                "parameters_best_estimate": [2.5e3, 5e3, 6.409e-2, 9.938e3],
                "parameters_samples": [{"c": row, "w": float(wi)} for row, wi in zip(theta, w)],
                "target_reliability": 0.65},
    "solver": {"name": "ds-ns",
               "settings": {"efp_evaluation": {"constraints_func_ptr": owner.g},
                            "points_schedule": [(0.0, args.n_live, n_rep)], "stop_criteria": [{"inside_fraction": 1.0}]},
               "algorithms": {"sampling": {"settings": {"prng_seed": 1989, "stop_criteria": [{"max_iterations": args.iterations}]}}}},
}
t0 = time.time()
res = ns.DEUS(form).solve()
wall = time.time() - t0
if rank == 0:
    print(json.dumps({"config": "C5", "n_gpus": world, "n_live": args.n_live, "n_theta": args.n_theta, "proposals_per_iteration": n_rep,
                      "iterations": res["iterations"], "seconds_per_iteration": res["seconds_per_iteration"],
                      "evals_per_iteration": n_rep * args.n_theta,
                      "initial_live_set_seconds": wall - res["seconds_per_iteration"] * res["iterations"],
                      "inside_fraction": res["inside_fraction"], "evaluations": res["evaluations"]}), flush=True)
if world > 1:
    dist.destroy_process_group()
