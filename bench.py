#!/usr/bin/env python
"""bench.py -- (d,theta) constraint evaluations per second, fp64 (BASELINE.json's metric).

A *step* is one pass of the hot path over one batch: the full design-point x theta grid of config
C3 (synthetic reaction kinetics, 4 design variables, 1e4 live points x 1e4 theta samples, no
recourse) per GPU -- state solve (Radau collocation Newton), CQAs, big-M indicator and the
weighted P(feasible) reduction.  With N GPUs every rank owns its own 1e4 design points (weak
scaling, theta replicated) and the step ends with the all-gather of the per-point probabilities.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n-d ND] [--n-theta NT]

Prints ONE JSON line (rank 0).  ``--impl reference`` times the CPU restatement of the reference
path (the reference itself -- Pyomo + GAMS/CONOPT + DEUS -- cannot be installed here, see
DESIGN.md) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "(d,theta) constraint evaluations per second (fp64)"
UNIT = "evals/s"

# stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner under torchrun) are sent
# to stderr, the line itself is written to the saved descriptor
_STDOUT_FD = None


def claim_stdout():
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, (json.dumps(line) + "\n").encode())


def make_inputs(n_d, n_t, rank=0, rows=None):
    """Config C3 inputs: d = (t_f, T, cA0, Fbar) uniform in the box, seed 1989 (+rank);
    theta as run_twostage.py:113-116 with size n_t, seed 12341234; w = 1/n_t.  ``rows`` keeps only the
    first rows of the SAME n_d-point design batch (bounded CPU samples)."""
    rng = np.random.default_rng(1989 + rank)
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d), rng.uniform(1500, 2500, n_d),
                  rng.uniform(0, 2.5, n_d)], axis=1)
    if rows is not None:
        d = d[:rows]
    ran = np.random.default_rng(12341234)
    theta = np.stack([ran.normal(2.5e3, 2.5e1, n_t), ran.normal(5e3, 5e1, n_t), ran.normal(6.409e-2, 6.409e-4, n_t),
                      ran.normal(9.938e3, 9.938e1, n_t)], axis=1)
    w = np.full(n_t, 1.0 / n_t)
    return d, theta, w


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(parts[1])); smax = float(parts[2])
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:      # region shorter than the sampling period: fall back to all samples
            for ts, line in self.lines:
                parts = [p.strip() for p in line.split(",")]
                try:
                    sm.append(float(parts[1])); smax = float(parts[2])
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(n_rows, n_t, threads, tol=1e-10, n_d=10000):
    """Time the C restatement (oracle, kind 'port') on `threads` host threads over n_rows x n_t."""
    from oracle import cbind
    d, theta, w = make_inputs(n_d, n_t, rows=n_rows)
    cbind.kinetics_grid(d[:8], theta[:64], w[:64], tol=tol, want_g=False, threads=threads)      # warm
    t0 = time.perf_counter()
    _, P, it = cbind.kinetics_grid(d, theta, w, tol=tol, want_g=False, threads=threads)
    dt = time.perf_counter() - t0
    return n_rows * n_t / dt, dt, P, it


class _EfpOwner:
    """Adapter handed to the nested-sampling driver: ``g`` is never materialised at this size, the driver calls
    the fused ``efp`` (every rank passes the same proposals; rows are sharded, probabilities all-gathered)."""

    def __init__(self, sharded):
        self.sharded = sharded

    def efp(self, d, p, w):
        return self.sharded.efp_global(d, p, w)

    def g(self, d, p):
        raise NotImplementedError("the bench uses the fused efp path")


def ns_iteration_seconds(sharded, n_live, theta, w, iterations):
    """Seconds per nested-sampling iteration (BASELINE.json's second metric) of pyds_b200.ns at the bench
    configuration: n_live live points, n_live/10 proposals per iteration drawn from the enlarged bounding
    ellipsoid of the live set (host, numpy), their P(feasible) over all theta through the host-buffer path
    (H2D, kernels, D2H, all-gather), live-set update."""
    from pyds_b200 import ns
    owner = _EfpOwner(sharded)
    form = {
        "activity_type": "ds",
        "problem": {"design_variables": [{"t_f": [250.0, 350.0]}, {"T": [250.0, 300.0]}, {"cA0": [1500.0, 2500.0]},
                                         {"Fbar": [0.0, 2.5]}],
                    "parameters_best_estimate": [2.5e3, 5e3, 6.409e-2, 9.938e3],
                    "parameters_samples": [{"c": row, "w": float(wi)} for row, wi in zip(theta, w)],
                    "target_reliability": 0.65},
        "solver": {"name": "ds-ns",
                   "settings": {"efp_evaluation": {"constraints_func_ptr": owner.g},
                                "points_schedule": [(0.0, n_live, max(n_live // 10, 1))],
                                "stop_criteria": [{"inside_fraction": 1.0}]},
                   "algorithms": {"sampling": {"settings": {"prng_seed": 1989, "stop_criteria": [{"max_iterations": iterations}]}}}},
    }
    res = ns.DEUS(form).solve()
    return res["seconds_per_iteration"], res["iterations"], max(n_live // 10, 1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_t = args.n_theta
    rows = args.ref_rows
    from oracle import cbind  # noqa: F401  (builds the C oracle if needed)
    for _ in range(args.warmup):
        cpu_baseline(max(rows // 8, 1), n_t, threads, n_d=args.n_d)
    times = []
    for _ in range(args.steps):
        v, dt, _, _ = cpu_baseline(rows, n_t, threads, n_d=args.n_d)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = rows * n_t / (ms * 1e-3)
    sample = f"{rows} of the {args.n_d} design points x all {n_t} theta per step (same seeds as the GPU arm)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C3 synthetic kinetics, 4 design vars, {args.n_d} x {n_t} (d,theta), no recourse; "
                               f"bounded sample of {rows} design points per step",
                   "reference_impl": "CPU restatement in C + OpenMP (oracle/kinetics_c.c); the Pyomo+GAMS/CONOPT+DEUS "
                                     "reference is not installable in this image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-d", type=int, default=10000, help="design points per GPU")
    ap.add_argument("--n-theta", type=int, default=10000)
    ap.add_argument("--ref-rows", type=int, default=200, help="design points per step of the CPU reference arm")
    ap.add_argument("--cpu-rows", type=int, default=1000, help="design points of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ns-iterations", type=int, default=5, help="nested-sampling iterations timed for the secondary metric")
    ap.add_argument("--dense9", action="store_true", help="keep cC in the Newton system (9 unknowns, SURVEY A.5 flop count)")
    ap.add_argument("--coupled", action="store_true", help="one coupled 6-unknown Newton system instead of the block-triangular solve")
    ap.add_argument("--generic-blocks", action="store_true", help="block-triangular solve through the generic point-op / NS path (no POLY1)")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from pyds_b200 import capi, models
    from pyds_b200.parallel import ShardedEfp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    kw = {}
    if args.coupled:
        kw["block_triangular"] = False
    if args.generic_blocks:
        kw["polynomial_blocks"] = False
    lm = models.kinetics("design4", dense9=args.dense9, **kw).lower()
    eng = capi.Engine(lm.blob, local_rank)
    sharded = ShardedEfp(eng, world, rank)
    n_d, n_t = args.n_d, args.n_theta
    d_np, th_np, w_np = make_inputs(n_d, n_t, rank)
    d = torch.from_numpy(d_np).to(dev)
    th = torch.from_numpy(th_np).to(dev)
    w = torch.from_numpy(w_np).to(dev)
    P_all = torch.empty(world * n_d, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)       # 256 MiB > 126 MB L2

    def step():
        sharded.efp_device(d, th, w, P_all)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    peak_tflops, _ = capi.dfma_peak(local_rank, 5)
    for _ in range(args.warmup):
        step()
    barrier()
    eng.counters(reset=True)

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t0 = time.time()
    for k in range(args.steps):
        flush.fill_(float(k))                         # L2 flush between timed iterations (untimed)
        ev[k][0].record()
        kev[k][0].record()
        sharded.launch_local(d, th, w)                # the feasibility kernel + partial reduction
        kev[k][1].record()
        sharded.gather(P_all)                         # NCCL all-gather of the per-point probabilities
        ev[k][1].record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    cnt = eng.counters()
    evals_per_step = world * n_d * n_t
    value = evals_per_step / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (feas_kernel): algorithmic flops per launch / its event time
    I = cnt["newton_iters"] / max(cnt["scenarios"], 1) / lm.nfe
    f_alg = lm.flops_per_evaluation(I, newton_flops_per_eval=cnt["newton_flops"] / max(cnt["scenarios"], 1))
    k_ms = float(np.mean(kern_ms))
    achieved = n_d * n_t * f_alg / (k_ms * 1e-3) / 1e12
    alg_bytes = 8 * (n_d * lm.d_dim + n_t * (lm.p_dim + 1) + n_d)
    kernel_name = f"feas_kernel<{max(len(s) for s in lm.blocks)},{eng.info['scenarios_per_thread']}>"
    traffic, ncu_pipes = None, None     # per launch, from the committed ncu --set full capture of this config
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_feas_kernel.json")))
        if tj["kernel"] == kernel_name and tj["n_d"] == n_d and tj["n_theta"] == n_t:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            ncu_pipes = {k: tj[k] for k in ("fp64_pipe_active_pct", "issue_active_pct", "shared_wavefronts_pct") if k in tj}
    except (OSError, KeyError, ValueError):
        pass

    # end to end through the host-buffer C-ABI call (pinned staging, H2D, kernels, D2H inside the timed region)
    P_host = None
    for _ in range(2):
        P_host = sharded.efp_host(d_np, th_np, w_np, equal_shards=True)
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        P_host = sharded.efp_host(d_np, th_np, w_np, equal_shards=True)
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - te0) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = evals_per_step / float(e2e_s.item())
    h2d = 8 * (n_d * lm.d_dim + n_t * lm.p_dim + n_t)
    d2h = 8 * n_d
    same = bool(np.allclose(P_host[rank * n_d:(rank + 1) * n_d], P_all[rank * n_d:(rank + 1) * n_d].cpu().numpy(), atol=1e-12))

    ns_s, ns_it, ns_rep = ns_iteration_seconds(sharded, world * n_d, th_np, w_np, args.ns_iterations)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, dt, P_cpu, _ = cpu_baseline(min(args.cpu_rows, n_d), n_t, threads, n_d=n_d)
            agree = bool(np.allclose(P_cpu, P_all[:len(P_cpu)].cpu().numpy(), atol=1e-12))
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"first {args.cpu_rows} of the {n_d} design points x all {n_t} theta ({dt:.1f} s); "
                             f"P(feasible) equal to the GPU result on the sample: {agree}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"C3 synthetic kinetics (A->B->C semi-batch, Radau 8x3), 4 design vars, "
                                   f"{n_d} design points per GPU x {n_t} theta, no recourse; P(feasible) per point",
                       "n_d_per_gpu": n_d, "n_theta": n_t, "sharding": f"design points over {world} GPU(s), theta replicated",
                       "l2": "256 MiB buffer written between timed steps (flush); inputs are < 1 MB, the path is fp64-pipe bound",
                       "newton_iters_per_element": I, "flops_per_eval": f_alg,
                       "state_blocks": [{"states": s, "unknowns": len(s) * lm.ncp, "linear": bool(l)} for s, l in zip(lm.blocks, lm.block_linear)]},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / peak_tflops, "traffic": traffic, "kernel": kernel_name, "ncu": ncu_pipes,
                         "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes,
                         "peak_source": "measured in this run: pydsb_dfma_peak (independent-chain DFMA microkernel, best of 5); "
                                        "MEASURED_PEAKS.json carries no fp64 figure"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "matches_device_path": same},
            "ns_iteration": {"seconds": ns_s, "iterations_timed": ns_it, "n_live": world * n_d, "proposals_per_iteration": ns_rep,
                             "what": "pyds_b200.ns: ellipsoid proposals (host) + P(feasible) of the proposals over all theta "
                                     "through the host-buffer path + live-set update"},
            "gpu_launches": cnt["launches"],
            "clocks": clocks,
            "counters": cnt,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
