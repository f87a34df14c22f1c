#!/usr/bin/env python
"""bench.py -- (d,theta) constraint evaluations per second, fp64 (BASELINE.json's metric).

A *step* is one pass of the hot path over one batch: the full design-point x theta grid of config
C3 (synthetic reaction kinetics, 4 design variables, 1e4 live points x 1e4 theta samples, no
recourse) per GPU -- state solve (Radau collocation Newton), CQAs, big-M indicator and the
weighted P(feasible) reduction.  With N GPUs every rank owns its own 1e4 design points (weak
scaling, theta replicated) and the step ends with the all-gather of the per-point probabilities.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n-d ND] [--n-theta NT]

Prints ONE JSON line (rank 0).  Besides the contract keys it carries
  configs[]    the other BASELINE.json configs measured in the same run: C4 (1e3 x 1e4, 8 recourse controls per scenario:
               sens_kernel + lm_update_kernel with their own roofline), C1 (run_twostage.py's g(150 x 50) through the
               manager), C3 at FIXED total work (strong scaling) and the C5 nested-sampling iteration (6 design variables,
               1e5 live points x 1e5 theta sharded over the ranks);
  ns_iteration seconds per nested-sampling iteration at config C5 (BASELINE.json's second metric).
``--impl reference`` times the CPU restatement of the reference path (the reference itself -- Pyomo + GAMS/CONOPT +
DEUS -- cannot be installed here) on the host cores, on a bounded sample of the same workload, in the
formulation ``--ref-formulation`` names (blocktri: the equations as the GPU solves them; dense9: the 9-unknown Newton
system SURVEY.md A.5 counts).
"""
import argparse
import hashlib
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


def workload(n_d, n_t):
    """The workload string of BOTH arms (the driver compares them)."""
    return (f"C3 synthetic kinetics (A->B->C semi-batch, Radau 8x3), 4 design vars, {n_d} design points per GPU x "
            f"{n_t} theta, no recourse; P(feasible) per point")


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


FEAS_KERNEL_SOURCES = ["chain_shapes.h", "common.cuh", "feas_kernel.cuh", "poly_chain.cuh", "tape_vm.cuh", "xvm_ops.cuh"]


def source_hash():
    """sha256 over the kernel sources: the committed ncu figures (profiles/traffic_feas_kernel.json) are only quoted
    while the kernels they were captured from are the ones that run."""
    h = hashlib.sha256()
    src = os.path.join(ROOT, "pyds_b200", "csrc")
    for name in FEAS_KERNEL_SOURCES:                   # device code of feas_kernel (the host side may change freely)
        h.update(name.encode())
        h.update(open(os.path.join(src, name), "rb").read())
    return h.hexdigest()


def committed_ncu(kernel_name, n_d, n_t, path=None, current_hash=None):
    """(traffic bytes per launch, pipe figures, note) from the committed ncu capture, or (None, None, why) when it was
    taken from other kernel sources or another shape."""
    path = path or os.path.join(ROOT, "profiles", "traffic_feas_kernel.json")
    try:
        tj = json.load(open(path))
    except (OSError, ValueError):
        return None, None, "no committed capture"
    if tj.get("kernel") != kernel_name or tj.get("n_d") != n_d or tj.get("n_theta") != n_t:
        return None, None, "committed capture is of another kernel / shape"
    if tj.get("source_sha256") != (current_hash or source_hash()):
        return None, None, "kernel sources changed since the committed capture (stale: not quoted)"
    pipes = {k: tj[k] for k in ("fp64_pipe_active_pct", "issue_active_pct", "shared_wavefronts_pct") if k in tj}
    return tj["dram_bytes_read"] + tj["dram_bytes_write"], pipes, tj.get("source")


def sens_flops_per_solve(lm, newton_iters_per_solve):
    """Algorithmic flops of ONE state + control-sensitivity solve of sens_kernel (SURVEY.md 8d convention applied to the
    algorithm that runs: FMA = 2, a division 1).  Per Newton iteration of an element: the point tape at the three
    points, the linear collocation residual 2 n (ncp + 1), the factorisation and one solve; once more per element at the
    converged point (the factors the sensitivities reuse); per element and control that has acted so far: a
    right-hand side (2 n) and a solve, plus the quadrature sensitivities.  The Newton system of the kinetics is block
    lower triangular (ns systems of 3 unknowns with the coupling substituted)."""
    fl = lm.tapes_sens.stats["flops"]
    ns, ncp, nfe = lm.ns, lm.ncp, lm.nfe
    n = ns * ncp
    factor = ns * (2 * 27) // 3
    solve = ns * 2 * 9 + ns * (ns - 1) // 2 * 2 * ncp
    per_iter = ncp * fl["pt0"] + 2 * n * (ncp + 1) + factor + solve
    per_elem = ncp * fl["pt0"] + 2 * n * (ncp + 1) + factor
    live = sum(len(set(int(z) for z in lm.cmap[:e + 1].reshape(-1))) for e in range(nfe))     # controls acted so far
    per_sens = 2 * n + solve + 2 * ncp * (lm.nq + 1)
    return fl["pro"] + fl["g"] + nfe * (fl["elem"] + per_elem) + newton_iters_per_solve * per_iter + live * per_sens


def rsens_flops_per_solve(lm, newton_iters_per_solve):
    """The same count for the chain evaluator (rsens_kernel, csrc/recourse_chain.cuh): the state solve of the fused chain
    (Newton iterations of the nonlinear blocks x their algorithmic flops, one inverse per scenario and a matrix-vector
    product per element for a constant-decay block), the LU factors at the converged points (2/3 n^3 + the diagonal, once
    per element and Newton block), and per element and control that has acted so far, per block: the derivative of the
    right-hand side (4 flops per point), the right-hand side (2 n) and a solve (2 n^2); the quadrature sensitivities
    (2 ncp + 2 each); the gradient assembly 2 ng nz (nb + nq)."""
    fl = (lm.tapes_chain or lm.tapes).stats["flops"]
    ncp, nfe, nb, nq, n = lm.ncp, lm.nfe, len(lm.blocks), lm.nq, lm.ncp
    newton_blocks = [b for b in range(nb) if not lm.block_linear[b]]
    const_blocks = nb - len(newton_blocks)
    state = newton_iters_per_solve * lm.block_flops[newton_blocks[0]] if newton_blocks else 0.0
    state += len(newton_blocks) * nfe * ((2 * n ** 3) // 3 + 2 * n) + const_blocks * (72 + nfe * (2 * n + 2 * n * n))
    state += nfe * (fl["elem"] + sum(lm.block_fixed_flops) + ncp * fl.get("ptq", 0))
    live = sum(len(set(int(z) for z in lm.cmap[:e + 1].reshape(-1))) for e in range(nfe))
    per_pair = nb * (4 * ncp + 2 * n + 2 * n * n) + nq * (2 * ncp + 2)
    return fl["pro"] + fl["g"] + state + live * per_pair + 2 * lm.ng * lm.nz * (nb + nq)


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


def cpu_baseline(n_rows, n_t, threads, tol=1e-10, n_d=10000, formulation="blocktri"):
    """Time the C restatement (oracle, kind 'port') on `threads` host threads over n_rows x n_t.
    Returns (evals/s, seconds, P, Newton iterations per evaluation)."""
    from oracle import cbind
    d, theta, w = make_inputs(n_d, n_t, rows=n_rows)
    cbind.kinetics_grid(d[:8], theta[:64], w[:64], tol=tol, want_g=False, threads=threads, formulation=formulation)      # warm
    t0 = time.perf_counter()
    _, P, it = cbind.kinetics_grid(d, theta, w, tol=tol, want_g=False, threads=threads, formulation=formulation)
    dt = time.perf_counter() - t0
    return n_rows * n_t / dt, dt, P, abs(it) / float(n_rows * n_t)


class _EfpOwner:
    """Adapter handed to the nested-sampling driver: ``g`` is never materialised at this size, the driver calls
    the fused ``efp`` (every rank passes the same proposals; rows are sharded, probabilities all-gathered)."""

    def __init__(self, sharded):
        self.sharded = sharded

    def efp(self, d, p, w):
        return self.sharded.efp_global(d, p, w)

    def efp_device(self, d, p, w):
        return self.sharded.efp_global_device(d, p, w)

    def g(self, d, p):
        raise NotImplementedError("the bench uses the fused efp path")


C5_BOX = [{"t_f": [250.0, 350.0]}, {"T": [250.0, 300.0]}, {"cA0": [1500.0, 2500.0]}, {"Fbar": [0.0, 2.5]},
          {"pB": [90.0, 110.0]}, {"pA": [15.0, 25.0]}]


def ns_iteration_seconds(sharded, design_box, n_live, theta, w, iterations, device=None):
    """Seconds per nested-sampling iteration (BASELINE.json's second metric) of pyds_b200.ns: n_live live points,
    n_live/10 proposals per iteration from the enlarged bounding ellipsoid of the live set, their P(feasible) over all
    theta sharded over the ranks, live-set update.  Returns (median s per iteration, iterations, proposals, s of the
    initial live-set evaluation)."""
    from pyds_b200 import ns
    owner = _EfpOwner(sharded)
    n_rep = max(n_live // 10, 1)
    form = {
        "activity_type": "ds",
        "problem": {"design_variables": design_box,
                    "parameters_best_estimate": [2.5e3, 5e3, 6.409e-2, 9.938e3],
                    "parameters_samples": [{"c": row, "w": float(wi)} for row, wi in zip(theta, w)],
                    "target_reliability": 0.65},
        "solver": {"name": "ds-ns",
                   "settings": {"efp_evaluation": {"constraints_func_ptr": owner.g},
                                "points_schedule": [(0.0, n_live, n_rep)],
                                "stop_criteria": [{"inside_fraction": 1.0}]},
                   "algorithms": {"sampling": {"settings": {"prng_seed": 1989, "stop_criteria": [{"max_iterations": iterations}]}}}},
    }
    if device is not None:
        form["solver"]["algorithms"]["sampling"]["settings"]["device"] = device
    res = ns.DEUS(form).solve()
    # the median: the first iteration of a run also pays one-time allocator growth
    per_it = float(np.median(res["iteration_seconds"])) if res["iteration_seconds"] else 0.0
    return per_it, res["iterations"], n_rep, res["initial_live_set_seconds"]


def reference_line(args, value, ms, threads, iters_per_eval):
    from oracle import cbind
    rows, n_t = args.ref_rows, args.n_theta
    sample = f"{rows} of the {args.n_d} design points x all {n_t} theta per step (same seeds as the GPU arm)"
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload(args.n_d, n_t), "n_d_per_gpu": args.n_d, "n_theta": n_t,
                   "sample_per_step": sample,
                   "reference_impl": "CPU restatement in C + OpenMP (oracle/kinetics_c.c); the Pyomo+GAMS/CONOPT+DEUS "
                                     "reference is not installable in this image",
                   "formulation": args.ref_formulation,
                   "newton_iters_per_eval": iters_per_eval,
                   "flops_per_eval": cbind.flops_per_eval(args.ref_formulation, iters_per_eval)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_t = args.n_theta
    rows = args.ref_rows
    from oracle import cbind  # noqa: F401  (builds the C oracle if needed)
    for _ in range(args.warmup):
        cpu_baseline(max(rows // 8, 1), n_t, threads, n_d=args.n_d, formulation=args.ref_formulation)
    times, its = [], []
    for _ in range(args.steps):
        v, dt, _, it = cpu_baseline(rows, n_t, threads, n_d=args.n_d, formulation=args.ref_formulation)
        times.append(dt); its.append(it)
    ms = 1e3 * float(np.mean(times))
    emit(reference_line(args, rows * n_t / (ms * 1e-3), ms, threads, float(np.mean(its))))


def build_line(m):
    """The JSON line of the B200 arm from the measurements in ``m`` (a pure function: tests/test_bench_contract_cpu.py
    runs it on stub numbers)."""
    n_d, n_t, world = m["n_d"], m["n_t"], m["world"]
    evals_per_step = world * n_d * n_t
    value = evals_per_step / (m["ms_per_step"] * 1e-3)
    achieved = n_d * n_t * m["f_alg"] / (m["kernel_ms"] * 1e-3) / 1e12
    pipe = (m.get("ncu_pipes") or {}).get("fp64_pipe_active_pct")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": m["steps"], "warmup": m["warmup"],
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload(n_d, n_t),
                   "n_d_per_gpu": n_d, "n_theta": n_t, "sharding": f"design points over {world} GPU(s), theta replicated",
                   "l2": "256 MiB buffer written between timed steps (flush); inputs are < 1 MB, the path is fp64-pipe bound",
                   "newton_iters_per_element": m["newton_iters_per_element"], "flops_per_eval": m["f_alg"],
                   "formulation": m["formulation"], "state_blocks": m["state_blocks"]},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": m["peak_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / m["peak_tflops"], "frac_fp64_pipe_ncu": None if pipe is None else pipe / 100.0,
                     "traffic": m.get("traffic"), "kernel": m["kernel_name"], "ncu": m.get("ncu_pipes"),
                     "ncu_note": m.get("ncu_note"), "kernel_ms": m["kernel_ms"],
                     "algorithmic_bytes": 8 * (n_d * m["d_dim"] + n_t * (m["p_dim"] + 1) + n_d),
                     "peak_source": "measured in this run: pydsb_dfma_peak (independent-chain DFMA microkernel, best of 5); "
                                    "MEASURED_PEAKS.json carries no fp64 figure"},
        "cpu_baseline": m.get("cpu"),
        "e2e": {"value": evals_per_step / m["e2e_s"], "unit": UNIT,
                "h2d_bytes_per_step": 8 * (n_d * m["d_dim"] + n_t * m["p_dim"] + n_t) * world,
                "d2h_bytes_per_step": 8 * n_d * world, "matches_device_path": m["e2e_matches"]},
        "configs": m.get("configs", []),
        "ns_iteration": m.get("ns_iteration"),
        "gpu_launches": m["counters"]["launches"],
        "clocks": m["clocks"],
        "counters": m["counters"],
    }
    return line


def c4_section(torch, capi, models, dev, peak_tflops, n_d=1000, n_t=10000):
    """Config C4: synthetic two-stage model, 8 recourse controls re-optimised PER SCENARIO, 1e3 design points x 1e4 theta
    (batched Levenberg-Marquardt on the smoothed hinge: sens_kernel + lm_update_kernel), then the shared-control
    (reference-shaped) solve of the same grid."""
    lm = models.kinetics("twostage").lower()
    eng = capi.Engine(lm.blob, dev.index)
    rng = np.random.default_rng(1989)
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    _, p, w = make_inputs(10, n_t)
    td, tp, tw = (torch.from_numpy(a).to(dev) for a in (d, p, w))
    eng.recourse_solve_device(td[:4].contiguous(), tp[:256].contiguous(), per_scenario=True)             # warm-up
    eng.recourse_solve_device(td, tp, per_scenario=True, max_iter=2)     # ... and the workspace of the full batch (kept by the model)
    torch.cuda.synchronize()
    eng.set_profiling(True)
    eng.counters(reset=True); eng.profile(reset=True)
    t0 = time.perf_counter()
    r = eng.recourse_solve_device(td, tp, per_scenario=True, max_iter=50)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    c, pr = eng.counters(), eng.profile()
    solves = c["scenarios"]
    iters = c["newton_iters"] / max(solves, 1)
    chain = eng.info.get("recourse_chain_blocks", 0) > 0
    f_alg = rsens_flops_per_solve(lm, iters) if chain else sens_flops_per_solve(lm, iters)
    kernel = f"rsens_kernel<{eng.info['recourse_chain_blocks']}> (chain, sensitivities in registers)" if chain else f"sens_kernel<{lm.ns},tri>"
    achieved = solves * f_alg / (pr["sens_ms"] * 1e-3) / 1e12
    # classification at the optimised controls, and how many scenarios sit inside the relaxed big-M tolerance band
    g = torch.empty((n_d, n_t, lm.ng), dtype=torch.float64, device=dev)
    eng.eval_device(td, tp, None, z=r["z"].view(n_d, n_t, lm.nz).contiguous(), z_mode=capi.Z_PER_SCENARIO, g_out=g)
    cmax = (g / torch.tensor(lm.big_m, device=dev)).amax(dim=-1)
    band = int(((cmax > 0) & (cmax <= 1e-7)).sum().item())
    feas0 = int((cmax <= 0).sum().item())
    out = {"name": "C4", "workload": f"synthetic two-stage kinetics, 8 recourse controls per scenario, {n_d} x {n_t} (d,theta), "
                                     "batched Levenberg-Marquardt on the smoothed hinge, iteration cap 50",
           "value": n_d * n_t / secs, "unit": "scenario optimisations/s", "seconds": secs,
           "state_sensitivity_solves": solves, "solves_per_scenario": solves / (n_d * n_t),
           "solves_per_s": solves / secs, "newton_iters_per_solve": iters,
           "kernels": {"evaluator": kernel, "sens_kernel_ms": pr["sens_ms"], "sens_launches": int(pr["sens_launches"]),
                       "lm_update_kernel_ms": pr["lm_ms"], "lm_launches": int(pr["lm_launches"]),
                       "lm_share_of_kernel_time": pr["lm_ms"] / max(pr["sens_ms"] + pr["lm_ms"], 1e-12),
                       "kernel_time_share_of_wall": (pr["sens_ms"] + pr["lm_ms"]) * 1e-3 / secs},
           "roofline": {"bound": "fp64", "kernel": kernel, "kernel_ms": pr["sens_ms"],
                        "flops_per_solve": f_alg, "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                        "frac": achieved / peak_tflops},
           "classification": {"feasible_at_optimum": feas0, "inside_1e-7_band": band, "scenarios": n_d * n_t,
                              "status_counts": {int(k): int(v) for k, v in zip(*np.unique(r["status"].cpu().numpy(), return_counts=True))}}}
    del g, r
    # the reference's own formulation: ONE control vector per design point, shared by every theta
    eng.counters(reset=True); eng.profile(reset=True)
    z0 = torch.zeros((1, lm.nz), dtype=torch.float64, device=dev)
    t0 = time.perf_counter()
    rs = eng.recourse_solve_device(td, tp, tw, z0)
    torch.cuda.synchronize()
    secs_s = time.perf_counter() - t0
    cs, ps = eng.counters(), eng.profile()
    out["shared_control"] = {"what": "one control vector per design point (reference pyds/run.py:55-67), same grid",
                             "seconds": secs_s, "design_points_per_s": n_d / secs_s, "evaluations_per_design_point": float(rs["iters"].double().mean().item()),
                             "state_sensitivity_solves": cs["scenarios"], "sens_kernel_ms": ps["sens_ms"], "lm_update_kernel_ms": ps["lm_ms"]}
    eng.set_profiling(False)
    return out


def c4_sharded_section(torch, capi, models, dev, world, rank, barrier, max_over_ranks, n_d=1000, n_t=10000):
    """Config C4 with its design points sharded over the ranks (every (d,theta) problem is independent: no exchange), FIXED
    total work: max-over-ranks wall time of the sharded solve against this rank's own solve of the whole grid."""
    from pyds_b200.parallel import shard_rows
    lm = models.kinetics("twostage").lower()
    eng = capi.Engine(lm.blob, dev.index)
    rng = np.random.default_rng(1989)
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    _, p, _ = make_inputs(10, n_t)
    a, b = shard_rows(n_d, world, rank)
    td, tp = torch.from_numpy(d).to(dev), torch.from_numpy(p).to(dev)
    tds = td[a:b].contiguous()
    eng.recourse_solve_device(td, tp, per_scenario=True, max_iter=2)      # warm-up + the workspace of the full batch
    torch.cuda.synchronize()

    def timed(rows):
        barrier()
        t0 = time.perf_counter()
        eng.recourse_solve_device(rows, tp, per_scenario=True, max_iter=50)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    full = min(timed(td) for _ in range(2))
    part = max_over_ranks(min(timed(tds) for _ in range(2)))
    full = max_over_ranks(full)
    return {"name": "C4-fixed", "workload": f"C4 with FIXED total work: {n_d} x {n_t} per-scenario 8-control problems in all, design "
                                            f"points sharded over {world} GPU(s), no exchange (wall time, max over ranks)",
            "seconds": part, "value": n_d * n_t / part, "unit": "scenario optimisations/s", "scaling": "strong",
            "single_gpu_seconds": full, "strong_efficiency": full / (world * part)}


def c1_section(np_, tempfile_mod):
    """Config C1: the shipped two-stage script's call shape -- g(150 design points x 50 theta) through TwoStageManager,
    control optimisation included (what DEUS calls once per nested-sampling iteration)."""
    from examples import twostage_reactor as ex2
    from examples.reactor_model import theta_samples
    from pyds_b200.run import TwoStageManager
    m = TwoStageManager(ex2.build_config(tempfile_mod.mkdtemp()))
    _, p_samples = theta_samples()
    rng = np_.random.default_rng(1989)
    d150 = np_.stack([rng.uniform(250, 350, 150), rng.uniform(250, 300, 150)], axis=1)
    p50 = np_.array([s["c"] for s in p_samples])
    m.g(d150[:10], p50)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); m.g(d150, p50); ts.append(time.perf_counter() - t0)
    te = []
    for _ in range(5):
        t0 = time.perf_counter(); P = m.efp(d150, p50, np_.full(50, 0.02)); te.append(time.perf_counter() - t0)
    res = getattr(m.solver, "result", {}) or {}
    return {"name": "C1", "workload": "run_twostage.py as shipped: TwoStageManager.g(150 design points x 50 theta), one control "
                                      "vector optimised per design point",
            "g_call_ms": 1e3 * float(np_.median(ts)), "efp_call_ms": 1e3 * float(np_.median(te)),
            "evals_per_call": 7500, "P_mean": float(np_.mean(P)),
            "inside_band": res.get("inside_band"), "lm_evaluations_mean": float(np_.mean(res["iters"])) if "iters" in res else None}


def c2_section(np_, tempfile_mod):
    """Config C2: run_threestage.py's call shape -- ThreeStageManager.g(100 design points x 50 theta), ONE control vector for
    the whole tree optimised over all 5 000 scenarios (reference pyds/run.py:76-98) -- and the same at 2 000 design points."""
    from examples import threestage_reactor as ex3
    from examples.reactor_model import theta_samples
    from pyds_b200.run import ThreeStageManager
    m = ThreeStageManager(ex3.build_config(tempfile_mod.mkdtemp()))
    _, p_samples = theta_samples()
    p50 = np_.array([s["c"] for s in p_samples])
    rng = np_.random.default_rng(1989)
    out = {"name": "C2", "workload": "run_threestage.py as repaired in examples/threestage_reactor.py: ThreeStageManager.g(n_d design points "
                                     "x 50 theta), one control vector for the whole tree (global group)"}
    for n_d in (100, 2000):
        d = np_.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
        m.g(d[:8], p50)
        tg, te = [], []
        for _ in range(3):
            t0 = time.perf_counter(); m.g(d, p50); tg.append(time.perf_counter() - t0)
        for _ in range(3):
            t0 = time.perf_counter(); P = m.efp(d, p50, np_.full(50, 0.02)); te.append(time.perf_counter() - t0)
        res = getattr(m.solver, "result", {}) or {}
        out[f"n_d_{n_d}"] = {"g_call_ms": 1e3 * float(np_.median(tg)), "efp_call_ms": 1e3 * float(np_.median(te)),
                             "scenarios": n_d * 50, "P_mean": float(np_.mean(P)),
                             "lm_evaluations": int(np_.max(res["iters"])) if "iters" in res else None}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-d", type=int, default=10000, help="design points per GPU")
    ap.add_argument("--n-theta", type=int, default=10000)
    ap.add_argument("--ref-rows", type=int, default=1000, help="design points per step of the CPU reference arm")
    ap.add_argument("--ref-formulation", default="blocktri", choices=["blocktri", "dense9"],
                    help="CPU arm: the equations as the GPU solves them (block triangular) or the dense 9-unknown Newton system")
    ap.add_argument("--cpu-rows", type=int, default=1000, help="design points of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C2 / C4 / strong-scaling / C5 sections")
    ap.add_argument("--ns-iterations", type=int, default=5, help="nested-sampling iterations timed (config C5; the median is reported)")
    ap.add_argument("--ns-live", type=int, default=100000, help="live points of the C5 nested-sampling section")
    ap.add_argument("--ns-theta", type=int, default=100000)
    ap.add_argument("--ns-host", action="store_true", help="nested sampling with the host-side sampler instead of the device-resident one")
    ap.add_argument("--dense9", action="store_true", help="keep cC in the Newton system (9 unknowns, SURVEY A.5 flop count)")
    ap.add_argument("--coupled", action="store_true", help="one coupled 6-unknown Newton system instead of the block-triangular solve")
    ap.add_argument("--generic-blocks", action="store_true", help="block-triangular solve through the generic point-op / NS path (no POLY1)")
    ap.add_argument("--interpreted-loop", action="store_true", help="POLY1 per block and element instead of the fused chain")
    args = ap.parse_args()
    claim_stdout()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args)
        return

    import tempfile

    import torch
    import torch.distributed as dist
    from pyds_b200 import capi, models
    from pyds_b200.parallel import ShardedEfp, shard_rows

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
    if args.interpreted_loop:
        kw["fused_chain"] = False
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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    ms_per_step = max_over_ranks(sum(step_ms)) / args.steps
    cnt = eng.counters()

    # roofline of the dominant kernel (feas_kernel): algorithmic flops per launch / its event time
    I = cnt["newton_iters"] / max(cnt["scenarios"], 1) / lm.nfe
    f_alg = lm.flops_per_evaluation(I, newton_flops_per_eval=cnt["newton_flops"] / max(cnt["scenarios"], 1))
    k_ms = float(np.mean(kern_ms))
    info = capi.assemble(lm.blob)[1]
    kernel_name = f"feas_kernel<{max(len(s) for s in lm.blocks)},{eng.info['scenarios_per_thread']}>" + \
                  (f" fused chain, shape {'registered' if info['chain_shape_registered'] else 'generic'}" if info["chain_blocks"] else "")
    traffic, ncu_pipes, ncu_note = committed_ncu(kernel_name, n_d, n_t)

    # end to end through the host-buffer C-ABI call (pinned staging, H2D, kernels, D2H inside the timed region)
    P_host = None
    for _ in range(2):
        P_host = sharded.efp_host(d_np, th_np, w_np, equal_shards=True)
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        P_host = sharded.efp_host(d_np, th_np, w_np, equal_shards=True)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - te0) / args.steps)
    same = bool(np.allclose(P_host[rank * n_d:(rank + 1) * n_d], P_all[rank * n_d:(rank + 1) * n_d].cpu().numpy(), atol=1e-12))

    configs = []
    ns_entry = None
    if not args.no_configs:
        # ---- C3 at FIXED total work: the same n_d x n_t batch with its rows sharded over the ranks (strong scaling) ----
        a, b = shard_rows(n_d, world, rank)
        d0 = torch.from_numpy(make_inputs(n_d, n_t, 0)[0][a:b].copy()).to(dev)
        Pl = torch.empty(max(b - a, 1), dtype=torch.float64, device=dev)
        for _ in range(3):
            eng.eval_device(d0, th, w, P_out=Pl)
        barrier()
        fe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k in range(args.steps):
            flush.fill_(float(k))
            fe[k][0].record(); eng.eval_device(d0, th, w, P_out=Pl); fe[k][1].record()
        barrier()
        fixed_ms = max_over_ranks(sum(x.elapsed_time(y) for x, y in fe)) / args.steps
        configs.append({"name": "C3-fixed", "workload": f"C3 with FIXED total work: {n_d} x {n_t} (d,theta) in all, rows sharded over "
                                                       f"{world} GPU(s) (kernel + partial reduction, max over ranks)",
                        "ms_per_step": fixed_ms, "value": n_d * n_t / (fixed_ms * 1e-3), "unit": UNIT, "scaling": "strong",
                        "strong_efficiency": k_ms / (world * fixed_ms) if world > 1 else 1.0,
                        "strong_efficiency_note": "single-GPU kernel time of the same batch (this run, rank-local) / (N x sharded time)"})
        # ---- C3 with the DEUS-shaped return materialised on the device: g (n_d, n_t, n_g) and -indicator (n_d, n_t) ----
        if world == 1:
            try:
                g_buf = torch.empty((n_d, n_t, lm.ng), dtype=torch.float64, device=dev)
                i_buf = torch.empty((n_d, n_t), dtype=torch.float64, device=dev)
                eng.eval_device(d, th, w, P_out=P_all[:n_d], g_out=g_buf, ind_out=i_buf)
                torch.cuda.synchronize()
                ge = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
                for k in range(args.steps):
                    flush.fill_(float(k))
                    ge[k][0].record(); eng.eval_device(d, th, w, P_out=P_all[:n_d], g_out=g_buf, ind_out=i_buf); ge[k][1].record()
                torch.cuda.synchronize()
                g_ms = sum(x.elapsed_time(y) for x, y in ge) / args.steps
                wbytes = g_buf.numel() * 8 + i_buf.numel() * 8
                configs.append({"name": "C3-g", "workload": "C3 with g and -indicator of every (d,theta) written to device memory "
                                                            "(what the list-of-arrays contract of DEUS needs before any copy)",
                                "ms_per_step": g_ms, "value": n_d * n_t / (g_ms * 1e-3), "unit": UNIT,
                                "bytes_written_per_step": wbytes, "extra_ms_over_P_only": g_ms - k_ms,
                                "write_GBps": wbytes / (g_ms * 1e-3) / 1e9,
                                "note": "the stores run under the FP64 work (a few percent of the HBM write bandwidth): the kernel "
                                        "stays FP64 bound; what the DEUS contract costs is the copy of these bytes to the host"})
                del g_buf, i_buf
            except Exception as e:
                configs.append({"name": "C3-g", "error": repr(e)})
        if world > 1:
            try:
                configs.append(c4_sharded_section(torch, capi, models, dev, world, rank, barrier, max_over_ranks))
            except Exception as e:
                configs.append({"name": "C4-fixed", "error": repr(e)})
        if world == 1:
            configs.append(c4_section(torch, capi, models, dev, peak_tflops))
            try:
                configs.append(c1_section(np, tempfile))
            except Exception as e:                     # the manager path must not take the headline down with it
                configs.append({"name": "C1", "error": repr(e)})
            try:
                configs.append(c2_section(np, tempfile))
            except Exception as e:
                configs.append({"name": "C2", "error": repr(e)})
        # ---- C5: nested sampling, 6 design variables, n_live x n_theta sharded over the ranks ----
        eng6 = capi.Engine(models.kinetics("design6").lower().blob, local_rank)
        sh6 = ShardedEfp(eng6, world, rank)
        _, th5, w5 = make_inputs(10, args.ns_theta)
        sh6.efp_global(np.array([[300.0, 275.0, 2000.0, 1.0, 100.0, 20.0]] * (2 * world)), th5[:512], w5[:512])       # warm-up
        barrier()
        ns_dev = None if args.ns_host else f"cuda:{local_rank}"
        ns_iteration_seconds(sh6, C5_BOX, 2000, th5[:2000], np.full(2000, 1.0 / 2000), 2, ns_dev)       # warm-up: library handles, sort / rand kernels
        barrier()
        ns_s, ns_it, ns_rep, ns_init = ns_iteration_seconds(sh6, C5_BOX, args.ns_live, th5, w5, args.ns_iterations, ns_dev)
        ns_s, ns_init = max_over_ranks(ns_s), max_over_ranks(ns_init)
        ns_entry = {"seconds": ns_s, "iterations_timed": ns_it, "n_live": args.ns_live, "n_theta": args.ns_theta,
                    "proposals_per_iteration": ns_rep, "evals_per_iteration": ns_rep * args.ns_theta,
                    "evals_per_s": ns_rep * args.ns_theta / max(ns_s, 1e-12), "initial_live_set_seconds": ns_init,
                    "sampler": "host (numpy / native passes)" if args.ns_host else "device resident (live set, proposals, P on the GPU)",
                    "what": "config C5 (6 design vars): pyds_b200.ns iteration = ellipsoid proposals + P(feasible) of the "
                            "proposals over all theta sharded over the ranks + live-set update; FIXED total work (strong scaling)"}
        configs.append(dict(ns_entry, name="C5-ns"))

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            from oracle import cbind
            rows = min(args.cpu_rows, n_d)
            v, dt, P_cpu, it_cpu = cpu_baseline(rows, n_t, threads, n_d=n_d, formulation="blocktri")
            v9, dt9, _, it9 = cpu_baseline(max(rows // 5, 1), n_t, threads, n_d=n_d, formulation="dense9")
            agree = bool(np.allclose(P_cpu, P_all[:len(P_cpu)].cpu().numpy(), atol=1e-12))
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "formulation": "blocktri",
                   "flops_per_eval": cbind.flops_per_eval("blocktri", it_cpu),
                   "sample": f"first {rows} of the {n_d} design points x all {n_t} theta ({dt:.1f} s); "
                             f"P(feasible) equal to the GPU result on the sample: {agree}",
                   "dense9": {"value": v9, "flops_per_eval": cbind.flops_per_eval("dense9", it9),
                              "sample": f"first {max(rows // 5, 1)} design points ({dt9:.1f} s), the 9-unknown Newton system of SURVEY A.5"}}
        m = {"n_d": n_d, "n_t": n_t, "world": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
             "kernel_ms": k_ms, "f_alg": f_alg, "newton_iters_per_element": I, "peak_tflops": peak_tflops,
             "kernel_name": kernel_name, "traffic": traffic, "ncu_pipes": ncu_pipes, "ncu_note": ncu_note,
             "d_dim": lm.d_dim, "p_dim": lm.p_dim, "cpu": cpu, "e2e_s": e2e_s, "e2e_matches": same, "configs": configs,
             "ns_iteration": ns_entry, "counters": cnt, "clocks": clocks,
             "formulation": "dense9" if args.dense9 else ("coupled" if args.coupled else "blocktri"),
             "state_blocks": [{"states": s, "unknowns": len(s) * lm.ncp, "linear": bool(l)} for s, l in zip(lm.blocks, lm.block_linear)]}
        emit(build_line(m))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
