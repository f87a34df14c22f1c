import sys, time, tempfile
sys.path.insert(0, ".")
import numpy as np, torch
from examples import twostage_reactor as ex2
from examples.reactor_model import theta_samples
from pyds_b200.run import TwoStageManager
m = TwoStageManager(ex2.build_config(tempfile.mkdtemp()))
p_best, p_samples = theta_samples()
rng = np.random.default_rng(1989)
d = np.stack([rng.uniform(250, 350, 150), rng.uniform(250, 300, 150)], axis=1)
p = np.array([s["c"] for s in p_samples]); w = np.full(50, 0.02)
m.efp(d[:10], p, w)
eng = m.simulator.engine
def T(f, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
z0 = m.simulator.default_controls()[None, :]
print("efp total ms", T(lambda: m.efp(d, p, w)))
print("recourse_solve ms", T(lambda: eng.recourse_solve(d, p, w, z0=z0)))
td, tp, tw, tz = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (d, p, w, z0))
print("recourse_solve_device ms", T(lambda: eng.recourse_solve_device(td, tp, tw, tz)))
print("recourse_solve_device check_every=100 ms", T(lambda: eng.recourse_solve_device(td, tp, tw, tz, check_every=100)))
r = eng.recourse_solve(d, p, w, z0=z0)
print("iters", r["iters"].mean(), r["iters"].max())
print("eval_host g ms", T(lambda: eng.eval_host(d, p, None, z=r["z"], want_g=True, want_P=False)))
print("h2d 4 tensors ms", T(lambda: [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (d, p, w, z0)]))
eng.set_profiling(True); eng.profile(reset=True)
for _ in range(10): eng.recourse_solve_device(td, tp, tw, tz)
pr = eng.profile(); print("per call: evaluator kernel ms", pr["sens_ms"] / 10, "launches", pr["sens_launches"] / 10, "lm ms", pr["lm_ms"] / 10)
eng.set_profiling(False)
print("g total ms", T(lambda: m.g(d, p), 10))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10): m.g(d, p)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
