"""cProfile of TwoStageManager.efp / g at 150 x 50 (config C1) -- development tool."""
import cProfile, os, pstats, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from examples import twostage_reactor as ex2
from examples.reactor_model import theta_samples
from pyds_b200.run import TwoStageManager
m = TwoStageManager(ex2.build_config(tempfile.mkdtemp()))
_, ps = theta_samples()
p = np.array([s["c"] for s in ps]); w = np.full(50, 0.02)
rng = np.random.default_rng(1989)
d = np.stack([rng.uniform(250, 350, 150), rng.uniform(250, 300, 150)], axis=1)
for f, name in ((lambda: m.efp(d, p, w), "efp"), (lambda: m.g(d, p), "g")):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(20): f()
    print(name, "ms", (time.perf_counter() - t0) / 20 * 1e3)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20): f()
    pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(14)
