"""cProfile of ThreeStageManager.g at 2 000 x 50 (config C2, scaled) -- development tool."""
import cProfile, os, pstats, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from examples import threestage_reactor as ex3
from examples.reactor_model import theta_samples
from pyds_b200.run import ThreeStageManager
m = ThreeStageManager(ex3.build_config(tempfile.mkdtemp()))
_, ps = theta_samples()
p = np.array([s["c"] for s in ps])
rng = np.random.default_rng(1989)
d = np.stack([rng.uniform(250, 350, 2000), rng.uniform(250, 300, 2000)], axis=1)
m.g(d[:8], p); m.g(d, p)
t0 = time.perf_counter(); m.g(d, p); print("g ms", (time.perf_counter() - t0) * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): m.g(d, p)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
