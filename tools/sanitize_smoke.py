"""Small pass over every kernel class with finite-result checks: fused chain (plain, PACK, g / indicator outputs, ragged
tiles), trajectories, the three optimiser group shapes with both evaluators and both loops.  Written for compute-sanitizer,
which is closed on this pool (gpurun_out/r02_memcheck.log); kept as the quickest all-kernels smoke run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_inputs  # noqa: E402
from pyds_b200 import capi, models  # noqa: E402

rng = np.random.default_rng(3)
d4, th, w = make_inputs(7, 300)
eng = capi.Engine(models.kinetics("design4").lower().blob, 0)
g, ind, P = eng.eval_host(d4, th, w, want_g=True, want_ind=True)
g2, ind2, P2 = eng.eval_host(d4, th[:50], None, want_g=True, want_ind=True)          # PACK tiles
assert np.isfinite(g).all() and np.isfinite(P).all() and np.isfinite(P2).all()
x = eng.state_solve(d4[:2], th[:33])
assert np.isfinite(x["x"] if isinstance(x, dict) else x[0]).all()
e2 = capi.Engine(models.kinetics("twostage").lower().blob, 0)
d2 = np.stack([rng.uniform(250, 350, 5), rng.uniform(250, 300, 5)], axis=1)
for env in ({}, {"PYDSB_NO_RCHAIN": "1"}, {"PYDSB_NO_FUSED_LM": "1"}):
    os.environ.update(env)
    a = e2.recourse_solve(d2, th[:50], np.full(50, 0.02))                              # one tile per group (fused loop)
    b = e2.recourse_solve(d2[:2], th[:300], np.full(300, 1 / 300))                     # several tiles per group
    c = e2.recourse_solve(d2[:2], th[:70], per_scenario=True)                          # group = scenario
    gg = e2.recourse_solve(d2[:3], th[:50], np.full(50, 0.02), global_group=True)      # one group for the grid
    for r in (a, b, c, gg):
        assert np.isfinite(r["phi"]).all()
    for k in env:
        del os.environ[k]
print("sanitize_smoke ok")
