"""Pin the oracle on the REAL reference -- for a machine that has what this image lacks:
pyomo, GAMS/CONOPT (or another NLP solver Pyomo can drive) and the reference checkout.

It runs the reference's own ``TwoStageManager.g`` (reference pyds/run.py:47-69) on a fixed design grid with the
shipped theta samples (reference run_twostage.py:110-120) and writes ``tests/golden/reference_twostage_g.json``:
the design points, theta, and per design point the returned ``-indicator`` column plus -- when the solver output is
available -- the controls F_in and the CQAs g1, g2 of every scenario.  ``tests/test_reference_golden.py`` compares
the GPU path with that file when it exists (identical classification except |g| < 1e-8, SURVEY.md / BASELINE.json).

    python tools/dump_reference.py /path/to/reference [--solver gams|ipopt] [--n 5]

Nothing here can run in the build image (no pyomo, no solver, no network); the file it writes is the missing golden
vector set.  Until it exists every report says "parity unpinned" for the optimiser and for Pyomo's transcription.
"""
import argparse
import json
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference", help="checkout of cornelmarck/pyds")
    ap.add_argument("--solver", default="gams")
    ap.add_argument("--n", type=int, default=5, help="design grid is n x n over the box of run_twostage.py:138-141")
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    try:
        import pyomo.environ  # noqa: F401
    except ImportError as e:
        raise SystemExit(f"pyomo is required to run the reference: {e}")
    import run_twostage as ref                    # builds config_form, theta samples and the manager at import
    cfg = ref.config_form
    cfg["solver"]["name"] = args.solver
    cfg["solver"]["tee"] = False
    cfg["solver"]["save output"] = True
    cfg["solver"]["save solution state"] = True
    from pyds.run import TwoStageManager
    m = TwoStageManager(cfg)
    t_f = np.linspace(250.0, 350.0, args.n)
    T = np.linspace(250.0, 300.0, args.n)
    d = np.array([[a, b] for a in t_f for b in T])
    p = np.array([s["c"] for s in ref.p_samples])
    g_list = m.g(d, p)
    out = {"source": "cornelmarck/pyds run_twostage.py, TwoStageManager.g", "solver": args.solver,
           "d": d.tolist(), "theta": p.tolist(), "minus_indicator": [np.asarray(g)[:, 0].tolist() for g in g_list]}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_twostage_g.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, "design points:", len(d), "theta:", len(p))


if __name__ == "__main__":
    main()
