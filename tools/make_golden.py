"""Generate tests/golden/*.json -- the committed fixtures the parity tests are pinned on.

Sources (this is test infrastructure; it imports oracle/):
* theta fixture: the reference's own sampling code, run_twostage.py:110-116, seed 12341234
  (numpy PCG64 stream is version-stable) -- sha256 recorded in SURVEY.md section 4;
* stages_to_update known answer: printed output of reference pyds/test.py;
* g1, g2 at a fixed set of (d, theta, F) points: 50-digit mpmath restatement
  (oracle/kinetics_mp.py), rounded to fp64 -- independent of both the numpy and the C oracle.

Run from the repo root:  python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import kinetics as K, kinetics_mp as MP, radau  # noqa: E402

out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(out_dir, exist_ok=True)

p = K.theta_samples()
theta = {
    "source": "run_twostage.py:110-116, default_rng(12341234), n=50, draw order E1,E2,k1,k2",
    "sha256": hashlib.sha256(p.tobytes()).hexdigest(),
    "row0": p[0].tolist(),
    "col_sums": p.sum(axis=0).tolist(),
    "p_best": list(K.P_BEST),
}
json.dump(theta, open(os.path.join(out_dir, "theta_fixture.json"), "w"), indent=1)

json.dump({"source": "pyds/test.py:3-11 printed output", "input": [1, 2, 0, 1, 0, 0],
           "output": [[1, 2, 0, 1], [1, 2, 0, 1, 0], [1, 2, 0, 1, 0, 0]]},
          open(os.path.join(out_dir, "stages_to_update.json"), "w"), indent=1)

json.dump({"source": "Lagrange-Radau ncp=3, 50-digit mpmath (oracle/radau.py)", "tau": radau.TAU, "adot": radau.ADOT,
           "bw": radau.BW}, open(os.path.join(out_dir, "radau3.json"), "w"), indent=1)

rng = np.random.default_rng(20261019)
cases = []
pts = [(300.0, 275.0), (350.0, 282.5), (250.0, 250.0), (350.0, 300.0), (340.0, 285.0), (275.0, 292.0)]
profiles = [np.zeros(8), np.full(8, 2.5), np.array([2.5, 0, 0, 0, 0, 0, 0, 0.0]), rng.uniform(0, 2.5, 8)]
for (tf, T) in pts:
    for F in profiles:
        for th in (np.array(K.P_BEST), p[0], p[17]):
            g1, g2, xe = MP.g_values_mp(tf, T, *th, F)
            cases.append({"d": [tf, T], "theta": th.tolist(), "F": F.tolist(), "g1": float(g1), "g2": float(g2),
                          "c_end": [float(v) for v in xe]})
# the synthetic families of SURVEY.md 8d (cA0, Fbar, pB, pA as design variables)
for _ in range(12):
    tf, T, cA0, Fb, pB, pA = rng.uniform(250, 350), rng.uniform(250, 300), rng.uniform(1500, 2500), rng.uniform(0, 2.5), \
        rng.uniform(90, 110), rng.uniform(15, 25)
    th = p[rng.integers(0, 50)]
    g1, g2, xe = MP.g_values_mp(tf, T, *th, np.full(8, Fb), cA0=cA0, pB=pB, pA=pA)
    cases.append({"d": [tf, T, cA0, Fb, pB, pA], "theta": th.tolist(), "F": [Fb] * 8, "g1": float(g1), "g2": float(g2),
                  "c_end": [float(v) for v in xe]})
json.dump({"source": "oracle/kinetics_mp.py (mpmath, 50 digits), equations run_twostage.py:49-65, Radau 8x3",
           "cases": cases}, open(os.path.join(out_dir, "kinetics_g_mp.json"), "w"), indent=1)
print(len(cases), "cases written")

# state profiles at the 25 collocation times (pins pydsb_state_solve and the oracles' trajectories)
traj_cases = []
for (tf, T), F, th in (((300.0, 275.0), profiles[3], np.array(K.P_BEST)), ((350.0, 282.5), profiles[1], p[0]),
                       ((275.0, 292.0), profiles[2], p[17]), ((340.0, 285.0), profiles[0], p[5])):
    g1, g2, xe, traj = MP.g_values_mp(tf, T, *th, F, return_traj=True)
    traj_cases.append({"d": [tf, T], "theta": th.tolist(), "F": F.tolist(), "columns": ["cA", "cB", "cC", "u"],
                       "traj": [[float(v) for v in row] for row in traj]})
json.dump({"source": "oracle/kinetics_mp.py (mpmath, 50 digits): states at tau = 0 and the 24 Radau points, 8 x 3",
           "cases": traj_cases}, open(os.path.join(out_dir, "kinetics_traj_mp.json"), "w"), indent=1)
print(len(traj_cases), "trajectory cases written")
