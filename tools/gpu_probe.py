"""First-contact GPU probe: parity of the CUDA path against the oracle on small grids, the DFMA
peak, and a first timing at config C3's shape.  Development tool (uses oracle/ as the checker)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from oracle import cbind, kinetics as K, tape_vm  # noqa: E402
from pyds_b200 import capi, models  # noqa: E402

out = {}
print("device:", torch.cuda.get_device_name(0), "cpus:", os.cpu_count())
tf, ms = capi.dfma_peak(0, 10)
print(f"DFMA peak: {tf:.2f} TFLOP/s ({ms:.3f} ms)")
out["dfma_tflops"] = tf

rng = np.random.default_rng(7)
for variant, nd_cols in (("twostage", 2), ("design4", 4), ("design6", 6)):
    lm = models.kinetics(variant).lower()
    eng = capi.Engine(lm.blob, 0)
    print(variant, eng.info)
    n_d, n_t = 37, 301
    cols = [rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d), rng.uniform(1500, 2500, n_d), rng.uniform(0, 2.5, n_d),
            rng.uniform(90, 110, n_d), rng.uniform(15, 25, n_d)]
    d = np.stack(cols[:nd_cols], axis=1)
    p = K.theta_samples(n_t)
    w = rng.uniform(0.5, 1.5, n_t); w /= w.sum()
    z = rng.uniform(0, 2.5, 8) if variant == "twostage" else None
    g, ind, P = eng.eval_host(d, p, w, z=z, want_g=True, want_ind=True, want_P=True)
    go, Po, it = cbind.kinetics_grid(d, p, w, F=z, tol=1e-13)
    e1 = np.abs(g[..., 0] - go[..., 0]).max()
    e2 = np.abs(g[..., 1] - go[..., 1]).max() / 1e5
    feas_o = (go[..., 0] <= 0) & (go[..., 1] <= 0)
    feas_g = (ind == 0)
    print(f"  max|dg1|={e1:.3e} max|dg2|/1e5={e2:.3e} class mismatches={int((feas_o != feas_g).sum())} "
          f"max|dP|={np.abs(P - Po).max():.3e} P.mean={P.mean():.4f}")
    print("  counters", eng.counters(reset=True))
    out[variant] = dict(e1=float(e1), e2=float(e2), mism=int((feas_o != feas_g).sum()), dP=float(np.abs(P - Po).max()))

# timing at C3 shape (device-resident)
lm = models.kinetics("design4").lower()
eng = capi.Engine(lm.blob, 0)
for n_d, n_t in ((1000, 10000), (10000, 10000)):
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d), rng.uniform(1500, 2500, n_d), rng.uniform(0, 2.5, n_d)], 1)
    p = K.theta_samples(n_t)
    dd = torch.from_numpy(d).cuda(); pp = torch.from_numpy(p).cuda()
    P = torch.empty(n_d, dtype=torch.float64, device="cuda")
    eng.eval_device(dd, pp, P_out=P); torch.cuda.synchronize()
    eng.counters(reset=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.eval_device(dd, pp, P_out=P); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c = eng.counters()
    I = c["newton_iters"] / c["scenarios"] / lm.nfe
    fl = lm.flops_per_evaluation(I, newton_flops_per_eval=c['newton_flops'] / c['scenarios'])
    print(f"C3-shape {n_d}x{n_t}: {ms:.2f} ms  {n_d * n_t / ms * 1e3:.3e} evals/s  I={I:.2f}/elem  F_alg={fl:.0f} flop/eval "
          f"-> {n_d * n_t * fl / ms / 1e9:.2f} TFLOP/s ({n_d * n_t * fl / ms / 1e9 / tf * 100:.1f}% of DFMA peak)  counters={c}")
    out[f"c3_{n_d}"] = dict(ms=ms, evals_per_s=n_d * n_t / ms * 1e3, I=I, flops=fl)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
