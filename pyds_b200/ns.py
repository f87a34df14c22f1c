"""A compact nested-sampling design-space driver with DEUS's activity-form interface.

The reference hands ``Manager.g`` to DEUS (``DEUS(the_activity_form).solve()``, reference
``run_twostage.py:212-214``); DEUS is a third-party package that is not installed here, so the
example scripts fall back to this driver when ``import deus`` fails.  It reads the same nested dict
(reference run_twostage.py:122-210): design box, theta samples + weights, target reliability,
``points_schedule`` = (reliability level, n_live, n_replacements) rows, ``inside_fraction`` stop,
``prng_seed``; proposals come from the enlarged bounding ellipsoid of the live points
("suob-ellipsoid").  PARITY with DEUS's own sampler is UNPINNED -- DEUS's algorithmic details are
restated from memory; what is exact is the contract: the constraint function is called as
``g(d, p) -> list of (n_p, n_g)`` and the probability of feasibility is
``sum_j w_j 1[all_k g_jk >= 0]``.  If the function's owner offers ``efp(d, p, w)`` (pyds_b200's
managers do), that fused GPU path is used instead and g is never materialised.

``parameters_best_estimate`` (reference run_twostage.py:145-150) is read and kept as ``p_best`` but NOT used: DEUS
starts with a deterministic phase that scores points by the worst constraint value at the nominal parameters before it
switches to the probabilistic phase; this driver samples the probabilistic phase from the first iteration (the
feasibility probability over the whole theta set is cheap here), so the nominal point never enters.
"""
from __future__ import annotations

import heapq
import math
import os
import time

import numpy as np


def efp_from_g(g_list, w):
    """DEUS's reduction on the g-function output (note: -0.0 >= 0 counts as feasible)."""
    return np.array([float(np.sum(w * np.all(np.asarray(gm) >= 0.0, axis=1))) for gm in g_list])


class DEUS:
    _blas = None                                  # threadpoolctl controller, created on first use

    def __init__(self, form):
        if form.get("activity_type") != "ds":
            raise NotImplementedError("only the design-space ('ds') activity is implemented")
        self.form = form
        prob = form["problem"]
        self.names = [list(v.keys())[0] for v in prob["design_variables"]]
        self.lo = np.array([list(v.values())[0][0] for v in prob["design_variables"]], dtype=np.float64)
        self.hi = np.array([list(v.values())[0][1] for v in prob["design_variables"]], dtype=np.float64)
        self.p = np.array([s["c"] for s in prob["parameters_samples"]], dtype=np.float64)
        self.w = np.array([s["w"] for s in prob["parameters_samples"]], dtype=np.float64)
        self.p_best = np.array([prob["parameters_best_estimate"]], dtype=np.float64)
        self.target = float(prob["target_reliability"])
        sol = form["solver"]
        if sol.get("name") != "ds-ns":
            raise NotImplementedError("only the 'ds-ns' solver is implemented")
        st = sol["settings"]
        self.g_func = st["efp_evaluation"]["constraints_func_ptr"]
        self.schedule = sorted(st["points_schedule"])
        self.inside_fraction = 1.0
        for c in st.get("stop_criteria", []):
            self.inside_fraction = float(c.get("inside_fraction", self.inside_fraction))
        alg = sol["algorithms"]["sampling"]["settings"]
        self.seed = int(alg.get("prng_seed", 0))
        self.f0 = float(alg.get("f0", 0.1))
        self.alpha = float(alg.get("alpha", 0.3))
        self.max_iterations = 100000
        for c in alg.get("stop_criteria", []):
            self.max_iterations = int(c.get("max_iterations", self.max_iterations))
        self.use_fast_path = bool(st["efp_evaluation"].get("use_fused_efp", True))
        # device-resident sampling (settings 'device': a torch device): live set, proposals and probabilities stay on
        # the GPU; needs an owner with ``efp_device(d, p, w) -> tensor`` (solve_device below)
        self.device = alg.get("device", None)
        self.native_update = True                 # live-set update in libpyds_b200 (False: the Python loop)
        self.result = None

    # ----------------------------------------------------------------------------------------------
    def efp(self, d):
        owner = getattr(self.g_func, "__self__", None)
        if self.use_fast_path and owner is not None and hasattr(owner, "efp"):
            return np.asarray(owner.efp(d, self.p, self.w), dtype=np.float64)
        return efp_from_g(self.g_func(d, self.p), self.w)

    def _propose(self, rng, live, n, enlarge, method=None):
        """Uniform samples from the enlarged bounding ellipsoid of ``live``, clipped to the box.  ``method``: None picks
        the rejection scheme with the higher acceptance, "ellipsoid" / "box" force one (tests)."""
        # numpy path only: its products have an inner dimension of a few design variables, on which a multi-threaded
        # BLAS spends tens of milliseconds spinning its pool (measured: 70 ms with 8 OpenBLAS threads, 2 ms with one, for
        # 1e5 x 6).  The controller is created once: building it walks every loaded shared library (25 ms with torch).
        import contextlib
        ctx = contextlib.nullcontext()
        if not self.native_update:
            try:
                if DEUS._blas is None:
                    from threadpoolctl import ThreadpoolController
                    DEUS._blas = ThreadpoolController()
                ctx = DEUS._blas.limit(limits=1, user_api="blas")
            except ImportError:                           # optional dependency: without it the calls are just slower
                pass
        with ctx:
            n_pts, dim = live.shape
            span = self.hi - self.lo
            if self.native_update:
                # mean, covariance factor and largest Mahalanobis radius in two native passes over the live points
                from .capi import ns_ellipsoid
                mu, L, r = ns_ellipsoid(live, self.lo, self.hi)
            else:
                # numpy, dimension-major layout (dim, n): every reduction runs along the long axis and every product has
                # the long axis last, which is what numpy and BLAS are fast at
                uT = (live.T - self.lo[:, None]) / span[:, None]
                mu = uT.mean(axis=1)
                xc = uT - mu[:, None]
                cov = (xc @ xc.T) / (n_pts - 1) + 1e-12 * np.eye(dim) if n_pts > dim else np.eye(dim) * 0.25
                L = np.linalg.cholesky(np.atleast_2d(cov))
                y = np.linalg.inv(L) @ xc
                r = np.sqrt((y * y).sum(axis=0).max())
            r = r * (1.0 + enlarge)                       # the smallest ellipsoid of that shape holding all points, enlarged
            # proposals per batch: cache-sized (one batch of 2e5 x 6 for a wide early live set, of which ~5 % lies in the
            # box, measured 6x slower than 15 batches of 1.3e4: page faults on fresh 10 MB temporaries)
            m = min(int(1.3 * n) + 16, 16384)
            # Uniform samples of (ellipsoid intersected with the unit box), by rejection from whichever is tighter: the
            # ellipsoid (directions from normals; rejects what falls outside the box -- 95 % for a wide early live set in
            # 6 dimensions, whose ellipsoid contains the whole box) or the ellipsoid's bounding box clipped to the unit
            # box (rejects what falls outside the ellipsoid -- 92 % for a concentrated set in 6 dimensions).  The first
            # batch uses the ellipsoid and measures its acceptance; the volumes then tell which one wins.
            ext = r * np.sqrt((L * L).sum(axis=1))
            blo, bhi = np.maximum(mu - ext, 0.0), np.minimum(mu + ext, 1.0)
            vol_box = float(np.prod(np.maximum(bhi - blo, 0.0)))
            vol_ell = math.pi ** (dim / 2.0) / math.gamma(dim / 2.0 + 1.0) * r ** dim * float(np.prod(np.diag(L)))
            Linv = None
            chunks, have, from_box = [], 0, method == "box"
            if from_box:
                Linv = np.linalg.inv(L)
            while have < n:
                if from_box:
                    cand = blo[:, None] + (bhi - blo)[:, None] * rng.uniform(size=(dim, m))
                    y = Linv @ (cand - mu[:, None])
                    ok = (y * y).sum(axis=0) <= r * r
                else:
                    z = rng.normal(size=(dim, m))
                    z *= rng.uniform(size=m) ** (1.0 / dim) * r / np.sqrt((z * z).sum(axis=0))
                    cand = mu[:, None] + L @ z
                    ok = (cand.min(axis=0) >= 0.0) & (cand.max(axis=0) <= 1.0)
                chunks.append(cand[:, ok])                # arrays, not rows: 1e4 proposals per iteration at config C5
                have += chunks[-1].shape[1]
                if method is None and len(chunks) == 1 and have < n and vol_box > 0.0:
                    a_ell = max(have / m, 1.0 / m)        # share of the ellipsoid inside the box
                    if a_ell * vol_ell / vol_box > a_ell:  # = share of the clipped bounding box inside the ellipsoid
                        from_box, Linv = True, np.linalg.inv(L)
            return self.lo + np.concatenate(chunks, axis=1)[:, :n].T * span

    def _replace_worst_py(self, live, live_p, cand, cand_p, n_live, heap=None):
        """The update rule in plain Python (growth phase; also the cross-check of the native pass in the tests)."""
        if heap is None:
            heap = [(float(v), k) for k, v in enumerate(live_p)]
            heapq.heapify(heap)
        grown = []
        order = range(len(cand))
        if len(live_p) >= n_live:
            # the worst live value only rises during the pass: a proposal below the worst at the start of the pass
            # can never be accepted -- drop those with one vectorised comparison, keep the order of the rest
            order = np.flatnonzero(cand_p >= heap[0][0])
        for ci in order:
            c, cp = cand[ci], float(cand_p[ci])
            if len(live_p) + len(grown) < n_live:
                if cp >= heap[0][0]:
                    grown.append((c, cp))
                    heapq.heappush(heap, (cp, len(live_p) + len(grown) - 1))
                continue
            pk, k = heap[0]
            if cp > pk or (cp == pk and cp >= self.target):
                heapq.heapreplace(heap, (cp, k))
                if k < len(live_p):
                    live[k], live_p[k] = c, cp
                else:
                    grown[k - len(live_p)] = (c, cp)
        if grown:
            live = np.vstack([live] + [g[0][None, :] for g in grown])
            live_p = np.append(live_p, [g[1] for g in grown])
        return live, live_p, heap

    def solve(self):
        if self.device is not None:
            return self.solve_device()
        rng = np.random.default_rng(self.seed)
        dim = len(self.lo)
        level, n_live, n_rep = self.schedule[0]
        live = self.lo + rng.uniform(size=(n_live, dim)) * (self.hi - self.lo)
        t_start = time.time()
        live_p = self.efp(live)
        t_initial = time.time() - t_start
        n_evals = len(live)
        history, iter_times = [(live.copy(), live_p.copy())], []
        heap = None
        it = 0
        while it < self.max_iterations:
            inside = float(np.mean(live_p >= self.target))
            if inside >= self.inside_fraction:
                break
            t0 = time.time()
            worst = live_p.min()
            for lv, nl, nr in self.schedule:                 # the last schedule row whose level has been reached
                if worst >= lv:
                    level, n_live, n_rep = lv, nl, nr
            enlarge = self.f0 * max(1.0 - inside, 0.05) ** self.alpha
            cand = self._propose(rng, live, n_rep, enlarge)
            cand_p = self.efp(cand)
            n_evals += len(cand)
            history.append((cand.copy(), cand_p.copy()))
            # replace-the-worst, candidate by candidate.  Full live set: one native pass (pydsb_ns_replace_worst: a heap of
            # (P, slot) reproduces np.argmin's choice -- lowest slot among equal minima -- in O(log n_live) per candidate;
            # 1e4 proposals against 1e5 live points, config C5, cost 0.5 s per iteration with an argmin scan and 60 ms
            # in a Python loop over heapq).  While the set still grows towards the schedule's n_live: the same rule in
            # Python.
            if len(live_p) >= n_live and self.native_update:
                from .capi import ns_replace_worst
                live_p = np.ascontiguousarray(live_p, dtype=np.float64)
                slots = ns_replace_worst(live_p, cand_p, self.target)
                taken = slots >= 0
                live[slots[taken]] = cand[taken]          # repeated slots: the last candidate written stays
                heap = None
            else:
                live, live_p, heap = self._replace_worst_py(live, live_p, cand, cand_p, n_live, heap)
            iter_times.append(time.time() - t0)
            it += 1
        self.result = {
            "live_points": live, "live_efp": live_p, "iterations": it, "evaluations": n_evals,
            "inside_fraction": float(np.mean(live_p >= self.target)), "seconds": time.time() - t_start,
            "seconds_per_iteration": float(np.mean(iter_times)) if iter_times else 0.0,
            "iteration_seconds": list(iter_times),
            "initial_live_set_seconds": t_initial,
            "samples": np.vstack([h[0] for h in history]), "samples_efp": np.concatenate([h[1] for h in history]),
        }
        return self.result

    # ---- device-resident variant ---------------------------------------------------------------------
    def solve_device(self):
        """The same loop with the live set, the proposals and their probabilities resident on ``self.device`` (torch
        tensors): per iteration nothing crosses the host boundary but a handful of scalars.

        * bounding ellipsoid: mean / covariance / largest Mahalanobis radius by reductions on the device (the 6 x 6
          Cholesky factor goes through the host: 36 doubles);
        * proposals: rejection sampling from the tighter of (ellipsoid, its bounding box clipped to the design box) with
          the device's counter-based generator -- every rank seeds it alike and therefore draws the SAME global proposal
          set without any exchange, then evaluates only its own row block (``owner.efp_device``);
        * live-set update: the sequential replace-the-worst pass leaves the ``n_live`` best of (live, accepted
          candidates); here that set is selected in one stable descending sort of the concatenated probabilities
          (ties: live points before candidates, candidates in proposal order).  Same set as the sequential rule whenever
          no candidate ties with the final worst live value below the target; DEUS's own tie handling is unpinned.

        Parity with DEUS is unpinned like ``solve``; the host variant stays the default."""
        import torch
        dev = torch.device(self.device)
        owner = getattr(self.g_func, "__self__", None)
        if owner is None or not hasattr(owner, "efp_device"):
            raise ValueError("device-resident sampling needs a constraint-function owner with efp_device(d, p, w)")
        f64 = dict(dtype=torch.float64, device=dev)
        gen = torch.Generator(device=dev)
        gen.manual_seed(self.seed)
        lo, hi = torch.tensor(self.lo, **f64), torch.tensor(self.hi, **f64)
        span = hi - lo
        p_t, w_t = torch.tensor(self.p, **f64), torch.tensor(self.w, **f64)
        dim = len(self.lo)
        level, n_live, n_rep = self.schedule[0]
        sync = (lambda: torch.cuda.synchronize(dev)) if dev.type == "cuda" else (lambda: None)
        live = lo + torch.rand((n_live, dim), generator=gen, **f64) * span
        sync(); t_start = time.time()
        live_p = owner.efp_device(live, p_t, w_t)
        sync(); t_initial = time.time() - t_start
        n_evals = n_live
        iter_times, eval_times, it = [], [], 0
        vol_ball = math.pi ** (dim / 2.0) / math.gamma(dim / 2.0 + 1.0)
        while it < self.max_iterations:
            # (scalars the host needs leave the device in as few transfers as possible: each one is a synchronisation)
            inside, worst = torch.stack([(live_p >= self.target).double().mean(), live_p.min()]).tolist()
            if inside >= self.inside_fraction:
                break
            sync(); t0 = time.time()
            trace = [] if os.environ.get("PYDSB_NS_TRACE") else None
            def mark(name):
                if trace is not None:
                    sync(); trace.append((name, time.time()))
            for lv, nl, nr in self.schedule:
                if worst >= lv:
                    level, n_live, n_rep = lv, nl, nr
            enlarge = self.f0 * max(1.0 - inside, 0.05) ** self.alpha
            # ---- bounding ellipsoid in box-scaled coordinates ----
            u = (live - lo) / span
            mu = u.mean(dim=0)
            xc = u - mu
            n_pts = u.shape[0]
            # the dim x dim factor goes through the host (a few dozen doubles, with the mean in the same transfer): no
            # dense-solver library on the device path
            mc = torch.cat([mu, ((xc.T @ xc) / max(n_pts - 1, 1)).reshape(-1)]).cpu().numpy()
            mu_h = mc[:dim]
            cov_h = mc[dim:].reshape(dim, dim) + 1e-12 * np.eye(dim) if n_pts > dim else 0.25 * np.eye(dim)
            L_h = np.linalg.cholesky(cov_h)
            L, Linv = torch.tensor(L_h, **f64), torch.tensor(np.linalg.inv(L_h), **f64)
            y = Linv @ xc.T
            r = float((y * y).sum(dim=0).max().sqrt()) * (1.0 + enlarge)
            mark("ellipsoid")
            # bounding box of the ellipsoid and the two volumes: on the host, from the factor that is there already
            ext_h = r * np.sqrt((L_h * L_h).sum(axis=1))
            blo_h, bhi_h = np.maximum(mu_h - ext_h, 0.0), np.minimum(mu_h + ext_h, 1.0)
            blo, bhi = torch.tensor(blo_h, **f64), torch.tensor(bhi_h, **f64)
            vol_box = float(np.prod(np.maximum(bhi_h - blo_h, 0.0)))
            vol_ell = vol_ball * r ** dim * float(np.prod(np.diagonal(L_h)))
            # ---- proposals ----
            m = max(int(4 * n_rep) + 64, 4096)
            from_box, chunks, have, first = False, [], 0, True
            while have < n_rep:
                if from_box:
                    cand = blo + (bhi - blo) * torch.rand((m, dim), generator=gen, **f64)
                    yy = Linv @ (cand - mu).T
                    ok = (yy * yy).sum(dim=0) <= r * r
                else:
                    z = torch.randn((m, dim), generator=gen, **f64)
                    rad = torch.rand(m, generator=gen, **f64) ** (1.0 / dim) * r
                    z = z * (rad / z.norm(dim=1))[:, None]
                    cand = mu + z @ L.T
                    ok = (cand.min(dim=1).values >= 0.0) & (cand.max(dim=1).values <= 1.0)
                sel = cand[ok]
                chunks.append(sel)
                have += sel.shape[0]
                if first and have < n_rep and vol_box > 0.0:
                    a_ell = max(sel.shape[0] / m, 1.0 / m)
                    from_box = vol_ell > vol_box              # same test as the host sampler: the tighter region wins
                    m = min(max(int(1.5 * (n_rep - have) / max(a_ell if not from_box else a_ell * vol_ell / vol_box, 1e-3)), 4096), 1 << 20)
                first = False
            cand = lo + torch.cat(chunks)[:n_rep] * span
            mark("proposals")
            sync(); t_e = time.time()
            cand_p = owner.efp_device(cand, p_t, w_t)
            sync(); eval_times.append(time.time() - t_e)
            n_evals += n_rep
            # ---- live-set update: the n_live best of (live, candidates that beat the worst) ----
            allp = torch.cat([live_p, cand_p])
            allx = torch.cat([live, cand])
            keep = torch.sort(allp, descending=True, stable=True).indices[:n_live]
            keep = torch.sort(keep).values                    # keep the surviving live points in their slots' order
            live, live_p = allx[keep], allp[keep]
            mark("update")
            if trace:
                print("ns trace:", [(n, round(1e3 * (t - (trace[i - 1][1] if i else t0)), 2)) for i, (n, t) in enumerate(trace)], flush=True)
            sync(); iter_times.append(time.time() - t0)
            it += 1
        self.result = {
            "live_points": live.cpu().numpy(), "live_efp": live_p.cpu().numpy(), "iterations": it, "evaluations": n_evals,
            "inside_fraction": float((live_p >= self.target).double().mean()), "seconds": time.time() - t_start,
            "seconds_per_iteration": float(np.mean(iter_times)) if iter_times else 0.0,
            "iteration_seconds": list(iter_times),
            # of which the evaluation of the proposals (kernel + gather); the rest is the sampler itself
            "seconds_per_iteration_evaluation": float(np.mean(eval_times)) if eval_times else 0.0,
            "initial_live_set_seconds": t_initial,
        }
        return self.result
