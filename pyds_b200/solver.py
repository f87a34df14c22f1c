"""``Solver``: the control (recourse) optimisation + indicator logic, batched on the GPU.

Same constructor and entry points as the reference's ``pyds/solver.py`` (``Solver(parent, config)``,
``solve()``, ``_get_indicator_var_values()``, the ``no_infeasible`` counter and the
``warn infeasible`` message), with the GAMS/CONOPT process launch per design point
(reference pyds/solver.py:50-55) replaced by ``pydsb_recourse_solve`` over ALL design points at
once, followed by one pass of the feasibility kernel at the optimised controls.

Classification rule (reference pyds/solver.py:71-80): the relaxed indicator is snapped ``value == 0
-> 0 else 1``.  CONOPT returns an indicator sitting on its bound as exactly 0 while satisfying the
big-M rows to its feasibility tolerance; the equivalent here is ``max_k g_k/M_k <= feasibility
tolerance`` at the optimised controls (default 1e-7 in big-M-scaled units; 0 when the controls are
not optimised).  A scenario whose state solve fails is reported infeasible (indicator 1), as the
reference does for an infeasible NLP (pyds/solver.py:39-44, 63-69).  That tolerance is a guess at the solver's
behaviour (parity with CONOPT is unpinned), so every call reports how much rests on it: ``result['inside_band']`` counts
the scenarios with ``0 < max_k g_k/M_k <= tolerance`` (``efp``: ``result['band_weight']``, the theta weight per design
point that the tolerance turns feasible).

``io options`` beyond the reference's:
  'optimise controls' (True), 'feasibility tolerance' (1e-7), the optimiser's 'mu0' 'mu_min' 'lam0' 'step_tol' 'max_iter',
  'warm start carry-over' (False): the reference's sequencing -- Pyomo keeps the variable values between solves
      (``warmstart=True``, pyds/solver.py:17) and the warm-start simulator rewrites the states only (pyds/simulator.py:70-94),
      so the control profile of design point i starts from the optimum of design point i-1 (pyds/run.py:55-67), across calls
      too.  Serial by nature: one optimiser run per design point; a mode, not the default.
  'check state bounds' (None = whenever 'save output' computes the profiles anyway; True: always, one extra state solve
      per call): count the scenarios whose solved profiles leave the variable bounds of the user's model (``c`` in
      [0, 1e8], reference run_twostage.py:44), which the state solve does not enforce; ``result['bound_violations']``
      (None when not checked).
"""
from __future__ import annotations

import numpy as np


class _LazyResult(dict):
    """``Solver.result`` with entries that are computed when they are first read."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        self.lazy = {}

    def __missing__(self, key):
        if key in self.lazy:
            self[key] = self.lazy.pop(key)()
            return self[key]
        raise KeyError(key)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self.lazy


class Solver:
    def __init__(self, parent, config):
        self.parent = parent
        self.tee = config["tee"]
        config.setdefault("use relaxation", True)
        self.use_relaxation = config["use relaxation"]
        self.io_options = config["io options"]
        self.io_options.setdefault("warmstart", True)
        self.solve_trajectories = config["solve trajectories"]
        self.warn_infeasible = config["warn infeasible"]
        self.save_output = config["save output"]
        self.save_solution_states = config["save solution state"]
        self.no_infeasible = 0
        self.name = config["name"]                       # 'gams' in the reference scripts: informational
        opt = self.io_options
        self.optimise_controls = bool(opt.get("optimise controls", True))
        self.feas_tol = float(opt.get("feasibility tolerance", 1e-7))
        self.lm_options = {k: opt[k] for k in ("mu0", "mu_min", "lam0", "step_tol", "max_iter", "check_every") if k in opt}
        self.carry_over = bool(opt.get("warm start carry-over", False))
        self.check_bounds = opt.get("check state bounds", None)
        self._z_carry = None                             # the last optimised control profile (carry-over mode)
        self.result = None
        self.output = {"relaxation": None, "trajectories": None}
        self._indicator = None
        self.controls = None

    # ------------------------------------------------------------------------------------------------
    def solve(self, d=None, p=None, w=None, mode="shared", n_d_total=None, group=None):
        """Optimise the controls for the current batch and classify every scenario.

        ``mode``: ``"shared"`` one control vector per design point (two-stage, reference run.py:55-67),
        ``"global"`` one for the whole grid (three-stage, reference run.py:76-91), ``"per_scenario"``.
        ``n_d_total`` (global mode under torch.distributed): ``d`` holds only this rank's rows of a batch of
        ``n_d_total`` design points; the ranks of ``group`` optimise the one control vector together (an all-reduce of
        the optimiser statistics per iteration) and every rank must make this call, with zero rows if it has none.
        Returns the indicator matrix (n_d, n_p) in {0, 1}."""
        sim = self.parent.simulator
        eng = sim.engine
        if d is None:
            d, p = self.parent._current_inputs
        d = np.asarray(d, dtype=np.float64)
        d = d.reshape(-1, sim.lowered.d_dim) if d.size == 0 else np.atleast_2d(d)
        p = np.atleast_2d(np.asarray(p, dtype=np.float64))
        if n_d_total is not None and mode != "global":
            raise ValueError("only the global-control mode couples the ranks")
        self.output = {"relaxation": None, "trajectories": None}
        n_free = int((~sim.lowered.z_fixed).sum()) if sim.lowered.nz else 0
        z, z_mode, tol = None, None, 0.0
        self.result = {"status": None, "iters": None, "phi": None}
        if self.use_relaxation and self.optimise_controls and n_free > 0:
            z0 = sim.default_controls()[None, :]
            if n_d_total is not None:
                res = self._solve_global_sharded(eng, d, int(n_d_total), p, w, z0, group)
            elif self.carry_over and mode == "shared" and len(d):
                res = self._solve_carry_over(eng, d, p, w, z0)
            else:
                res = eng.recourse_solve(d, p, w, z0=z0, per_scenario=(mode == "per_scenario"),
                                         global_group=(mode == "global"), **self.lm_options)
            z = res["z"]
            self.result = {"status": res["status"], "iters": res["iters"], "phi": res["phi"]}
            tol = self.feas_tol
        elif n_free > 0:
            z = sim.default_controls()
        self.controls = z
        if d.shape[0] == 0:                              # a rank without rows (sharded global mode)
            self._indicator = np.zeros((0, p.shape[0]))
            self._g = np.zeros((0, p.shape[0], sim.lowered.ng))
            return self._indicator
        g, ind, _ = eng.eval_host(d, p, None, z=z, want_g=True, want_ind=False, want_P=False)
        c = g / sim.lowered.big_m[None, None, :]
        # (reductions over the short constraint axis, one column at a time: numpy's axis=-1 reduce over 2 entries is
        # ten times slower per scenario than the column-wise form)
        with np.errstate(invalid="ignore"):
            cmax, finite = c[..., 0], np.isfinite(c[..., 0])
            for k in range(1, c.shape[-1]):
                cmax = np.maximum(cmax, c[..., k])         # NaN propagates, as in ndarray.max
                finite = finite & np.isfinite(c[..., k])
            y = np.maximum(0.0, cmax)
            indicator = np.where(y <= tol, 0.0, 1.0)
        failed = ~finite
        indicator[failed] = 1.0
        with np.errstate(invalid="ignore"):
            self.result["inside_band"] = int(((y > 0.0) & (y <= tol) & ~failed).sum())
        self.result["band"] = tol
        if failed.any():
            self.no_infeasible += int(failed.sum())
            if self.warn_infeasible:
                print("Warning: Infeasible relaxation. Total number of infeasible solves: {}".format(self.no_infeasible))
        self._indicator = indicator
        self._g = g
        self._traj = None
        check = self.save_output if self.check_bounds is None else bool(self.check_bounds)
        traj_failed = None
        if self.save_output or check:
            # the solved state profiles the reference reads back with utils.parse_value (pyds/solver.py:100-117)
            self._traj, traj_failed = eng.state_solve(d, p, z=z)
            self._time = eng.time_grid()
        self.result["bound_violations"] = self._count_bound_violations(self._traj, traj_failed) if check else None
        if self.save_output:
            self.output["relaxation"] = self._collect_output()
            if self.solve_trajectories:
                # reference pyds/solver.py:46-48,57-61: a second NLP solve with the snapped indicators fixed -- its
                # objective is then constant, so the controls found by the relaxation remain a solution: same profiles
                self.output["trajectories"] = self.output["relaxation"]
        return indicator

    def efp(self, d, p, w, mode="shared", n_d_total=None, group=None):
        """P(feasible) per design point with everything on the device: control optimisation, one feasibility pass at
        the optimised controls with the same classification rule as ``solve`` (``g_k / M_k <= feasibility tolerance``),
        weighted reduction; only the n_d probabilities come back.  ``solve`` + a host-side sum gives the same numbers
        (tested) but copies g for every scenario to the host -- 1.6 GB at 1e4 x 1e4."""
        import torch
        from . import capi
        sim = self.parent.simulator
        eng = sim.engine
        dev = torch.device("cuda", eng.device)
        d = np.asarray(d, dtype=np.float64)
        d = d.reshape(-1, sim.lowered.d_dim) if d.size == 0 else np.atleast_2d(d)
        p = np.atleast_2d(np.asarray(p, dtype=np.float64))
        n_d, n_p = d.shape[0], p.shape[0]
        if n_d_total is not None and mode != "global":
            raise ValueError("only the global-control mode couples the ranks")
        up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        td, tp, tw = up(d), up(p), up(w)
        n_free = int((~sim.lowered.z_fixed).sum()) if sim.lowered.nz else 0
        tz, z_mode, tol = None, capi.Z_MODEL, 0.0
        self.result = {"status": None, "iters": None, "phi": None}
        if self.use_relaxation and self.optimise_controls and n_free > 0:
            tz0 = up(sim.default_controls()[None, :])
            if n_d_total is not None:
                opts = {k: v for k, v in self.lm_options.items() if k != "check_every"}
                r = eng.recourse_solve_global_dist(td, int(n_d_total), tp, tw, tz0, group=group, **opts)
            elif n_d == 0:
                r = None
            else:
                r = eng.recourse_solve_device(td, tp, tw, tz0, per_scenario=(mode == "per_scenario"),
                                              global_group=(mode == "global"), **self.lm_options)
            if r is not None:
                tz = r["z"].contiguous()
                z_mode = {"shared": capi.Z_PER_DESIGN, "global": capi.Z_SHARED, "per_scenario": capi.Z_PER_SCENARIO}[mode]
                self.result = {k: r[k].cpu().numpy() for k in ("status", "iters", "phi")}
                self.controls = tz.cpu().numpy().reshape((-1,) if mode == "global" else (n_d, n_p, -1) if mode == "per_scenario" else (n_d, -1))
                tol = self.feas_tol
        elif n_free > 0:
            tz = up(sim.default_controls())
            z_mode = capi.Z_SHARED
            self.controls = sim.default_controls()
        if n_d == 0:
            return np.empty(0)
        P = torch.empty(n_d, dtype=torch.float64, device=dev)
        eng.counters(reset=True)
        eng.eval_device(td, tp, tw, z=tz, z_mode=z_mode, P_out=P, feas_tol=tol)
        out = P.cpu().numpy()
        failed = eng.counters()["failed"]                # one synchronising read (the call above is the only work since the reset)
        if tol > 0.0:
            # how much of the answer rests on the tolerance: the same pass with the exact rule g <= 0.  It costs a second
            # evaluation of the grid, so it runs when the figure is first read (the inputs stay on the device until the
            # next call replaces ``result``)
            def band_weight():
                P0 = torch.empty(n_d, dtype=torch.float64, device=dev)
                eng.eval_device(td, tp, tw, z=tz, z_mode=z_mode, P_out=P0, feas_tol=0.0)
                return (P - P0).cpu().numpy()
            self.result = _LazyResult(self.result)
            self.result.lazy["band_weight"] = band_weight
            self.result["band"] = tol
        if failed:
            self.no_infeasible += int(failed)
            if self.warn_infeasible:
                print("Warning: Infeasible relaxation. Total number of infeasible solves: {}".format(self.no_infeasible))
        return out

    def _solve_carry_over(self, eng, d, p, w, z0):
        """The reference's sequencing (pyds/run.py:55-67 with ``warmstart=True``): design point i starts its control
        optimisation from the optimum of design point i-1; the first point of the first call from the warm-start profile."""
        z_prev = self._z_carry if self._z_carry is not None else np.asarray(z0, dtype=np.float64).reshape(-1)
        out = {"z": [], "phi": [], "status": [], "iters": []}
        for i in range(d.shape[0]):
            r = eng.recourse_solve(d[i:i + 1], p, w, z0=z_prev[None, :], **self.lm_options)
            z_prev = r["z"][0].copy()
            for k in out:
                out[k].append(r[k][0])
        self._z_carry = z_prev
        return {k: np.asarray(v) for k, v in out.items()}

    def _count_bound_violations(self, x, failed):
        """Scenarios whose solved state profiles ``x`` (pydsb_state_solve) leave the bounds the user's model declares (not
        enforced by the state solve; the reference hands them to the NLP solver)."""
        lm = self.parent.simulator.lowered
        lo, hi = lm.state_lo, lm.state_hi
        if lo is None or not (np.isfinite(lo).any() or np.isfinite(hi).any()):
            return 0
        with np.errstate(invalid="ignore"):
            bad = ((x < lo[None, None, None, :] - 1e-9 * np.maximum(1.0, np.abs(lo))) |
                   (x > hi[None, None, None, :] + 1e-9 * np.maximum(1.0, np.abs(hi)))).any(axis=(2, 3)) & ~failed
        n = int(bad.sum())
        if n and self.warn_infeasible:
            print(f"Warning: {n} scenario(s) leave the variable bounds of the model (not enforced by the state solve)")
        return n

    def _solve_global_sharded(self, eng, d_local, n_d_total, p, w, z0, group):
        import torch
        dev = torch.device("cuda", eng.device)
        up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        opts = {k: v for k, v in self.lm_options.items() if k != "check_every"}
        r = eng.recourse_solve_global_dist(up(d_local), n_d_total, up(p), up(w), up(z0), group=group, **opts)
        return dict(z=r["z"].cpu().numpy().reshape(-1), phi=r["phi"].cpu().numpy(), status=r["status"].cpu().numpy(),
                    iters=r["iters"].cpu().numpy())

    def _get_indicator_var_values(self):
        """Indicators of the final-stage scenarios in leaf (product) order (reference solver.py:91-98)."""
        return self._indicator.reshape(-1).copy()

    def _scenario_values(self, i, j):
        """``{name: {index + (t,): value}}`` of scenario (i, j) for the last stage's names of the output map -- what
        ``utils.parse_value(scen, name)`` returns for an indexed Var (reference pyds/utils.py:149-155).  Time keys
        are the discretised ContinuousSet points, rounded to 6 decimals as Pyomo's collocation does [from memory]."""
        sim = self.parent.simulator
        lm = sim.lowered
        names = lm.dyn_states + lm.quad_states
        keys = getattr(sim.dae_model, "state_keys", {})
        out_map = self.parent.output_map
        wanted = out_map[max(out_map)] if out_map else []
        v = {}
        for name in wanted:
            vals = None
            for col, sname in enumerate(names):
                var, key, (t0, t1) = keys.get(sname, (sname, (), (0.0, 1.0)))
                if var != name:
                    continue
                vals = {} if vals is None else vals
                for k, tau in enumerate(self._time):
                    t = round(t0 + (t1 - t0) * float(tau), 6)
                    vals[tuple(key) + (t,) if key else t] = float(self._traj[i, j, k, col])
            v[name] = vals                                   # None: no such component (parse_value's answer too)
        return v

    def _record(self, rows):
        """One output container over the design points ``rows`` (leaf order: design point major, theta minor)."""
        rows = list(rows)
        n_p = self._g.shape[1]
        phi = self.result["phi"]
        data = []
        for i in rows:
            for j in range(n_p):
                v = self._scenario_values(i, j) if self._traj is not None else {}
                v["g"] = self._g[i, j].copy()
                data.append(v)
        z = None if self.controls is None else np.asarray(self.controls)
        if z is not None and z.ndim > 1:
            z = z[list(rows)] if len(rows) > 1 else z[list(rows)[0]]
        if phi is None:
            objective = None
        else:
            objective = np.asarray(phi).reshape(-1)[list(rows)] if np.size(phi) == self._g.shape[0] else np.asarray(phi).copy()
        return {"data": data, "objective": objective, "controls": None if z is None else z.copy(), "user time": None,
                "infeasible": bool((self._indicator[list(rows)] == 1).all())}

    def _collect_output(self):
        return self._record(range(self._g.shape[0]))

    def output_for(self, i):
        """The solver output of design point ``i`` alone (the two-stage manager writes one record per design point,
        reference pyds/run.py:55-67)."""
        if not self.save_output or self.output["relaxation"] is None:
            return self.output
        rec = self._record([i])
        return {"relaxation": rec, "trajectories": rec if self.solve_trajectories else None}
