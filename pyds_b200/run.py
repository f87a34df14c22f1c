"""``Manager`` / ``TwoStageManager`` / ``ThreeStageManager``: the drop-in boundary.

``g(d, p)`` has the contract of the reference (``pyds/run.py:47-69, 74-98``): ``d`` is
``(n_d, d_dim)``, ``p`` is ``(n_p, p_dim)`` float64, the result is a ``list`` of ``n_d`` arrays
``(n_p, 1)`` holding ``-indicator`` (feasible <=> value >= 0; note ``-0.0 >= 0``), i.e. what DEUS
consumes through ``constraints_func_ptr`` (reference run_twostage.py:145-165).  Construction takes the
same nested config dict (reference run.py:11-26, run_twostage.py:81-106).

What changed underneath: the reference loops over design points in Python and launches one GAMS/CONOPT
process each; here one call evaluates the whole (d, theta) grid on the GPU.  ``efp(d, p, w)`` is the
fused fast path (P(feasible) per design point without materialising g), the only form that exists at
1e5 x 1e5.  With ``torch.distributed`` initialised, ``efp`` shards the design points over the ranks.
"""
from __future__ import annotations

import os

import numpy as np

from . import utils
from .output_manager import OutputManager
from .simulator import Simulator
from .solver import Solver


class Manager:
    shards_by_design_point = True          # rows can be split over the ranks
    couples_design_points = False          # True: the ranks optimise one control vector together (three-stage)

    def __init__(self, config):
        problem = config["problem"]
        self.stage_rules = problem["stage rules"]
        self.model_transformation = problem["model transformation"]
        self.n_stages = len(self.stage_rules)
        self.input_map = problem["input map"]
        self.output_map = problem["output map"]
        if problem["output folder"] is None:
            self.output_folder = os.getcwd()
        else:
            self.output_folder = problem["output folder"]
        self.output_manager = OutputManager(self.output_folder)
        self.solver = Solver(self, config["solver"])
        self.simulator = Simulator(self, config["simulator"])
        self.model = None
        self.output = {}
        self._current_inputs = None
        self._stages = sorted(self.input_map.keys())

    def g(self, d, p):
        raise NotImplementedError

    def _build_model(self, BFs):
        self.model = self.simulator.model
        object.__setattr__(self.model, "BFs", [int(b) for b in BFs])
        object.__setattr__(self.model, "n_stages", len(BFs))

    def scenario(self, idx):
        return utils.get_scenario(self.model, idx)

    def write_output_to_disk(self):
        self.output_manager.write_data_to_disk()

    def _split_inputs(self, input_values):
        """{stage: array} -> (d (n_d, d_dim), p (n_p, p_dim)); design stages are concatenated."""
        theta_stage = self._stages[-1]
        p = np.asarray(input_values[theta_stage], dtype=np.float64)
        ds = [np.asarray(input_values[s], dtype=np.float64) for s in self._stages[:-1]]
        d = np.concatenate(ds, axis=1) if len(ds) > 1 else ds[0]
        return np.atleast_2d(d), np.atleast_2d(p)

    def _check(self, d, p):
        d = np.asarray(d, dtype=np.float64)
        p = np.asarray(p, dtype=np.float64)
        lm = self.simulator.lowered
        if d.ndim != 2 or d.shape[1] != lm.d_dim:
            raise ValueError(f"d must be (n_d, {lm.d_dim})")
        if p.ndim != 2 or p.shape[1] != lm.p_dim:
            raise ValueError(f"p must be (n_p, {lm.p_dim})")
        return np.ascontiguousarray(d), np.ascontiguousarray(p)

    # -- fused fast path -----------------------------------------------------------------------------------
    def efp(self, d, p, w=None):
        """P(feasible) per design point: sum_j w_j 1[scenario (i,j) feasible] after control optimisation."""
        d, p = self._check(d, p)
        world, rank = 1, 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world, rank = dist.get_world_size(), dist.get_rank()
        except ImportError:
            pass
        from .parallel import ShardedEfp, shard_rows, shard_sizes
        w_arr = np.full(len(p), 1.0 / max(len(p), 1)) if w is None else np.asarray(w, dtype=np.float64)
        if world > 1 and not self.shards_by_design_point:
            # a manager that cannot split its rows: each rank evaluates the whole batch (identical results, no speed-up)
            return self._efp_rows(d, p, w_arr)
        a, b = shard_rows(len(d), world, rank)
        if world > 1 and self.couples_design_points:
            # one control vector couples every design point (three-stage): the ranks optimise it together -- an
            # all-reduce of the {Phi, grad, H} statistics per iteration -- so every rank calls, rows or not
            P = self._efp_rows(d[a:b], p, w_arr, n_d_total=len(d))
        elif b > a:
            P = self._efp_rows(d[a:b], p, w_arr)
        else:
            P = np.empty(0)
        if world == 1:
            return P
        return ShardedEfp(None, world, rank).all_gather_host(P, shard_sizes(len(d), world))

    def _efp_rows(self, d, p, w, **kw):
        """P(feasible) of these rows.  Default: from the indicator matrix; the shipped managers reduce on the device."""
        ind = self._indicators(d, p, w, **kw)
        return ((1.0 - ind) * w[None, :]).sum(axis=1)

    def _indicators(self, d, p, w):
        raise NotImplementedError


class TwoStageManager(Manager):
    """Stage 0 = design variables + the control profile shared across theta; stage 1 = theta."""

    def __init__(self, config):
        super().__init__(config)

    def _indicators(self, d, p, w):
        return self.solver.solve(d, p, w, mode="shared")

    def _efp_rows(self, d, p, w):
        return self.solver.efp(d, p, w, mode="shared")

    def g(self, d, p):
        d, p = self._check(d, p)
        n_p = p.shape[0]
        if self.model is None or self.model.BFs[1] != n_p:
            self._build_model([1, n_p])
        w = np.full(n_p, 1.0 / n_p)
        inputs = utils.load_input(self.model, self.input_map, {self._stages[0]: d[:1], self._stages[1]: p}) if len(d) else {}
        self._current_inputs = (d, p)
        if self.simulator.save_output and len(d):            # warm-start simulation records (reference run.py:61-62)
            self.simulator.simulate_all_scenarios(self.model, {self._stages[0]: d, self._stages[1]: p})
        indicator = self.solver.solve(d, p, w, mode="shared") if len(d) else np.zeros((0, n_p))
        if not self.simulator.save_output and not self.solver.save_output:
            # nothing but the inputs to log: the same records (one per design point, reference run.py:67), written at once;
            # DEUS uses g >= 0 as "feasible" -- the (n_p, 1) matrices are views of one negated array
            self.output_manager.write_input_records(self._stages[0], self._stages[1], d, p, self.solver.output)
            return list((-indicator)[:, :, None])
        g_list, records = [], []
        for i in range(d.shape[0]):
            g_mat = np.empty((n_p, 1))
            g_mat[:, 0] = -indicator[i]            # DEUS uses g >= 0 as "feasible"
            g_list.append(g_mat)
            self.output_manager.clear_buffer()
            self.output_manager.add_input({self._stages[0]: d[i:i + 1], self._stages[1]: p})
            self.output_manager.add_solver_solution(self.solver.output_for(i))
            if self.simulator.save_output:
                self.output_manager.add_simulator_solution(self.simulator.output_for([i]))
            records.append(self.output_manager.data)         # one record per design point (reference run.py:67)
        if records:
            self.output_manager.write_records_to_disk(records)
        return g_list


class ThreeStageManager(Manager):
    """Stage 0 = the control profile shared by every design point and theta; stage 1 = design variables;
    stage 2 = theta (reference pyds/run.py:71-98; see examples/threestage_reactor.py for the repaired
    example script)."""

    couples_design_points = True           # the stage-0 control couples all design points

    def __init__(self, config):
        super().__init__(config)

    def _indicators(self, d, p, w, n_d_total=None):
        return self.solver.solve(d, p, w, mode="global", n_d_total=n_d_total)

    def _efp_rows(self, d, p, w, n_d_total=None):
        return self.solver.efp(d, p, w, mode="global", n_d_total=n_d_total)

    def g(self, d, p):
        d, p = self._check(d, p)
        n_d, n_p = d.shape[0], p.shape[0]
        BFs = [1, n_d, n_p]
        if self.model is None or list(self.model.BFs) != BFs:
            self._build_model(BFs)
        w = np.full(n_p, 1.0 / n_p)
        self._current_inputs = (d, p)
        if self.simulator.save_output and n_d:               # reference run.py:89-90
            self.simulator.simulate_all_scenarios(self.model, {self._stages[0]: d, self._stages[1]: p})
        indicator = self.solver.solve(d, p, w, mode="global") if n_d else np.zeros((0, n_p))
        # DEUS uses g >= 0 as "feasible": one (n_p, 1) matrix per design point (reference run.py:92-96), views of one array
        g_list = list((-indicator.reshape(n_d, n_p))[:, :, None])
        self.output_manager.clear_buffer()
        self.output_manager.add_input({self._stages[0]: d, self._stages[1]: p})
        self.output_manager.add_solver_solution(self.solver.output)
        if self.simulator.save_output and n_d:
            self.output_manager.add_simulator_solution(self.simulator.output)
        self.output_manager.write_data_to_disk()            # one record per call (reference run.py:97)
        return g_list
