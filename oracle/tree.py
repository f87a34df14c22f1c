"""Restatement of the reference's scenario-tree bookkeeping (oracle; test infrastructure only).

Pure-Python mirrors of the Pyomo-free logic in the reference, used to pin the ordering and
slicing conventions of the CUDA path:

* ``get_all_idx``            -- pyds/utils.py:57-66  (itertools.product order, row-major in BFs)
* ``scenario_indices_at_stage`` -- pyds/utils.py:68-76
* ``stages_to_update``       -- pyds/simulator.py:64-68 (copy in pyds/test.py:3-7)
* ``load_input_plan``        -- pyds/utils.py:123-134 (which row of the input array lands in which block)
* ``three_stage_rows``       -- pyds/run.py:85-96 (np.tile layout and per-d slicing of the indicator vector)
* ``snap_indicator`` / ``infeasible_indicator`` -- pyds/solver.py:63-80
"""
from __future__ import annotations

from itertools import product

import numpy as np


def get_all_idx(BFs):
    return list(product(*[range(int(i)) for i in BFs]))


def scenario_indices_at_stage(BFs, stage):
    return get_all_idx(BFs[0:stage + 1])


def stages_to_update(idx):
    a = np.array(idx)
    mask = [bool(np.all(a[i + 1:] == 0)) for i in range(len(a))]
    return [tuple(idx[:i + 1]) for i in range(len(a)) if mask[i]]


def load_input_plan(BFs, input_map, input_values):
    """Return {(scenario idx tuple, param name): value} exactly as utils.load_input assigns."""
    plan = {}
    for stage, values in input_values.items():
        if stage not in input_map:
            raise ValueError("Undefined input binding: " + str(stage))
        values = np.asarray(values)
        for param_idx, name in enumerate(input_map[stage]):
            for scen_count, sidx in enumerate(scenario_indices_at_stage(BFs, stage)):
                plan[(sidx, name)] = values[scen_count, param_idx]
    return plan


def three_stage_rows(n_d, n_p):
    """Leaf k (product order over [1, n_d, n_p]) reads row k of np.tile(p, (n_d*n_p, 1)),
    i.e. p[k mod n_p]; design point i owns leaves [n_p*i, n_p*(i+1)) (run.py:85, 93-96)."""
    leaves = get_all_idx([1, n_d, n_p])
    return [(leaf[1], k % n_p) for k, leaf in enumerate(leaves)]


def snap_indicator(y):
    """solver.py:71-80: value == 0 -> 0, anything else -> 1."""
    y = np.asarray(y, dtype=np.float64)
    return np.where(y == 0.0, 0.0, 1.0)


def infeasible_indicator(n):
    """solver.py:63-69: NLP reported infeasible -> every leaf indicator is 1."""
    return np.ones(n)
