"""Scenario-tree helpers and big-M bookkeeping -- same names and argument meaning as the reference's
``pyds/utils.py``, re-thought for a scenario grid that lives on the GPU.

The reference materialises the extensive form as a recursive tree of Pyomo ``Block`` objects
(``create_EF``, reference pyds/utils.py:92-112) and walks it for every parameter write
(``load_input``, :123-134).  Here the tree is implicit: leaf ``(0, i, j)`` of branching factors
``[1, n_d, n_p]`` is thread ``j`` of tile row ``i`` in the kernel grid, in the same
``itertools.product`` order (reference :57-66), so these helpers only validate and reorder arrays.
"""
from __future__ import annotations

from itertools import product
from math import prod

import numpy as np

from . import environ as pyo
from .tape import Relational


def add_BigMConstraint(model, name, bigM_constant, *args, **kwargs):
    """Big-M form of an (in)equality: ``body <= ub + M*y`` / ``body >= lb - M*y`` with ONE relaxed
    indicator ``y`` per scenario block shared by all its big-M rows (reference pyds/utils.py:7-44).
    Recorded symbolically; the kernel evaluates ``y = max(0, max_k g_k/M_k)`` directly."""
    constraint = pyo.Constraint(*args, **kwargs)
    constraint.construct()
    if not hasattr(model, "_indicator_var"):
        model._indicator_var = pyo.Var(initialize=0, within=pyo.Binary, bounds=(0, 1))
    M = float(bigM_constant)
    lower, upper, body = constraint.lower, constraint.upper, constraint.body
    g = model._graph
    rows = []
    if upper is not None:
        rows.append(g.sub(g.wrap(body), g.wrap(upper)))          # body - ub <= M*y
    if lower is not None:
        rows.append(g.sub(g.wrap(lower), g.wrap(body)))          # lb - body <= M*y
    if not rows:
        return
    for k, r in enumerate(rows):
        model._bigm.append((name if len(rows) == 1 else f"{name}[{k + 1}]", r, M))


def _create_objective(m):
    """max (1/N) sum_s (1 - y_s)  (reference pyds/utils.py:46-51); kept as a marker -- the reduced form
    min sum_s w_s max(0, max_k g_ks/M_k) is what pyds_b200.solver minimises."""
    m.obj = pyo.Objective(sense=pyo.maximize)


def get_all_idx(BFs):
    """All scenario index tuples of a tree with branching factors ``BFs``, in product (row-major) order."""
    return product(*[range(int(i)) for i in BFs])


def scenarios_at_stage(m, stage):
    """Index tuples of the scenarios at ``stage`` (the reference returns the Pyomo blocks themselves)."""
    return list(get_all_idx(m.BFs[0:stage + 1]))


def get_scenario(m, idx):
    return tuple(idx)


def get_final_scenarios(m):
    return scenarios_at_stage(m, m.n_stages - 1)


def create_flattened_model(stage_rules):
    """One model on which every stage rule has written its components (reference pyds/utils.py:114-121)."""
    m = pyo.ConcreteModel()
    object.__setattr__(m, "BFs", [1] * len(stage_rules))
    object.__setattr__(m, "n_stages", len(stage_rules))
    dummy_parent_models = [m] * (len(stage_rules) - 1)
    for rule in stage_rules:
        rule(m, dummy_parent_models)
    return m


def create_EF(stage_rules, BFs):
    """The extensive form.  The symbolic model is the same for every scenario, so this is the flattened
    model tagged with the branching factors (reference pyds/utils.py:92-112 builds n_scenarios copies)."""
    if len(BFs) == 1:
        raise ValueError("Multistage model must contain more than one stage")
    m = create_flattened_model(stage_rules)
    object.__setattr__(m, "BFs", list(int(b) for b in BFs))
    object.__setattr__(m, "n_stages", len(BFs))
    _create_objective(m)
    return m


def load_input(model, input_map, input_values):
    """Validate and split ``{stage: 2-D array (n_scenarios_at_stage, n_params)}`` into per-stage arrays
    (reference pyds/utils.py:123-134 writes every value into its block)."""
    out = {}
    for stage, values in input_values.items():
        if stage not in input_map.keys():
            raise ValueError("Undefined input binding: " + str(stage))
        values = np.asarray(values, dtype=np.float64)
        n_scen = prod(int(b) for b in model.BFs[0:stage + 1])
        if values.ndim != 2 or values.shape[0] < n_scen or values.shape[1] < len(input_map[stage]):
            raise ValueError(f"input for stage {stage} must be (>= {n_scen}, {len(input_map[stage])})")
        out[stage] = np.ascontiguousarray(values[:n_scen, :len(input_map[stage])])
    return out


def parse_value(m, name):
    c = m.component(name)
    if c is None:
        return None
    if isinstance(c, pyo.Var):
        if c.is_indexed():
            return dict(c.values)
        return c.value
    if isinstance(c, pyo.Param):
        if c.is_indexed():
            return dict(c.items())
        return c.value
    return None
