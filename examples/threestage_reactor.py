"""Three-stage problem of the reference (``run_threestage.py``), REPAIRED so that it can run.

The shipped script cannot execute: its config has no ``problem['output map']``
(reference run_threestage.py:80-85 vs pyds/run.py:18), and ``stage0_rule`` uses ``m.u``, ``m.F_in``,
``m.dudt`` and ``m.t_f`` without declaring them (reference run_threestage.py:12-20).  Repairs made here,
and nothing else:

1. ``'output map'`` moved under ``problem``;
2. the feed F_in (with its integral u) is declared at stage 0 -- it is the control shared by every
   design point and every theta, as the warm-start Suffix of the shipped stage0_rule implies;
3. t_f and T stay at stage 1 as shipped; the stage-0 equation ``F_in/t_f == dudt`` refers to them,
   which only the flattened model can express -- the kernel evaluates exactly that flattened model;
4. the shipped ``reduce_collocation_points(var=m.u)`` (which would freeze the *state* u) is dropped.

Branching factors [1, n_d, n_p], scenario order and the per-design-point slicing follow reference
pyds/run.py:76-96.  Warm start F_in = 2.5 on [0, 0.2), 0 afterwards (reference run_threestage.py:19).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pyds_b200.environ as pyo                              # noqa: E402
from examples.reactor_model import (apply_collocation, declare_feed, declare_kinetics,   # noqa: E402
                                    solver_and_simulator_config, theta_samples)
from pyds_b200.run import ThreeStageManager                  # noqa: E402

try:
    from deus import DEUS
except ImportError:
    from pyds_b200.ns import DEUS


def stage0_rule(m, stage=None):
    declare_feed(m, t_f=lambda: m.t_f, warm_start={0: 2.5, .2: 0})     # m.t_f is written by stage 1 (flattened model)


def stage1_rule(m, stage):
    m.T = pyo.Param(initialize=0, mutable=True)
    m.t_f = pyo.Param(initialize=0, mutable=True)


def stage2_rule(m, stage):
    s0, s1 = stage[0], stage[1]
    declare_kinetics(m, s0.t, t_f=lambda: s1.t_f, T=lambda: s1.T, u=s0.u, F_in=s0.F_in)


def build_config(output_folder=None, **io_options):
    solver, simulator = solver_and_simulator_config(**io_options)
    return {"problem": {"stage rules": [stage0_rule, stage1_rule, stage2_rule], "model transformation": apply_collocation,
                        "input map": {1: ["t_f", "T"], 2: ["E1", "E2", "k1", "k2"]}, "output map": {2: ["c"]},
                        "output folder": output_folder or os.getcwd()},
            "solver": solver, "simulator": simulator}


def activity_form(manager, p_best, p_samples):
    from examples.twostage_reactor import activity_form as two
    form = two(manager, p_best, p_samples)
    form["problem"]["design_variables"] = [{"t_f": [250, 400]}, {"T": [250, 300]}]          # run_threestage.py:136-139
    form["solver"]["settings"]["points_schedule"] = [(.0, 100, 10), (.8, 130, 13)]          # run_threestage.py:171-174
    return form


if __name__ == "__main__":
    m = ThreeStageManager(build_config())
    p_best, p_samples = theta_samples()
    the_deus = DEUS(activity_form(m, p_best, p_samples))
    t0 = time.time()
    the_deus.solve()
    print(time.time() - t0)
