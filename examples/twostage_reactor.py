"""Two-stage design-space problem of the reference (``run_twostage.py``), on pyds_b200.

Stage 0: design variables (t_f, T) and the feed profile F_in shared by all theta scenarios.
Stage 1: uncertain kinetics theta = (E1, E2, k1, k2), 50 samples.
Run:  python examples/twostage_reactor.py            (needs a CUDA device)
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pyds_b200.environ as pyo                              # noqa: E402
from examples.reactor_model import (apply_collocation, declare_feed, declare_kinetics,   # noqa: E402
                                    solver_and_simulator_config, theta_samples)
from pyds_b200.run import TwoStageManager                    # noqa: E402

try:
    from deus import DEUS                                     # the reference's sampler, if installed
except ImportError:
    from pyds_b200.ns import DEUS                             # built-in nested-sampling driver


def stage0_rule(m, stage=None):
    m.T = pyo.Param(initialize=0, mutable=True)
    m.t_f = pyo.Param(initialize=0, mutable=True)
    declare_feed(m, t_f=lambda: m.t_f, warm_start={0: 0})   # F_in = 0 (reference run_twostage.py:26)


def stage1_rule(m, stage):
    s0 = stage[0]
    declare_kinetics(m, s0.t, t_f=lambda: s0.t_f, T=lambda: s0.T, u=s0.u, F_in=s0.F_in)


def build_config(output_folder=None, **io_options):
    solver, simulator = solver_and_simulator_config(**io_options)
    return {"problem": {"stage rules": [stage0_rule, stage1_rule], "model transformation": apply_collocation,
                        "input map": {0: ["t_f", "T"], 1: ["E1", "E2", "k1", "k2"]}, "output map": {1: ["c"]},
                        "output folder": output_folder or os.getcwd()},
            "solver": solver, "simulator": simulator}


def activity_form(manager, p_best, p_samples):
    """The DEUS 'ds' activity of reference run_twostage.py:122-210."""
    return {
        "activity_type": "ds",
        "activity_settings": {"case_name": "DS_Test1", "case_path": os.getcwd(), "resume": False, "save_period": 10},
        "problem": {"user_script_filename": "ds_test2_user_script", "constraints_func_name": "g_func_cat2",
                    "parameters_best_estimate": p_best, "parameters_samples": p_samples, "target_reliability": 0.65,
                    "design_variables": [{"t_f": [250, 350]}, {"T": [250, 300]}]},
        "solver": {"name": "ds-ns",
                   "settings": {"score_evaluation": {"method": "serial", "constraints_func_ptr": manager.g, "store_constraints": False},
                                "efp_evaluation": {"method": "serial", "constraints_func_ptr": manager.g, "store_constraints": False,
                                                   "acceleration": False},
                                "points_schedule": [(.0, 100, 10), (.1, 110, 11), (.2, 120, 12), (.4, 130, 13), (.55, 140, 14),
                                                    (.65, 150, 15)],
                                "stop_criteria": [{"inside_fraction": 1.0}]},
                   "algorithms": {"sampling": {"algorithm": "mc_sampling-ns_global",
                                               "settings": {"nlive": 10, "nreplacements": 5, "prng_seed": 1989, "f0": 0.1, "alpha": 0.3,
                                                            "stop_criteria": [{"max_iterations": 100000}], "debug_level": 0,
                                                            "monitor_performance": False},
                                               "algorithms": {"replacement": {"sampling": {"algorithm": "suob-ellipsoid"}}}}}},
    }


if __name__ == "__main__":
    m = TwoStageManager(build_config())
    p_best, p_samples = theta_samples()
    the_deus = DEUS(activity_form(m, p_best, p_samples))
    t0 = time.time()
    the_deus.solve()
    print(time.time() - t0)
