"""The reference's OWN stage rules, verbatim: lines 12-79 of /root/reference/run_twostage.py (stage0_rule, stage1_rule,
apply_collocation) are executed against pyds_b200's modelling layer -- nothing but the two import names ``pyo`` / ``dae``
and ``add_BigMConstraint`` is substituted -- lowered, and compared with the oracle's restatement of the same equations.
Pins the lowering on the reference's text rather than on this repository's copy of the model (examples/reactor_model.py).
Skipped where the reference tree is absent (the GPU box)."""
import os

import numpy as np
import pytest

import pyds_b200.environ as pyo
from oracle import kinetics as K, tape_vm
from pyds_b200 import dae, utils
from pyds_b200.simulator import lower_stage_model

REF = "/root/reference/run_twostage.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


def _reference_rules():
    lines = open(REF).read().splitlines()
    src = "\n".join(lines[11:79])                            # lines 12-79: the three functions, nothing else
    assert src.lstrip().startswith("def stage0_rule") and "def apply_collocation" in src and "config_form" not in src
    ns = {"pyo": pyo, "dae": dae, "add_BigMConstraint": utils.add_BigMConstraint, "np": np}
    exec(compile(src, REF, "exec"), ns)
    return ns["stage0_rule"], ns["stage1_rule"], ns["apply_collocation"]


def test_verbatim_reference_rules_lower_to_the_oracle_equations():
    s0, s1, transform = _reference_rules()
    m = utils.create_flattened_model([s0, s1])
    transform(m)
    input_map = {0: ["t_f", "T"], 1: ["E1", "E2", "k1", "k2"]}       # reference run_twostage.py:85
    lm = lower_stage_model(m, input_map, m.component("F_in_default")).lower()
    assert (lm.ns, lm.nq, lm.ng, lm.nz, lm.d_dim, lm.p_dim, lm.nfe, lm.ncp) == (2, 2, 2, 8, 2, 4, 8, 3)
    np.testing.assert_array_equal(lm.big_m, [1e2, 1e7])                # bigM_constant, reference :69
    np.testing.assert_array_equal(lm.z_lo, 0.0)                         # F_in bounds (0, 2.5), reference :19
    np.testing.assert_array_equal(lm.z_hi, 2.5)
    assert [e.value for e in lm.z_init] == [0.0] * 8                    # F_in_default = {0: 0}, reference :25-26
    vm = tape_vm.TapeVM(lm.blob)
    rng = np.random.default_rng(2026)
    d = np.stack([rng.uniform(250, 350, 12), rng.uniform(250, 300, 12)], axis=1)
    p = K.theta_samples()
    for F in (np.zeros(8), rng.uniform(0, 2.5, 8)):
        r = vm.evaluate(d, p, z=F, tol=1e-13)
        g1, g2, _ = K.g_grid(d, p, F=F)
        assert np.abs(r["g"][..., 0] - g1).max() <= 1e-13             # |g1| < 1
        assert np.all(np.abs(r["g"][..., 1] - g2) <= 1e-10 * np.maximum(np.abs(g2), 1e5))
    # the fused element loop is what this model assembles to (the registered kinetics chain with per-element feed)
    from pyds_b200 import capi
    info = capi.assemble(lm.blob)[1]
    assert info["chain_blocks"] == 2 and info["chain_shape_registered"] == 1


def test_verbatim_reference_rules_build_the_reference_scenario_tree():
    """utils.create_EF on the reference's rules: the tree [1, n_p] is implicit here (one symbolic model, scenarios are
    index tuples in the reference's product order, utils.py:53-76), the big-M rows of reference :69-71 are recorded."""
    s0, s1, transform = _reference_rules()
    m = utils.create_EF([s0, s1], [1, 5])
    transform(m)
    assert list(m.BFs) == [1, 5] and m.n_stages == 2
    assert utils.get_final_scenarios(m) == [(0, j) for j in range(5)]
    assert [(name, M) for name, _, M in m._bigm] == [("cqa1", 1e2), ("cqa2", 1e7)]
