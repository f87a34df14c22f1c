"""CPU oracle for the pyds feasibility hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the mathematics that the reference
(cornelmarck/pyds) delegates to Pyomo + GAMS/CONOPT + DEUS for the path
``TwoStageManager.g`` / ``ThreeStageManager.g`` (reference ``pyds/run.py:47-98``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under ``pyds_b200/`` does.

PARITY STATUS: **parity unpinned** for everything that lives in absent third-party
code (Pyomo's collocation transcription, the CONOPT local optimum, DEUS's reduction):
the reference ships no tests, golden vectors or stored outputs and cannot run in this
image (no pyomo / deus / gams).  What *is* pinned against the reference's own
artefacts:

* ``tree.stages_to_update`` against the printed answer of ``pyds/test.py:3-11``;
* the theta sample fixture of ``run_twostage.py:110-120`` (sha256 in ``tests/golden``);
* scenario ordering = ``itertools.product`` order (``pyds/utils.py:57-66``).

Everything else is cross-checked between independent restatements (numpy fp64,
mpmath 50-digit, plain C) and against scipy ODE integration under grid refinement.
"""
