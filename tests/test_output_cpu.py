"""The append-only pickle log (reference pyds/output_manager.py): the one-write path of the two-stage manager must produce
the byte stream of one ``pickle.dump`` per design point."""
import pickle

import numpy as np

from pyds_b200.output_manager import OutputManager


def _reference_stream(folder, key_d, key_p, d, p, solver_output):
    om = OutputManager(str(folder))
    for i in range(len(d)):                                   # reference pyds/run.py:55-67, save output off
        om.clear_buffer()
        om.add_input({key_d: d[i:i + 1], key_p: p})
        om.add_solver_solution(solver_output)
        om.write_data_to_disk()
    with open(om.output_path_fnx, "rb") as f:
        return f.read()


def _fast_stream(folder, key_d, key_p, d, p, solver_output):
    om = OutputManager(str(folder))
    om.write_input_records(key_d, key_p, d, p, solver_output)
    with open(om.output_path_fnx, "rb") as f:
        return f.read()


def test_one_write_equals_one_dump_per_design_point(tmp_path):
    rng = np.random.default_rng(5)
    out = {"relaxation": None, "trajectories": None}
    for n_d, n_p in ((150, 50), (3, 1), (2, 7), (1, 4)):
        d, p = rng.uniform(250, 350, (n_d, 2)), rng.uniform(0, 1, (n_p, 4))
        (tmp_path / f"a{n_d}").mkdir(); (tmp_path / f"b{n_d}").mkdir()
        a = _reference_stream(tmp_path / f"a{n_d}", 0, 1, d, p, out)
        b = _fast_stream(tmp_path / f"b{n_d}", 0, 1, d, p, out)
        assert a == b
    # the records load back as the reference's: one per design point, the design point's own row
    import io
    f = io.BytesIO(b)
    rec = pickle.load(f)
    assert set(rec) == {"input", "solver", "simulator"} and rec["input"][0].shape == (1, 2) and rec["simulator"] is None


def test_ambiguous_row_bytes_fall_back_to_pickling_every_record(tmp_path):
    # the first design point's bytes also occur inside p: the template cannot be patched blindly
    d = np.array([[1.5, 2.5], [3.0, 4.0], [5.0, 6.0], [7.0, 8.0]])
    p = np.array([[1.5, 2.5, 0.0, 0.0], [9.0, 9.0, 9.0, 9.0]])
    out = {"relaxation": None, "trajectories": None}
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    assert _reference_stream(tmp_path / "a", 0, 1, d, p, out) == _fast_stream(tmp_path / "b", 0, 1, d, p, out)
    # rows with equal values, empty batches
    d2 = np.zeros((5, 2))
    (tmp_path / "c").mkdir(); (tmp_path / "e").mkdir()
    assert _reference_stream(tmp_path / "c", 0, 1, d2, p, out) == _fast_stream(tmp_path / "e", 0, 1, d2, p, out)
    om = OutputManager(str(tmp_path))
    om.write_input_records(0, 1, np.zeros((0, 2)), p, out)
    import os
    assert not os.path.exists(om.output_path_fnx)
