"""Host side of the feasibility interpreter, without a GPU: ``pydsb_assemble`` turns a model blob into the
pre-decoded per-scenario program (csrc/assembler.h).  Checks the program structure the kernel relies on."""
import numpy as np
import pytest

from pyds_b200 import capi, models
from pyds_b200.model import DaeModel


def _names(ops):
    return [o[1] for o in ops]


def _elem_loop(ops):
    """(index of the first op of the element loop, index of ELEM_END)."""
    end = _names(ops).index("ELEM_END")
    target = int(ops[end][2][1])
    start = [i for i, o in enumerate(ops) if o[0] == target]
    assert len(start) == 1
    return start[0], end


def test_polynomial_blocks_program_c3():
    lm = models.kinetics("design4").lower()
    ops, info = capi.assemble(lm.blob)
    names = _names(ops)
    assert names[-1] == "END" and names.count("INIT") == 1 and names.count("XEND") == 1
    # both one-state blocks run as in-register polynomial solves: no NB / NS, two scenarios per thread
    assert names.count("POLY1") == 2 and not any(n.startswith(("NB", "NS")) for n in names)
    assert info["scenarios_per_thread"] == 2
    a, b = _elem_loop(ops)
    # Fbar is one control entry for every element: its load and the elem tape are hoisted out of the loop
    assert "LDCTRL" in names[:a] and "LDCTRL" not in names[a:b]
    # k * cA^2 (the coefficient c0 of the cB block) is formed inside its POLY1 (flag 16, a third word with the narrow and
    # the wide operand); the cC integrand k * cB is folded into its ELEM_END entry as a scale factor: 3 dispatches
    assert names[a:b + 1] == ["POLY1", "POLY1", "ELEM_END"]
    # ELEM_END advances the two quadratures (u and cC) itself: two uint4 per entry, the second with a scale
    assert int(ops[b][2][2]) == 2 and len(ops[b][2]) == 4 * (1 + 2 * 2)
    assert [int(ops[b][2][4 + 8 * q + 5]) for q in range(2)] == [0, 1]
    # POLY1 of cA: nonlinear, c1 absent (-> zero slot); of cB: linear (flag bit 3), c0 wide (flag bit 0)
    pa, pb = [o for o in ops if o[1] == "POLY1"]
    assert int(pa[2][7]) & 8 == 0 and int(pb[2][7]) & 8 == 8 and int(pb[2][7]) & 1 == 1
    assert int(pa[2][7]) & 16 == 0 and len(pa[2]) == 8 and int(pb[2][7]) & 16 == 16 and len(pb[2]) == 12
    assert int(pb[2][9]) == int(pa[2][1])                      # the wide operand is the cA block's state (its x offset)
    assert int(pa[2][3]) == lm.newton_max_it and int(pa[2][2]) == lm.block_flops[0]
    # program length in 16-byte words == position after END
    assert info["program_length"] == ops[-1][0] + 2


def test_per_element_controls_stay_inside_the_loop():
    ops, _ = capi.assemble(models.kinetics("twostage").lower().blob)
    names = _names(ops)
    a, b = _elem_loop(ops)
    assert names[a] == "LDCTRL" and "LDCTRL" not in names[:a]      # F_in has one entry per finite element


def test_generic_paths_and_kernel_instantiations():
    for kw, spt, newton in (({"polynomial_blocks": False}, 1, ["NB1", "NS1", "NB1", "NS1L"]),
                            ({"block_triangular": False}, 1, ["NB2", "NS2"]),
                            ({"dense9": True}, 1, ["NB3", "NS3"])):
        ops, info = capi.assemble(models.kinetics("design4", **kw).lower().blob)
        names = _names(ops)
        assert [n for n in names if n.startswith(("NB", "NS"))] == newton
        assert info["scenarios_per_thread"] == spt and "POLY1" not in names
        # an NS instruction loops back to the first point op of its block
        for i, o in enumerate(ops):
            if o[1].startswith("NS"):
                nb = max(j for j in range(i) if names[j].startswith("NB"))
                assert int(o[2][1]) == ops[nb + 1][0]
                assert (int(o[2][3]) >> 16) == 25                  # iteration cap travels in the instruction


def test_point_ops_are_pattern_specialised_and_commutative_operands_ordered():
    m = DaeModel(["a"], ["b"], nfe=4)
    g = m.graph
    x, y = m.state("x", 1.0), m.state("y", 0.5)
    a, b = m.d["a"], m.p["b"]
    m.ode(x, g.tanh(y) * a - x * y + b)            # wide*narrow written wide first; coupled 2-state block
    m.ode(y, x - g.exp(-y) / (1.0 + x * x))
    m.constraint("c", m.end(x) - m.end(y), 10.0)
    ops, info = capi.assemble(m.lower().blob)
    names = _names(ops)
    assert "NB2" in names and "NS2" in names and "COLD_W" in names
    wide = [n for n in names if n.endswith(("_NW", "_WN", "_WW", "_NWN", "_NWW", "_WWN", "_WWW", "_W"))]
    assert wide and not any(n.endswith("_NN") for n in names)


def test_bad_blobs_fail_loudly():
    lm = models.kinetics("design4").lower()
    with pytest.raises(capi.PydsbError):
        capi.assemble(lm.blob[:100])
    bad = bytearray(lm.blob)
    bad[0] ^= 0xFF
    with pytest.raises(capi.PydsbError):
        capi.assemble(bytes(bad))
