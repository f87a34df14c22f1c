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
    """The interpreted element loop (fused_chain=False): POLY1 per block and element."""
    lm = models.kinetics("design4", fused_chain=False).lower()
    ops, info = capi.assemble(lm.blob)
    names = _names(ops)
    assert names[-1] == "END" and names.count("INIT") == 1 and names.count("XEND") == 1
    # both one-state blocks run as in-register polynomial solves: no NB / NS, two scenarios per thread
    assert names.count("POLY1") == 2 and not any(n.startswith(("NB", "NS")) for n in names)
    assert info["scenarios_per_thread"] == 2 and info["chain_blocks"] == 0
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
    ops, _ = capi.assemble(models.kinetics("twostage", fused_chain=False).lower().blob)
    names = _names(ops)
    a, b = _elem_loop(ops)
    assert names[a] == "LDCTRL" and "LDCTRL" not in names[:a]      # F_in has one entry per finite element


def _chain_fields(pc):
    """Decode a PCHAIN instruction (csrc/poly_chain.cuh): blocks [(x_off, flops, kind, minv, [(n_off, meta)] * 3)], quadratures."""
    w = [int(x) for x in pc]
    nb, nq = w[1] & 0xFF, (w[1] >> 8) & 0xFF
    blocks = [(w[8 + 12 * b], w[9 + 12 * b], w[10 + 12 * b], w[11 + 12 * b],
               [(w[12 + 12 * b + 2 * k], w[13 + 12 * b + 2 * k]) for k in range(3)]) for b in range(nb)]
    q0 = 8 + 12 * nb
    quads = [(w[q0 + 4 * q], w[q0 + 4 * q + 1], w[q0 + 4 * q + 2]) for q in range(nq)]
    return nb, nq, w[2], blocks, quads


def _meta(m):
    return dict(present=m & 1, has_n=(m >> 1) & 1, has_ctrl=(m >> 2) & 1, ctrl=(m >> 4) & 15, p=(m >> 8) & 3, src=(m >> 12) & 3)


def test_fused_chain_program_c3():
    """design4 (config C3): the whole element loop is one PCHAIN between two sections of scalar ops; no INIT, no POLY1."""
    lm = models.kinetics("design4").lower()
    ops, info = capi.assemble(lm.blob)
    names = _names(ops)
    assert names.count("PCHAIN") == 1 and names[-1] == "END" and "INIT" not in names and "POLY1" not in names
    assert set(names) <= {"FMAG", "DIV1", "RCP1", "COLD1", "LDCTRL", "PCHAIN", "XEND", "END"}
    k = names.index("PCHAIN")
    assert names[k + 1] == "XEND" and "LDCTRL" in names[:k]                    # Fbar: element invariant, loaded before the chain
    assert info["chain_blocks"] == 2 and info["chain_shape_registered"] == 1
    assert (info["chain_shape_sb"], info["chain_shape_sq"]) == (0x18B980062, 0x5C1A)        # csrc/chain_shapes.h
    nb, nq, max_it, blocks, quads = _chain_fields(ops[k][2])
    assert (nb, nq, max_it) == (2, 2, lm.newton_max_it)
    (xa, fa, ka, _, ca), (xb, fb, kb, minv, cb) = blocks
    assert ka == 0 and fa == lm.block_flops[0]                                 # cA: Newton
    assert [_meta(m)["present"] for _, m in ca] == [1, 0, 1] and all(_meta(m)["p"] == 0 for _, m in ca)
    assert kb == 2 and fb == 24                                                # cB: linear, constant decay -> one inverse
    m0 = _meta(cb[0][1])
    assert (m0["present"], m0["has_n"], m0["p"], m0["src"]) == (1, 1, 2, 0)    # c0 = narrow * cA^2
    assert _meta(cb[1][1])["p"] == 0 and _meta(cb[2][1])["present"] == 0
    unit = 2 * 1024 * 8 // 8                                                   # bytes per unit, two scenarios per thread
    assert minv % unit == 0 and minv // unit + 9 <= info["units"]              # scratch: nine consecutive units
    assert info["units"] < lm.tapes.layout.n_units                             # renumbered by liveness (compact_units)
    assert [(_meta(m)["p"], _meta(m)["src"]) for _, _, m in quads] == [(0, 0), (1, 1)]      # u: element constant; cC: k * cB
    # every non-FMAG scalar op carries the length of the FMAG run that follows it
    for sec in (range(0, k), range(k + 2, len(ops) - 1)):
        run = 0
        for i in reversed(list(sec)):
            if names[i] == "FMAG":
                run += 1
            else:
                assert int(ops[i][2][7]) == run
                run = 0


def _run_chain_program(lm, par, monkeypatch, compact):
    """Execute the scalar dataflow of an assembled chain program in Python (one scenario): the two scalar sections as
    the kernel runs them, PCHAIN as a black box that reads its inputs, poisons its scratch and then writes values that
    depend on every input.  Returns the constraint values and the number of units."""
    import math
    from pyds_b200.tape import OPS
    if compact: monkeypatch.delenv("PYDSB_NO_COMPACT", raising=False)
    else: monkeypatch.setenv("PYDSB_NO_COMPACT", "1")
    blob, consts = lm.blob, lm.tapes.consts
    ops, info = capi.assemble(blob)
    cd = capi.chain_descriptor(blob)
    spt = info["scenarios_per_thread"]
    unit, cw = spt * 1024, 8 * spt
    S = [float("nan")] * info["units"]

    def cst(off):
        i = off // cw
        return (1.0, -0.0, 0.0)[i] if i < 3 else float(consts[i - 3])

    def ld(off, is_c):
        if is_c: return cst(off)
        assert off % unit == 0 and off // unit < info["units"]
        return S[off // unit]

    assert len(cd["par_off"]) == len(par) and len(set(cd["par_off"])) == len(par)      # all written at the tile start
    for o, v in zip(cd["par_off"], par): S[o // unit] = v
    names = [n for _, n, _ in ops]
    k = names.index("PCHAIN")
    assert ops[k][0] == cd["pre_end"] and ops[k + 2][0] == cd["g_begin"] if k + 2 < len(ops) - 1 else True

    def scalar(sec):
        for _, name, w in sec:
            w = [int(x) for x in w]
            fl = w[5]
            if name == "LDCTRL":                                       # z_mode 0: the control's initial / fixed value
                zi = int(lm.cmap[0, w[6]])
                r = ld(cd["z_off"][zi], cd["z_const"][zi])
            else:
                a, b = ld(w[2], fl & 1), ld(w[3], fl & 2)              # the kernel loads a and b of every op
                if name == "FMAG":
                    c = ld(w[4], fl & 4)
                    r = (-a if fl & 8 else a) * b + (-c if fl & 16 else c)
                elif name == "DIV1": r = a / b
                elif name == "RCP1": r = 1.0 / a
                elif name == "COLD1": r = {"EXP": math.exp, "LOG": math.log, "SQRT": math.sqrt}[OPS[w[6]]](a)
                else: raise AssertionError(name)
            S[w[1] // unit] = r

    scalar(ops[:k])
    ins = []
    for b in cd["blocks"]:
        ins += [ld(o, False) for o, m in b["terms"] if (m & 1) and (m & 2)]
        ins.append(ld(b["x0_off"], b["x0_const"]))
    for q in cd["quads"]:
        if (q["meta"] & 1) and (q["meta"] & 2): ins.append(ld(q["n_off"], False))
        ins.append(ld(q["q0_off"], q["q0_const"]))
    if any(m & 4 for b in cd["blocks"] for _, m in b["terms"]) or any(q["meta"] & 4 for q in cd["quads"]):
        ins += [ld(cd["z_off"][int(zi)], cd["z_const"][int(zi)]) for zi in lm.cmap.ravel()]    # per-element controls
    assert all(v == v for v in ins)                                    # no input was overwritten before the chain
    for b in cd["blocks"]:
        if b["kind"] == 2:
            for i in range(9): S[b["minv_off"] // unit + i] = float("nan")
    mix = sum((i + 1) * math.atan(v) for i, v in enumerate(ins))
    outs = [b["xe_off"] for b in cd["blocks"]] + [q["q_off"] for q in cd["quads"]]
    assert len(set(outs)) == len(outs)
    for i, o in enumerate(outs): S[o // unit] = mix + i
    scalar(ops[k + 2:-1])
    return [ld(o, False) for o in cd["g_off"]], info["units"]


def _three_block_chain():
    """An unregistered chain: three blocks, a bimolecular term, a quadrature without narrow factor, a control-only one,
    per-element controls with initial values (tests/test_gpu_chain.py runs the same model on the GPU)."""
    m = DaeModel(["a"], ["k1", "k2"], nfe=6)
    a, k1, k2 = m.d["a"], m.p["k1"], m.p["k2"]
    cA, cB, cC = m.state("cA", 1.0 + a), m.state("cB", 0.1), m.state("cC", 0.0)
    q1, q2 = m.state("q1", 0.0), m.state("q2", 0.5)
    F = m.control("F", 0.0, 1.0, init=[0.1, 0.3, 0.0, 0.2, 0.5, 0.4])
    m.ode(cA, F - k1 * cA ** 2)
    m.ode(cB, k1 * cA ** 2 - k2 * cA * cB)
    m.ode(cC, 0.7 * cB - (0.2 + a) * cC)
    m.ode(q1, cC * cC)
    m.ode(q2, F)
    m.constraint("c1", m.end(cC) - 0.4, 10.0)
    m.constraint("c2", m.end(q1) + m.end(q2) - m.end(cB) - 0.55, 10.0)
    return m


@pytest.mark.parametrize("variant", ["design4", "design6", "twostage", "three-block"])
def test_unit_compaction_preserves_the_dataflow(variant, monkeypatch):
    """The liveness-based renumbering of a chain program's units (assembler.h, compact_units) must not change what the
    program computes: same g from the renumbered and the original program, fewer units, nothing read after its unit was
    given to another value (such a read would see the other value or the poison the scratch is filled with)."""
    lm = (_three_block_chain() if variant == "three-block" else models.kinetics(variant)).lower()
    rng = np.random.default_rng(7)
    n_par = len(lm.tapes.layout.param_slots)
    par = list(rng.uniform(0.5, 2.0, n_par))
    g1, u1 = _run_chain_program(lm, par, monkeypatch, compact=True)
    g0, u0 = _run_chain_program(lm, par, monkeypatch, compact=False)
    assert all(v == v for v in g0) and g1 == g0
    assert u1 < u0 and u1 <= 20                      # 4 CTAs of 128 threads x 2 scenarios per SM need <= 25 units


@pytest.mark.parametrize("seed", range(24))
def test_unit_compaction_on_random_chains(seed, monkeypatch):
    """The same check over randomly drawn chain-shaped models (tests/random_chains.py): every combination of term kinds,
    free and fixed controls, 1-3 blocks, 0-2 quadratures."""
    from random_chains import random_chain_model
    m, nb = random_chain_model(seed)
    lm = m.lower()
    if capi.assemble(lm.blob)[1]["chain_blocks"] == 0:
        pytest.skip("not chain shaped (more than one term in a coefficient)")
    rng = np.random.default_rng(seed)
    par = list(rng.uniform(0.5, 2.0, len(lm.tapes.layout.param_slots)))
    g1, u1 = _run_chain_program(lm, par, monkeypatch, compact=True)
    g0, u0 = _run_chain_program(lm, par, monkeypatch, compact=False)
    assert all(v == v for v in g0) and g1 == g0 and u1 <= u0


def test_fused_chain_per_element_controls():
    """twostage: F_in differs between the elements, so no LDCTRL / elem ops remain -- the chain multiplies the narrow
    factor by the element's control itself (c0 of cA = t_f * F_in, u integrand = F_in / t_f)."""
    ops, info = capi.assemble(models.kinetics("twostage").lower().blob)
    names = _names(ops)
    assert "LDCTRL" not in names and names.count("PCHAIN") == 1
    assert info["chain_shape_registered"] == 1 and info["chain_shape_sb"] == 0x18B980062 | 4
    _, _, _, blocks, quads = _chain_fields(ops[names.index("PCHAIN")][2])
    m = _meta(blocks[0][4][0][1])
    assert (m["has_n"], m["has_ctrl"], m["ctrl"], m["p"]) == (1, 1, 0, 0)
    mq = _meta(quads[0][2])
    assert (mq["has_ctrl"], mq["ctrl"], mq["p"]) == (1, 0, 0)
    assert all(_meta(mm)["has_ctrl"] == 0 for _, mm in blocks[1][4]) and _meta(quads[1][2])["has_ctrl"] == 0


def test_models_outside_the_chain_shape_keep_the_interpreted_loop():
    # a right-hand side with two mass-action terms in c0 (k3 cD + k1 cA^2) is not a single term
    m = DaeModel(["a"], ["k"], nfe=4)
    xa, xd, xb = m.state("xa", 1.0), m.state("xd", 0.5), m.state("xb", 0.0)
    k = m.p["k"]
    m.ode(xa, -k * xa * xa)
    m.ode(xd, -0.3 * xd)
    m.ode(xb, k * xa * xa + 0.3 * xd - m.d["a"] * xb)
    m.constraint("c", m.end(xb) - 0.4, 1.0)
    ops, info = capi.assemble(m.lower().blob)
    assert info["chain_blocks"] == 0 and _names(ops).count("POLY1") == 3 and "INIT" in _names(ops)


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
