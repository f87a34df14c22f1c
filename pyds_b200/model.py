"""Process-model description and its lowering to the binary tape blob the C-ABI consumes.

This is the Pyomo-free core behind ``pyds_b200.simulator.Simulator``: it describes exactly
the model class the reference's stage rules build (reference ``run_twostage.py:12-79``):

* mutable parameters bound through the ``input map`` (design variables ``d`` and uncertain
  parameters ``theta``)                                          -- ``run_twostage.py:14-15, 38-41, 85``
* differential states with fixed initial values                  -- ``:17-21, 44-45, 72-74``
* piecewise-constant controls, one value per finite element      -- ``:19, 79``
* ODE right-hand sides                                           -- ``:22-23, 49-58``
* critical quality attributes ``g_k <= 0`` in big-M form         -- ``:60-71``, ``pyds/utils.py:7-44``
* Radau collocation ``nfe`` x 3                                  -- ``:76-79``

``DaeModel.lower()`` partitions the expression graph by dependency level, differentiates the
right-hand sides symbolically and serialises everything (tapes, reference tables, collocation
constants) into one blob; see ``include/pydsb.h`` (``pydsb_model_create``) for the layout.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

from . import tape as T
from .collocation import radau_tables

BLOB_MAGIC = 0x42534450  # 'PDSB'
BLOB_VERSION = 1

_DT = {"u32": 0, "f64": 1, "u64": 2, "u16": 3}
_NP = {"u32": np.uint32, "f64": np.float64, "u64": np.uint64, "u16": np.uint16}


@dataclass
class _State:
    name: str
    x0: object
    rhs: object = None
    lo: float = -np.inf            # variable bounds of the user's model (reference run_twostage.py:44: 0 <= c <= 1e8);
    hi: float = np.inf             # not enforced by the state solve -- violations are detected (Solver, 'check state bounds')


@dataclass
class _Control:
    name: str
    lo: float
    hi: float
    init: object          # float, list of floats (per piece) or Expr of parameters
    pieces: int           # number of z entries (nfe, or 1 for a constant-in-time control)
    z0: int = 0           # first index in the z vector
    fixed: bool = False   # fixed controls are parameter expressions, never optimised


class DaeModel:
    def __init__(self, d_names, p_names, nfe=8, ncp=3, newton_rtol=1e-12, newton_max_it=25,
                 quadratures="auto", block_triangular=True, polynomial_blocks=True, second_derivatives=False,
                 fused_chain=True):
        self.graph = T.Graph()
        self.d_names = list(d_names)
        self.p_names = list(p_names)
        self.nfe, self.ncp = int(nfe), int(ncp)
        self.newton_rtol = float(newton_rtol)
        self.newton_max_it = int(newton_max_it)
        # "auto": every state that appears in no right-hand side becomes a quadrature (exact); or an explicit
        # list of state names, e.g. ["u"] keeps cC in the Newton system (the 9-unknown form of SURVEY.md A.3)
        self.quadratures = quadratures
        # solve the strongly connected components of the state dependency graph one after the other
        # (A -> B -> C kinetics: cA alone, then cB, which is linear given cA) instead of one coupled system
        self.block_triangular = bool(block_triangular)
        # a one-state block whose right-hand side is a polynomial of degree <= 2 in its own state (mass-action
        # kinetics) is lowered to its coefficients: the kernel runs that block's Newton loop in registers
        self.polynomial_blocks = bool(polynomial_blocks)
        # analytic second derivatives in the sensitivity program (opt-in: the batched optimiser is Gauss-Newton and
        # does not consume them, so by default the tapes stay short): d2g_k/de_a de_b over the end values
        # e = (x(1), q(1)) after the g tape, d2f_r/dx_a dx_b in the point tape
        self.second_derivatives = bool(second_derivatives)
        # when every block is a polynomial one-state block whose coefficients are single mass-action terms, the assembler
        # fuses the whole finite-element loop into one instruction with the states in registers (csrc/poly_chain.cuh);
        # False keeps the interpreted loop (POLY1 per block and element) -- both are parity-tested
        self.fused_chain = bool(fused_chain)
        self.states: list[_State] = []
        self.controls: list[_Control] = []
        self.constraints = []          # (name, Expr (<= 0 feasible), bigM)
        g = self.graph
        self.d = {n: g.param(i) for i, n in enumerate(self.d_names)}
        self.p = {n: g.param(len(self.d_names) + i) for i, n in enumerate(self.p_names)}
        self._end_leaf = {}

    # -- declaration ------------------------------------------------------------------------------
    @property
    def n_param(self):
        return len(self.d_names) + len(self.p_names)

    def param(self, name):
        return self.d[name] if name in self.d else self.p[name]

    def state(self, name, x0=0.0, bounds=None):
        i = len(self.states)
        lo, hi = (-np.inf, np.inf) if bounds is None else (-np.inf if bounds[0] is None else float(bounds[0]),
                                                           np.inf if bounds[1] is None else float(bounds[1]))
        self.states.append(_State(name, x0, lo=lo, hi=hi))
        return self.graph._mk("x", (), i)            # provisional index = declaration order

    def control(self, name, lo, hi, init=0.0, per_element=True):
        k = len(self.controls)
        self.controls.append(_Control(name, float(lo), float(hi), init, self.nfe if per_element else 1))
        return self.graph.ctrl(k)

    def fixed_control(self, name, value):
        """A control profile that is a function of the parameters only (no recourse)."""
        k = len(self.controls)
        self.controls.append(_Control(name, -np.inf, np.inf, value, 1, fixed=True))
        return self.graph.ctrl(k)

    def ode(self, state, rhs):
        assert state.op == "x"
        self.states[state.value].rhs = self.graph.wrap(rhs)

    def end(self, state):
        """Value of a state at the final time (tau = 1)."""
        assert state.op == "x"
        return self.graph._mk("xe", (), 1000 + state.value)     # provisional, remapped in lower()

    def initial(self, state):
        assert state.op == "x"
        return self.graph.wrap(self.states[state.value].x0)

    def constraint(self, name, expr_le0, big_m):
        self.constraints.append((name, self.graph.wrap(expr_le0), float(big_m)))

    # -- lowering -----------------------------------------------------------------------------------
    def _rebuild(self, e, leaf_map, memo):
        """Re-create ``e`` with leaves substituted (runs the simplifier again)."""
        g = self.graph
        for n in T.topo_order([e]):
            if n.id in memo:
                continue
            if not n.args:
                memo[n.id] = leaf_map.get((n.op, n.value), n)
                continue
            a = [memo[c.id] for c in n.args]
            op = n.op
            if op == "add": r = g.add(*a)
            elif op == "sub": r = g.sub(*a)
            elif op == "mul": r = g.mul(*a)
            elif op == "div": r = g.div(*a)
            elif op == "neg": r = g.neg(*a)
            elif op == "sqr": r = g.mul(a[0], a[0])
            elif op == "rcp": r = g.div(1.0, a[0])
            elif op == "pow": r = g.pow(*a)
            elif op == "min": r = g.minimum(*a)
            elif op == "max": r = g.maximum(*a)
            else: r = getattr(g, op)(*a)
            memo[n.id] = r
        return memo[e.id]

    def lower(self):
        g = self.graph
        n_states = len(self.states)
        for s in self.states:
            if s.rhs is None:
                raise ValueError(f"state '{s.name}' has no ODE")
        # quadrature states: appear in no right-hand side (their end value is a weighted sum)
        used = set()
        for s in self.states:
            for n in T.topo_order([s.rhs]):
                if n.op == "x":
                    used.add(n.value)
        if self.quadratures == "auto":
            quad = [i for i in range(n_states) if i not in used]
        else:
            quad = [i for i in range(n_states) if self.states[i].name in self.quadratures]
            bad = [self.states[i].name for i in quad if i in used]
            if bad:
                raise ValueError(f"states {bad} appear in a right-hand side and cannot be quadratures")
        dyn = [i for i in range(n_states) if i not in quad]
        ns, nq = len(dyn), len(quad)
        if ns == 0:
            raise ValueError("model has no dynamic state")
        leaf_map = {}
        for new, old in enumerate(dyn):
            leaf_map[("x", old)] = g.x(new) if new != old else g._mk("x", (), new)
            leaf_map[("xe", 1000 + old)] = g.xe(new)
        for new, old in enumerate(quad):
            leaf_map[("xe", 1000 + old)] = g.qe(new)
        memo = {}
        rb = lambda e: self._rebuild(g.wrap(e), leaf_map, memo)

        def ctrl_linear(e):
            """An element-level value that is LINEAR in one control, ``alpha(parameters) * z``, rewritten to exactly that
            product: however the user's equations spell it (``F_in[t] / t_f == dudt[t]`` reaches here as
            ``-((z * (1 / t_f)) / -1)``), the elem tape then holds one multiplication and the fused element loop
            (csrc/poly_chain.cuh) can take the control straight from z.  Anything else is returned unchanged."""
            if e is None or e.level != T.LEVEL_CTRL:
                return e
            leaves = {n.value: n for n in T.topo_order([e]) if n.op == "ctrl"}
            if len(leaves) != 1:
                return e
            z = next(iter(leaves.values()))
            pc = g.poly_coeffs(e, z, 1)
            if pc is None or len(pc) != 2 or not pc[0].is_const(0.0) or pc[1].level > T.LEVEL_PARAM:
                return e
            return g.mul(pc[1], z)

        f = [rb(self.states[i].rhs) for i in dyn]
        fq = [ctrl_linear(rb(self.states[i].rhs)) for i in quad]
        x0 = [rb(self.states[i].x0) for i in dyn]
        q0 = [rb(self.states[i].x0) for i in quad]
        for e in x0 + q0:
            if e.level > T.LEVEL_PARAM:
                raise ValueError("initial values may depend on parameters only")
        J = [g.diff(f[r], g.x(c)) for r in range(ns) for c in range(ns)]
        gk = [rb(e) for _, e, _ in self.constraints]
        ng = len(gk)
        if ng == 0:
            raise ValueError("model has no constraint")

        # controls -> z vector
        nce = len(self.controls)
        z_exprs, z_lo, z_hi, z_fixed = [], [], [], []
        cmap = np.zeros((self.nfe, max(nce, 1)), dtype=np.uint16)
        for k, c in enumerate(self.controls):
            c.z0 = len(z_exprs)
            for piece in range(c.pieces):
                if isinstance(c.init, (list, tuple, np.ndarray)):
                    v = c.init[piece if c.pieces > 1 else 0]
                else:
                    v = c.init
                v = rb(v)
                if v.level > T.LEVEL_PARAM:
                    raise ValueError("control values may depend on parameters only")
                z_exprs.append(v)
                z_lo.append(c.lo); z_hi.append(c.hi); z_fixed.append(1 if c.fixed else 0)
            for e in range(self.nfe):
                cmap[e, k] = c.z0 + (e if c.pieces > 1 else 0)
        nz = len(z_exprs)

        for e in gk:
            for n in T.topo_order([e]):
                if n.op in ("ctrl", "x"):
                    raise ValueError("constraints may use end-point values m.end(state) and parameters only")
        # ---- blocks of the state solve: strongly connected components, dependencies first ----------------
        dep = [[c for c in range(ns) if not J[r * ns + c].is_const(0.0)] for r in range(ns)]
        if self.block_triangular:
            blocks = _tarjan_scc(dep)
        else:
            blocks = [list(range(ns))]
        for blk in blocks:
            if len(blk) > 3:
                raise ValueError(f"a block of {len(blk)} mutually coupled dynamic states exceeds the kernel's limit of 3")
        roots = {"x0": x0, "q0": q0, "z": z_exprs, "g": gk, "fq": fq}
        groups = []
        blk_linear = []
        blk_poly = []                 # per block: None, or the references [c0, c1, c2] table name
        for b, blk in enumerate(blocks):
            co = None
            if self.polynomial_blocks and len(blk) == 1:
                co = g.poly_coeffs(f[blk[0]], g.x(blk[0]), 2)
            if co is not None:
                co = [None if c.is_const(0.0) else ctrl_linear(c) for c in co] + [None] * (3 - len(co))
                roots[f"c{b}"] = co
                groups.append((f"pt{b}", [f"c{b}"]))       # the coefficients that depend on earlier blocks' states
                blk_poly.append(f"c{b}")
                blk_linear.append(1 if co[2] is None else 0)
                continue
            blk_poly.append(None)
            roots[f"f{b}"] = [f[i] for i in blk]
            roots[f"J{b}"] = [J[r * ns + c] for r in blk for c in blk]
            groups.append((f"pt{b}", [f"f{b}", f"J{b}"]))
            # linear block: its Jacobian does not depend on its own states -> Newton is exact in one step
            own = {("x", i) for i in blk}
            lin = all(not any((n.op, n.value) in own for n in T.topo_order([e])) for e in roots[f"J{b}"])
            blk_linear.append(1 if lin else 0)
        groups.append(("ptq", ["fq"]))
        # program 1 ("sens"): one coupled block with everything forward control sensitivities need:
        # df/dz, dfq/dx, dfq/dz in the point tape, dg/dxe, dg/dqe after the g tape
        all_blk = list(range(ns))
        roots_s = {"x0": x0, "q0": q0, "z": z_exprs, "g": gk, "fq": fq, "f0": f, "J0": J,
                   "fqx": [g.diff(fq[q], g.x(c)) for q in range(nq) for c in range(ns)],
                   "fz": [g.diff(f[r], g.ctrl(k)) for r in range(ns) for k in range(nce)],
                   "fqz": [g.diff(fq[q], g.ctrl(k)) for q in range(nq) for k in range(nce)],
                   "gx": [g.diff(gk[k], g.xe(c)) for k in range(ng) for c in range(ns)],
                   "gq": [g.diff(gk[k], g.qe(q)) for k in range(ng) for q in range(nq)]}
        groups_s = [("pt0", ["f0", "J0", "fq", "fqx", "fz", "fqz"])]
        if self.second_derivatives:
            ends = [g.xe(c) for c in range(ns)] + [g.qe(q) for q in range(nq)]
            roots_s["gee"] = [g.diff(g.diff(gk[k], ea), eb) for k in range(ng) for ea in ends for eb in ends]
            roots_s["fxx"] = [g.diff(J[r * ns + a], g.x(b)) for r in range(ns) for a in range(ns) for b in range(ns)]
            groups_s = [("pt0", ["f0", "J0", "fq", "fqx", "fz", "fqz", "fxx"])]
        tau, adot, bw = radau_tables(self.ncp)
        secs = []

        def sec(tag, dtype, arr):
            arr = np.ascontiguousarray(np.asarray(arr, dtype=_NP[dtype]).reshape(-1))
            secs.append((tag, dtype, arr))

        def tab(low, name, n):
            return low.refs[name] if n else [T.REF_NONE]

        lows = []
        blk_flops = []
        programs = [(roots, groups, blocks, blk_linear, blk_poly)]
        if ns <= 3:          # the sensitivity program is one coupled block: only for <= 3 dynamic states
            programs.append((roots_s, groups_s, [all_blk], [0], [None]))
            if all(blk_poly) and self.fused_chain and any(not c.fixed for c in self.controls):
                # program 2: the feasibility program again, its g tape extended by dg/dxe and dg/dqe -- when it assembles
                # to a fused chain, the control sensitivities are propagated through the chain in registers
                # (csrc/recourse_chain.cuh) instead of through the coupled-block interpreter of program 1
                roots_c = dict(roots)
                roots_c["gx"] = roots_s["gx"]
                roots_c["gq"] = roots_s["gq"]
                programs.append((roots_c, groups, blocks, blk_linear, blk_poly))
        blk_fixed = []
        for prog, (rts, grp, blks, lin, pol) in enumerate(programs):
            low = T.lower(g, n_param=self.n_param, n_ctrl_elem=nce, ns=ns, nq=nq, roots=rts, pt_groups=grp, ncp=self.ncp)
            lows.append(low)
            lay = low.layout
            if prog:
                sec("PSEP", "u32", [prog])
            tape_names = [name for name, _ in grp]
            pt_code = np.concatenate([low.code[n] for n in tape_names]) if tape_names else np.zeros(0, np.uint64)
            blkt = [len(low.code[f"pt{b}"]) for b in range(len(blks))] + [len(low.code["ptq"]) if "ptq" in low.code else 0]
            blks_tab = []
            flops = []
            fixed = []                # per block: point-tape flops that run once per element (polynomial blocks)
            for b, blk in enumerate(blks):
                # kind: bit 0 linear, bit 1 polynomial (RF__/RJ__/RP2_ hold the references of c0/c1/c2)
                blks_tab += [len(blk), lin[b] | (2 if pol[b] else 0)] + list(blk) + [0] * (3 - len(blk))
                n_b = len(blk) * self.ncp
                solve = 2 * n_b * (self.ncp + 1) + (2 * n_b ** 3) // 3 + 2 * n_b * n_b
                if pol[b]:
                    # Horner f = (c2 x + c1) x + c0 and J = c2 x + (c2 x + c1): three FMAs per point
                    flops.append(6 * self.ncp + solve)
                    fixed.append(low.stats["flops"][f"pt{b}"] * self.ncp)
                else:
                    flops.append(low.stats["flops"][f"pt{b}"] * self.ncp + solve)
                    fixed.append(0)
            if not prog:
                blk_flops = flops
                blk_fixed = fixed
            sec("HEAD", "u32", [ns, nq, ng, len(self.d_names), len(self.p_names), nz, self.nfe, self.ncp, nce,
                                lay.n_units, lay.ctrl_slot, lay.xe_slot, lay.qe_slot, lay.x_slot, self.newton_max_it,
                                0 if self.fused_chain else 1, lay.fq_state_dep, 0, len(blks)])      # HEAD[17]: reserved
            sec("RPAR", "u16", lay.param_slots)
            sec("DBLS", "f64", [self.newton_rtol])
            sec("CNST", "f64", low.consts if len(low.consts) else [0.0])
            sec("ADOT", "f64", adot)
            sec("BWTS", "f64", bw)
            sec("BIGM", "f64", [m for _, _, m in self.constraints])
            sec("ZLOW", "f64", z_lo if nz else [0.0])
            sec("ZHIG", "f64", z_hi if nz else [0.0])
            sec("ZFIX", "u16", z_fixed if nz else [0])
            sec("TPRO", "u64", low.code["pro"])
            sec("TELM", "u64", low.code["elem"])
            sec("TPNT", "u64", pt_code)
            sec("TGEE", "u64", low.code["g"])
            sec("BLKS", "u16", blks_tab)
            sec("BLKT", "u32", blkt)
            sec("BLKF", "u32", flops)
            sec("RX0_", "u16", low.refs["x0"])
            sec("RQ0_", "u16", tab(low, "q0", nq))
            sec("RZ__", "u16", tab(low, "z", nz))
            sec("RF__", "u16", np.concatenate([low.refs[pol[b]][0:1] if pol[b] else low.refs[f"f{b}"] for b in range(len(blks))]))
            sec("RJ__", "u16", np.concatenate([low.refs[pol[b]][1:2] if pol[b] else low.refs[f"J{b}"] for b in range(len(blks))]))
            sec("RP2_", "u16", [int(low.refs[pol[b]][2]) if pol[b] else T.REF_NONE for b in range(len(blks))])
            sec("RFQ_", "u16", tab(low, "fq", nq))
            sec("RG__", "u16", low.refs["g"])
            sec("CMAP", "u16", cmap)
            if prog == 2:
                sec("RGX_", "u16", tab(low, "gx", ng * ns))
                sec("RGQ_", "u16", tab(low, "gq", ng * nq))
            if prog == 1:
                sec("RFQX", "u16", tab(low, "fqx", nq * ns))
                sec("RFZ_", "u16", tab(low, "fz", ns * nce))
                sec("RFQZ", "u16", tab(low, "fqz", nq * nce))
                sec("RGX_", "u16", tab(low, "gx", ng * ns))
                sec("RGQ_", "u16", tab(low, "gq", ng * nq))
                if self.second_derivatives and prog == 1:
                    sec("RGEE", "u16", low.refs["gee"])          # (ng, ns+nq, ns+nq)
                    sec("RFXX", "u16", low.refs["fxx"])          # (ns, ns, ns), wide references
        low = lows[0]
        lay = low.layout

        blob = bytearray(struct.pack("<IIII", BLOB_MAGIC, BLOB_VERSION, len(secs), 0))
        for tag, dtype, arr in secs:
            raw = arr.tobytes()
            pad = (-len(raw)) % 8
            blob += struct.pack("<4sIII", tag.encode(), _DT[dtype], arr.size, 0)
            blob += raw + b"\0" * pad
        dyn_names = [self.states[i].name for i in dyn]
        info = LoweredModel(
            blob=bytes(blob), ns=ns, nq=nq, ng=ng, nz=nz, d_dim=len(self.d_names), p_dim=len(self.p_names),
            nfe=self.nfe, ncp=self.ncp, n_units=lay.n_units, tapes=low, dyn_states=[self.states[i].name for i in dyn],
            quad_states=[self.states[i].name for i in quad], z_lo=np.asarray(z_lo), z_hi=np.asarray(z_hi),
            tapes_sens=lows[1] if len(lows) > 1 else None, tapes_chain=lows[2] if len(lows) > 2 else None, z_init=z_exprs, blocks=[[dyn_names[i] for i in blk] for blk in blocks], block_linear=blk_linear,
            block_flops=blk_flops, block_fixed_flops=blk_fixed, block_poly=[bool(x) for x in blk_poly],
            z_fixed=np.asarray(z_fixed, dtype=bool), big_m=np.asarray([m for _, _, m in self.constraints]),
            constraint_names=[n for n, _, _ in self.constraints], cmap=cmap, newton_rtol=self.newton_rtol,
            newton_max_it=self.newton_max_it,
            state_lo=np.asarray([self.states[i].lo for i in dyn + quad]), state_hi=np.asarray([self.states[i].hi for i in dyn + quad]))
        return info


@dataclass
class LoweredModel:
    blob: bytes
    ns: int
    nq: int
    ng: int
    nz: int
    d_dim: int
    p_dim: int
    nfe: int
    ncp: int
    n_units: int
    tapes: T.LoweredTapes
    tapes_sens: T.LoweredTapes
    z_init: list
    blocks: list
    block_linear: list
    block_flops: list
    block_fixed_flops: list
    block_poly: list
    dyn_states: list
    quad_states: list
    z_lo: np.ndarray
    z_hi: np.ndarray
    z_fixed: np.ndarray
    big_m: np.ndarray
    constraint_names: list
    cmap: np.ndarray
    newton_rtol: float
    newton_max_it: int
    tapes_chain: T.LoweredTapes = None   # program 2: the feasibility tapes with dg/dxe, dg/dqe (chain evaluator)
    state_lo: np.ndarray = None        # bounds of the states in trajectory column order (dynamic, then quadrature)
    state_hi: np.ndarray = None

    def flops_per_newton_iteration(self, block=0):
        """Algorithmic flops of one Newton iteration of one block on one finite element:
        point tape (x ncp) + linear collocation residual + dense LU + two triangular solves."""
        return self.block_flops[block]

    def fixed_flops_per_evaluation(self):
        """Tape work that does not depend on the Newton iteration count."""
        fl = self.tapes.stats["flops"]
        return fl["pro"] + fl["g"] + self.nfe * (fl["elem"] + self.ncp * fl.get("ptq", 0) + sum(self.block_fixed_flops))

    def flops_per_evaluation(self, mean_newton_iters_per_element, newton_flops_per_eval=None):
        """F_alg per (d,theta) evaluation.  With block-triangular solves pass the measured
        ``newton_flops_per_eval`` (device counter: sum over blocks of iterations x block flops)."""
        if newton_flops_per_eval is None:
            newton_flops_per_eval = self.nfe * mean_newton_iters_per_element * sum(self.block_flops)
        return self.fixed_flops_per_evaluation() + newton_flops_per_eval


def _tarjan_scc(dep):
    """Strongly connected components of ``i -> dep[i]``, emitted dependencies first."""
    n = len(dep)
    index, low, on, stack, out = [None] * n, [0] * n, [False] * n, [], []
    counter = [0]

    def visit(v):
        index[v] = low[v] = counter[0]; counter[0] += 1
        stack.append(v); on[v] = True
        for w in dep[v]:
            if w == v:
                continue
            if index[w] is None:
                visit(w); low[v] = min(low[v], low[w])
            elif on[w]:
                low[v] = min(low[v], index[w])
        if low[v] == index[v]:
            comp = []
            while True:
                w = stack.pop(); on[w] = False; comp.append(w)
                if w == v:
                    break
            out.append(sorted(comp))

    for v in range(n):
        if index[v] is None:
            visit(v)
    return out


def parse_blob(blob: bytes):
    """Decode a blob back into {tag: ndarray}; the sensitivity program's sections sit under key "sens"."""
    magic, ver, nsec, _ = struct.unpack_from("<IIII", blob, 0)
    if magic != BLOB_MAGIC or ver != BLOB_VERSION:
        raise ValueError("not a pyds_b200 model blob")
    off = 16
    out = {}
    cur = out
    inv = {v: k for k, v in _DT.items()}
    for _ in range(nsec):
        tag, dt, cnt, _ = struct.unpack_from("<4sIII", blob, off)
        off += 16
        dtype = _NP[inv[dt]]
        nbytes = cnt * np.dtype(dtype).itemsize
        arr = np.frombuffer(blob, dtype=dtype, count=cnt, offset=off).copy()
        off += nbytes + ((-nbytes) % 8)
        if tag == b"PSEP":                               # payload: the program's number (1 sensitivities, 2 chain + dg)
            cur = out.setdefault("sens" if int(arr[0]) == 1 else "chain", {})
            continue
        cur[tag.decode()] = arr
    return out
