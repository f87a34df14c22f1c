"""Expression graph -> flat fp64 tape lowering (host side of the tape interpreter kernel).

The reference hands its model to Pyomo, which expands every constraint of every scenario
block into a GAMS file per solve (reference ``pyds/utils.py:92-112``, ``pyds/solver.py:50-55``).
Here the model is lowered ONCE into straight-line tapes; ``pydsb_model_create`` assembles them into the
pre-decoded per-scenario program (``csrc/assembler.h``) that the CUDA interpreter (``csrc/feas_kernel.cuh``)
runs per (design point, theta) scenario:

==========  ========================================  =========================================
tape        depends on                                runs
==========  ========================================  =========================================
``pro``     parameters (d, theta) only                once per scenario
``elem``    + the controls active in a finite element  once per finite element
``pt``      + the states at one collocation point      per Newton iteration, 3 points at once
``g``       + end-point states / quadratures           once per scenario
==========  ========================================  =========================================

Every node is hash-consed (CSE), constant-folded, differentiated symbolically (first
derivatives for the Newton Jacobian and control sensitivities, second derivatives on
request), fused into FMA forms and assigned a slot of the per-thread register file that
the kernel keeps in shared memory.  An instruction is one 64-bit word::

    op[7:0] | dst[21:8] | a[35:22] | b[49:36] | c[63:50]

and an operand reference is 14 bits: bit 13 = constant-pool index, bit 12 = "wide" slot
(three consecutive slots, one per collocation point), bits 11:0 = slot / constant index.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

import numpy as np

# ---------------------------------------------------------------------------------------------
# opcodes -- must match csrc/tape_vm.cuh
# ---------------------------------------------------------------------------------------------
OPS = ["MOV", "NEG", "ADD", "SUB", "MUL", "DIV", "FMA", "FMS", "FNMA", "SQR", "RCP", "EXP", "LOG",
       "SQRT", "SIN", "COS", "TANH", "POW", "ABS", "MIN", "MAX", "BC3"]
OPCODE = {n: i for i, n in enumerate(OPS)}

REF_CONST = 1 << 13
REF_WIDE = 1 << 12
REF_MASK = 0xFFF
REF_NONE = 0x3FFF          # "structurally zero / absent" in reference tables

LEVEL_CONST, LEVEL_PARAM, LEVEL_CTRL, LEVEL_STATE, LEVEL_END = 0, 1, 2, 3, 4


class Expr:
    """A node of the expression DAG.  Create leaves through :class:`Graph`."""
    __slots__ = ("g", "op", "args", "value", "level", "id")

    def __init__(self, g, op, args, value, level, id_):
        self.g, self.op, self.args, self.value, self.level, self.id = g, op, args, value, level, id_

    # arithmetic ------------------------------------------------------------------------------
    def __add__(self, o): return self.g.add(self, o)
    def __radd__(self, o): return self.g.add(o, self)
    def __sub__(self, o): return self.g.sub(self, o)
    def __rsub__(self, o): return self.g.sub(o, self)
    def __mul__(self, o): return self.g.mul(self, o)
    def __rmul__(self, o): return self.g.mul(o, self)
    def __truediv__(self, o): return self.g.div(self, o)
    def __rtruediv__(self, o): return self.g.div(o, self)
    def __neg__(self): return self.g.neg(self)
    def __pos__(self): return self
    def __pow__(self, o): return self.g.pow(self, o)
    def __rpow__(self, o): return self.g.pow(o, self)

    # relations (used by the modelling layer: ``lhs == rhs``, ``body <= ub``) ---------------------------
    def __eq__(self, o): return Relational(self, o, "==")
    def __le__(self, o): return Relational(self, o, "<=")
    def __ge__(self, o): return Relational(self, o, ">=")
    def __hash__(self): return id(self)

    def is_const(self, v=None):
        return self.op == "const" and (v is None or self.value == v)

    def __repr__(self):
        if self.op == "const":
            return repr(self.value)
        if self.op in ("param", "ctrl", "x", "xe", "qe"):
            return f"{self.op}{self.value}"
        return f"{self.op}({', '.join(map(repr, self.args))})"


class Relational:
    """``lhs (==|<=|>=) rhs`` recorded symbolically (never evaluated to a bool)."""
    __slots__ = ("lhs", "rhs", "op")

    def __init__(self, lhs, rhs, op):
        self.lhs, self.rhs, self.op = lhs, rhs, op

    def __bool__(self):
        raise TypeError("a symbolic relation has no truth value")


class Graph:
    """Hash-consing expression factory with algebraic simplification."""

    # "v" (variable reference), "dv" (time derivative) and "alg" (algebraic variable) are placeholders
    # of the modelling layer; they are substituted before lowering.
    LEAF_LEVEL = {"param": LEVEL_PARAM, "ctrl": LEVEL_CTRL, "x": LEVEL_STATE, "xe": LEVEL_END, "qe": LEVEL_END,
                  "v": LEVEL_STATE, "dv": LEVEL_STATE, "alg": LEVEL_END, "pn": LEVEL_PARAM}

    def __init__(self):
        self._tab = {}
        self._n = 0

    # construction ----------------------------------------------------------------------------
    def _mk(self, op, args=(), value=None):
        key = (op, tuple(a.id for a in args), value if op != "const" else struct.pack("<d", value))
        n = self._tab.get(key)
        if n is None:
            lvl = self.LEAF_LEVEL.get(op, 0) if not args else max(a.level for a in args)
            n = Expr(self, op, tuple(args), value, lvl, self._n)
            self._n += 1
            self._tab[key] = n
        return n

    def const(self, v): return self._mk("const", (), float(v))
    def param(self, i): return self._mk("param", (), int(i))
    def ctrl(self, i): return self._mk("ctrl", (), int(i))
    def x(self, i): return self._mk("x", (), int(i))
    def xe(self, i): return self._mk("xe", (), int(i))
    def qe(self, i): return self._mk("qe", (), int(i))

    def wrap(self, v):
        if hasattr(v, "_as_expr"):
            v = v._as_expr()
        if isinstance(v, Expr):
            if v.g is not self:
                raise ValueError("expression belongs to another graph")
            return v
        if isinstance(v, (int, float, np.floating, np.integer)):
            return self.const(float(v))
        raise TypeError(f"cannot use {type(v).__name__} in a model expression")

    # simplifying builders --------------------------------------------------------------------
    def add(self, a, b):
        a, b = self.wrap(a), self.wrap(b)
        if a.is_const() and b.is_const(): return self.const(a.value + b.value)
        if a.is_const(0.0): return b
        if b.is_const(0.0): return a
        if b.op == "neg": return self.sub(a, b.args[0])
        if a.op == "neg": return self.sub(b, a.args[0])
        if a.id > b.id and not a.is_const(): a, b = b, a          # canonical order (commutative)
        return self._mk("add", (a, b))

    def sub(self, a, b):
        a, b = self.wrap(a), self.wrap(b)
        if a.is_const() and b.is_const(): return self.const(a.value - b.value)
        if b.is_const(0.0): return a
        if a.is_const(0.0): return self.neg(b)
        if a is b: return self.const(0.0)
        if b.op == "neg": return self.add(a, b.args[0])
        return self._mk("sub", (a, b))

    @staticmethod
    def _factors(n, out):
        """Flatten a product tree into (constant, [non-constant factors])."""
        c = 1.0
        stack = [n]
        while stack:
            m = stack.pop()
            if m.op == "const": c *= m.value
            elif m.op == "mul": stack.extend(m.args)
            elif m.op == "neg":
                c = -c
                stack.append(m.args[0])
            elif m.op == "sqr" and m.args[0].op != "sqr":
                out.append(m.args[0]); out.append(m.args[0])
            else: out.append(m)
        return c

    def mul(self, a, b):
        a, b = self.wrap(a), self.wrap(b)
        if a.is_const() and b.is_const(): return self.const(a.value * b.value)
        if a.is_const(0.0) or b.is_const(0.0): return self.const(0.0)
        if a.is_const(1.0): return b
        if b.is_const(1.0): return a
        # distribute a strictly lower-level factor over a sum so that it can be hoisted and the
        # sum fuses into FMAs:   p*(u + v) -> p*u + p*v
        if b.op in ("add", "sub") and a.level < b.level:
            l, r = self.mul(a, b.args[0]), self.mul(a, b.args[1])
            return self.add(l, r) if b.op == "add" else self.sub(l, r)
        if a.op in ("add", "sub") and b.level < a.level:
            l, r = self.mul(a.args[0], b), self.mul(a.args[1], b)
            return self.add(l, r) if a.op == "add" else self.sub(l, r)
        # canonical product: constant first, then factors by ascending level (hoistable prefix)
        fs = []
        c = self._factors(a, fs) * self._factors(b, fs)
        if c == 0.0: return self.const(0.0)
        fs.sort(key=lambda n: (n.level, n.id))
        terms = []
        i = 0
        while i < len(fs):
            if i + 1 < len(fs) and fs[i] is fs[i + 1]:
                terms.append(self._mk("sqr", (fs[i],))); i += 2
            else:
                terms.append(fs[i]); i += 1
        terms.sort(key=lambda n: (n.level, n.id))
        negate = False
        if c == -1.0:
            negate, acc = True, None
        elif c == 1.0:
            acc = None
        else:
            acc = self.const(c)
        for t in terms:
            if acc is None:
                acc = t
            elif negate and acc.level < t.level:
                # fold the sign into the hoistable prefix
                acc = self._mk("mul", (self.neg(acc), t)); negate = False
            else:
                acc = self._mk("mul", (acc, t))
        if acc is None:
            return self.const(c)
        return self._mk("neg", (acc,)) if negate else acc

    def div(self, a, b):
        a, b = self.wrap(a), self.wrap(b)
        if b.is_const():
            if b.value == 0.0: raise ZeroDivisionError("division by constant zero in model expression")
            if a.is_const(): return self.const(a.value / b.value)
            if b.value == 1.0: return a
        if a.is_const(0.0): return self.const(0.0)
        if a.op == "neg": return self.neg(self.div(a.args[0], b))
        # dividing a high-level value by a lower-level one: multiply by a hoistable reciprocal
        if b.level < a.level and b.level <= LEVEL_CTRL and not b.is_const():
            return self.mul(self._mk("rcp", (b,)), a)
        if a.is_const(1.0): return self._mk("rcp", (b,))
        return self._mk("div", (a, b))

    def neg(self, a):
        a = self.wrap(a)
        if a.is_const(): return self.const(-a.value)
        if a.op == "neg": return a.args[0]
        if a.op == "sub": return self.sub(a.args[1], a.args[0])
        return self._mk("neg", (a,))

    def pow(self, a, b):
        a, b = self.wrap(a), self.wrap(b)
        if b.is_const():
            e = b.value
            if a.is_const(): return self.const(a.value ** e)
            if e == 0.0: return self.const(1.0)
            if e == 1.0: return a
            if e == 2.0: return self.mul(a, a)
            if e == 3.0: return self.mul(self.mul(a, a), a)
            if e == 4.0:
                s = self.mul(a, a)
                return self.mul(s, s)
            if e == -1.0: return self.div(1.0, a)
            if e == -2.0: return self.div(1.0, self.mul(a, a))
            if e == 0.5: return self.sqrt(a)
        return self._mk("pow", (a, b))

    def _unary(self, name, fn, a):
        a = self.wrap(a)
        if a.is_const(): return self.const(fn(a.value))
        return self._mk(name, (a,))

    def exp(self, a): return self._unary("exp", math.exp, a)
    def log(self, a): return self._unary("log", math.log, a)
    def sqrt(self, a): return self._unary("sqrt", math.sqrt, a)
    def sin(self, a): return self._unary("sin", math.sin, a)
    def cos(self, a): return self._unary("cos", math.cos, a)
    def tanh(self, a): return self._unary("tanh", math.tanh, a)
    def abs(self, a): return self._unary("abs", abs, a)

    def minimum(self, a, b):
        a, b = self.wrap(a), self.wrap(b)
        if a.is_const() and b.is_const(): return self.const(min(a.value, b.value))
        return self._mk("min", (a, b))

    def maximum(self, a, b):
        a, b = self.wrap(a), self.wrap(b)
        if a.is_const() and b.is_const(): return self.const(max(a.value, b.value))
        return self._mk("max", (a, b))

    # differentiation -------------------------------------------------------------------------
    def diff(self, e, wrt, _memo=None):
        """Symbolic d e / d wrt (``wrt`` is a leaf).  Forward accumulation over the DAG."""
        memo = {} if _memo is None else _memo
        order = topo_order([e])
        for n in order:
            if n.id in memo:
                continue
            memo[n.id] = self._dnode(n, wrt, memo)
        return memo[e.id]

    def _dnode(self, n, wrt, memo):
        if n is wrt: return self.const(1.0)
        if not n.args: return self.const(0.0)
        a = n.args[0]
        da = memo[a.id]
        op = n.op
        if len(n.args) == 2:
            b = n.args[1]
            db = memo[b.id]
        if op == "add": return self.add(da, db)
        if op == "sub": return self.sub(da, db)
        if op == "mul": return self.add(self.mul(da, b), self.mul(a, db))
        if op == "div": return self.div(self.sub(da, self.mul(n, db)), b)
        if op == "neg": return self.neg(da)
        if op == "sqr": return self.mul(self.mul(2.0, a), da)
        if op == "rcp": return self.neg(self.mul(self.mul(n, n), da))
        if op == "exp": return self.mul(n, da)
        if op == "log": return self.div(da, a)
        if op == "sqrt": return self.div(da, self.mul(2.0, n))
        if op == "sin": return self.mul(self.cos(a), da)
        if op == "cos": return self.neg(self.mul(self.sin(a), da))
        if op == "tanh": return self.mul(self.sub(1.0, self.mul(n, n)), da)
        if op == "pow":
            # d(a^b) = a^b * (db*log a + b*da/a); constant exponent avoids the log
            if b.is_const():
                return self.mul(self.mul(b, self.pow(a, self.const(b.value - 1.0))), da)
            return self.mul(n, self.add(self.mul(db, self.log(a)), self.div(self.mul(b, da), a)))
        if op in ("abs", "min", "max"):
            # almost-everywhere derivatives with the subgradient 0 at the kink:
            #   sign(v) = v / max(|v|, tiny);  min/max(a,b) = (a + b -/+ |a - b|)/2
            if da.is_const(0.0) and (len(n.args) == 1 or db.is_const(0.0)):
                return self.const(0.0)
            sign = lambda v: self.div(v, self.maximum(self.abs(v), 1e-300))
            if op == "abs":
                return self.mul(sign(a), da)
            s = sign(self.sub(a, b))
            half_sum = self.mul(0.5, self.add(da, db))
            half_dif = self.mul(0.5, self.mul(s, self.sub(da, db)))
            return self.sub(half_sum, half_dif) if op == "min" else self.add(half_sum, half_dif)
        raise NotImplementedError(op)


    # polynomial structure ---------------------------------------------------------------------------
    def poly_coeffs(self, e, x, maxdeg=2):
        """Coefficients ``[c0, c1, ...]`` (expressions free of the leaf ``x``) with ``e == sum_k c_k x**k``, or
        ``None`` when ``e`` is not a polynomial of degree <= ``maxdeg`` in ``x``.  Mass-action kinetics are: the
        kernel then solves a one-state block with its Newton loop entirely in registers (csrc/feas_kernel.cuh, POLY1)."""
        memo = {}
        zero = self.const(0.0)

        def trim(p):
            while len(p) > 1 and p[-1].is_const(0.0):
                p = p[:-1]
            return p

        def pmul(p, q):
            if len(p) + len(q) - 2 > maxdeg:
                return None
            out = [zero] * (len(p) + len(q) - 1)
            for i, a in enumerate(p):
                for j, b in enumerate(q):
                    out[i + j] = self.add(out[i + j], self.mul(a, b))
            return trim(out)

        for n in topo_order([e]):
            if n is x:
                memo[n.id] = [zero, self.const(1.0)]
                continue
            if not n.args:
                memo[n.id] = [n]
                continue
            ps = [memo[a.id] for a in n.args]
            if any(p is None for p in ps):
                memo[n.id] = None
            elif all(len(p) == 1 for p in ps):
                memo[n.id] = [n]                              # free of x: stays one node
            elif n.op in ("add", "sub"):
                p, q = ps
                m = max(len(p), len(q))
                p = p + [zero] * (m - len(p)); q = q + [zero] * (m - len(q))
                memo[n.id] = trim([self.add(a, b) if n.op == "add" else self.sub(a, b) for a, b in zip(p, q)])
            elif n.op == "neg":
                memo[n.id] = [self.neg(a) for a in ps[0]]
            elif n.op == "mul":
                memo[n.id] = pmul(ps[0], ps[1])
            elif n.op == "sqr":
                memo[n.id] = pmul(ps[0], ps[0])
            elif n.op == "div" and len(ps[1]) == 1:
                memo[n.id] = [self.div(a, ps[1][0]) for a in ps[0]]
            else:
                memo[n.id] = None
        return memo[e.id]


def topo_order(roots):
    """Post-order (dependencies first) list of the nodes reachable from ``roots``."""
    seen, out = set(), []
    for r in roots:
        if r.id in seen:
            continue
        stack = [(r, 0)]
        while stack:
            n, i = stack.pop()
            if i == 0:
                if n.id in seen:
                    continue
                seen.add(n.id)
            if i < len(n.args):
                stack.append((n, i + 1))
                c = n.args[i]
                if c.id not in seen:
                    stack.append((c, 0))
            else:
                out.append(n)
    return out


# ---------------------------------------------------------------------------------------------
# lowering
# ---------------------------------------------------------------------------------------------
@dataclass
class TapeLayout:
    """Slot assignment shared between the lowering and the kernel (units of one double per thread)."""
    n_param: int
    n_ctrl_elem: int
    ns: int
    nq: int
    param_slots: list = field(default_factory=list)   # slot of every parameter (d first, then theta)
    ctrl_slot: int = 0
    xe_slot: int = 0
    qe_slot: int = 0
    x_slot: int = 0            # wide: slot x_slot + 3*i + point
    j_mask: int = 0            # bit r*ns+c set when dJ[r][c] is not structurally zero
    fq_state_dep: int = 0      # 1 when some quadrature integrand depends on the dynamic states
    n_units: int = 0           # total narrow-equivalent slots


@dataclass
class LoweredTapes:
    layout: TapeLayout
    consts: np.ndarray
    code: dict                 # name -> np.ndarray(uint64)
    refs: dict                 # name -> np.ndarray(uint16)
    stats: dict = field(default_factory=dict)


_BINOPS = {"add": "ADD", "sub": "SUB", "mul": "MUL", "div": "DIV", "pow": "POW", "min": "MIN", "max": "MAX"}
_UNOPS = {"neg": "NEG", "sqr": "SQR", "rcp": "RCP", "exp": "EXP", "log": "LOG", "sqrt": "SQRT", "sin": "SIN",
          "cos": "COS", "tanh": "TANH", "abs": "ABS"}
_FMA_OPS = ("FMA", "FMS", "FNMA")


def encode(op, dst, a=0, b=0, c=0):
    return (OPCODE[op] & 0xFF) | ((dst & 0x3FFF) << 8) | ((a & 0x3FFF) << 22) | ((b & 0x3FFF) << 36) | ((c & 0x3FFF) << 50)


def decode(word):
    word = int(word)
    return (OPS[word & 0xFF], (word >> 8) & 0x3FFF, (word >> 22) & 0x3FFF, (word >> 36) & 0x3FFF, (word >> 50) & 0x3FFF)


class _SlotPool:
    """Linear-scan slot allocator over a list of free units (optionally growing at the end)."""

    def __init__(self, width, free=None, grow_from=None):
        self.width = width
        self.free = list(free or [])
        self.next = grow_from
        self.high = grow_from if grow_from is not None else 0

    def take(self):
        if self.free:
            return self.free.pop(0)
        if self.next is None:
            raise RuntimeError("slot pool exhausted")
        s = self.next
        self.next += self.width
        self.high = self.next
        return s

    def give(self, s):
        self.free.insert(0, s)


def lower(graph: Graph, *, n_param, n_ctrl_elem, ns, nq, roots: dict, pt_groups=None, ncp: int = 3) -> LoweredTapes:
    """Lower root expressions to tapes.

    ``roots`` maps a table name to a list of :class:`Expr` (``None`` = absent).  Tables read by the kernel
    through reference tables: ``x0``, ``q0``, ``z`` (<= PARAM level), ``g``, ``gx``, ``gq`` (<= END level) and the
    point-level tables named in ``pt_groups``.

    ``pt_groups`` is an ordered list ``[(tape name, [table names]), ...]``: every group becomes one point tape
    holding the STATE-level nodes its tables need.  The kernel runs the groups one after the other inside a
    finite element (one group per block of the block-triangular state solve, then the quadrature integrands),
    so a STATE-level node needed by two groups is *recomputed* in each -- the states it depends on have moved
    in between -- and all groups share the same temporaries.  Default: one tape ``"pt"`` with ``f``, ``J``, ``fq``.

    Register-file layout (one unit = one double per thread)::

        [ persistent narrow: live parameters | ctrl | xe | qe | constants used by point tapes |
          pro / elem values consumed by later tapes ]
        [ wide: X | point-tape outputs and temporaries (shared by all point tapes) ]     (3 units per value)

    Short-lived values -- parameters read only by the prologue, prologue temporaries, g-tape temporaries --
    are overlaid on wide units that are dead while they live.
    """
    if ncp != 3:
        raise NotImplementedError("the kernel evaluates the point tapes 3-wide (ncp = 3)")
    if pt_groups is None:
        pt_groups = [("pt", [t for t in ("f", "J", "fq") if t in roots])]
    slot_tables = set(t for _, tabs in pt_groups for t in tabs) | {"gx", "gq", "gee"}      # read through per-thread slots
    pt_names = [name for name, _ in pt_groups]
    all_roots = [e for lst in roots.values() for e in lst if e is not None]
    order = topo_order(all_roots)

    users = {}
    for n in order:
        for a in n.args:
            users.setdefault(a.id, []).append(n)
    root_ids = {e.id for e in all_roots}

    # ---- constant pool ---------------------------------------------------------------------------
    const_index, consts = {}, []

    def cref(v):
        key = struct.pack("<d", float(v))
        i = const_index.get(key)
        if i is None:
            i = len(consts)
            if i >= REF_MASK:
                raise ValueError("constant pool overflow")
            const_index[key] = i
            consts.append(float(v))
        return REF_CONST | i

    # ---- FMA fusion: add/sub with a single-use mul of the same level --------------------------------
    fused_into, fma_form = {}, {}
    for n in order:
        if n.op not in ("add", "sub") or n.level == LEVEL_CONST:
            continue
        l, r = n.args
        cand = None
        if l.op == "mul" and len(users.get(l.id, [])) == 1 and l.id not in root_ids and l.level == n.level \
                and l.id not in fused_into:
            cand = ("FMA" if n.op == "add" else "FMS", l, r)
        elif r.op == "mul" and len(users.get(r.id, [])) == 1 and r.id not in root_ids and r.level == n.level \
                and r.id not in fused_into:
            cand = ("FMA" if n.op == "add" else "FNMA", r, l)
        if cand is not None:
            opn, m, other = cand
            fused_into[m.id] = n.id
            fma_form[n.id] = (opn, m.args[0], m.args[1], other)

    def eff_args(n):
        return fma_form[n.id][1:] if n.id in fma_form else n.args

    def reachable(root_list, level):
        """Nodes of ``level`` needed by ``root_list`` (dependencies first), fused muls folded away."""
        seen, out = set(), []
        stack = [(e, 0) for e in reversed([e for e in root_list if e is not None])]
        while stack:
            n, i = stack.pop()
            if i == 0:
                if n.id in seen or not n.args or n.level != level:
                    continue
                seen.add(n.id)
            args = eff_args(n)
            if i < len(args):
                stack.append((n, i + 1))
                stack.append((args[i], 0))
            else:
                out.append(n)
        return out

    sched = {"pro": [], "elem": [], "g": []}
    level_tape = {LEVEL_PARAM: "pro", LEVEL_CTRL: "elem", LEVEL_END: "g"}
    for n in order:
        if n.level in level_tape and n.args and n.id not in fused_into:
            sched[level_tape[n.level]].append(n)
    for name, tabs in pt_groups:
        sched[name] = reachable([e for t in tabs for e in roots.get(t, [])], LEVEL_STATE)

    lay = TapeLayout(n_param=n_param, n_ctrl_elem=n_ctrl_elem, ns=ns, nq=nq)
    extra = {"pro": []}                          # constants -> per-thread slots, appended to the prologue
    const_roots = []
    for tname in slot_tables:
        for e in roots.get(tname, []):
            if e is not None and e.op == "const" and not e.is_const(0.0):
                const_roots.append(e)
    for e in roots.get("fq", []):
        if e is not None and e.level == LEVEL_STATE:
            lay.fq_state_dep = 1
    for i, e in enumerate(roots.get("J", [])):
        if e is not None and not e.is_const(0.0):
            lay.j_mask |= 1 << i

    # ---- who reads what, and where ------------------------------------------------------------------
    home = {}                                     # node id -> set of tapes that compute it
    read_in = {}                                  # node id -> set of tapes that read it
    for name, lst in sched.items():
        for n in lst:
            home.setdefault(n.id, set()).add(name)
            for a in eff_args(n):
                read_in.setdefault(a.id, set()).add(name)
    table_read = {e.id for lst in roots.values() for e in lst if e is not None}

    def persistent(n):
        """A narrow value that must outlive the tape producing it (parameters: the prologue)."""
        if n.id in table_read:
            return True
        h = home.get(n.id, {"pro"})               # parameter leaves behave like prologue values
        return bool(read_in.get(n.id, set()) - h)

    # ---- persistent narrow region -------------------------------------------------------------------
    cur = 0
    slot_of = {}                                  # narrow tapes: node id -> slot
    params = [graph.param(i) for i in range(n_param)]
    lay.param_slots = [None] * n_param
    for i, pn in enumerate(params):
        if persistent(pn):
            lay.param_slots[i] = cur; cur += 1
    lay.ctrl_slot = cur; cur += n_ctrl_elem
    lay.xe_slot = cur; cur += ns
    lay.qe_slot = cur; cur += nq
    const_slot = {}
    for a in [a for name in pt_names for n in sched[name] for a in eff_args(n)] + const_roots:
        if a.op == "const":
            key = struct.pack("<d", a.value)
            if key not in const_slot:
                const_slot[key] = cur; cur += 1
                extra["pro"].append(a)
    for name in ("pro", "elem"):
        for n in sched[name]:
            if persistent(n) or name == "elem":
                slot_of[n.id] = cur; cur += 1

    # ---- wide region ---------------------------------------------------------------------------------
    lay.x_slot = cur; cur += 3 * ns
    wide_fixed_end = cur

    def leaf_ref(n, in_pt=False):
        if n.op == "const":
            return const_slot[struct.pack("<d", n.value)] if in_pt else cref(n.value)
        if n.op == "param": return lay.param_slots[n.value]
        if n.op == "ctrl": return lay.ctrl_slot + n.value
        if n.op == "xe": return lay.xe_slot + n.value
        if n.op == "qe": return lay.qe_slot + n.value
        if n.op == "x": return REF_WIDE | (lay.x_slot + 3 * n.value)
        raise KeyError(n.op)

    def assign(lst, pool, wide, slots):
        """Linear-scan assignment of the not-yet-placed nodes of one tape into ``slots``."""
        last_use = {}
        for pos, n in enumerate(lst):
            for a in eff_args(n):
                last_use[a.id] = pos
        release_at = {}
        for pos, n in enumerate(lst):
            for s_ in release_at.pop(pos, []):
                pool.give(s_)
            if n.id in slots:
                continue
            s_ = pool.take()
            slots[n.id] = (REF_WIDE | s_) if wide else s_
            if n.id not in table_read:                     # values the kernel reads afterwards stay put
                release_at.setdefault(last_use.get(n.id, pos) + 1, []).append(s_)

    pt_slots = {}
    wide_end = cur
    for name in pt_names:                                   # every point tape starts from the same base
        pool = _SlotPool(3, grow_from=wide_fixed_end)
        pt_slots[name] = {}
        assign(sched[name], pool, True, pt_slots[name])
        wide_end = max(wide_end, pool.high)
    cur = wide_end
    pt_temp_units = list(range(wide_fixed_end, cur))

    x_units = list(range(lay.x_slot, lay.x_slot + 3 * ns))
    pro_pool = _SlotPool(1, free=x_units + pt_temp_units, grow_from=cur)
    for i, pn in enumerate(params):
        if lay.param_slots[i] is None:
            lay.param_slots[i] = pro_pool.take()
    # trailing constant moves read nothing but the pool; prologue temporaries may use the scratch
    assign(sched["pro"], pro_pool, False, slot_of)
    cur = max(cur, pro_pool.high)
    g_pool = _SlotPool(1, free=list(range(lay.x_slot, cur)), grow_from=cur)
    assign(sched["g"], g_pool, False, slot_of)
    cur = max(cur, g_pool.high)
    if cur - 1 > REF_MASK:
        raise ValueError("register file overflow (4096 slots)")
    lay.n_units = cur

    # ---- emit ------------------------------------------------------------------------------------------
    def ref(n, tape=None):
        if not n.args:
            return leaf_ref(n, tape in pt_slots)
        if tape in pt_slots and n.level == LEVEL_STATE:
            return pt_slots[tape][n.id]
        return slot_of[n.id]

    code = {}
    for name in ["pro", "elem", "g"] + pt_names:
        words = []
        for n in sched[name]:
            dst = ref(n, name)
            if n.id in fma_form:
                opn, a, b, c = fma_form[n.id]
                words.append(encode(opn, dst, ref(a, name), ref(b, name), ref(c, name)))
            elif n.op in _BINOPS:
                words.append(encode(_BINOPS[n.op], dst, ref(n.args[0], name), ref(n.args[1], name)))
            elif n.op in _UNOPS:
                words.append(encode(_UNOPS[n.op], dst, ref(n.args[0], name)))
            else:
                raise NotImplementedError(n.op)
        for src in extra.get(name, []):                              # constants -> per-thread slots
            words.append(encode("MOV", const_slot[struct.pack("<d", src.value)], cref(src.value)))
        code[name] = np.asarray(words, dtype=np.uint64)

    table_tape = {t: name for name, tabs in pt_groups for t in tabs}
    ref_tables = {}
    for name, lst in roots.items():
        out = []
        for e in lst:
            if e is None or e.is_const(0.0):
                out.append(REF_NONE)
            elif e.op == "const" and name in slot_tables:
                out.append(const_slot[struct.pack("<d", e.value)])
            else:
                out.append(ref(e, table_tape.get(name)))
        ref_tables[name] = np.asarray(out, dtype=np.uint16)

    stats = {k: int(len(v)) for k, v in code.items()}
    stats["flops"] = {k: int(sum(2 if decode(w)[0] in _FMA_OPS else (0 if decode(w)[0] in ("MOV", "BC3") else 1)
                                 for w in v)) for k, v in code.items()}
    return LoweredTapes(layout=lay, consts=np.asarray(consts, dtype=np.float64), code=code, refs=ref_tables, stats=stats)
