"""A small Pyomo-shaped modelling layer: just enough of ``pyomo.environ`` for the reference's stage
rules (reference ``run_twostage.py:12-79``, ``run_threestage.py:12-77``) to run UNCHANGED apart from
their import lines::

    import pyds_b200.environ as pyo
    from pyds_b200 import dae

Pyomo itself is not installed in this environment, and the CUDA path does not want a block tree
anyway: components only *record* the model symbolically (expressions are nodes of
``pyds_b200.tape.Graph``); ``pyds_b200.simulator`` then lowers the record to a ``DaeModel`` and to
the tape blob.  Supported: ``ConcreteModel``/``Block``, ``Param`` (scalar mutable, or indexed
constants), ``Var`` (scalar, or indexed by ordinary sets and at most one ``dae.ContinuousSet``),
``Set``/``RangeSet``, ``Constraint`` (scalar, or indexed by the continuous set), ``ConstraintList``,
``Suffix``, ``Objective``, ``exp/log/sqrt/sin/cos/tanh``, ``TransformationFactory`` for
``dae.collocation`` (``apply_to`` + ``reduce_collocation_points``) and ``core.relax_integer_vars``.
"""
from __future__ import annotations

from . import tape as T

__all__ = ["ConcreteModel", "Block", "Param", "Var", "Set", "RangeSet", "Constraint", "ConstraintList", "Suffix",
           "Objective", "maximize", "minimize", "Binary", "Reals", "NonNegativeReals", "exp", "log", "sqrt", "sin", "cos",
           "tanh", "TransformationFactory", "value"]

Binary, Reals, NonNegativeReals = "Binary", "Reals", "NonNegativeReals"
maximize, minimize = -1, 1


class _TimeSymbol:
    """The 'for all t' index handed to a rule of a Constraint indexed by a ContinuousSet."""

    def __repr__(self):
        return "t"


TIME = _TimeSymbol()


class _Component:
    name = None
    model = None

    def _attach(self, model, name):
        self.model, self.name = model, name

    @property
    def local_name(self):
        return self.name


class ConcreteModel:
    """Attribute registry + owner of the expression graph.  Also plays Pyomo's ``Block``."""

    def __init__(self):
        object.__setattr__(self, "_components", {})
        object.__setattr__(self, "_graph", T.Graph())
        object.__setattr__(self, "_tokens", [])          # token -> (kind, component, key, time class)
        object.__setattr__(self, "_token_index", {})
        object.__setattr__(self, "_bigm", [])            # (name, Relational, M)
        object.__setattr__(self, "_fixed", {})           # (var name, key incl. time) -> value
        object.__setattr__(self, "_collocation", None)
        object.__setattr__(self, "_reduced", {})         # var name -> ncp

    def __setattr__(self, name, value):
        if isinstance(value, _Component):
            if value.model is not None and value.model is not self:
                raise ValueError(f"component '{name}' already belongs to another model")
            value._attach(self, name)
            self._components[name] = value
        object.__setattr__(self, name, value)

    def add_component(self, name, comp):
        if comp is not None:
            setattr(self, name, comp)

    def component(self, name):
        return self._components.get(name)

    def component_objects(self, ctype=None):
        return [c for c in self._components.values() if ctype is None or isinstance(c, ctype)]

    def _token(self, kind, comp, key, tclass):
        k = (kind, id(comp), key, tclass)                # components overload ==, so index by identity
        i = self._token_index.get(k)
        if i is None:
            i = len(self._tokens)
            self._tokens.append((kind, comp, key, tclass))
            self._token_index[k] = i
        return i


Block = ConcreteModel


class Set(_Component):
    def __init__(self, initialize=None, **kw):
        self.members = list(initialize or [])

    def __iter__(self):
        return iter(self.members)

    def __len__(self):
        return len(self.members)


class RangeSet(Set):
    def __init__(self, a, b=None):
        lo, hi = (1, a) if b is None else (a, b)
        super().__init__(initialize=range(lo, hi + 1))


class _Operand:
    """Mixin: arithmetic and relations through the symbolic expression of the object."""

    def _as_expr(self):
        raise NotImplementedError

    def __add__(self, o): return self._as_expr() + o
    def __radd__(self, o): return o + self._as_expr()
    def __sub__(self, o): return self._as_expr() - o
    def __rsub__(self, o): return o - self._as_expr()
    def __mul__(self, o): return self._as_expr() * o
    def __rmul__(self, o): return o * self._as_expr()
    def __truediv__(self, o): return self._as_expr() / o
    def __rtruediv__(self, o): return o / self._as_expr()
    def __neg__(self): return -self._as_expr()
    def __pow__(self, o): return self._as_expr() ** o
    def __rpow__(self, o): return o ** self._as_expr()
    def __eq__(self, o): return T.Relational(self._as_expr(), o, "==")
    def __le__(self, o): return T.Relational(self._as_expr(), o, "<=")
    def __ge__(self, o): return T.Relational(self._as_expr(), o, ">=")
    __hash__ = object.__hash__


class Param(_Component, _Operand):
    def __init__(self, *index, initialize=None, mutable=False, **kw):
        self.index = index
        self.mutable = mutable
        self.initialize = initialize
        self._value = initialize if not index else None

    # scalar ---------------------------------------------------------------------------------
    @property
    def value(self):
        return self._value

    def set_value(self, v):
        self._value = v

    def is_indexed(self):
        return bool(self.index)

    def _as_expr(self):
        if self.index:
            raise TypeError(f"indexed Param '{self.name}' needs an index")
        if self.model is None:
            raise ValueError("Param used before it was attached to a model")
        return self.model._graph._mk("pn", (), self.model._token("param", self, None, None))

    # indexed: plain constants ------------------------------------------------------------------
    def __getitem__(self, key):
        init = self.initialize
        return float(init[key]) if isinstance(init, dict) else float(init)

    def items(self):
        return self.initialize.items()


class VarRef(_Operand):
    """One element of a Var: ``m.c['A', t]``, ``m.c['B', 1]``, ``m.u[0]`` or a scalar Var itself."""

    def __init__(self, var, key, tclass):
        self.var, self.key, self.tclass = var, key, tclass

    def _as_expr(self):
        m = self.var.model
        kind = "alg" if self.var.contset is None else "v"
        return m._graph._mk(kind, (), m._token("var", self.var, self.key, self.tclass))

    def fix(self, value):
        self.var.model._fixed[(self.var.name, self.key, self.tclass)] = value

    def set_value(self, value):
        self.var.values[(self.key, self.tclass)] = value

    @property
    def value(self):
        return self.var.values.get((self.key, self.tclass), self.var.initialize)


class Var(_Component, _Operand):
    def __init__(self, *index_sets, initialize=0, bounds=None, within=None, **kw):
        from .dae import ContinuousSet
        self.index_sets = index_sets
        self.contset = None
        for s in index_sets:
            if isinstance(s, ContinuousSet):
                if self.contset is not None:
                    raise NotImplementedError("a Var may be indexed by one ContinuousSet")
                self.contset = s
        self.initialize = initialize
        self.bounds = bounds
        self.within = within
        self.values = {}
        self.fixed = False
        self.derivative_of = None      # set for DerivativeVar

    def is_indexed(self):
        return bool(self.index_sets)

    def _split(self, key):
        key = key if isinstance(key, tuple) else (key,)
        if len(key) != len(self.index_sets):
            raise KeyError(f"{self.name}{list(key)}: expected {len(self.index_sets)} indices")
        other, tclass = [], None
        for k, s in zip(key, self.index_sets):
            if s is self.contset:
                if k is TIME:
                    tclass = "t"
                else:
                    lo, hi = s.bounds
                    if k == lo: tclass = "first"
                    elif k == hi: tclass = "last"
                    else:
                        raise NotImplementedError(f"{self.name}: only t, {lo} and {hi} can be referenced symbolically")
            else:
                other.append(k)
        return tuple(other), tclass

    def __getitem__(self, key):
        other, tclass = self._split(key)
        return VarRef(self, other, tclass)

    def _as_expr(self):
        if self.index_sets:
            raise TypeError(f"indexed Var '{self.name}' needs an index")
        return VarRef(self, (), None)._as_expr()

    # scalar conveniences
    def fix(self, value):
        VarRef(self, (), None).fix(value)

    def unfix(self):
        self.fixed = False

    def set_value(self, v):
        self.values[((), None)] = v

    @property
    def value(self):
        return self.values.get(((), None), self.initialize)

    def keys(self):
        from itertools import product
        sets = [list(s) if s is not self.contset else ["t"] for s in self.index_sets]
        return [k for k in product(*sets)]


class Constraint(_Component):
    def __init__(self, *index_sets, rule=None, expr=None, **kw):
        self.index_sets, self.rule, self.expr = index_sets, rule, expr
        self._relations = None

    def construct(self):
        pass

    @property
    def relations(self):
        """Rules are evaluated on first use (at lowering), when every stage has declared its components:
        a stage-0 equation may then refer to a later stage's Param, as the flattened three-stage model does."""
        if self._relations is None:
            self._build()
        return self._relations

    def _build(self):
        from .dae import ContinuousSet
        if self.expr is not None:
            rel = [((), self.expr)]
        elif not self.index_sets:
            rel = [((), self.rule(self.model))]
        elif len(self.index_sets) == 1 and isinstance(self.index_sets[0], ContinuousSet):
            rel = [((TIME,), self.rule(self.model, TIME))]
        else:
            raise NotImplementedError("Constraint index must be empty or one ContinuousSet")
        for _, r in rel:
            if not isinstance(r, T.Relational):
                raise TypeError(f"constraint '{self.name}' rule must return a relation (==, <=, >=)")
        self._relations = rel

    # the three views add_BigMConstraint reads (reference pyds/utils.py:27-40)
    def _single(self):
        return self.relations[0][1]

    @property
    def body(self):
        return self._single().lhs

    @property
    def upper(self):
        r = self._single()
        return r.rhs if r.op in ("<=", "==") else None

    @property
    def lower(self):
        r = self._single()
        return r.rhs if r.op in (">=", "==") else None


class ConstraintList(_Component):
    def __init__(self):
        self.items = []

    def construct(self):
        pass

    def add(self, expr):
        self.items.append(expr)


class Objective(_Component):
    def __init__(self, rule=None, expr=None, sense=minimize):
        self.rule, self.expr, self.sense = rule, expr, sense


class Suffix(_Component, dict):
    LOCAL, EXPORT, IMPORT = 0, 1, 2

    def __init__(self, direction=0, **kw):
        dict.__init__(self)
        self.direction = direction

    def __setitem__(self, comp, profile):            # keyed by component identity
        dict.__setitem__(self, id(comp), (comp, profile))

    def profile_of(self, comp):
        rec = dict.get(self, id(comp))
        return None if rec is None else rec[1]

    __hash__ = object.__hash__


def _fn(name):
    def f(x):
        if hasattr(x, "_as_expr"):
            x = x._as_expr()
        if isinstance(x, T.Expr):
            return getattr(x.g, name)(x)
        import math
        return getattr(math, name)(x)
    f.__name__ = name
    return f


exp, log, sqrt, sin, cos, tanh = (_fn(n) for n in ("exp", "log", "sqrt", "sin", "cos", "tanh"))


def value(x):
    return x.value if hasattr(x, "value") else x


class _Collocation:
    def apply_to(self, m, nfe=10, ncp=3, scheme="LAGRANGE-RADAU", wrt=None, **kw):
        if scheme != "LAGRANGE-RADAU":
            raise NotImplementedError("only the LAGRANGE-RADAU scheme (Pyomo's default) is supported")
        m._collocation = {"nfe": int(nfe), "ncp": int(ncp)}

    def reduce_collocation_points(self, m, var=None, ncp=None, contset=None):
        if m._collocation is None:
            raise RuntimeError("apply_to must be called before reduce_collocation_points")
        m._reduced[var.name] = int(ncp)


class _Relax:
    def apply_to(self, m, **kw):
        object.__setattr__(m, "_relaxed_integer_vars", True)


def TransformationFactory(name):
    if name == "dae.collocation":
        return _Collocation()
    if name == "core.relax_integer_vars":
        return _Relax()
    raise NotImplementedError(f"transformation '{name}'")
