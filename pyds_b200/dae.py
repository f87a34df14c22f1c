"""The slice of ``pyomo.dae`` the reference's stage rules use (``ContinuousSet``, ``DerivativeVar``);
see ``pyds_b200.environ``.  Discretisation itself happens inside the CUDA kernel (Radau collocation,
reference ``run_twostage.py:76-79``), so these classes only record the structure."""
from __future__ import annotations

from .environ import Var, _Component


class ContinuousSet(_Component):
    def __init__(self, bounds=(0, 1), **kw):
        self.bounds = tuple(bounds)

    def first(self):
        return self.bounds[0]

    def last(self):
        return self.bounds[1]

    def __iter__(self):
        return iter(self.bounds)


class DerivativeVar(Var):
    def __init__(self, sVar, wrt=None, initialize=0, **kw):
        super().__init__(*sVar.index_sets, initialize=initialize)
        if sVar.contset is None:
            raise ValueError("DerivativeVar needs a Var indexed by a ContinuousSet")
        self.derivative_of = sVar
        sVar.derivative = self


class DAE_Error(Exception):
    pass
