"""``Simulator``: lowers the user's stage-rule model ONCE into the fp64 tape blob and owns the engine.

In the reference this class is the warm-start initialiser: it integrates every scenario with
``pyomo.dae.Simulator`` (scipy) and copies the interpolated trajectories into the NLP's variables
(reference ``pyds/simulator.py:34-94``).  BASELINE.json re-purposes the name: here the constructor
(same signature ``Simulator(parent, config)``, reference :12) builds the flattened model from the
stage rules (reference :131-134), applies the user's model transformation, and lowers the result

    modelling record  ->  DaeModel  ->  tapes + sensitivity program  ->  blob  ->  libpyds_b200 handle

and ``simulate_all_scenarios(model, input_values)`` (same name, reference :34) evaluates the state
solve of every scenario at the warm-start control profile on the GPU -- the role the scipy
integration played, without the n_p Python-callback ODE solves per design point.
"""
from __future__ import annotations

import time

import numpy as np

from . import environ as pyo
from . import tape as T
from . import utils
from .model import DaeModel


class LoweringError(ValueError):
    pass


def lower_stage_model(model, input_map, f_profile_suffix=None, **model_kw):
    """Translate the symbolic record written by the stage rules into a :class:`DaeModel`.

    ``input_map`` is the reference's ``{stage: [param names]}`` (reference run_twostage.py:85): the last
    stage's names are the uncertain parameters theta, all earlier stages' names the design variables d.
    """
    stages = sorted(input_map.keys())
    p_names = list(input_map[stages[-1]])
    d_names = [n for s in stages[:-1] for n in input_map[s]]
    if model._collocation is None:
        raise LoweringError("the model transformation must apply 'dae.collocation'")
    nfe, ncp = model._collocation["nfe"], model._collocation["ncp"]
    dm = DaeModel(d_names, p_names, nfe=nfe, ncp=ncp, **model_kw)
    dm.state_keys = {}                 # state name -> (Var name, index key, (t0, t1)): for the output records
    g_src, g = model._graph, dm.graph

    tok = model._tokens
    # --- classify variables --------------------------------------------------------------------------
    time_vars = [c for c in model.component_objects(pyo.Var) if c.contset is not None and c.derivative_of is None]
    deriv_vars = [c for c in model.component_objects(pyo.Var) if c.derivative_of is not None]
    state_vars = [dv.derivative_of for dv in deriv_vars]
    control_vars = [v for v in time_vars if not any(v is s for s in state_vars)]

    def keys_of(var):
        from itertools import product
        sets = [list(s) for s in var.index_sets if s is not var.contset]
        return [tuple(k) for k in product(*sets)] if sets else [()]

    # states
    state_leaf = {}
    for sv in state_vars:
        for key in keys_of(sv):
            fx = model._fixed.get((sv.name, key, "first"))
            if fx is None:
                raise LoweringError(f"state {sv.name}{list(key)} has no fixed initial value (use .fix at t = {sv.contset.first()})")
            state_leaf[(id(sv), key)] = (sv, key, fx)
    # controls: piecewise constant per finite element
    ctrl_leaf = {}
    for cv in control_vars:
        if model._reduced.get(cv.name) != 1:
            raise LoweringError(f"control '{cv.name}' must be made piecewise constant with reduce_collocation_points(ncp=1)")
        for key in keys_of(cv):
            lo, hi = cv.bounds if cv.bounds is not None else (-np.inf, np.inf)
            init = cv.initialize
            if f_profile_suffix is not None:
                prof = f_profile_suffix.profile_of(cv)
                if prof is not None:
                    t0, t1 = cv.contset.bounds
                    pts = sorted(prof.items())
                    init = []
                    for e in range(nfe):
                        tm = t0 + (t1 - t0) * (e + 0.5) / nfe             # element mid-point
                        val = pts[0][1]
                        for tb, vb in pts:
                            if tb <= tm:
                                val = vb
                        init.append(float(val))
            ctrl_leaf[(id(cv), key)] = dm.control(f"{cv.name}{list(key) if key else ''}", lo if lo is not None else -np.inf,
                                                  hi if hi is not None else np.inf, init=init)

    # --- leaf translation ---------------------------------------------------------------------------
    param_expr = {}

    def param_leaf(comp):
        if comp.name in dm.d:
            return dm.d[comp.name]
        if comp.name in dm.p:
            return dm.p[comp.name]
        if comp.value is None:
            raise LoweringError(f"Param '{comp.name}' is neither bound by the input map nor initialised")
        return g.const(float(comp.value))

    state_handle = {}

    def translate(e, alg_defs, allow_deriv=None):
        """Rebuild ``e`` (graph of the modelling layer) inside the DaeModel graph."""
        memo = {}
        for n in T.topo_order([e]):
            if n.op == "const":
                r = g.const(n.value)
            elif n.op == "pn":
                r = param_leaf(tok[n.value][1])
            elif n.op == "v":
                _, comp, key, tclass = tok[n.value]
                if comp.derivative_of is not None:
                    if allow_deriv is None or tclass != "t":
                        raise LoweringError(f"derivative '{comp.name}' may only appear in its own differential equation")
                    r = allow_deriv(comp, key)
                elif (id(comp), key) in state_handle:
                    h = state_handle[(id(comp), key)]
                    r = {"t": h, "last": dm.end(h), "first": dm.initial(h)}[tclass]
                elif (id(comp), key) in ctrl_leaf:
                    if tclass != "t":
                        raise LoweringError(f"control '{comp.name}' can only be referenced at the running time t")
                    r = ctrl_leaf[(id(comp), key)]
                else:
                    raise LoweringError(f"time-indexed Var '{comp.name}' is neither a state nor a control")
            elif n.op == "alg":
                _, comp, key, _ = tok[n.value]
                if id(comp) not in alg_defs:
                    raise LoweringError(f"algebraic Var '{comp.name}' has no defining equation 'var == expression'")
                r = alg_defs[id(comp)]
            else:
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

    for (vid, key), (sv, _, fx) in state_leaf.items():
        x0 = fx
        if hasattr(fx, "_as_expr") or isinstance(fx, T.Expr):
            x0 = translate(g_src.wrap(fx), {})
        state_handle[(vid, key)] = dm.state(f"{sv.name}{list(key) if key else ''}", x0, bounds=sv.bounds)
        dm.state_keys[f"{sv.name}{list(key) if key else ''}"] = (sv.name, key, sv.contset.bounds)

    # --- equations -------------------------------------------------------------------------------------
    alg_defs = {}
    scalar_cons = [c for c in model.component_objects(pyo.Constraint) if c.relations and c.relations[0][0] == ()]
    time_cons = [c for c in model.component_objects(pyo.Constraint) if c.relations and c.relations[0][0] != ()]
    pending = list(scalar_cons)
    progress = True
    while pending and progress:                      # algebraic definitions may reference one another
        progress = False
        for c in list(pending):
            rel = c.relations[0][1]
            if rel.op != "==":
                raise LoweringError(f"scalar constraint '{c.name}' must be an equality defining an algebraic Var")
            lhs, rhs = g_src.wrap(rel.lhs), g_src.wrap(rel.rhs)
            target, body = (lhs, rhs) if lhs.op == "alg" else (rhs, lhs) if rhs.op == "alg" else (None, None)
            if target is None:
                raise LoweringError(f"constraint '{c.name}' is not of the form 'var == expression'")
            try:
                alg_defs[id(tok[target.value][1])] = translate(body, alg_defs)
                pending.remove(c)
                progress = True
            except LoweringError:
                continue
    if pending:
        raise LoweringError("could not resolve algebraic definitions: " + ", ".join(c.name for c in pending))

    defined = set()
    for c in time_cons:
        rel = c.relations[0][1]
        if rel.op != "==":
            raise LoweringError(f"path constraint '{c.name}' must be a differential equation (==)")
        resid_src = g_src.sub(g_src.wrap(rel.lhs), g_src.wrap(rel.rhs))
        derivs = [n for n in T.topo_order([resid_src]) if n.op == "v" and tok[n.value][1].derivative_of is not None]
        if len(derivs) != 1:
            raise LoweringError(f"constraint '{c.name}' must contain exactly one time derivative")
        _, dcomp, dkey, _ = tok[derivs[0].value]
        marker = g._mk("dv", (), len(defined))

        resid = translate(resid_src, alg_defs, allow_deriv=lambda comp, key: marker)
        coef = g.diff(resid, marker)                                  # residual = coef * dx/dt + rest
        if any(n is marker for n in T.topo_order([coef])):
            raise LoweringError(f"'{c.name}' is not linear in the time derivative")
        memo = {}
        rest = dm._rebuild(resid, {("dv", marker.value): g.const(0.0)}, memo)
        sv = dcomp.derivative_of
        dm.ode(state_handle[(id(sv), dkey)], g.div(g.neg(rest), coef))
        defined.add((id(sv), dkey))
    for k in state_handle:
        if k not in defined:
            raise LoweringError("a differential state has no differential equation")

    if not model._bigm:
        raise LoweringError("the model has no big-M constraint (utils.add_BigMConstraint)")
    for name, expr, M in model._bigm:
        dm.constraint(name, translate(g_src.wrap(expr), alg_defs), M)
    return dm


def _default_device() -> int:
    """The GPU of this process: under torchrun every rank owns one device (LOCAL_RANK, or the device the caller made
    current), so that N managers do not all land on GPU 0; a single process uses device 0."""
    import os
    try:
        import torch
        if torch.cuda.is_available():
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                if "LOCAL_RANK" in os.environ:
                    return int(os.environ["LOCAL_RANK"]) % max(torch.cuda.device_count(), 1)
                return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


class Simulator:
    def __init__(self, parent, config):
        self.parent = parent
        self.stage_rules = parent.stage_rules
        self.model_transformation = parent.model_transformation
        self.enabled = config["enabled"]
        self.package = config["package"]                 # 'scipy' in the reference; informational here
        self.save_output = config["save output"]
        self.suffix_name = config["suffix name"]
        self.kwargs = config["kwargs"]
        self.model = None
        self.dae_model = None
        self.lowered = None
        self._engine = None
        self.user_time = None
        self.output = []
        self._traj = None
        self.device = int(config["device"]) if "device" in config else _default_device()
        self._build_model()
        self.input_names = []
        for stage, names in self.parent.input_map.items():
            self.input_names.extend(names)

    # -- lowering ------------------------------------------------------------------------------------
    def _build_model(self):
        self.model = utils.create_flattened_model(self.stage_rules)
        if self.model_transformation is not None:
            self.model_transformation(self.model)
        suffix = self.model.component(self.suffix_name) if self.suffix_name is not None else None
        self.dae_model = lower_stage_model(self.model, self.parent.input_map, suffix, **self.kwargs.get("lowering", {}))
        self.lowered = self.dae_model.lower()

    @property
    def engine(self):
        """The device handle (created on first use: needs a CUDA device, there is no CPU fallback)."""
        if self._engine is None:
            from . import capi
            self._engine = capi.Engine(self.lowered.blob, self.device)
        return self._engine

    def default_controls(self):
        """Warm-start control profile (the ``Suffix`` of reference run_twostage.py:25-26) as a (nz,) array."""
        z = np.zeros(self.lowered.nz)
        for k, e in enumerate(self.lowered.z_init):
            if self.lowered.z_fixed[k]:
                continue
            if e.op != "const":
                raise LoweringError("warm-start control values must be constants")
            z[k] = e.value
        return z

    # -- evaluation at the warm start --------------------------------------------------------------------
    def simulate_all_scenarios(self, output_model, input_values):
        """g of every scenario at the warm-start controls.  ``input_values`` as in the reference:
        ``{stage: 2-D array}`` with the design stage(s) before the theta stage."""
        t0 = time.time()
        self.output = []
        if not self.enabled:
            return None
        d, p = self.parent._split_inputs(input_values)
        g, ind, _ = self.engine.eval_host(d, p, None, z=self.default_controls() if self.lowered.nz and not self.lowered.z_fixed.all() else None,
                                          want_g=True, want_ind=True, want_P=False)
        self.last_g = g
        self._traj = None
        if self.save_output:
            # the simulated profiles of every scenario at the warm-start controls (reference simulator.py:49-51,112-119)
            z = self.default_controls() if self.lowered.nz and not self.lowered.z_fixed.all() else None
            self._traj = self.engine.state_solve(d, p, z=z)[0]
            self._tsim = self.engine.time_grid()
        self.user_time = time.time() - t0
        if self.save_output:
            self.output = self.output_for(range(d.shape[0]))
        return g

    def output_for(self, rows):
        """The per-scenario containers (reference simulator.py:112-119: user time, tsim, var_names, simsolution) of the
        design points ``rows``, theta scenarios in order; ``simsolution`` is (n_times, n_vars) at the collocation times."""
        if not self.save_output or self._traj is None:
            return []
        names = self.lowered.dyn_states + self.lowered.quad_states
        return [{"user time": self.user_time, "tsim": self._tsim.copy(), "var_names": list(names),
                 "simsolution": self._traj[i, j].copy(), "g": self.last_g[i, j].copy()}
                for i in rows for j in range(self._traj.shape[1])]

    def get_sim_result(self):
        return getattr(self, "last_g", None)
