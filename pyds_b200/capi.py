"""ctypes binding of libpyds_b200.so (the C ABI declared in ``include/pydsb.h``).

PyTorch tensors are used only as device buffers: the library receives raw ``data_ptr()`` values
and a CUDA stream handle.  There is no CPU fallback -- if the library is missing it must be built
(``python -m pyds_b200.build``), and on a machine without a CUDA device ``Engine`` raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

_LIB = None

Z_MODEL, Z_SHARED, Z_PER_DESIGN, Z_PER_SCENARIO = 0, 1, 2, 3
INFO_NAMES = ["ns", "nq", "ng", "d_dim", "p_dim", "nz", "nfe", "ncp", "n_units", "smem_bytes", "blocks_per_sm",
              "num_sms", "regs_per_thread", "scenarios_per_thread", "program_length", "chain_blocks", "recourse_chain_blocks"]
COUNTER_NAMES = ["scenarios", "newton_iters", "newton_warp_iters", "failed", "launches", "newton_flops"]

EXPORTS = ["pydsb_last_error", "pydsb_version", "pydsb_model_create", "pydsb_model_destroy", "pydsb_model_info",
           "pydsb_eval", "pydsb_eval_tol", "pydsb_eval_g", "pydsb_efp", "pydsb_eval_host", "pydsb_state_solve", "pydsb_recourse_solve",
           "pydsb_recourse_solve_global_dist", "pydsb_get_counters", "pydsb_dfma_peak", "pydsb_assemble",
           "pydsb_ns_replace_worst", "pydsb_ns_ellipsoid", "pydsb_set_profiling", "pydsb_get_profile"]
PROFILE_NAMES = ["feas_ms", "feas_launches", "sens_ms", "sens_launches", "lm_ms", "lm_launches"]
# int (*pydsb_allreduce_fn)(double* device_buf, int n, void* user)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p)


class PydsbError(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB


def load():
    """Load (building if the sources are newer) the shared library and declare its prototypes."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("PYDSB_LIB")          # development: a variant built by tools/build_variant.py
    if not path:
        path = _build.LIB
        if _build.is_stale():
            _build.build()
    L = ctypes.CDLL(path)
    vp, ll, dp, ci = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int
    L.pydsb_last_error.restype = ctypes.c_char_p
    L.pydsb_last_error.argtypes = []
    L.pydsb_version.restype = ci
    L.pydsb_model_create.restype = ci
    L.pydsb_model_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ci, ctypes.POINTER(vp)]
    L.pydsb_assemble.restype = ci
    L.pydsb_assemble.argtypes = [ctypes.c_char_p, ctypes.c_size_t, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t),
                                 ctypes.POINTER(ll), ci]
    L.pydsb_ns_ellipsoid.restype = ci
    L.pydsb_ns_ellipsoid.argtypes = [vp, ll, ci, vp, vp, vp, vp, vp]
    L.pydsb_ns_replace_worst.restype = ci
    L.pydsb_ns_replace_worst.argtypes = [vp, ll, vp, ll, ctypes.c_double, vp]
    L.pydsb_model_destroy.restype = ci
    L.pydsb_model_destroy.argtypes = [vp]
    L.pydsb_model_info.restype = ci
    L.pydsb_model_info.argtypes = [vp, ctypes.POINTER(ll), ci]
    L.pydsb_eval.restype = ci
    L.pydsb_eval.argtypes = [vp, dp, ll, dp, dp, ll, dp, ci, dp, dp, dp, vp]
    L.pydsb_eval_g.restype = ci
    L.pydsb_eval_g.argtypes = [vp, dp, ll, dp, ll, dp, vp]
    L.pydsb_efp.restype = ci
    L.pydsb_efp.argtypes = [vp, dp, ll, dp, dp, ll, dp, vp]
    L.pydsb_eval_host.restype = ci
    L.pydsb_eval_host.argtypes = [vp, dp, ll, dp, dp, ll, dp, ci, dp, dp, dp]
    L.pydsb_eval_tol.restype = ci
    L.pydsb_eval_tol.argtypes = [vp, dp, ll, dp, dp, ll, dp, ci, ctypes.c_double, dp, dp, dp, vp]
    L.pydsb_state_solve.restype = ci
    L.pydsb_state_solve.argtypes = [vp, dp, ll, dp, ll, dp, ci, dp, dp, vp]
    L.pydsb_recourse_solve.restype = ci
    L.pydsb_recourse_solve.argtypes = [vp, ci, dp, ll, dp, dp, ll, dp, ll, dp, dp, dp, vp, vp, vp]
    L.pydsb_recourse_solve_global_dist.restype = ci
    L.pydsb_recourse_solve_global_dist.argtypes = [vp, dp, ll, ll, dp, dp, ll, dp, dp, dp, dp, vp, vp, dp, ALLREDUCE_FN, vp, vp]
    L.pydsb_get_counters.restype = ci
    L.pydsb_get_counters.argtypes = [vp, ctypes.POINTER(ctypes.c_ulonglong), ci, ci]
    L.pydsb_set_profiling.restype = ci
    L.pydsb_set_profiling.argtypes = [vp, ci]
    L.pydsb_get_profile.restype = ci
    L.pydsb_get_profile.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ci, ci]
    L.pydsb_dfma_peak.restype = ci
    L.pydsb_dfma_peak.argtypes = [ci, ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise PydsbError(f"libpyds_b200 error {rc}: {load().pydsb_last_error().decode()}")


def _np_ptr(a):
    return None if a is None else a.ctypes.data


def _np64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class Engine:
    """One uploaded model on one device (wraps a ``pydsb_model`` handle)."""

    def __init__(self, blob: bytes, device: int = 0):
        self._h = ctypes.c_void_p()
        self._lib = load()
        self.device = int(device)
        _check(self._lib.pydsb_model_create(blob, len(blob), self.device, ctypes.byref(self._h)))
        buf = (ctypes.c_longlong * len(INFO_NAMES))()
        _check(self._lib.pydsb_model_info(self._h, buf, len(INFO_NAMES)))
        self.info = {k: int(v) for k, v in zip(INFO_NAMES, buf)}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pydsb_model_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host (numpy) path: what the drop-in Manager.g uses --------------------------------------
    def eval_host(self, d, theta, w=None, z=None, z_mode=None, want_g=False, want_ind=False, want_P=True):
        d = _np64(d)
        d = d.reshape(-1, self.info["d_dim"]) if self.info["d_dim"] else d.reshape(len(d), 0)
        theta = _np64(theta).reshape(-1, self.info["p_dim"])
        n_d, n_t = d.shape[0], theta.shape[0]
        w = _np64(w)
        if w is not None and w.shape != (n_t,):
            raise ValueError("w must have one weight per theta sample")
        z, z_mode = self._norm_z(z, z_mode, n_d, n_t)
        g = np.empty((n_d, n_t, self.info["ng"])) if want_g else None
        ind = np.empty((n_d, n_t)) if want_ind else None
        P = np.empty(n_d) if want_P else None
        _check(self._lib.pydsb_eval_host(self._h, _np_ptr(d), n_d, _np_ptr(theta), _np_ptr(w), n_t, _np_ptr(z), z_mode,
                                         _np_ptr(g), _np_ptr(ind), _np_ptr(P)))
        return g, ind, P

    def _norm_z(self, z, z_mode, n_d, n_t):
        nz = self.info["nz"]
        if z is None or nz == 0:
            return None, Z_MODEL
        z = _np64(z)
        if z_mode is None:
            z_mode = {1: Z_SHARED, 2: Z_PER_DESIGN, 3: Z_PER_SCENARIO}[z.ndim]
        want = {Z_SHARED: (nz,), Z_PER_DESIGN: (n_d, nz), Z_PER_SCENARIO: (n_d, n_t, nz)}[z_mode]
        if z.shape != want:
            raise ValueError(f"z has shape {z.shape}, expected {want}")
        return z, z_mode

    # -- device (torch tensor) path ----------------------------------------------------------------
    def eval_device(self, d, theta, w=None, z=None, z_mode=Z_MODEL, g_out=None, ind_out=None, P_out=None, stream=None,
                    feas_tol=0.0):
        """All arguments are contiguous fp64 CUDA tensors (or None); asynchronous on ``stream``.  ``feas_tol`` > 0:
        a scenario counts as feasible when ``g_k / M_k <= feas_tol`` for every k (the relaxed big-M rule)."""
        import torch
        for name, t in (("d", d), ("theta", theta), ("w", w), ("z", z), ("g_out", g_out), ("ind_out", ind_out), ("P_out", P_out)):
            if t is None:
                continue
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.device.index == self.device):
                raise ValueError(f"{name} must be a contiguous float64 CUDA tensor on device {self.device}")
        n_d = d.shape[0]
        n_t = theta.shape[0]
        if theta.shape[1] != self.info["p_dim"] or (d.dim() != 2 or d.shape[1] != self.info["d_dim"]):
            raise ValueError("d/theta column count does not match the model")
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        ptr = lambda t: None if t is None else t.data_ptr()
        _check(self._lib.pydsb_eval_tol(self._h, ptr(d), n_d, ptr(theta), ptr(w), n_t, ptr(z),
                                        z_mode if z is not None else Z_MODEL, float(feas_tol), ptr(g_out), ptr(ind_out),
                                        ptr(P_out), ctypes.c_void_p(stream.cuda_stream)))

    def state_solve(self, d, theta, z=None, z_mode=None):
        """The collocation solution behind g (numpy in / numpy out): ``x`` (n_d, n_t, 1 + ncp * nfe, ns + nq) -- tau = 0
        then the Radau points of every finite element; dynamic states first, then quadrature states -- and ``failed``
        (n_d, n_t) bool.  ``time_grid()`` gives the tau values of axis 2."""
        import torch
        dev = torch.device("cuda", self.device)
        d = _np64(d)
        d = d.reshape(-1, self.info["d_dim"]) if self.info["d_dim"] else d.reshape(len(d), 0)
        theta = _np64(theta).reshape(-1, self.info["p_dim"])
        n_d, n_t = d.shape[0], theta.shape[0]
        z, z_mode = self._norm_z(z, z_mode, n_d, n_t)
        n_time = 1 + self.info["ncp"] * self.info["nfe"]
        n_col = self.info["ns"] + self.info["nq"]
        if n_d == 0 or n_t == 0:
            return np.zeros((n_d, n_t, n_time, n_col)), np.zeros((n_d, n_t), dtype=bool)
        up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        td, tt, tz = up(d), up(theta), up(z)
        x = torch.empty((n_d, n_t, n_time, n_col), dtype=torch.float64, device=dev)
        failed = torch.empty((n_d, n_t), dtype=torch.float64, device=dev)
        ptr = lambda t: None if t is None or t.numel() == 0 else t.data_ptr()
        stream = torch.cuda.current_stream(self.device)
        _check(self._lib.pydsb_state_solve(self._h, ptr(td), n_d, ptr(tt), n_t, ptr(tz), z_mode, ptr(x), ptr(failed),
                                           ctypes.c_void_p(stream.cuda_stream)))
        stream.synchronize()
        return x.cpu().numpy(), failed.cpu().numpy() != 0.0

    def time_grid(self):
        """Normalised time of the trajectory rows: 0, then (e + c_j) / nfe for the Radau roots c_j."""
        from .collocation import radau_tables
        nfe, ncp = self.info["nfe"], self.info["ncp"]
        c = radau_tables(ncp)[0][1:]
        return np.concatenate([[0.0], ((np.arange(nfe)[:, None] + c[None, :]) / nfe).reshape(-1)])

    def recourse_solve_device(self, d, theta, w=None, z0=None, per_scenario=False, global_group=False, mu0=0.0, mu_min=0.0,
                              lam0=0.0, step_tol=0.0, max_iter=0, check_every=0):
        """Optimise the controls; contiguous fp64 CUDA tensors in, CUDA tensors out (z (G, nz), phi, status, iters with
        G = n_d * n_t groups in per-scenario mode, 1 with ``global_group``, else n_d).  Synchronous."""
        import torch
        dev = torch.device("cuda", self.device)
        n_d, n_t, nz = d.shape[0], theta.shape[0], self.info["nz"]
        mode = 1 if per_scenario else (2 if global_group else 0)
        G = n_d * n_t if per_scenario else (1 if global_group else n_d)
        rows = 0 if z0 is None else z0.shape[0]
        if z0 is not None and rows not in (1, G):
            raise ValueError(f"z0 must have 1 or {G} rows")
        z = torch.empty((max(G, 1), nz), dtype=torch.float64, device=dev)
        phi = torch.empty(max(G, 1), dtype=torch.float64, device=dev)
        st = torch.zeros(max(G, 1), dtype=torch.int32, device=dev)
        it = torch.zeros(max(G, 1), dtype=torch.int32, device=dev)
        opts = np.array([mu0, mu_min, lam0, step_tol, max_iter, check_every], dtype=np.float64)
        ptr = lambda t: None if t is None else t.data_ptr()
        stream = torch.cuda.current_stream(self.device)
        torch.cuda.synchronize(self.device)
        _check(self._lib.pydsb_recourse_solve(self._h, mode, ptr(d), n_d, ptr(theta), ptr(w), n_t, ptr(z0),
                                              rows, opts.ctypes.data, ptr(z), ptr(phi), ptr(st), ptr(it),
                                              ctypes.c_void_p(stream.cuda_stream)))
        return dict(z=z[:G], phi=phi[:G], status=st[:G], iters=it[:G])

    def recourse_solve_global_dist(self, d_local, n_d_total, theta, w=None, z0=None, group=None, mu0=0.0, mu_min=0.0,
                                   lam0=0.0, step_tol=0.0, max_iter=0):
        """The global-control solve with the design points sharded over the ranks of ``group`` (torch.distributed, NCCL):
        ``d_local`` are this rank's rows (CUDA fp64, may be empty), ``n_d_total`` the rows of all ranks; theta / w / z0
        are replicated.  Every rank gets the same dict(z (1, nz), phi, status, iters).  One all-reduce of the
        6 x NV statistics per optimiser iteration is the only exchange."""
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.device)
        n_d, n_t, nz = d_local.shape[0], theta.shape[0], self.info["nz"]
        z = torch.empty((1, nz), dtype=torch.float64, device=dev)
        phi = torch.empty(1, dtype=torch.float64, device=dev)
        st = torch.zeros(1, dtype=torch.int32, device=dev)
        it = torch.zeros(1, dtype=torch.int32, device=dev)
        nv = 2 + nz + nz * (nz + 1) // 2
        buf = torch.zeros(6 * nv, dtype=torch.float64, device=dev)
        opts = np.array([mu0, mu_min, lam0, step_tol, max_iter, 1], dtype=np.float64)
        ptr = lambda t: None if t is None or t.numel() == 0 else t.data_ptr()
        stream = torch.cuda.current_stream(self.device)
        err = []

        def reduce(p, n, _user):
            try:
                if n != buf.numel():                     # the library and this wrapper must agree on the layout
                    raise RuntimeError(f"reduction of {n} doubles, buffer holds {buf.numel()}")
                if dist.get_backend(group) == "nccl":
                    dist.all_reduce(buf, group=group)
                    torch.cuda.current_stream(self.device).synchronize()
                else:                                    # host-side backends (gloo): 276 doubles through the host
                    h = buf.cpu()
                    dist.all_reduce(h, group=group)
                    buf.copy_(h)
                    torch.cuda.synchronize(self.device)
                return 0
            except Exception as e:                       # an exception must not unwind through the C frames
                err.append(e)
                return 1

        # every rank must enter the loop with the same problem, or the all-reduce inside it never completes: agree on the
        # sizes first, so that an inconsistent call fails on ALL ranks here instead of hanging some of them
        local_ok = int(d_local.dim() == 2 and d_local.shape[1] == self.info["d_dim"] and theta.dim() == 2 and
                       theta.shape[1] == self.info["p_dim"] and (w is None or w.numel() == n_t) and nz > 0)
        on = dev if dist.get_backend(group) == "nccl" else torch.device("cpu")
        mine = torch.tensor([float(n_d_total), float(n_t), float(nz), float(local_ok)], dtype=torch.float64, device=on)
        lo_, hi_, tot = mine.clone(), mine.clone(), torch.tensor([float(n_d)], dtype=torch.float64, device=on)
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(tot, group=group)
        if not torch.equal(lo_, hi_) or lo_[3].item() != 1.0 or int(tot.item()) != int(n_d_total):
            raise ValueError(f"inconsistent distributed solve: (n_d_total, n_t, nz, ok) ranges {lo_.tolist()} .. {hi_.tolist()}, "
                             f"rows of all ranks {int(tot.item())}")
        cb = ALLREDUCE_FN(reduce)
        torch.cuda.synchronize(self.device)
        rc = self._lib.pydsb_recourse_solve_global_dist(self._h, ptr(d_local), n_d, int(n_d_total), ptr(theta), ptr(w), n_t,
                                                        ptr(z0), opts.ctypes.data, ptr(z), ptr(phi), ptr(st), ptr(it),
                                                        ptr(buf), cb, None, ctypes.c_void_p(stream.cuda_stream))
        if err:
            raise err[0]
        _check(rc)
        return dict(z=z, phi=phi, status=st, iters=it)

    def recourse_solve(self, d, theta, w=None, z0=None, per_scenario=False, global_group=False, **opts):
        """Optimise the controls (numpy in / numpy out).  Returns dict(z, phi, status, iters); ``z`` is
        (n_d, nz) in shared mode and (n_d, n_t, nz) in per-scenario mode."""
        import torch
        dev = torch.device("cuda", self.device)
        d = _np64(d).reshape(-1, self.info["d_dim"])
        theta = _np64(theta).reshape(-1, self.info["p_dim"])
        n_d, n_t, nz = d.shape[0], theta.shape[0], self.info["nz"]
        G = n_d * n_t if per_scenario else (1 if global_group else n_d)
        td = torch.from_numpy(d).to(dev)
        tt = torch.from_numpy(theta).to(dev)
        tw = None if w is None else torch.from_numpy(_np64(w)).to(dev)
        tz0 = None
        if z0 is not None:
            z0 = _np64(z0).reshape(-1, nz)
            if z0.shape[0] not in (1, G):
                raise ValueError(f"z0 must have 1 or {G} rows")
            tz0 = torch.from_numpy(z0).to(dev)
        r = self.recourse_solve_device(td, tt, tw, tz0, per_scenario=per_scenario, global_group=global_group, **opts)
        shape = (n_d, n_t, nz) if per_scenario else ((nz,) if global_group else (n_d, nz))
        out = dict(z=r["z"].cpu().numpy().reshape(shape), phi=r["phi"].cpu().numpy(), status=r["status"].cpu().numpy(),
                   iters=r["iters"].cpu().numpy())
        if per_scenario:
            for k in ("phi", "status", "iters"):
                out[k] = out[k].reshape(n_d, n_t)
        return out

    def set_profiling(self, on=True):
        """Per-kernel device time (CUDA events around every launch); read with :meth:`profile`."""
        _check(self._lib.pydsb_set_profiling(self._h, 1 if on else 0))

    def profile(self, reset=False):
        buf = (ctypes.c_double * len(PROFILE_NAMES))()
        _check(self._lib.pydsb_get_profile(self._h, buf, len(PROFILE_NAMES), 1 if reset else 0))
        return {k: float(v) for k, v in zip(PROFILE_NAMES, buf)}

    def counters(self, reset=False):
        buf = (ctypes.c_ulonglong * len(COUNTER_NAMES))()
        _check(self._lib.pydsb_get_counters(self._h, buf, len(COUNTER_NAMES), 1 if reset else 0))
        return {k: int(v) for k, v in zip(COUNTER_NAMES, buf)}


def ns_ellipsoid(live, lo, hi):
    """Bounding ellipsoid of the live points in box-scaled coordinates (host only): returns (mu, L, r)."""
    live = np.ascontiguousarray(live, dtype=np.float64)
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    n, dim = live.shape
    mu, L, r = np.empty(dim), np.empty((dim, dim)), ctypes.c_double(0.0)
    _check(load().pydsb_ns_ellipsoid(live.ctypes.data, n, dim, lo.ctypes.data, hi.ctypes.data, mu.ctypes.data, L.ctypes.data,
                                     ctypes.addressof(r)))
    return mu, L, r.value


def ns_replace_worst(live_p, cand_p, target):
    """The live-set update of one nested-sampling iteration in native code (host only, no device needed): ``live_p`` (a
    contiguous float64 array) is updated in place; returns the slot each candidate took, or -1."""
    if not (isinstance(live_p, np.ndarray) and live_p.dtype == np.float64 and live_p.flags.c_contiguous):
        raise ValueError("live_p must be a contiguous float64 array (it is updated in place)")
    cand_p = np.ascontiguousarray(cand_p, dtype=np.float64)
    slots = np.empty(len(cand_p), dtype=np.int64)
    _check(load().pydsb_ns_replace_worst(live_p.ctypes.data, len(live_p), cand_p.ctypes.data, len(cand_p), float(target),
                                         slots.ctypes.data))
    return slots


# opcode numbering of the assembled program (csrc/feas_kernel.cuh, enum XCode)
XCODES = ["POLY1", "MUL_NW", "NS1", "ELEM_END",
          "FMA_NWN", "FMA_NWW", "SQR_W", "FMS_NWW", "MUL_WW", "MULSQ_NW", "MUL1", "FMA1", "ADD1", "NS1L", "NS2", "NB1",
          "ADD_NW", "FMA_WWN", "FMA_WWW", "FMS_NWN", "FMS_WWN", "FMS_WWW", "FNMA_NWN", "FNMA_NWW", "FNMA_WWN", "FNMA_WWW", "ADD_WW",
          "SUB_WW", "SUB_NW", "SUB_WN", "DIV_NW", "DIV_WN", "DIV_WW",
          "NEG_W", "RCP_W", "ABS_W", "MOV_W", "BC_N", "MIN_NW", "MIN_WW", "MAX_NW", "MAX_WW", "COLD_W",
          "FMS1", "FNMA1", "SUB1", "DIV1", "SQR1", "NEG1", "RCP1", "ABS1", "MOV1", "MIN1", "MAX1", "COLD1",
          "INIT", "LDCTRL", "NB2", "NB3", "NS3", "NS2L", "NS3L", "XEND", "END", "PCHAIN", "FMAG"]


def assemble(blob: bytes):
    """Host-only: the pre-decoded program the feasibility kernel runs for ``blob`` (no GPU needed).  Returns
    ``(ops, info)``: ``ops`` is the list of ``(pcw, name, words)`` in program order, ``info`` the layout figures."""
    L = load()
    n = ctypes.c_size_t()
    info = (ctypes.c_longlong * 8)()
    _check(L.pydsb_assemble(blob, len(blob), None, 0, ctypes.byref(n), info, 8))
    words = np.zeros(n.value, dtype=np.uint32)
    _check(L.pydsb_assemble(blob, len(blob), words.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n), info, 8))
    ns_len = lambda k: 1 + (k + 3 * k + 3 * k * k + 3) // 4
    ops, pcw = [], 0
    while 4 * pcw < len(words):
        code = int(words[4 * pcw])
        name = XCODES[code]
        if name in ("NS1", "NS1L"): length = ns_len(1)
        elif name in ("NS2", "NS2L"): length = ns_len(2)
        elif name in ("NS3", "NS3L"): length = ns_len(3)
        elif name == "ELEM_END": length = 1 + 2 * int(words[4 * pcw + 2])
        elif name == "POLY1": length = 3 if int(words[4 * pcw + 7]) & 16 else 2      # c0 = narrow * wide^2: a third word
        elif name == "PCHAIN": length = 2 + 3 * (int(words[4 * pcw + 1]) & 0xFF) + ((int(words[4 * pcw + 1]) >> 8) & 0xFF)
        else: length = 2
        ops.append((pcw, name, words[4 * pcw:4 * (pcw + length)].copy()))
        pcw += length
    return ops, dict(zip(("scenarios_per_thread", "units", "smem_bytes", "program_length", "chain_blocks", "chain_shape_sb",
                          "chain_shape_sq", "chain_shape_registered"), (int(v) for v in info)))


def chain_descriptor(blob: bytes):
    """Host-only: the descriptor of the fused chain of ``blob``'s feasibility program (csrc/poly_chain.cuh, ChainDesc) as
    a dict of byte offsets into a thread's register file, or None when the element loop is interpreted."""
    L = load()
    n = ctypes.c_size_t()
    N = 8 + 130
    info = (ctypes.c_longlong * N)()
    _check(L.pydsb_assemble(blob, len(blob), None, 0, ctypes.byref(n), info, N))
    if int(info[4]) == 0:
        return None
    w = [int(v) for v in info[8:]]
    head = dict(zip(("nb", "nq", "max_it", "pre_end", "g_begin", "g_end", "n_par", "ng"), w[:8]))
    blocks = []
    for b in range(3):
        o = 8 + 14 * b
        blocks.append(dict(x_off=w[o], flops=w[o + 1], kind=w[o + 2], minv_off=w[o + 3],
                           terms=[(w[o + 4 + 2 * k], w[o + 5 + 2 * k]) for k in range(3)],      # (n_off, meta)
                           x0_off=w[o + 10], x0_const=w[o + 11], xe_off=w[o + 12]))
    quads = []
    for q in range(4):
        o = 8 + 42 + 8 * q
        quads.append(dict(q_off=w[o], n_off=w[o + 1], meta=w[o + 2], q0_off=w[o + 3], q0_const=w[o + 4]))
    o = 8 + 42 + 32
    head.update(blocks=blocks[:head["nb"]], quads=quads[:head["nq"]], pre_run0=w[o], g_run0=w[o + 1],
                par_off=w[o + 4:o + 4 + head["n_par"]], g_off=w[o + 20:o + 20 + head["ng"]])
    nz = w[o + 45]
    head.update(z_off=w[o + 28:o + 28 + nz], z_const=[(w[o + 44] >> i) & 1 for i in range(nz)])
    return head


def dfma_peak(device: int = 0, repeats: int = 5):
    """Measured FP64 FMA throughput (TFLOP/s, best of ``repeats``) and the kernel time in ms."""
    t, ms = ctypes.c_double(), ctypes.c_double()
    _check(load().pydsb_dfma_peak(device, repeats, ctypes.byref(t), ctypes.byref(ms)))
    return t.value, ms.value
