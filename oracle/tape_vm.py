"""numpy executor for a lowered model blob (oracle; test infrastructure only).

Runs the SAME blob the CUDA kernel consumes, with plain numpy arrays as the register file and
LAPACK (partial pivoting) for the Newton step, so that (a) the host-side lowering can be tested
on a CPU-only box against the hand-written restatement in ``oracle/kinetics.py`` and (b) any
model the DSL can express has a CPU oracle for the GPU parity tests.

Algorithm restated: per finite element solve  sum_k adot[k][j] x_k = h f(x_j)  (j = 1..3) by full
Newton from x_j = x_0; quadrature states advance by  h * sum_j bw_j fq(x_j);  afterwards evaluate
the CQAs g_k, the relaxed big-M indicator y = max(0, max_k g_k/M_k) (reference pyds/utils.py:23-29,
SURVEY.md A.4) and snap it as pyds/solver.py:71-80 does.
"""
from __future__ import annotations

import struct

import numpy as np

_OPS = ["MOV", "NEG", "ADD", "SUB", "MUL", "DIV", "FMA", "FMS", "FNMA", "SQR", "RCP", "EXP", "LOG",
        "SQRT", "SIN", "COS", "TANH", "POW", "ABS", "MIN", "MAX", "BC3"]
_REF_CONST, _REF_WIDE, _REF_MASK, _REF_NONE = 1 << 13, 1 << 12, 0xFFF, 0x3FFF
_DT = {0: np.uint32, 1: np.float64, 2: np.uint64, 3: np.uint16}


def parse_blob(blob: bytes):
    magic, ver, nsec, _ = struct.unpack_from("<IIII", blob, 0)
    assert magic == 0x42534450 and ver == 1
    off, out = 16, {}
    cur = out
    for _ in range(nsec):
        tag, dt, cnt, _ = struct.unpack_from("<4sIII", blob, off)
        off += 16
        nbytes = cnt * np.dtype(_DT[dt]).itemsize
        arr = np.frombuffer(blob, dtype=_DT[dt], count=cnt, offset=off).copy()
        off += nbytes + ((-nbytes) % 8)
        if tag == b"PSEP":                               # payload: the program's number (1 sensitivities, 2 chain + dg)
            cur = out.setdefault("sens" if int(arr[0]) == 1 else "chain", {})
            continue
        cur[tag.decode()] = arr
    return out


class TapeVM:
    def __init__(self, blob: bytes, program: int = 0):
        s = parse_blob(blob)
        if program:
            s = s["sens"]
        self.sec = s
        (self.ns, self.nq, self.ng, self.d_dim, self.p_dim, self.nz, self.nfe, self.ncp, self.nce, self.n_units,
         self.ctrl_slot, self.xe_slot, self.qe_slot, self.x_slot, self.max_it,
         _, self.fq_state_dep, _, self.n_blocks) = [int(v) for v in s["HEAD"]]
        # blocks of the block-triangular state solve: (states, linear flag, tape) + the quadrature tape
        B, Tn = s["BLKS"], s["BLKT"]
        self.blocks, pc, fpos, jpos = [], 0, 0, 0
        for b in range(self.n_blocks):
            nsb = int(B[5 * b])
            st = [int(x) for x in B[5 * b + 2:5 * b + 2 + nsb]]
            kind = int(B[5 * b + 1])                      # bit 0: linear, bit 1: polynomial in its own single state
            blk = dict(states=st, linear=kind & 1, code=s["TPNT"][pc:pc + int(Tn[b])],
                       f=s["RF__"][fpos:fpos + nsb], J=s["RJ__"][jpos:jpos + nsb * nsb], poly=None)
            if kind & 2:                                  # f = c0 + c1 x + c2 x^2: references of the coefficients
                blk["poly"] = [int(s["RF__"][fpos]), int(s["RJ__"][jpos]), int(s["RP2_"][b])]
            self.blocks.append(blk)
            pc += int(Tn[b]); fpos += nsb; jpos += nsb * nsb
        self.q_code = s["TPNT"][pc:pc + int(Tn[self.n_blocks])]
        self.param_slots = [int(v) for v in s["RPAR"]]
        self.rtol = float(s["DBLS"][0])
        self.consts = s["CNST"]
        self.adot = s["ADOT"].reshape(self.ncp + 1, self.ncp + 1)
        self.bw = s["BWTS"]
        self.big_m = s["BIGM"]
        self.cmap = s["CMAP"].reshape(self.nfe, -1)

    # register file: R[slot] is an array (S,) ; wide slots are three consecutive entries ---------
    def _ld(self, R, ref, w):
        ref = int(ref)
        if ref == _REF_NONE:
            return 0.0
        if ref & _REF_CONST:
            return self.consts[ref & _REF_MASK]
        idx = ref & _REF_MASK
        return R[idx + (w if ref & _REF_WIDE else 0)]

    def _run(self, R, code, width):
        for word in code:
            word = int(word)
            op = _OPS[word & 0xFF]
            dst, a, b, c = (word >> 8) & 0x3FFF, (word >> 22) & 0x3FFF, (word >> 36) & 0x3FFF, (word >> 50) & 0x3FFF
            if op == "BC3":
                for w in range(3):
                    R[(dst & _REF_MASK) + w] = np.broadcast_to(self._ld(R, a, 0), R[0].shape).copy()
                continue
            for w in range(width):
                A = self._ld(R, a, w)
                if op == "MOV": v = A
                elif op == "NEG": v = -A
                elif op == "ADD": v = A + self._ld(R, b, w)
                elif op == "SUB": v = A - self._ld(R, b, w)
                elif op == "MUL": v = A * self._ld(R, b, w)
                elif op == "DIV": v = A / self._ld(R, b, w)
                elif op == "FMA": v = A * self._ld(R, b, w) + self._ld(R, c, w)
                elif op == "FMS": v = A * self._ld(R, b, w) - self._ld(R, c, w)
                elif op == "FNMA": v = self._ld(R, c, w) - A * self._ld(R, b, w)
                elif op == "SQR": v = A * A
                elif op == "RCP": v = 1.0 / A
                elif op == "EXP": v = np.exp(A)
                elif op == "LOG": v = np.log(A)
                elif op == "SQRT": v = np.sqrt(A)
                elif op == "SIN": v = np.sin(A)
                elif op == "COS": v = np.cos(A)
                elif op == "TANH": v = np.tanh(A)
                elif op == "POW": v = np.power(A, self._ld(R, b, w))
                elif op == "ABS": v = np.abs(A)
                elif op == "MIN": v = np.minimum(A, self._ld(R, b, w))
                elif op == "MAX": v = np.maximum(A, self._ld(R, b, w))
                else: raise ValueError(op)
                idx = dst & _REF_MASK
                R[idx + (w if dst & _REF_WIDE else 0)] = np.broadcast_to(v, R[0].shape).copy()

    def evaluate(self, d, p, z=None, tol=None, max_it=None, return_traj=False, want_sens=False):
        """g over the (n_d, n_p) grid.  ``z``: None (model defaults), (nz,), (n_d, nz) or (n_d, n_p, nz).
        Returns dict(g=(n_d,n_p,ng), y, feasible, iters=(n_d,n_p) total Newton iterations, status)."""
        d = np.atleast_2d(np.asarray(d, dtype=np.float64))
        p = np.atleast_2d(np.asarray(p, dtype=np.float64))
        n_d, n_p = d.shape[0], p.shape[0]
        S = n_d * n_p
        tol = self.rtol if tol is None else tol
        max_it = self.max_it if max_it is None else max_it
        R = [np.zeros(S) for _ in range(self.n_units)]
        for i in range(self.d_dim):
            R[self.param_slots[i]] = np.repeat(d[:, i], n_p)
        for i in range(self.p_dim):
            R[self.param_slots[self.d_dim + i]] = np.tile(p[:, i], n_d)
        s = self.sec
        self._run(R, s["TPRO"], 1)
        one = np.ones(S)
        if z is None:
            Z = [self._ld(R, r, 0) * one for r in s["RZ__"][:self.nz]]
        else:
            z = np.asarray(z, dtype=np.float64)
            if z.ndim == 1: z = np.broadcast_to(z, (n_d, n_p, self.nz))
            elif z.ndim == 2: z = np.broadcast_to(z[:, None, :], (n_d, n_p, self.nz))
            z = z.reshape(S, self.nz)
            # fixed (parameter-defined) controls always come from the tape
            Z = [(self._ld(R, s["RZ__"][k], 0) * one) if s["ZFIX"][k] else z[:, k].copy() for k in range(self.nz)]
        ns, nq, ncp = self.ns, self.nq, self.ncp
        x0 = np.stack([self._ld(R, r, 0) * one for r in s["RX0_"][:ns]], axis=1)
        q = np.stack([self._ld(R, r, 0) * one for r in s["RQ0_"][:nq]], axis=1) if nq else np.zeros((S, 0))
        h = 1.0 / self.nfe
        A = self.adot
        iters = np.zeros(S, dtype=np.int64)
        failed = np.zeros(S, dtype=bool)
        traj = np.zeros((S, self.nfe * ncp + 1, ns))
        traj[:, 0] = x0
        n = ns * ncp
        if want_sens:
            if "RFZ_" not in s:
                raise ValueError("sensitivities need the blob's program 1: TapeVM(blob, program=1)")
            Sx = np.zeros((S, ns, self.nz))
            Sq = np.zeros((S, nq, self.nz))
        for e in range(self.nfe):
            for k in range(self.nce):
                R[self.ctrl_slot + k] = Z[int(self.cmap[e, k])]
            self._run(R, s["TELM"], 1)
            X = np.repeat(x0[:, None, :], ncp, axis=1)            # (S, pt, state)

            def put_states(Xc):
                for i in range(ns):
                    for w in range(ncp):
                        R[self.x_slot + 3 * i + w] = Xc[:, w, i]

            put_states(X)
            for blk in self.blocks:
                st = blk["states"]
                nb = len(st)
                nbn = nb * ncp
                active = ~failed
                if blk["poly"] is not None:
                    # the block tape computes the coefficients once (they depend on earlier blocks' states only)
                    put_states(X)
                    self._run(R, blk["code"], ncp)
                    coef = [[self._ld(R, r, w) * one for w in range(ncp)] for r in blk["poly"]]
                for it in range(max_it):
                    if not active.any():
                        break
                    if blk["poly"] is not None:
                        f = np.zeros((S, ncp, 1))
                        J = np.zeros((S, ncp, 1, 1))
                        for w in range(ncp):
                            xw = X[:, w, st[0]]
                            f[:, w, 0] = (coef[2][w] * xw + coef[1][w]) * xw + coef[0][w]
                            J[:, w, 0, 0] = 2.0 * coef[2][w] * xw + coef[1][w]
                    else:
                        put_states(X)
                        self._run(R, blk["code"], ncp)
                        f = np.stack([np.stack([self._ld(R, r, w) * one for r in blk["f"]], axis=1) for w in range(ncp)], axis=1)
                        J = np.zeros((S, ncp, nb, nb))
                        for w in range(ncp):
                            for r_ in range(nb):
                                for c_ in range(nb):
                                    J[:, w, r_, c_] = self._ld(R, blk["J"][r_ * nb + c_], w)
                    Xb = X[:, :, st]
                    lin = np.einsum("kj,skn->sjn", A[1:, 1:], Xb) + A[0, 1:][None, :, None] * x0[:, None, st]
                    Rres = lin - h * f
                    M = np.zeros((S, ncp, nb, ncp, nb))
                    eye = np.eye(nb)
                    for j in range(ncp):
                        for k in range(ncp):
                            M[:, j, :, k, :] += A[k + 1, j + 1] * eye
                        M[:, j, :, j, :] -= h * J[:, j]
                    idx = np.nonzero(active)[0]
                    with np.errstate(all="ignore"):
                        try:
                            dX = np.linalg.solve(M.reshape(S, nbn, nbn)[idx], -Rres.reshape(S, nbn, 1)[idx]).reshape(-1, ncp, nb)
                        except np.linalg.LinAlgError:
                            dX = np.full((idx.size, ncp, nb), np.nan)
                    Xn = X[idx]
                    Xn[:, :, st] += dX
                    X[idx] = Xn
                    iters[idx] += 1
                    scale = np.maximum(np.abs(X[idx][:, :, st]).max(axis=(1, 2)), 1.0)
                    dmax = np.abs(dX).max(axis=(1, 2))
                    bad = ~np.isfinite(dmax)
                    failed[idx[bad]] = True
                    done = (dmax <= tol * scale) | bad
                    active[idx[done]] = False
                failed |= active                                     # iteration cap reached
            # quadrature integrands (and, in the sensitivity program, everything else) at the converged points
            put_states(X)
            for blk in self.blocks:
                if want_sens or not len(self.q_code):
                    self._run(R, blk["code"], ncp)
            self._run(R, self.q_code, ncp)
            for k in range(nq):
                acc = 0.0
                for w in range(ncp):
                    acc = acc + self.bw[w] * self._ld(R, s["RFQ_"][k], w)
                q[:, k] = q[:, k] + h * acc
            if want_sens:
                # forward sensitivities: M dX_c = -dR/dx_prev S[:,c] - dR/dz_c at the converged points
                J = np.zeros((S, ncp, ns, ns))
                for w in range(ncp):
                    for r_ in range(ns):
                        for c_ in range(ns):
                            J[:, w, r_, c_] = self._ld(R, s["RJ__"][r_ * ns + c_], w)
                M = np.zeros((S, ncp, ns, ncp, ns))
                eye = np.eye(ns)
                for j in range(ncp):
                    for k in range(ncp):
                        M[:, j, :, k, :] += A[k + 1, j + 1] * eye
                    M[:, j, :, j, :] -= h * J[:, j]
                M = M.reshape(S, n, n)
                rhs = np.zeros((S, ncp, ns, self.nz))
                for c in range(self.nz):
                    for j in range(ncp):
                        rhs[:, j, :, c] = -A[0, j + 1] * Sx[:, :, c]
                        for k in range(self.nce):
                            if int(self.cmap[e, k]) == c:
                                for r_ in range(ns):
                                    rhs[:, j, r_, c] += h * self._ld(R, s["RFZ_"][r_ * self.nce + k], j)
                with np.errstate(all="ignore"):
                    dXc = np.linalg.solve(M, rhs.reshape(S, n, self.nz)).reshape(S, ncp, ns, self.nz)
                Sx = dXc[:, ncp - 1].copy()
                for qq in range(nq):
                    for c in range(self.nz):
                        acc = 0.0
                        for j in range(ncp):
                            t = 0.0
                            for r_ in range(ns):
                                t = t + self._ld(R, s["RFQX"][qq * ns + r_], j) * dXc[:, j, r_, c]
                            for k in range(self.nce):
                                if int(self.cmap[e, k]) == c:
                                    t = t + self._ld(R, s["RFQZ"][qq * self.nce + k], j)
                            acc = acc + self.bw[j] * t
                        Sq[:, qq, c] += h * acc
            traj[:, e * ncp + 1:(e + 1) * ncp + 1] = X
            x0 = X[:, ncp - 1, :].copy()
        for i in range(ns):
            R[self.xe_slot + i] = x0[:, i]
        for k in range(nq):
            R[self.qe_slot + k] = q[:, k]
        self._run(R, s["TGEE"], 1)
        g = np.stack([self._ld(R, r, 0) * one for r in s["RG__"][:self.ng]], axis=1)
        g[failed] = np.nan
        with np.errstate(invalid="ignore"):
            y = np.maximum(0.0, np.max(g / self.big_m[None, :], axis=1))
        y[failed] = 1.0
        feas = (y == 0.0)
        out = dict(g=g.reshape(n_d, n_p, self.ng), y=y.reshape(n_d, n_p), feasible=feas.reshape(n_d, n_p),
                   iters=iters.reshape(n_d, n_p), status=failed.reshape(n_d, n_p).astype(np.int32),
                   x_end=x0.reshape(n_d, n_p, ns), q_end=q.reshape(n_d, n_p, nq))
        if "RGEE" in s and want_sens:
            ne = ns + nq                                 # second derivatives of the CQAs over e = (x(1), q(1))
            out["gee"] = np.stack([self._ld(R, r, 0) * one for r in s["RGEE"]], axis=1).reshape(n_d, n_p, self.ng, ne, ne)
            out["fxx_end"] = np.stack([self._ld(R, r, ncp - 1) * one for r in s["RFXX"]], axis=1).reshape(n_d, n_p, ns, ns, ns)
        if want_sens:
            dg = np.zeros((S, self.ng, self.nz))
            for k in range(self.ng):
                for r_ in range(ns):
                    dg[:, k, :] += (self._ld(R, s["RGX_"][k * ns + r_], 0) * one)[:, None] * Sx[:, r_, :]
                for qq in range(nq):
                    dg[:, k, :] += (self._ld(R, s["RGQ_"][k * nq + qq], 0) * one)[:, None] * Sq[:, qq, :]
            dg[:, :, s["ZFIX"][:self.nz].astype(bool)] = 0.0
            out["dg"] = dg.reshape(n_d, n_p, self.ng, self.nz)
        if return_traj:
            out["traj"] = traj.reshape(n_d, n_p, -1, ns)
        return out

    def prob_feasible(self, d, p, w, z=None, **kw):
        r = self.evaluate(d, p, z, **kw)
        return (r["feasible"] * np.asarray(w, dtype=np.float64)[None, :]).sum(axis=1)
