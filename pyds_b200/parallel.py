"""Multi-GPU sharding of the feasibility path: one process per GPU, design points sharded.

The path has no data-path exchange: each design point's P(feasible) depends only on its own
``d`` row and the replicated theta set, so every rank runs the same kernel on a
contiguous block of rows and the only communication is one small all-gather of the per-point
probabilities (NCCL over NVLink on the GPU box, gloo in the CPU tests).  The reference itself is
strictly serial (DEUS ``"method": "serial"``, reference ``run_twostage.py:146,160``).
"""
from __future__ import annotations

import numpy as np


def shard_rows(n: int, world: int, rank: int):
    """Contiguous, balanced row block of rank ``rank``: returns (start, stop)."""
    base, rem = divmod(int(n), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n: int, world: int):
    return [shard_rows(n, world, r)[1] - shard_rows(n, world, r)[0] for r in range(world)]


class ShardedEfp:
    """P(feasible) for a design-point batch spread over the ranks of a process group."""

    def __init__(self, engine, world: int = 1, rank: int = 0, group=None):
        self.engine, self.world, self.rank, self.group = engine, int(world), int(rank), group
        self._P_local = None

    # -- device-resident inputs (equal shard sizes: every rank passes its own rows) ----------------
    def launch_local(self, d, theta, w=None, z=None, z_mode=0):
        import torch
        n = d.shape[0]
        if self._P_local is None or self._P_local.shape[0] != n or self._P_local.device != d.device:
            self._P_local = torch.empty(n, dtype=torch.float64, device=d.device)
        self.engine.eval_device(d, theta, w=w, z=z, z_mode=z_mode, P_out=self._P_local)
        return self._P_local

    def gather(self, P_all):
        import torch.distributed as dist
        if self.world == 1:
            P_all.copy_(self._P_local)
        else:
            dist.all_gather_into_tensor(P_all, self._P_local, group=self.group)
        return P_all

    def efp_device(self, d, theta, w, P_all):
        self.launch_local(d, theta, w)
        return self.gather(P_all)

    # -- host inputs: the user-facing call (numpy in, numpy out) -------------------------------------
    def efp_host(self, d_local, theta, w=None, z=None, equal_shards=False):
        """``d_local`` are THIS rank's rows; returns the probabilities of all ranks' rows, rank order.
        ``equal_shards``: every rank passes the same number of rows (no size exchange needed)."""
        _, _, P = self.engine.eval_host(d_local, theta, w, z=z, want_P=True)
        if self.world == 1:
            return P
        return self.all_gather_host(P, [len(P)] * self.world if equal_shards else None)

    def all_gather_host(self, P_local, sizes=None):
        """Concatenate every rank's block.  ``sizes``: the block lengths when every rank can compute them (rows split
        with ``shard_rows``) -- saves the pickled all_gather_object round trip per call."""
        import torch
        import torch.distributed as dist
        backend = dist.get_backend(self.group)
        # the gather runs on the engine's device (a manager may have been given an explicit one)
        dev = torch.device("cuda", getattr(self.engine, "device", torch.cuda.current_device())) if backend == "nccl" \
            else torch.device("cpu")
        t = torch.from_numpy(np.ascontiguousarray(P_local, dtype=np.float64)).to(dev)
        if sizes is None:
            sizes = [None] * self.world
            dist.all_gather_object(sizes, int(t.numel()), group=self.group)
        elif sizes[self.rank] != t.numel():
            raise ValueError("all_gather_host: this rank's block does not have the announced length")
        if len(set(sizes)) == 1:
            out = torch.empty(self.world * t.numel(), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(out, t, group=self.group)
            return out.cpu().numpy()
        m = max(sizes)                                   # ragged shards: pad to the longest, trim after
        pad = torch.zeros(m, dtype=torch.float64, device=dev)
        pad[:t.numel()] = t
        out = torch.empty(self.world * m, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, pad, group=self.group)
        out = out.view(self.world, m).cpu().numpy()
        return np.concatenate([out[r, :sizes[r]] for r in range(self.world)])

    def efp_global(self, d_global, theta, w=None, z=None):
        """Every rank passes the SAME global batch; rows are sharded here and the full result returned."""
        a, b = shard_rows(len(d_global), self.world, self.rank)
        zl = None
        if z is not None:
            z = np.asarray(z)
            zl = z[a:b] if z.ndim >= 2 else z
        if b > a:
            P = self.engine.eval_host(np.asarray(d_global)[a:b], theta, w, z=zl, want_P=True)[2]
        else:
            P = np.empty(0)
        if self.world == 1:
            return P
        return self.all_gather_host(P, shard_sizes(len(d_global), self.world))

    def efp_global_device(self, d_global, theta, w=None):
        """``efp_global`` with everything resident: ``d_global`` (n, d_dim), ``theta``, ``w`` are fp64 tensors on the
        engine's device and identical on every rank (the nested-sampling driver draws the proposals from a generator every
        rank seeds alike); this rank evaluates its row block, one ``all_gather_into_tensor`` returns the probabilities of
        all rows as a device tensor.  No host round trip: the only transfers are the gather over NVLink.  Works with CPU
        tensors / gloo when the engine does (tests use a stand-in engine)."""
        import torch
        n = d_global.shape[0]
        a, b = shard_rows(n, self.world, self.rank)
        m = max(shard_sizes(n, self.world)) if self.world > 1 else n
        P_pad = torch.zeros(m, dtype=torch.float64, device=d_global.device)
        if b > a:
            self.engine.eval_device(d_global[a:b].contiguous(), theta, w=w, P_out=P_pad[:b - a])
        if self.world == 1:
            return P_pad
        import torch.distributed as dist
        out = torch.empty(self.world * m, dtype=torch.float64, device=d_global.device)
        dist.all_gather_into_tensor(out, P_pad, group=self.group)
        sizes = shard_sizes(n, self.world)
        if len(set(sizes)) == 1:
            return out
        out = out.view(self.world, m)
        return torch.cat([out[r, :sizes[r]] for r in range(self.world)])
