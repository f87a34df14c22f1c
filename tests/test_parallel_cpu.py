"""world_size-2 gloo tests of the N>1 host logic (sharding + gather).  The compute stand-in is the
oracle -- this is test infrastructure; on the GPU box the engine is the CUDA one."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partition():
    from pyds_b200.parallel import shard_rows, shard_sizes
    for n in (0, 1, 7, 8, 100, 12501):
        for world in (1, 2, 3, 8):
            blocks = [shard_rows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


class _OracleEngine:
    """Stand-in with the Engine.eval_host signature, backed by the C oracle (tests only)."""

    def eval_host(self, d, theta, w=None, z=None, z_mode=None, want_g=False, want_ind=False, want_P=True):
        from oracle import cbind
        n_t = len(theta)
        w = np.full(n_t, 1.0 / n_t) if w is None else w
        g, P, _ = cbind.kinetics_grid(d, theta, w, F=z, want_g=want_g)
        return g, None, P


def _worker(rank, world, port, n_d, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import kinetics as K
    from pyds_b200.parallel import ShardedEfp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(4)
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    p = K.theta_samples()
    w = np.full(50, 0.02)
    sh = ShardedEfp(_OracleEngine(), world, rank)
    P = sh.efp_global(d, p, w)
    q.put((rank, P))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_d", [10, 11, 1])
def test_sharded_efp_gloo_world2(n_d):
    from oracle import cbind, kinetics as K
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_d) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_d, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    rng = np.random.default_rng(4)
    d = np.stack([rng.uniform(250, 350, n_d), rng.uniform(250, 300, n_d)], axis=1)
    _, P_ref, _ = cbind.kinetics_grid(d, K.theta_samples(), np.full(50, 0.02), want_g=False)
    for r in range(2):
        np.testing.assert_allclose(res[r], P_ref, atol=1e-15)


def _manager_worker(rank, world, port, shards, q, n_d=11):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pyds_b200 import run
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Stub(run.Manager):
        shards_by_design_point = shards != False                          # noqa: E712  ("coupled" shards too)
        couples_design_points = shards == "coupled"

        def __init__(self):
            self.rows_seen = []

        def _check(self, d, p):
            return np.asarray(d, dtype=np.float64), np.asarray(p, dtype=np.float64)

        def _indicators(self, d, p, w, n_d_total=None):
            self.rows_seen.append(len(d) if n_d_total is None else (len(d), n_d_total))
            if n_d_total is not None:                                     # the ranks meet once per optimiser iteration
                import torch
                t = torch.tensor([float(len(d))], dtype=torch.float64)
                dist.all_reduce(t)
                assert int(t.item()) == n_d_total
            return (d[:, :1] > p[None, :, 0]).astype(np.float64)         # indicator 1 where d0 > theta0

    rng = np.random.default_rng(9)
    d, p = rng.uniform(size=(n_d, 2)), rng.uniform(size=(7, 1))
    w = np.full(7, 1.0 / 7)
    m = Stub()
    P = m.efp(d, p, w)
    q.put((rank, P, m.rows_seen))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shards", [True, False])
def test_manager_efp_sharding_rule_gloo_world2(shards):
    """Managers split the design points over the ranks unless they opt out (then every rank sees the whole batch)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + int(shards)) % 2000
    procs = [ctx.Process(target=_manager_worker, args=(r, 2, port, shards, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = {r: (P, rows) for r, P, rows in (q.get(timeout=180) for _ in range(2))}
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    rng = np.random.default_rng(9)
    d, p = rng.uniform(size=(11, 2)), rng.uniform(size=(7, 1))
    P_ref = ((1.0 - (d[:, :1] > p[None, :, 0])) / 7.0).sum(axis=1)
    for r in range(2):
        np.testing.assert_allclose(res[r][0], P_ref, atol=1e-15)
    assert sorted(res[r][1][0] for r in range(2)) == ([5, 6] if shards else [11, 11])


@pytest.mark.parametrize("n_d", [11, 1])
def test_manager_efp_coupled_design_points_gloo_world2(n_d):
    """The three-stage manager shards its rows AND calls the solver on every rank -- also on a rank without rows --
    because the ranks all-reduce the optimiser statistics of the one global control vector."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() + n_d) % 2000
    procs = [ctx.Process(target=_manager_worker, args=(r, 2, port, "coupled", q, n_d)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = {r: (P, rows) for r, P, rows in (q.get(timeout=180) for _ in range(2))}
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    rng = np.random.default_rng(9)
    d, p = rng.uniform(size=(n_d, 2)), rng.uniform(size=(7, 1))
    P_ref = ((1.0 - (d[:, :1] > p[None, :, 0])) / 7.0).sum(axis=1)
    for r in range(2):
        np.testing.assert_allclose(res[r][0], P_ref, atol=1e-15)
    assert sorted(res[r][1][0] for r in range(2)) == ([(5, 11), (6, 11)] if n_d == 11 else [(0, 1), (1, 1)])


class _TorchEngine:
    """Stand-in with the Engine.eval_device signature on CPU tensors: P = smooth function of the design point."""

    def eval_device(self, d, theta, w=None, z=None, z_mode=0, g_out=None, ind_out=None, P_out=None, stream=None, feas_tol=0.0):
        import torch
        P_out.copy_(torch.clamp(1.2 - ((d - 0.6) ** 2).sum(dim=1) * 4.0, 0.0, 1.0))


def _device_ns_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from pyds_b200 import ns
    from pyds_b200.parallel import ShardedEfp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = ShardedEfp(_TorchEngine(), world, rank)
    # the resident gather itself, ragged (7 rows over 2 ranks)
    d = torch.linspace(0.0, 1.0, 21, dtype=torch.float64).reshape(7, 3)
    P = sh.efp_global_device(d, None, None)
    ref = torch.clamp(1.2 - ((d - 0.6) ** 2).sum(dim=1) * 4.0, 0.0, 1.0)
    assert P.shape == (7,) and torch.equal(P, ref)

    class Owner:
        def g(self, d, p):
            raise NotImplementedError

        def efp_device(self, d, p, w):
            return sh.efp_global_device(d, p, w)

    o = Owner()
    form = {"activity_type": "ds",
            "problem": {"design_variables": [{"a": [0, 1]}, {"b": [0, 1]}, {"c": [0, 1]}], "parameters_best_estimate": [1.0],
                        "parameters_samples": [{"c": [1.0], "w": 1.0}], "target_reliability": 0.9},
            "solver": {"name": "ds-ns", "settings": {"efp_evaluation": {"constraints_func_ptr": o.g}, "points_schedule": [(0.0, 301, 31)],
                                                     "stop_criteria": [{"inside_fraction": 1.0}]},
                       "algorithms": {"sampling": {"settings": {"prng_seed": 7, "device": "cpu", "stop_criteria": [{"max_iterations": 80}]}}}}}
    res = ns.DEUS(form).solve()
    q.put((rank, res["iterations"], res["inside_fraction"], res["live_points"], res["live_efp"]))
    dist.barrier()
    dist.destroy_process_group()


def test_device_resident_nested_sampling_gloo_world2():
    """Every rank seeds the generator alike, draws the same proposals without any exchange, evaluates its own rows, and
    all-gathers the probabilities: both ranks must follow the identical trajectory and reach the target."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + os.getpid() % 2000
    procs = [ctx.Process(target=_device_ns_worker, args=(r, 2, port, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = {r: rest for r, *rest in (q.get(timeout=240) for _ in range(2))}
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    assert res[0][0] == res[1][0] and 0 < res[0][0] < 80 and res[0][1] == 1.0
    np.testing.assert_array_equal(res[0][2], res[1][2])
    np.testing.assert_array_equal(res[0][3], res[1][3])
    assert res[0][3].min() >= 0.9 and res[0][2].shape == (301, 3)
