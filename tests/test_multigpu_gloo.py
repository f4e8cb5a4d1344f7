"""N>1 path on CPU: world_size-2 gloo processes run the row-sharded sweep (emulated with the oracle:
each rank streams only its rows, the exact int64 block rhs is summed with an all-reduce, the chain is
replicated, ycorr shards are re-assembled) and must reproduce the unsharded sweep bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, p, seed, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from jwas_b200 import multigpu
    from oracle import pyoracle as orc
    from helpers import Problem, uniform_starts
    r, w = multigpu.init_process_group("gloo")
    assert (r, w) == (rank, world)
    uid = multigpu.broadcast_bytes(bytes(range(128)) if rank == 0 else b"", 128)
    assert uid == bytes(range(128))
    prob = Problem(orc, n, p, seed=seed, missing=0.02)
    bounds = multigpu.shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    starts = uniform_starts(p, 64)
    yc, al, be, de = prob.fresh_state()

    def allreduce(arr):
        t = torch.from_numpy(arr)
        dist.all_reduce(t)

    # marker statistics of a sharded handle (jwas_init_sharding): integer code counts over the local rows, summed
    # over the ranks, then the closed form -- must equal the statistics of the whole matrix
    codes = prob.codes[lo:hi]
    cnt = np.array([(codes == 1).sum(axis=0), (codes == 2).sum(axis=0), (~np.isin(codes, (0, 1, 2))).sum(axis=0)], dtype=np.int64)
    allreduce(cnt.reshape(-1))
    nn = n - cnt[2]; ssum = cnt[0] + 2 * cnt[1]
    mu = (ssum.astype(np.float32) / nn.astype(np.float32)).astype(np.float32)
    m64 = mu.astype(np.float64)
    g = (cnt[0] + 4 * cnt[1]).astype(np.float64) - m64 * ssum
    g = g - m64 * ssum
    g = g + (m64 * m64) * nn
    np.testing.assert_array_equal(mu, prob.means)
    np.testing.assert_array_equal(g.astype(np.float32), prob.xpx)

    ve = np.full(p, 0.02); pi = np.full(p, 0.85)
    for it in (1, 2, 3):
        rc, _ = orc.sweep_contract(prob.packed, n, prob.means, prob.xpx, starts, yc, al, be, de,
                                   vare=1.0, varEffects=ve, pi=pi, seed=5, it=it,
                                   row_range=(lo, hi), allreduce=allreduce)
        assert rc == 0
        # re-assemble ycorr: every rank broadcasts its shard
        for src in range(world):
            sl = torch.from_numpy(yc[bounds[src]:bounds[src + 1]])
            dist.broadcast(sl, src=src)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), y=yc, a=al, b=be, d=de)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [203, 512])
def test_row_sharded_sweep_equals_unsharded(tmp_path, oracle, n):
    import torch.multiprocessing as mp
    from helpers import Problem, uniform_starts
    p, seed, world = 150, 17, 2
    port = 29600 + (os.getpid() + n) % 300
    mp.spawn(_worker, args=(world, port, n, p, seed, str(tmp_path)), nprocs=world, join=True)
    prob = Problem(oracle, n, p, seed=seed, missing=0.02)
    yc, al, be, de = prob.fresh_state()
    ve = np.full(p, 0.02); pi = np.full(p, 0.85)
    for it in (1, 2, 3):
        oracle.sweep_contract(prob.packed, n, prob.means, prob.xpx, uniform_starts(p, 64), yc, al, be, de,
                              vare=1.0, varEffects=ve, pi=pi, seed=5, it=it)
    assert de.sum() > 0
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        np.testing.assert_array_equal(got["d"], de)
        np.testing.assert_array_equal(got["a"].view(np.uint32), al.view(np.uint32))
        np.testing.assert_array_equal(got["y"].view(np.uint32), yc.view(np.uint32))


def test_shard_bounds():
    from jwas_b200 import multigpu
    assert multigpu.shard_bounds(50000, 1) == [0, 50000]
    b = multigpu.shard_bounds(50000, 8)
    assert b[0] == 0 and b[-1] == 50000 and all(x % 64 == 0 for x in b[:-1]) and all(y > x for x, y in zip(b, b[1:]))
    assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) < 128
    assert multigpu.shard_bounds(20, 4) == [0, 20, 20, 20, 20]      # fewer 64-row words than ranks (the library refuses)
