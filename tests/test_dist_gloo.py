"""world_size-2 `gloo` tests (CPU) of the multi-process host logic: frame sharding, the
scalar all-reduce that accompanies the loss, and the gather of per-frame match results."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gga_b200 import dist as gd
        assert gd.world() == (rank, world)
        # contiguous shards covering every frame exactly once, sizes differ by <= 1
        n = 3712 + 1
        lo, hi = gd.shard_range(n)
        sizes = [gd.shard_range(n, r, world) for r in range(world)]
        assert sizes[0][0] == 0 and sizes[-1][1] == n and all(a[1] == b[0] for a, b in zip(sizes, sizes[1:]))
        assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1
        # loss scalars: sum over ranks of [loss_sum, weight_sum, n_pos]
        v = torch.tensor([1.0 + rank, 10.0, float(hi - lo)])
        s = gd.reduce_scalars(v)
        assert torch.allclose(s, torch.tensor([sum(1.0 + r for r in range(world)), 10.0 * world, float(n)]))
        m = gd.reduce_mean(torch.tensor([float(rank)]))
        assert torch.allclose(m, torch.tensor([(world - 1) / 2.0]))
        # matching pass: per-frame results gathered in global frame order (uneven shards)
        local = torch.arange(lo, hi, dtype=torch.int32).reshape(-1, 1).repeat(1, 4)
        allr = gd.gather_frames(local, n)
        assert allr.shape == (n, 4) and torch.equal(allr[:, 0], torch.arange(n, dtype=torch.int32))
        # bench.py's warm-up repetition count must be the SAME on every rank although the ranks time
        # their probe run differently (the repeated function holds collectives: a rank that looped once
        # more than another deadlocked the 8-GPU run)
        import sys
        import time
        import types
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench

        def max_over_ranks(v):
            t = torch.tensor([v], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        x = types.SimpleNamespace(max_over_ranks=max_over_ranks)
        reps = bench.ramp_reps(x, lambda: time.sleep(0.002 * (1 + 4 * rank)), seconds=0.05, sync=lambda: None)
        both = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(both, torch.tensor([reps], dtype=torch.int64))
        assert int(both[0]) == int(both[1]) and 3 <= reps <= 6, (reps, both)
        q.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharding_and_scalar_collectives_world2():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res


def test_single_process_defaults():
    from gga_b200 import dist as gd
    assert gd.world() == (0, 1)
    assert gd.shard_range(10) == (0, 10)
    t = torch.tensor([1.0, 2.0])
    assert gd.reduce_scalars(t) is t
    assert torch.equal(gd.gather_frames(t, 2), t)
