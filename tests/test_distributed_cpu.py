"""N > 1 host logic on CPU: gloo, world_size 2 (spawned processes, rendezvous on 127.0.0.1)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from knightvision_b200 import parallel as P


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # weight broadcast: every rank ends with rank 0's blob
        blob = torch.arange(1000, dtype=torch.float32) * (1.0 if rank == 0 else -1.0)
        P.broadcast_weights(blob, src=0)
        assert torch.equal(blob, torch.arange(1000, dtype=torch.float32))
        # record gather: ragged counts (rank 0: 3 records, rank 1: 5), global game order
        lo, hi = P.shard_range(9, rank, world)
        n = 3 if rank == 0 else 5
        lines = torch.full((n, 16), rank + 1, dtype=torch.int64)
        move = torch.arange(n, dtype=torch.int32) + 100 * rank
        reward = torch.full((n,), 0.2 if rank == 0 else -1.0)
        game = torch.arange(n, dtype=torch.int32) % (hi - lo)
        out = P.gather_records(lines, move, reward, lo, game, dst=0)
        if rank == 0:
            L, M, R, G = out
            assert L.shape == (8, 16) and M.tolist() == [0, 1, 2, 100, 101, 102, 103, 104]
            assert R[:3].tolist() == [0.2] * 3 or abs(R[0].item() - 0.2) < 1e-6
            assert G.tolist() == [0, 1, 2, 5, 6, 7, 8, 5]
            ret.put("ok")
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_all_games():
    for total, world in ((32768, 8), (10, 4), (5, 8)):
        spans = [P.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == "ok"
