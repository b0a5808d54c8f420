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


# ---- training loop under DistributedDataParallel (gloo, CPU): step alignment, no_sync accumulation, non-finite skip ----
class _CpuEngine:
    """Stand-in for Engine in CPU tests of learn.train_epochs: only `encode` (answered by the oracle), device, index."""
    device = torch.device("cpu")
    index = 0

    def encode(self, lines):
        import numpy as np
        from oracle import kv_oracle as O
        return torch.from_numpy(O.encode(lines.numpy().view(np.uint64)))


def _train_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["KV_TRAIN_NATIVE"] = "0"            # torch operators: this test is about the loop, not the kernels
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from knightvision_b200 import layout as L
        from knightvision_b200 import learn as LR
        from knightvision_b200.model import ChessNet
        torch.manual_seed(0)
        net = ChessNet(stem=64, tower=64, blocks=1, conv2=False)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3)
        data = LR.ReplayData(_CpuEngine())
        n = 40 if rank == 0 else 56                  # rank 1 holds one batch more: both must run the same number of steps
        lines = torch.from_numpy(np.stack([L.start_line()] * n).view(np.int64))
        g = torch.Generator().manual_seed(rank)
        reward = torch.tensor([1.0, 0.2, -1.0])[torch.randint(0, 3, (n,), generator=g)]
        if rank == 1:
            reward[:] = float("inf")                 # every loss of rank 1 is non-finite: all ranks must skip together
        data.extend_packed(lines, torch.randint(0, 4096, (n,), generator=g), reward)
        before = [p.detach().clone() for p in net.parameters()]
        LR.train_epochs(net, opt, data, epochs=1, batch_size=16, accumulate_steps=2)
        # every batch was skipped by both ranks: nothing moved, nobody hung
        assert all(torch.equal(a, b) for a, b in zip(before, net.parameters()))
        # now finite data everywhere: parameters move and stay identical across ranks
        data.reward[:] = 0.2
        loss = LR.train_epochs(net, opt, data, epochs=1, batch_size=16, accumulate_steps=2)
        assert loss == loss and any(not torch.equal(a, b) for a, b in zip(before, net.parameters()))
        flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        other = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(other, flat)
        assert torch.equal(other[0], other[1])
        # the wrapper is built once per network
        assert LR.training_graph(net, data.eng) is LR.training_graph(net, data.eng)
        if rank == 0:
            ret.put("ok")
    finally:
        dist.destroy_process_group()


def test_train_epochs_ddp_alignment_and_nonfinite_skip_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == "ok"
