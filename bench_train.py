"""bench.py --workload train: one optimisation step of the reference's training loop (scripts/train.py:126-196:
CE + MSE - 0.01 * entropy, grad-clip 1.0, Adam) on synthetic self-play records, BASELINE.json configs[4]
(20-block x 256-channel tower; KV_BENCH_TOWER=ref selects the reference 5 x 512 net).  The tower convolutions run forward,
dgrad and wgrad on the tcgen05 kernels (knightvision_b200/train_ops.py); the same step with cuDNN convolutions
(KV_TRAIN_NATIVE=0 semantics) is timed beside it.  One process per GPU, DistributedDataParallel over NCCL for N > 1.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

BATCH = int(os.getenv("KV_BENCH_TRAIN_BATCH", "2048"))        # scripts/learn.py BATCH_SIZE default
TOWER = os.getenv("KV_BENCH_TOWER", "20x256")
GRAPH = os.getenv("KV_BENCH_TRAIN_GRAPH", "1") != "0"


def _arch():
    if TOWER == "ref":
        return dict(stem=256, tower=512, blocks=5, conv2=True), "reference ChessNet (conv 12->256, 256->512, 5 x ResidualBlock(512))"
    return dict(stem=256, tower=256, blocks=20, conv2=False), "20-block x 256-channel tower (BASELINE config 5), reference heads"


def _tower_flops_per_pos(arch):
    c, c1 = arch["tower"], arch["stem"]
    f = 2.0 * 64 * 9 * (2 * arch["blocks"] * c * c + (c1 * c if arch["conv2"] else 0))
    return f        # one direction; a training step runs fprop + dgrad + wgrad = 3x


ACCUM = int(os.getenv("KV_BENCH_TRAIN_ACCUM", "2"))       # scripts/train.py:183-190 accumulates 2 batches per optimizer step


def _step_fn(LR, net, opt, graph):
    """One optimizer step of the reference's loop: ACCUM micro-batches (loss / ACCUM each), clip 1.0, Adam.  Under
    DistributedDataParallel only the last micro-batch exchanges gradients (no_sync on the others)."""
    import contextlib
    import torch
    import torch.nn.functional as F
    ddp = hasattr(graph, "no_sync")

    def step(boards, moves, outcomes):
        loss = None
        for k in range(ACCUM):
            with (graph.no_sync() if ddp and k + 1 < ACCUM else contextlib.nullcontext()):
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    pol, val = graph(boards)
                pol = pol.float()
                logp = F.log_softmax(pol, dim=1)
                loss = F.cross_entropy(pol, moves) + F.mse_loss(val.squeeze(1).float(), outcomes) \
                    - LR.ENTROPY_COEF * (-(logp.exp() * logp).sum(dim=1).mean())
                (loss / ACCUM).backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss
    return step


def run_reference(args):
    """CPU arm: the same step in plain PyTorch fp32 on the host cores, bounded batch."""
    import torch
    from knightvision_b200 import learn as LR
    from knightvision_b200.model import ChessNet
    arch, name = _arch()
    torch.manual_seed(0)
    net = ChessNet(**arch)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    graph = LR.TrainGraph(net, engine=None)
    b = 32
    x = torch.zeros(b, 12, 8, 8); x[:, 0, 7, 4] = 1
    mv = torch.randint(0, 4096, (b,)); rw = torch.ones(b)

    def step():
        pol, val = graph(x)
        logp = torch.log_softmax(pol, 1)
        loss = torch.nn.functional.cross_entropy(pol, mv) + torch.nn.functional.mse_loss(val.squeeze(1), rw) \
            - LR.ENTROPY_COEF * (-(logp.exp() * logp).sum(1).mean())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step(); opt.zero_grad()
    step()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        step()
    dt = time.perf_counter() - t0
    v = b * max(1, args.steps) / dt
    line = {"impl": "reference", "metric": "train_positions_per_s", "value": v, "unit": "positions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"training step, {name}", "batch": b},
            "cpu_baseline": {"value": v, "unit": "positions/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"batch {b} x {max(1, args.steps)} steps, torch CPU fp32, same graph and loss"},
            "e2e": {"value": v, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run(args, rank, world, local_rank):
    from knightvision_b200.engine import Engine
    eng = Engine(local_rank)
    arms = tuple(a for a in os.getenv("KV_BENCH_TRAIN_ARMS", "native,cudnn,cudnn_nhwc").split(",") if a)
    line = train_record(args, eng, rank, world, local_rank, steps=args.steps, arms=arms, graph=GRAPH)
    if rank == 0:
        print(json.dumps(line), flush=True)


def train_record(args, eng, rank, world, local_rank, steps, arms=("native",), graph=True):
    """The training-step record (rank 0 returns the dict, the other ranks None)."""
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    from bench import Clocks, measured_peaks
    from knightvision_b200 import learn as LR
    from knightvision_b200.model import ChessNet

    class _A:
        pass
    a_ = _A()
    a_.steps, a_.warmup = steps, args.warmup
    args = a_
    GRAPH_ = graph
    dev = eng.device
    arch, name = _arch()
    B = BATCH
    # synthetic records: random legal positions (rules kernels), random move targets, rewards in {1, 0.2, -1}
    lines = eng.random_positions(B, 40, 99 + rank)
    g = torch.Generator(device="cpu").manual_seed(5 + rank)
    moves_h = torch.randint(0, 4096, (B,), generator=g).pin_memory()
    rewards_h = torch.tensor([1.0, 0.2, -1.0])[torch.randint(0, 3, (B,), generator=g)].pin_memory()
    lines_h = lines.cpu().pin_memory()
    moves, rewards = moves_h.to(dev), rewards_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    results = {}
    clk = None
    if "native" not in arms:
        arms = ("native",) + arms
    for arm in arms:
        torch.manual_seed(0)
        net = ChessNet(**arch, max_batch=2).to(dev)
        if arm == "cudnn_nhwc":
            net = net.to(memory_format=torch.channels_last)
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3)
        if arm == "native":
            graph = LR.training_graph(net, eng)      # under torch.distributed: DDP, 25 MB buckets, one exchange per optimizer step
        else:
            graph = LR.TrainGraph(net, engine=None)
            if world > 1:
                graph = nn.parallel.DistributedDataParallel(graph, device_ids=[local_rank])
        step = _step_fn(LR, net, opt, graph)
        boards = eng.encode(lines)
        for _ in range(max(args.warmup, 3)):
            step(boards, moves, rewards)
        torch.cuda.synchronize()
        clocks = Clocks(local_rank) if arm == "native" else None
        if clocks:
            clocks.start()
        eng.profile(True); eng.profile_read()
        l0 = eng.launches
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step(eng.encode(lines), moves, rewards)       # encode kernel: packed records -> planes, every step
        e1.record()
        barrier()
        dev_ms = e0.elapsed_time(e1)
        prof = eng.profile_read(); eng.profile(False)
        launches = eng.launches - l0
        # e2e: pinned host records -> device -> step -> loss back on the host, copies inside the timed region
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ld = lines_h.to(dev, non_blocking=True)
            mv = moves_h.to(dev, non_blocking=True)
            rw = rewards_h.to(dev, non_blocking=True)
            lv = float(step(eng.encode(ld), mv, rw).item())
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        if clocks:
            clk = clocks.stop()
        loss_last = lv
        graph_ms = None
        if GRAPH_ and world == 1:
            # the whole step (encode -> forward -> backward -> clip -> Adam) as ONE CUDA graph: no tracing compiler, the
            # same kernels, launched by the GPU front end instead of ~1 500 host calls
            try:
                loss = lv = None          # drop the eager autograd graph (its AccumulateGrad nodes live on the default stream)
                opt.zero_grad(set_to_none=True)
                opt_g = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)
                step_g = _step_fn(LR, net, opt_g, graph)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        step_g(eng.encode(lines), moves, rewards)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg):
                    loss_g = step_g(eng.encode(lines), moves, rewards)
                cg.replay(); torch.cuda.synchronize()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(args.steps):
                    cg.replay()
                g1.record(); torch.cuda.synchronize()
                graph_ms = g0.elapsed_time(g1) / args.steps
                assert torch.isfinite(loss_g).item()
            except Exception as e:      # report, do not hide
                graph_ms = f"capture failed: {type(e).__name__}: {e}"[:200] + " | kv: " + str(eng._lib.kv_last_error(eng.ctx))
                torch.cuda.synchronize()
        t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[arm] = dict(dev_ms=float(t[0]), e2e_ms=float(t[1]), prof=prof, launches=launches, loss=loss_last, graph_ms=graph_ms)
        if hasattr(net, "_kv_train_graph"):
            object.__delattr__(net, "_kv_train_graph")
        del graph, net, opt
    if rank != 0:
        return None
    peaks = measured_peaks()
    r = results["native"]
    fl = _tower_flops_per_pos(arch)
    PB = B * ACCUM                       # positions per optimizer step and GPU
    value = world * PB * args.steps / (r["dev_ms"] * 1e-3)
    wg_ms, wg_n = r["prof"]["train_wgrad"]
    cv_ms, cv_n = r["prof"]["net_conv"]
    wg_tf = (PB * args.steps * fl) / (wg_ms * 1e-3) / 1e12 if wg_ms else 0.0
    cv_tf = (2 * PB * args.steps * fl) / (cv_ms * 1e-3) / 1e12 if cv_ms else 0.0
    line = {
        "metric": "train_positions_per_s", "value": value, "unit": "positions/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": r["dev_ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"optimizer step (scripts/train.py:126-196: loss, {ACCUM} accumulated batches, clip 1.0, Adam), {name}, {ACCUM} x batch {B} per GPU, "
                               "tower convolutions fprop/dgrad/wgrad on tcgen05 kernels, fused BatchNorm+ReLU(+residual) kernels, stem conv / heads / loss / Adam in PyTorch",
                   "batch_per_gpu": B, "accumulate": ACCUM,
                   "parallelism": (f"DDP over {world} GPU(s): one fp32 gradient all-reduce per optimizer step (no_sync on the "
                                   "accumulating micro-batch), 25 MB buckets") if world > 1 else "single GPU",
                   "l2": "activations of one step (> 1 GB) exceed the 126 MB L2; no flush needed"},
        "cudnn_arm": ({"value": world * PB * args.steps / (results["cudnn"]["dev_ms"] * 1e-3), "unit": "positions/s",
                       "ms_per_step": results["cudnn"]["dev_ms"] / args.steps,
                       "note": "same step in plain PyTorch: cuDNN convolutions + torch BatchNorm, bf16 autocast, channels_first "
                               "as the reference runs it"} if "cudnn" in results else None),
        "cudnn_nhwc_arm": ({"value": world * PB * args.steps / (results["cudnn_nhwc"]["dev_ms"] * 1e-3), "unit": "positions/s",
                            "ms_per_step": results["cudnn_nhwc"]["dev_ms"] / args.steps,
                            "note": "same, parameters and activations in channels_last memory (cuDNN's preferred layout)"}
                           if "cudnn_nhwc" in results else None),
        "cuda_graph": {"ms_per_step": {a: results[a]["graph_ms"] for a in results},
                       "value": (world * PB / (r["graph_ms"] * 1e-3)) if isinstance(r["graph_ms"], float) else None,
                       "unit": "positions/s", "note": "the whole step captured as one CUDA graph and replayed (same kernels)"},
        "clocks": clk, "gpu_launches": r["launches"],
        "e2e": {"value": world * PB * args.steps / (r["e2e_ms"] * 1e-3), "unit": "positions/s",
                "h2d_bytes_per_step": B * (128 + 8 + 4), "d2h_bytes_per_step": 4,
                "api": "pinned host packed records -> device -> kv_encode -> TrainGraph step -> loss.item()"},
        "roofline": {"kernel": "conv3x3_wgrad_kernel (tcgen05 MN-major split-K)", "bound": "tensor", "achieved": wg_tf,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": wg_tf / peaks["bf16_sustained"],
                     "traffic": None, "peak_source": peaks["source"] + ", sustained bf16 figure",
                     "kernel_ms_per_step": wg_ms / args.steps, "kernel_share_of_step": wg_ms / r["dev_ms"],
                     "launches": wg_n,
                     "fprop_dgrad": {"kernel": "tower_umma2_kernel<halo> (one-layer launches)", "achieved": cv_tf, "unit": "TFLOP/s",
                                     "kernel_ms_per_step": cv_ms / args.steps, "launches": cv_n}},
        "kernels_ms_per_step": {k: v[0] / args.steps for k, v in r["prof"].items() if v[1]},
        "loss": r["loss"], "loss_cudnn_arm": results["cudnn"]["loss"] if "cudnn" in results else None,
    }
    return line
