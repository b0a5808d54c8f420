"""bench.py --workload learn: BASELINE.json configs[4] — the scripts/learn.py loop (train -> self-play -> extend,
scripts/learn.py:152-209) with a 20-block x 256-channel tower: every rank plays its shard of the games on the device
search engine and trains on its own records with gradients averaged over NCCL (DistributedDataParallel).
One step = one loop iteration (training on everything recorded so far, weight hand-over, one generation of self-play).
"""
from __future__ import annotations

import json
import os

GAMES = int(os.getenv("KV_BENCH_GAMES", "4096"))
SIMS = int(os.getenv("KV_BENCH_SIMS", "800"))
PLIES = int(os.getenv("KV_BENCH_LEARN_PLIES", "2"))       # plies per generation (bounded so a step takes ~10 s)


def run(args, rank, world, local_rank):
    from knightvision_b200.engine import Engine
    eng = Engine(local_rank)
    line = learn_record(args, eng, rank, world, local_rank, warm=max(args.warmup, 1), steps=args.steps, plies=PLIES)
    if rank == 0:
        print(json.dumps(line), flush=True)


def learn_record(args, eng, rank, world, local_rank, warm, steps, plies):
    """The learn-loop record (rank 0 returns the dict, the other ranks None)."""
    import torch
    import torch.distributed as dist
    from bench import Clocks
    from knightvision_b200 import learn as LR
    from knightvision_b200 import selfplay as SP
    from knightvision_b200.model import ChessNet

    class _A:
        pass
    a_ = _A()
    a_.steps, a_.warmup = steps, warm
    args = a_
    PLIES_ = plies
    SP._engines[local_rank] = eng
    dev = eng.device
    arch = dict(stem=256, tower=256, blocks=20, conv2=False)
    torch.manual_seed(0)
    net = ChessNet(**arch, max_batch=GAMES)
    cfg = LR.build_cfg(num_iterations=max(args.warmup, 1) + args.steps, device=dev, arch=arch)
    cfg.selfplay.num_games, cfg.selfplay.max_moves, cfg.selfplay.sims = GAMES * world, PLIES_, SIMS
    cfg.selfplay.cache_log2 = int(os.getenv("KV_BENCH_CACHE_LOG2", "22"))
    cfg.selfplay.decisive_filter = False      # keep every record (the reference's decisive-only filter would shrink the
                                              # training phase to a handful of positions at this ply cap)
    cfg.selfplay.random_start_plies = 40      # every game starts after k in [0,40) random legal plies (no cross-game sharing)
    cfg.train.epochs, cfg.train.batch_size = 1, 2048
    clocks = Clocks(local_rank)
    clocks.start()
    l0 = eng.launches
    net, data, hist = LR.reinforcement_loop(cfg, net=net)
    clk = clocks.stop()
    timed = hist[max(args.warmup, 1):]
    t = torch.tensor([sum(h["total_s"] for h in timed), sum(h["train_s"] for h in timed),
                      sum(h["selfplay_s"] for h in timed), sum(h["weights_s"] for h in timed)], dtype=torch.float64, device=dev)
    c = torch.tensor([float(sum(h["records"] for h in timed)), float(sum(h["train_positions"] for h in timed)),
                      float(sum(h["evals"] for h in timed))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    tot_s, train_s, sp_s, w_s = (float(x) for x in t)
    recs, trained, evals = (float(x) for x in c)
    sims = recs * SIMS
    line = {
        "metric": "selfplay_positions_per_s_in_learn_loop", "value": recs / tot_s, "unit": "positions/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * tot_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (f"learn loop (scripts/learn.py:152-209): per iteration train 1 epoch on all records so far "
                                f"(batch 2048, tcgen05 fprop/dgrad/wgrad + fused BN), then {GAMES} games per GPU x {PLIES_} plies "
                                f"x {SIMS} PUCT sims/move with the 20-block x 256-channel tower, random init; games start from random positions "
                                "(k in [0,40) random legal plies) so that generations do not share their searches"),
                   "games_per_gpu": GAMES, "sims_per_move": SIMS, "plies_per_generation": PLIES_,
                   "parallelism": f"games sharded by id over {world} GPU(s); DDP gradient all-reduce over NCCL",
                   "l2": "working sets (search pools, activations) exceed the 126 MB L2; no flush needed"},
        "phases": {"train_s_per_step": train_s / args.steps, "weights_s_per_step": w_s / args.steps,
                   "selfplay_s_per_step": sp_s / args.steps,
                   "selfplay_sims_per_s": sims / sp_s if sp_s else None,
                   "selfplay_positions_per_s": recs / sp_s if sp_s else None,
                   "train_positions_per_s": trained / train_s if train_s else None,
                   "evals_per_sim": evals / sims if sims else None},
        "history": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in h.items()} for h in hist],
        "clocks": clk, "gpu_launches": eng.launches - l0,
        "e2e": {"value": recs / tot_s, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "api": "knightvision_b200.learn.reinforcement_loop (records stay on the device in packed form; wall clock "
                       "with device synchronisation around every phase)"},
        "roofline": None, "cpu_baseline": None,
    }
    return line
