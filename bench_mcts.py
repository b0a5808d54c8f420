"""bench.py's self-play workload (BASELINE.json configs[2] / configs[3]): 4 096 concurrent games per GPU, 800 PUCT
simulations per move, the reference policy/value net (random init, torch.manual_seed(0), BN folded, bf16 tensor-core
tower), all games from the initial position.  One step = one move for every game = 800 search waves.
Games shard over ranks by game id (no data-path collective); NCCL only broadcasts the weight blob before the
generation and gathers the per-rank record counts after it.
"""
from __future__ import annotations

import json
import os
import time

import numpy as np

GAMES_PER_GPU = int(os.getenv("KV_BENCH_GAMES", "4096"))
SIMS = int(os.getenv("KV_BENCH_SIMS", "800"))
MAX_PLIES = 512
INFLIGHT = int(os.getenv("KV_BENCH_INFLIGHT", "8"))         # phase E: G/K games x K simulations in flight (0/1 = skip)
POLICY_MODE_PLIES = int(os.getenv("KV_BENCH_POLICY_PLIES", "40"))   # phase F: plies of reference-rule self-play (0 = skip)
CACHE_LOG2 = int(os.getenv("KV_BENCH_CACHE_LOG2", "24"))   # evaluation cache: 2^24 x 640 B = 10.7 GB (0 = off)
CONV_FLOPS_PER_EVAL = 2.0 * 64 * 9 * (256 * 512 + 10 * 512 * 512)      # conv2 + 5 residual blocks (tcgen05 kernel)
NET_FLOPS_PER_EVAL = 2.0 * 1587872256                                   # whole net, SURVEY §8d


def _ncu_traffic():
    """Average DRAM bytes (read + write) per launch of the tower kernel from the committed ncu capture, or None."""
    import glob
    import re
    root = os.path.dirname(os.path.abspath(__file__))
    files = sorted(glob.glob(os.path.join(root, "profiles", "*ncu_tower_umma2*.txt")))
    if not files:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    txt = open(files[-1]).read()
    rd = [float(v) * unit[u] for v, u in re.findall(r"dram__bytes_read\.sum\s+([0-9.]+) (\w*byte)", txt)]
    wr = [float(v) * unit[u] for v, u in re.findall(r"dram__bytes_write\.sum\s+([0-9.]+) (\w*byte)", txt)]
    if not rd or len(rd) != len(wr):
        return None, None
    return (sum(rd) + sum(wr)) / len(rd), os.path.relpath(files[-1], root)


def _cpu_eval_rate(batch=256, iters=4):
    """fp32 CPU forward of the reference graph (torch CPU kernels, all threads): evals/s."""
    import torch
    from knightvision_b200.model import ChessNet, fp32_reference_forward
    torch.manual_seed(0)
    net = ChessNet().eval()
    x = torch.zeros(batch, 12, 8, 8)
    x[:, 0, 7, 4] = 1
    with torch.no_grad():
        fp32_reference_forward(net, x[:8])
        t0 = time.perf_counter()
        for _ in range(iters):
            fp32_reference_forward(net, x)
        dt = time.perf_counter() - t0
    return batch * iters / dt, torch.get_num_threads()


def _tree_worker(args):
    from knightvision_b200 import layout as L
    from oracle import kv_oracle as O
    gid, sims, n_moves = args
    cfg = O.mcts_cfg(sims, temp_plies=30, max_plies=n_moves, seed=42)
    m, lines, res = O.selfplay_game(cfg, L.start_line(), game_id=gid)
    return len(m) * sims


def _cpu_tree_rate(procs, sims, games_per_proc=4, n_moves=10):
    """Sequential PUCT oracle (oracle/kv_oracle.c) with the hash evaluator on `procs` processes: sims/s."""
    import multiprocessing as mp
    from oracle import kv_oracle as O
    O.build()
    ctx = mp.get_context("fork")
    jobs = [(g, sims, n_moves) for g in range(procs * games_per_proc)]
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        tot = sum(pool.map(_tree_worker, jobs))
    return tot / (time.perf_counter() - t0), len(jobs), n_moves


def cpu_baseline(sims):
    procs = os.cpu_count() or 1
    tree, n_games, n_moves = _cpu_tree_rate(procs, sims)
    ev, threads = _cpu_eval_rate()
    combined = 1.0 / (1.0 / tree + 1.0 / ev)
    return {"value": combined, "unit": "sims/s", "cores": procs, "kind": "port",
            "sample": (f"{n_games} games x {n_moves} moves x {sims} sims of the sequential PUCT oracle (oracle/kv_oracle.c, hash "
                       f"evaluator) on {procs} processes = {tree:.0f} sims/s tree-only; reference net graph fp32 on torch "
                       f"CPU kernels, batch 256, {threads} threads = {ev:.1f} evals/s; one eval per sim => combined"),
            "tree_sims_per_s": tree, "net_evals_per_s": ev}


def run_reference(args):
    vals = []
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        vals.append(cpu_baseline(SIMS))
    ms = 1e3 * (time.perf_counter() - t0) / max(1, args.steps)
    value = float(np.mean([v["value"] for v in vals]))
    base = vals[-1]
    base["value"] = value
    line = {"impl": "reference", "metric": "mcts_sims_per_s", "value": value, "unit": "sims/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"self-play, {SIMS} PUCT sims/move, reference net (random init), initial position",
                       "sims_per_move": SIMS},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from bench import Clocks, measured_peaks
    from knightvision_b200 import layout as L
    from knightvision_b200.engine import Engine, lines_to_device
    from knightvision_b200.model import ChessNet
    from knightvision_b200.selfplay import SelfPlay

    G = GAMES_PER_GPU
    eng = Engine(local_rank)
    dev = eng.device
    torch.manual_seed(0)
    net = ChessNet().eval()
    sp = SelfPlay(net, G, dev, sims=SIMS, max_plies=MAX_PLIES, seed=42, engine=eng)
    if world > 1:
        # generation start: rank 0's fp32 weight blob -> every rank's staging buffer over NCCL, then fold on device
        blob = eng.net_blob_tensor()
        if rank == 0:
            blob.copy_(net.weight_blob().to(dev))
        dist.broadcast(blob, src=0)
        eng.net_commit()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_moves(n_moves, profile=False):
        """n_moves moves for every game; returns (device ms, status before, status after, per-kernel profile, launches)."""
        s0 = eng.mcts_status()
        if profile:
            eng.profile(True)
            eng.profile_read()
        l0 = eng.launches
        barrier()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(n_moves):
            eng.mcts_run_move()
        b_.record()
        barrier()
        prof_ = eng.profile_read() if profile else None
        if profile:
            eng.profile(False)
        return a_.elapsed_time(b_), s0, eng.mcts_status(), prof_, eng.launches - l0

    warm = max(args.warmup, 3)
    clocks = Clocks(local_rank)
    # ---- A: BASELINE config 3 — all games from the initial position, evaluation cache on ------------------------
    eng.mcts_enable_cache(CACHE_LOG2)
    eng.mcts_reset(None, game_id_base=rank * G)
    for _ in range(warm):
        eng.mcts_run_move()
    torch.cuda.synchronize()
    snapshot_h = eng.mcts_roots().cpu().pin_memory()      # positions at the start of the timed region (for B and C)
    clocks.start()
    dev_ms, st0, st1, prof, launches = timed_moves(args.steps, profile=True)
    live = G - st0["done"]
    sims_done = live * SIMS * args.steps          # every live game runs SIMS simulations per move
    evals = st1["evals"] - st0["evals"]
    hits = st1["cache_hits"] - st0["cache_hits"]
    positions = st1["plies"] - st0["plies"]

    # ---- A2: the same moves with the pipelined schedule (two game groups on two streams, opt-in) ---------------------------
    alt = None
    if os.getenv("KV_BENCH_COMPARE", "1") != "0" and G >= 2048:
        eng.mcts_set_pipeline(1)
        eng.mcts_cache_clear()
        eng.mcts_reset(None, game_id_base=rank * G)
        for _ in range(warm):
            eng.mcts_run_move()
        s_ms, s0_, s1_, _, _ = timed_moves(args.steps)
        alt = {"ms_per_step": s_ms / args.steps, "sims": (G - s0_["done"]) * SIMS * args.steps,
               "evals": s1_["evals"] - s0_["evals"]}
        eng.mcts_set_pipeline(-1)

    # ---- B: e2e — the same positions through the public API with HOST buffers: pinned host lines -> device, cold
    # evaluation cache (a new generation means new weights), the same number of moves, records back on the host as the
    # reference's tuples.  Everything, copies included, is inside the timed region.
    e2e_steps = args.steps
    barrier()
    t0 = time.perf_counter()
    eng.mcts_cache_clear()
    d = snapshot_h.to(dev, non_blocking=True)
    eng.mcts_reset(d, game_id_base=rank * G)
    for _ in range(e2e_steps):
        eng.mcts_run_move()
    recs = sp.records()                            # D2H: float planes + move + reward
    rec = len(recs)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()

    # ---- C: the same positions with the evaluation cache OFF (one network evaluation per simulation) -----------------
    nocache_ms = None
    if CACHE_LOG2:
        eng.mcts_enable_cache(0)
        eng.mcts_reset(snapshot_h.to(dev), game_id_base=rank * G)
        eng.mcts_run_move()
        nocache_ms = timed_moves(1)[0]

    # ---- D: "random positions" variant: k in [0,40) random legal plies per game, seed 1234, cache on ----
    eng.mcts_enable_cache(CACHE_LOG2)
    eng.mcts_reset(eng.random_positions(G, 40, 1234 + rank), game_id_base=rank * G)
    for _ in range(warm):
        eng.mcts_run_move()
    r_ms, r0, r1, _, _ = timed_moves(max(1, min(args.steps, 2)))
    r_moves = max(1, min(args.steps, 2))
    r_sims = (G - r0["done"]) * SIMS * r_moves
    r_evals = r1["evals"] - r0["evals"]

    # ---- E: few games, K simulations in flight per game with virtual loss (G/K games x K = the same leaves per wave) ----
    K = INFLIGHT
    vl = None
    if K > 1 and G % K == 0:
        Gk = G // K
        eng.mcts_create(Gk, SIMS, MAX_PLIES, seed=42, eval_mode=1, inflight=K)
        eng.mcts_enable_cache(CACHE_LOG2)
        eng.mcts_reset(None, game_id_base=rank * Gk)
        for _ in range(warm):
            eng.mcts_run_move()
        w0 = eng.mcts_waves()
        v_moves = max(1, min(args.steps, 3))
        v_ms, v0, v1, _, _ = timed_moves(v_moves)
        vl = {"games_per_gpu": Gk, "inflight": K, "ms_per_step": v_ms / v_moves,
              "sims": (Gk - v0["done"]) * SIMS * v_moves, "evals": v1["evals"] - v0["evals"],
              "waves_per_move": (eng.mcts_waves() - w0) / v_moves}

    # ---- F: the reference's own move rule (scripts/self_play.py:147-167): no search, one network evaluation per ply,
    # the move sampled from softmax(policy) + Dirichlet noise over the legal moves (sims = 1).  positions/s here is
    # directly comparable with the reference's self-play loop (BASELINE configs[0]).
    pm = None
    if POLICY_MODE_PLIES > 0:
        eng.mcts_create(G, 1, MAX_PLIES, seed=42, eval_mode=1)
        eng.mcts_enable_cache(0)
        eng.mcts_reset(None, game_id_base=rank * G)
        for _ in range(warm):
            eng.mcts_run_move()
        p_ms, p0, p1, _, _ = timed_moves(POLICY_MODE_PLIES)
        pm = {"ms": p_ms, "positions": p1["plies"] - p0["plies"], "evals": p1["evals"] - p0["evals"]}

    t = torch.tensor([dev_ms, e2e_s * 1e3, nocache_ms or 0.0, r_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(sims_done), float(evals), float(positions), float(rec), float(hits), float(r_sims),
                        float(r_evals)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)     # record counts gathered to every rank
    dev_ms, e2e_ms, nocache_ms, r_ms = float(t[0]), float(t[1]), float(t[2]) or None, float(t[3])
    sims_all, evals_all, pos_all, rec_all, hits_all, r_sims_all, r_evals_all = (float(x) for x in cnt)
    if rank != 0:
        return
    peaks = measured_peaks()
    conv_ms, conv_n = prof["net_conv"]
    traffic, traffic_src = _ncu_traffic()
    boards_per_launch = G
    achieved = (evals * CONV_FLOPS_PER_EVAL) / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    value = sims_all / (dev_ms * 1e-3)
    e2e_sims = world * G * SIMS * e2e_steps
    line = {
        "metric": "mcts_sims_per_s", "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (f"self-play: {G} concurrent games per GPU, {SIMS} PUCT sims/move, reference ChessNet "
                                "(random init seed 0, BN folded, bf16 tower / fp32 accumulate), all games from the initial "
                                "position; one step = one move for every game"),
                   "games_per_gpu": G, "sims_per_move": SIMS, "parallelism": f"games sharded by id over {world} GPU(s)",
                   "l2": "working set (node/edge pools + activations, > 3 GB) exceeds the 126 MB L2; no flush needed"},
        "positions_per_s": pos_all / (dev_ms * 1e-3), "net_evals_per_s": evals_all / (dev_ms * 1e-3),
        "evals_per_sim": evals_all / sims_all if sims_all else None,
        "eval_cache": {"log2_slots": CACHE_LOG2, "bytes": (640 << CACHE_LOG2) if CACHE_LOG2 else 0,
                       "served_per_sim": hits_all / sims_all if sims_all else None,
                       "note": "keyed by the 12 bitboards (the net's whole input); search results are bit-identical "
                               "with the cache on or off (tests/test_gpu_mcts.py::test_eval_cache_is_transparent)"},
        "schedule": {"tower": ("one dependency-scheduled launch per evaluation batch (tower_umma2_kernel<halo>: depth-first "
                               "chunks, activation tiles fetched 3x instead of 9x per channel block)"),
                     "pipelined": False,
                     "pipelined_alt": ({"value": world * alt["sims"] / (alt["ms_per_step"] * args.steps * 1e-3),
                                        "unit": "sims/s", "ms_per_step": alt["ms_per_step"],
                                        "evals_per_sim": alt["evals"] / alt["sims"] if alt["sims"] else None,
                                        "note": "rank 0's figures x world: same moves, same cold cache and warm-up, with "
                                                "kv_mcts_set_pipeline(1): two game groups whose waves alternate on two "
                                                "streams (tree kernels of one group under the other group's tower); "
                                                "bit-identical results, opt-in because the step is power-bound"}
                                       if alt else None)},
        "no_cache": ({"value": world * G * SIMS / (nocache_ms * 1e-3), "unit": "sims/s", "ms_per_step": nocache_ms,
                      "evals_per_sim": 1.0, "note": "same positions, evaluation cache disabled"} if nocache_ms else None),
        "random_positions": {"value": r_sims_all / (r_ms * 1e-3), "unit": "sims/s",
                             "evals_per_sim": r_evals_all / r_sims_all if r_sims_all else None,
                             "ms_per_step": r_ms / r_moves,
                             "note": "every game starts after k in [0,40) uniformly random legal plies (seed 1234), "
                                     "cache on, same warm-up"},
        "virtual_loss": ({"value": world * vl["sims"] / (vl["ms_per_step"] * 1e-3 * max(1, min(args.steps, 3))), "unit": "sims/s",
                          "games_per_gpu": vl["games_per_gpu"], "inflight": vl["inflight"],
                          "ms_per_step": vl["ms_per_step"], "evals_per_sim": vl["evals"] / vl["sims"] if vl["sims"] else None,
                          "waves_per_move": vl["waves_per_move"],
                          "note": "rank 0's figures x world: G/K games with K simulations in flight per game and wave "
                                  "(virtual loss), initial position, cache on"} if vl else None),
        "policy_mode": ({"value": world * pm["positions"] / (pm["ms"] * 1e-3), "unit": "positions/s", "plies": POLICY_MODE_PLIES,
                         "evals_per_position": pm["evals"] / pm["positions"] if pm["positions"] else None,
                         "note": "rank 0's figures x world: KV_SIMS=1, the reference's own rule (no search: one network evaluation per "
                                 "ply, move sampled from the noisy policy over the legal moves); SURVEY section 6 measured 48.6 "
                                 "positions/s for the unmodified reference on 8 CPU threads"} if pm else None),
        "clocks": clk, "gpu_launches": launches,
        "e2e": {"value": e2e_sims / (e2e_ms * 1e-3), "unit": "sims/s", "h2d_bytes_per_step": G * 128 // e2e_steps,
                "d2h_bytes_per_step": G * (12 * 64 * 4 + 8), "records_returned": rec_all, "steps": e2e_steps,
                "api": ("SelfPlay public API on the timed region's own start positions: pinned host lines -> kv_mcts_reset -> "
                        "steps x kv_mcts_run_move -> records() as the reference's (planes, move, reward) tuples; "
                        "evaluation cache cleared first (new generation)")},
        "roofline": {"kernel": ("tower_umma2_kernel (tcgen05 cta_group::2 implicit GEMM, the 11 tower convolutions of one "
                                "evaluation batch in one dependency-scheduled launch)"), "bound": "tensor", "achieved": achieved,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                     "traffic": traffic, "traffic_source": (f"{traffic_src}: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                                             f"launch over {boards_per_launch} boards (a wave of the search "
                                                             "evaluates fewer: evals_per_sim x games).  If every layer went through "
                                                             "HBM the launch would move 7.1 GB (2.8 in + 2.95 out + 1.34 residual + "
                                                             "0.05 weights); depth-first chunks keep a chunk's activations in L2")
                     if traffic else None,
                     "peak_source": peaks["source"] + ", sustained bf16 figure",
                     "kernel_ms_per_step": conv_ms / args.steps, "kernel_share_of_step": conv_ms / dev_ms,
                     "flops_per_eval_in_kernel": CONV_FLOPS_PER_EVAL, "launches": conv_n},
        "kernels_ms_per_step": {k: v[0] / args.steps for k, v in prof.items() if v[1]},
        "cpu_baseline": cpu_baseline(SIMS),
    }
    print(json.dumps(line), flush=True)
