"""Schedule experiments on the self-play workload: whole-tower launch on/off x pipelined search on/off.
   python tools/bench_sched.py [games] [sims]   -> one JSON line per configuration"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from knightvision_b200.engine import Engine  # noqa: E402
from knightvision_b200.model import ChessNet  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
SIMS = int(sys.argv[2]) if len(sys.argv) > 2 else 800
CONFIGS = [c.split(":") for c in os.getenv("KV_SCHED_CONFIGS", "1:0,0:0,1:1").split(",")]

eng = Engine(0)
torch.manual_seed(0)
net = ChessNet().eval().attach(eng, max_batch=G)


def timed(n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0 = eng.mcts_status()
    a.record()
    for _ in range(n):
        eng.mcts_run_move()
    b.record()
    torch.cuda.synchronize()
    s1 = eng.mcts_status()
    return a.elapsed_time(b) / n, (s1["evals"] - s0["evals"]) / ((G - s0["done"]) * SIMS * n)


for fused, piped in CONFIGS:
    fused, piped = int(fused), int(piped)
    eng.mcts_create(G, SIMS, 512, seed=42, eval_mode=1)
    eng.net_set_tower_fused(int(fused))
    eng.mcts_set_pipeline(piped)
    eng.mcts_enable_cache(24)
    eng.mcts_reset(None, 0)
    for _ in range(3):
        eng.mcts_run_move()
    ms, eps = timed(2)
    roots = eng.mcts_roots().clone()
    eng.mcts_enable_cache(0)
    eng.mcts_reset(roots, 0)
    eng.mcts_run_move()
    ms_nc, _ = timed(1)
    print(json.dumps({"tower_fused": fused, "pipelined": piped, "ms_per_move": ms, "sims_per_s": G * SIMS / ms * 1e3,
                      "evals_per_sim": eps, "no_cache_ms": ms_nc, "no_cache_sims_per_s": G * SIMS / ms_nc * 1e3}), flush=True)
