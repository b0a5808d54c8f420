"""A/B of the evaluator schedule at bench state: 4 096 games from random positions, 800 sims/move, cache on."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knightvision_b200.engine import Engine
from knightvision_b200.model import ChessNet
from knightvision_b200.selfplay import SelfPlay
eng = Engine(0)
torch.manual_seed(0)
sp = SelfPlay(ChessNet().eval(), 4096, eng.device, sims=800, max_plies=512, seed=42, engine=eng)
eng.mcts_enable_cache(24)
start = eng.random_positions(4096, 40, 1234)
for split in (0, 1, 0, 1):
    eng.mcts_set_eval_split(split)
    eng.mcts_cache_clear()
    eng.mcts_reset(start, 0)
    eng.mcts_run_move()
    eng.profile(True); eng.profile_read()
    torch.cuda.synchronize(); t0 = time.perf_counter(); s0 = eng.mcts_status()
    for _ in range(2):
        eng.mcts_run_move()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0; s1 = eng.mcts_status()
    prof = eng.profile_read(); eng.profile(False)
    sims = (s1["plies"] - s0["plies"]) * 800
    print("split", split, "sims/s", round(sims / dt), "evals/sim", round((s1["evals"] - s0["evals"]) / sims, 4),
          {k: round(v[0] / 2, 1) for k, v in prof.items() if v[1]}, flush=True)
