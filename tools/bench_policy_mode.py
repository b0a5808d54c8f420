"""sims = 1 (the reference's own move rule) on 4 096 games: positions/s, with the per-kernel split."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knightvision_b200.engine import Engine
from knightvision_b200.model import ChessNet
from knightvision_b200.selfplay import SelfPlay
eng = Engine(0)
torch.manual_seed(0)
sp = SelfPlay(ChessNet().eval(), 4096, eng.device, sims=1, max_plies=512, seed=42, engine=eng)
for mix in (1, 0):
    eng.mcts_set_root_mix(mix)
    eng.mcts_reset(None, 0)
    for _ in range(5):
        eng.mcts_run_move()
    eng.profile(True); eng.profile_read()
    torch.cuda.synchronize(); t0 = time.perf_counter(); s0 = eng.mcts_status()
    for _ in range(40):
        eng.mcts_run_move()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0; s1 = eng.mcts_status()
    prof = eng.profile_read(); eng.profile(False)
    print("root_mix", mix, "positions/s", (s1["plies"] - s0["plies"]) / dt, "done", s1["done"],
          {k: round(v[0] / 40, 3) for k, v in prof.items() if v[1]})
