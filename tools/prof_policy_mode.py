"""ncu driver for the reference-rule mode (sims = 1): 4 096 games, a few plies; capture mcts_eval_rootmix_kernel.
  ncu --set full --clock-control none --import-source on -k regex:mcts_eval_rootmix_kernel -s 3 -c 1 -o gpurun_out/X python tools/prof_policy_mode.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knightvision_b200.engine import Engine
from knightvision_b200.model import ChessNet
from knightvision_b200.selfplay import SelfPlay
eng = Engine(0)
torch.manual_seed(0)
sp = SelfPlay(ChessNet().eval(), 4096, eng.device, sims=1, max_plies=512, seed=42, engine=eng)
eng.mcts_reset(None, 0)
for _ in range(6):
    eng.mcts_run_move()
torch.cuda.synchronize()
