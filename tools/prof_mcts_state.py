"""ncu driver for the tree kernels AT BENCH STATE: 4 096 games, 800 sims/move, evaluation cache on, every game from its
own random position (k in [0,40) random legal plies: 0.92 network evaluations per simulation, like the bench's timed
window — from the initial position the first moves of all games share one search and the cache serves everything);
460 waves into the search, so the captured launches work on trees 450 simulations deep.
  ncu --set full --clock-control none --import-source on -k regex:"mcts_select_kernel|mcts_eval_net_kernel|mcts_late_net_kernel|stem_kernel" \
      -s 1800 -c 8 -o gpurun_out/X python tools/prof_mcts_state.py        (4 matching launches per wave: 450 waves skipped)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from knightvision_b200.engine import Engine  # noqa: E402
from knightvision_b200.model import ChessNet  # noqa: E402
from knightvision_b200.selfplay import SelfPlay  # noqa: E402

eng = Engine(0)
torch.manual_seed(0)
sp = SelfPlay(ChessNet().eval(), 4096, eng.device, sims=800, max_plies=512, seed=42, engine=eng)
eng.mcts_enable_cache(24)
eng.mcts_reset(eng.random_positions(4096, 40, 1234), 0)
eng.mcts_run_sims(460)
torch.cuda.synchronize()
print(eng.mcts_status())
