"""ncu driver for the rules kernels: movegen / make-move over 2^20 boards (reference playout positions, tiled), then one
counts-only perft depth 3 over 65 536 boards.  No other kernels of this library are launched, so `-k regex:` + `-s/-c`
select launches directly:
  ncu --set full --clock-control none --import-source on -k regex:"movegen_kernel|make_moves_kernel" -c 4 -o gpurun_out/X python tools/prof_rules.py
  ncu --set full --clock-control none --import-source on -k regex:perft_level_kernel -s 40 -c 6 -o gpurun_out/Y python tools/prof_rules.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from knightvision_b200.engine import Engine, lines_to_device  # noqa: E402

eng = Engine(0)
rows = np.load(os.path.join(bench.ROOT, "tests", "golden", "playouts.npz"))
ok = rows["counts"] > 0
base = rows["line_in"][ok]
n = 1 << 20
lines = lines_to_device(base[np.arange(n) % len(base)], eng.device)
mv = torch.from_numpy(rows["played"][ok][np.arange(n) % len(base)].view(np.int16)).to(eng.device)
for _ in range(2):
    moves, counts, flags = eng.movegen(lines)
work = lines.clone()
for _ in range(2):
    work.copy_(lines)
    eng.make_moves(work, mv)
torch.cuda.synchronize()
roots = lines_to_device(bench.perft_roots(bench.PERFT_BOARDS), eng.device)
out = eng.perft(roots, 3, chunk=-bench.PERFT_BOARDS)
torch.cuda.synchronize()
print("nodes", int(out[:, 0].sum()))
