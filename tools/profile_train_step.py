"""torch.profiler breakdown of one native training step (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench_train as BT
from knightvision_b200 import learn as LR
from knightvision_b200.engine import Engine
from knightvision_b200.model import ChessNet

eng = Engine(0)
arch, name = BT._arch()
B = 2048
lines = eng.random_positions(B, 40, 99)
moves = torch.randint(0, 4096, (B,), device="cuda"); rewards = torch.ones(B, device="cuda")
torch.manual_seed(0)
net = ChessNet(**arch, max_batch=2).cuda(); net.train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
graph = LR.TrainGraph(net, engine=eng if os.getenv("ARM", "native") == "native" else None)
if os.getenv("ARM") == "cudnn_nhwc":
    net = net.to(memory_format=torch.channels_last)
step = BT._step_fn(LR, net, opt, graph)
for _ in range(3):
    step(eng.encode(lines), moves, rewards)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step(eng.encode(lines), moves, rewards)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
