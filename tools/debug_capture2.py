import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_train as BT
from knightvision_b200 import learn as LR
from knightvision_b200.engine import Engine
from knightvision_b200.model import ChessNet
eng = Engine(0)
B = 256
lines = eng.random_positions(B, 40, 99)
moves = torch.randint(0, 4096, (B,), device="cuda"); rewards = torch.ones(B, device="cuda")
net = ChessNet(stem=256, tower=256, blocks=2, conv2=False, max_batch=2).cuda(); net.train()
graph = LR.TrainGraph(net, engine=eng)
opt = torch.optim.Adam(net.parameters(), lr=1e-3, capturable=True)
step = BT._step_fn(LR, net, opt, graph)
def fwd_only():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        p, v = graph(eng.encode(lines))
    return p.float().sum() + v.float().sum()
def fwd_bwd():
    fwd_only().backward()
    opt.zero_grad(set_to_none=True)
def full():
    step(eng.encode(lines), moves, rewards)
for mode in ("global", "thread_local"):
    for name, f in (("fwd", fwd_only), ("fwd_bwd", fwd_bwd), ("full", full)):
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): f()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode=mode):
                f()
            g.replay(); torch.cuda.synchronize()
            print(mode, name, "capture ok", flush=True)
        except Exception as e:
            print(mode, name, "FAILED", str(e)[:160].replace("\n", " "), "| kv:", eng._lib.kv_last_error(eng.ctx), flush=True)
            torch.cuda.synchronize()
