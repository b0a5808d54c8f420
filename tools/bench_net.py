"""Times the network forward at a given batch with per-kernel CUDA-event breakdown (not the headline bench)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from knightvision_b200.engine import Engine, lines_to_device
from knightvision_b200.model import ChessNet
from knightvision_b200 import layout as L

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=4096); ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--tower", type=int, default=512); ap.add_argument("--blocks", type=int, default=5); ap.add_argument("--no-conv2", action="store_true"); ap.add_argument("--cta-group", type=int, default=2)
a = ap.parse_args()
eng = Engine(0)
torch.manual_seed(0)
net = ChessNet(tower=a.tower, blocks=a.blocks, conv2=not a.no_conv2, stem=256).eval().attach(eng, max_batch=a.batch)
eng.net_set_conv_mode(a.cta_group)
clusters4 = eng.net_tower_clusters4()
lines = lines_to_device(np.stack([L.start_line()] * a.batch), eng.device)
for _ in range(3):
    eng.net_forward(lines, want_policy=False)
torch.cuda.synchronize()
eng.profile(True); eng.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    eng.net_forward(lines, want_policy=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
prof = eng.profile_read()
macs = {(512, 5, True): 1587872256}.get((a.tower, a.blocks, not a.no_conv2), None)
flops = 2 * macs if macs else None
out = {"clusters4": clusters4, "cta_group": a.cta_group, "batch": a.batch, "ms_per_forward": ms, "evals_per_s": a.batch / ms * 1e3,
       "tflops": (a.batch * flops / ms / 1e9) if flops else None,
       "kernels_ms": {k: v[0] / a.iters for k, v in prof.items() if v[1]}}
print(json.dumps(out))
