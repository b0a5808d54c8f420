"""Quick GPU check of the training-side conv operators against torch fp32 (development aid; tests/test_gpu_train.py is the test)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from knightvision_b200.engine import Engine

eng = Engine(0)
torch.manual_seed(0)
for n, cin, cout in ((6, 256, 512), (101, 512, 512), (64, 256, 256)):
    x = (torch.randn(n, 8, 8, cin, device="cuda") * 0.5).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
    b = torch.randn(cout, device="cuda") * 0.1
    dy = (torch.randn(n, 8, 8, cout, device="cuda") * 0.5).to(torch.bfloat16)
    wb = w.to(torch.bfloat16).float()
    xf, dyf = x.float().permute(0, 3, 1, 2), dy.float().permute(0, 3, 1, 2)
    ref_y = F.conv2d(xf, wb, b, padding=1).permute(0, 2, 3, 1)
    ref_dx = torch.nn.grad.conv2d_input(xf.shape, wb, dyf, padding=1).permute(0, 2, 3, 1)
    ref_dw = torch.nn.grad.conv2d_weight(xf, w.shape, dyf, padding=1)
    y = eng.conv3x3_fprop(x, eng.conv3x3_pack(w), b)
    dx = eng.conv3x3_fprop(dy, eng.conv3x3_pack(w, flip_transpose=True))
    dw = eng.conv3x3_wgrad(x, dy)
    torch.cuda.synchronize()
    def rel(a, r):
        return ((a.float() - r).abs().max() / r.abs().max()).item()
    print(n, cin, cout, "fprop", rel(y, ref_y), "dgrad", rel(dx, ref_dx), "wgrad", rel(dw, ref_dw), flush=True)
# timing at a training-sized batch
n, cin, cout = 2048, 512, 512
x = (torch.randn(n, 8, 8, cin, device="cuda") * 0.5).to(torch.bfloat16)
dy = (torch.randn(n, 8, 8, cout, device="cuda") * 0.5).to(torch.bfloat16)
w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.02
wp = eng.conv3x3_pack(w)
for name, fn in (("fprop", lambda: eng.conv3x3_fprop(x, wp)), ("wgrad", lambda: eng.conv3x3_wgrad(x, dy))):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(name, "ms", ms, "TFLOP/s", 2 * n * 64 * 9 * cin * cout / ms / 1e9)
xc = x.permute(0, 3, 1, 2)  # channels_last view
wc = w.to(torch.bfloat16).to(memory_format=torch.channels_last)
dyc = dy.permute(0, 3, 1, 2)
for name, fn in (("cudnn fprop", lambda: F.conv2d(xc, wc, padding=1)),
                 ("cudnn wgrad", lambda: torch.nn.grad.conv2d_weight(xc, w.shape, dyc, padding=1))):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(name, "ms", ms, "TFLOP/s", 2 * n * 64 * 9 * cin * cout / ms / 1e9)
