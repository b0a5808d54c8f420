import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knightvision_b200.engine import Engine
from knightvision_b200 import train_ops as T
eng = Engine(0)
n, C = 256, 256
x = torch.randn(n, C, 8, 8, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()
w = (torch.randn(C, C, 3, 3, device="cuda") * 0.02).requires_grad_()
b = torch.zeros(C, device="cuda", requires_grad=True)
bn = torch.nn.BatchNorm2d(C).cuda(); bn.train()

def f_conv():
    y = T.conv3x3_b200(x, w, b, eng)
    y.float().sum().backward()
def f_bn():
    y = T.bn_relu_b200(x, bn, eng, residual=x)
    y.float().sum().backward()
def f_bn_fwd():
    with torch.no_grad():
        eng.bn_relu_fwd(x.detach().permute(0, 2, 3, 1), bn.weight, bn.bias, bn.running_mean, bn.running_var, 0.1, 1e-5)
def f_cs():
    eng.channel_sum(x.detach().permute(0, 2, 3, 1))
for name, f in (("conv", f_conv), ("cs", f_cs), ("bn_fwd", f_bn_fwd), ("bn", f_bn)):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): f()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            f()
        g.replay(); torch.cuda.synchronize()
        print(name, "capture ok", flush=True)
    except Exception as e:
        print(name, "FAILED", str(e)[:200], "| last kv error:", eng._lib.kv_last_error(eng.ctx), flush=True)
        break
