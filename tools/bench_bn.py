"""Times the fused BatchNorm operators alone (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from knightvision_b200.engine import Engine
eng = Engine(0)
for n, C in ((2048, 256), (2048, 512)):
    z = torch.randn(n, 8, 8, C, device="cuda").to(torch.bfloat16)
    res = torch.randn(n, 8, 8, C, device="cuda").to(torch.bfloat16)
    dy = torch.randn(n, 8, 8, C, device="cuda").to(torch.bfloat16)
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    rm = torch.zeros(C, device="cuda"); rv = torch.ones(C, device="cuda")
    y, mean, rstd = eng.bn_relu_fwd(z, g, b, rm, rv, 0.1, 1e-5, residual=res)
    def t(fn, it=20):
        """GPU time per call: `it` calls captured in one CUDA graph (no host launch gaps), replayed 10 times."""
        for _ in range(3): fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(it): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (10 * it) * 1e3
    mb = z.numel() * 2 / 1e6
    print(n, C, f"tensor {mb:.0f} MB",
          "fwd %.1f us" % t(lambda: eng.bn_relu_fwd(z, g, b, rm, rv, 0.1, 1e-5)),
          "fwd+res %.1f us" % t(lambda: eng.bn_relu_fwd(z, g, b, rm, rv, 0.1, 1e-5, residual=res)),
          "bwd %.1f us" % t(lambda: eng.bn_relu_bwd(dy, y, z, g, mean, rstd)),
          "bwd+dres %.1f us" % t(lambda: eng.bn_relu_bwd(dy, y, z, g, mean, rstd, want_dres=True)),
          "colsum %.1f us" % t(lambda: eng.channel_sum(dy)), flush=True)
