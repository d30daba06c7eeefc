"""ConvNeXt block MLP at stage-0 / stage-1 size (B = 4096 cutouts): two GEMMs vs the fused kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from applecider_b200 import ops  # noqa: E402

for C, M in ((96, 4096 * 225), (192, 4096 * 49)):
    torch.manual_seed(0)
    y = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    res = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, device="cuda") * C ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device="cuda") * (4 * C) ** -0.5).to(torch.bfloat16)
    b1, b2, g = torch.zeros(4 * C, device="cuda"), torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")

    def timeit(f, n=10):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            o = f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, o

    t2, a = timeit(lambda: ops.gemm(ops.gemm(y, w1, b1, act=ops.ACT_GELU), w2, b2, res=res, gamma=g, res_mode=ops.RES_ADD))
    tf, b = timeit(lambda: ops.convnext_mlp(y, res, w1, b1, w2, b2, g))
    fl = 2.0 * M * C * 4 * C * 2
    print(f"C={C} M={M}: two GEMMs {t2:.3f} ms, fused {tf:.3f} ms ({fl / tf / 1e9:.0f} TFLOP/s, {M * C * 2 * 3 / tf / 1e6:.0f} GB/s of y+res+out); "
          f"max|diff| {(a.float() - b.float()).abs().max().item():.4f}")
