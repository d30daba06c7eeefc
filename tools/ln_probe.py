"""Time LayerNorm(+GELU) forward on the SpectraNet / transformer shapes; checks small-vs-large row-count consistency."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from applecider_b200 import ops  # noqa: E402

for rows, C in [(4096 * 1024, 384), (4096 * 256, 768), (237482, 128), (4096 * 64, 1536)]:
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    w = torch.rand(C, device="cuda") + 0.5
    b = torch.randn(C, device="cuda") * 0.1
    for _ in range(3):
        y = ops.layernorm(x, w, b, 1e-5, post_act=ops.ACT_GELU)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = ops.layernorm(x, w, b, 1e-5, post_act=ops.ACT_GELU)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    small = ops.layernorm(x[:1000].contiguous(), w, b, 1e-5, post_act=ops.ACT_GELU)
    print(f"LN+GELU rows={rows} C={C}: {ms:.3f} ms = {2*rows*C*2/ms/1e6:.0f} GB/s; small/large identical: {torch.equal(small, y[:1000])}")
