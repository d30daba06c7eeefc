"""Time (and optionally profile under ncu) the depthwise 7x7 + LayerNorm kernel on the ConvNeXt stage shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from applecider_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for (H, C) in [(15, 96), (7, 192), (3, 384), (1, 768)]:
    x = torch.randn(B * H * H, C, device="cuda").to(torch.bfloat16)
    w = torch.randn(C, 1, 7, 7, device="cuda") * 0.1
    b = torch.randn(C, device="cuda") * 0.1
    g = torch.ones(C, device="cuda")
    be = torch.zeros(C, device="cuda")
    for _ in range(3):
        ops.dwconv7_ln(x, B, H, H, C, w, b, g, be, 1e-6)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.dwconv7_ln(x, B, H, H, C, w, b, g, be, 1e-6)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = 2 * x.numel() * 2 / 1e9
    print(f"dwconv7_ln B={B} {H}x{H}x{C}: {ms:.3f} ms  ({gb/ms*1e3:.0f} GB/s algorithmic, {B*H*H*C*49*2/ms/1e9:.1f} TFLOP/s fp32)")
