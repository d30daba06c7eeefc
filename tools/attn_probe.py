"""Time the varlen attention kernels (tcgen05 vs CUDA-core) on the synthetic light-curve length distribution."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from applecider_b200 import ops, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x, pad, lens = synth.photometry_batch(B, seed=1337)
cu, src = ops.photo_compact(pad.cuda())
T = int(cu[-1])
qkv = torch.randn(T, 384, device="cuda").to(torch.bfloat16)
outs = {}
for tc in (True, False):
    ops.USE_TC_ATTENTION = tc
    for _ in range(3):
        o = ops.attention_varlen(qkv, cu, B, 8, 16, 258)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = ops.attention_varlen(qkv, cu, B, 8, 16, 258)
    e1.record()
    torch.cuda.synchronize()
    outs[tc] = o.float()
    print("tcgen05" if tc else "cuda-core", "attention ms", e0.elapsed_time(e1) / 10, "tokens", T)
print("max |tc - cuda-core| =", (outs[True] - outs[False]).abs().max().item())
