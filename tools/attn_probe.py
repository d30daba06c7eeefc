"""Time the varlen attention kernels (packed tcgen05 / per-sequence tcgen05 / CUDA-core) on the synthetic light-curve lengths."""
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
plan = ops.attention_plan(cu, B, T)
print("tiles", int(plan[0][0]), "long", int(plan[0][1]), "tokens", T, "rows/tile", T / max(1, int(plan[0][0])))
outs = {}


def timeit(f, n=20):
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


ms, _ = timeit(lambda: ops.attention_plan(cu, B, T))
print("plan kernel ms", ms)
ms, outs["packed"] = timeit(lambda: ops.attention_varlen(qkv, cu, B, 8, 16, 258, plan=plan))
print("packed tcgen05 attention ms", ms, "->", (T * 384 * 2 + T * 128 * 2) / ms / 1e6, "GB/s of compulsory traffic")
ms, outs["tc"] = timeit(lambda: ops.attention_varlen(qkv, cu, B, 8, 16, 258))
print("per-sequence tcgen05 attention ms", ms)
ops.USE_TC_ATTENTION = False
ms, outs["cc"] = timeit(lambda: ops.attention_varlen(qkv, cu, B, 8, 16, 258))
print("cuda-core attention ms", ms)
print("max |packed - cuda-core| =", (outs["packed"].float() - outs["cc"].float()).abs().max().item(),
      " max |tc - cuda-core| =", (outs["tc"].float() - outs["cc"].float()).abs().max().item())
