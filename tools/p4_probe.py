"""Time / profile the P4 cutout normalisation alone (32 768 cutouts resident in HBM)."""
import sys

import torch

sys.path.insert(0, ".")
from applecider_b200 import preprocess as pp, synth  # noqa: E402

img = synth.cutouts(4096, seed=3, normalise=False).cuda().repeat(8, 1, 1, 1)
for _ in range(3):
    pp.normalize_cutouts(img, "median")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    pp.normalize_cutouts(img, "median")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"P4 {ms:.3f} ms / 32768 cutouts = {32768 / ms / 1e3:.2f} M cutouts/s, {img.numel() * 8 / ms / 1e6:.0f} GB/s")
