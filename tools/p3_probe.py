"""Time / profile the P3 spectrum resampling alone (50 k ragged spectra resident in HBM)."""
import sys

import torch

sys.path.insert(0, ".")
from applecider_b200 import preprocess as pp, synth  # noqa: E402

specs = synth.raw_spectra(2000, seed=2) * 25
wl, offs = pp.ragged([s[:, 0] for s in specs])
fx, _ = pp.ragged([s[:, 1] for s in specs])
grid = pp.wave_grid()
mx = int(max(len(s) for s in specs))
args = (wl, fx, offs, grid, mx)
for _ in range(3):
    pp.resample_spectra(*args)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    pp.resample_spectra(*args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"P3 {ms:.3f} ms / {len(specs)} spectra = {len(specs) / ms / 1e3:.2f} M spectra/s; mean length {wl.numel() / len(specs):.0f}, max {mx}")
