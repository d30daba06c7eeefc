"""One profiled fusion training step (for `ncu --profile-from-start off`): 2 warm steps, then cudaProfilerStart/Stop
around a single step.  python tools/train_probe.py [B] [workload: train|cnn_train]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import applecider_b200 as ab  # noqa: E402
from applecider_b200 import fn, synth  # noqa: E402
from applecider_b200.ddp import ddp_train_step  # noqa: E402
from applecider_b200.optim import fused_from_torch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
wl = sys.argv[2] if len(sys.argv) > 2 else "train"
if wl == "train":
    model = ab.AppleCider(ab.default_config(), hidden_dim=64, fusion="avg", compute_dtype="bf16")
else:
    cfg = ab.default_config()
    cfg["model"]["AstroMiNN"]["compute_dtype"] = "bf16"
    model = ab.AstroMiNN(cfg)
model.load_state_dict(synth.det_state_dict(model, 0), strict=True)
model = model.cuda().train()
topt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.01) if wl == "train" else model.this_optimizer
opt = fused_from_torch(topt, bf16_shadow=True)
x, pad, _ = synth.photometry_batch(B, seed=1337)
d = {"x": x.cuda(), "pad": pad.cuda(), "meta": synth.metadata(B, seed=1337).cuda(), "img": synth.cutouts(B, seed=1337).cuda(),
     "spec": synth.spectra(B, seed=1337, L=4096).cuda(), "tgt": torch.nn.functional.one_hot(synth.labels(B, seed=1337), 5).float().cuda()}


def loss_fn():
    out = model(d["x"], d["pad"], d["meta"], d["img"], d["spec"]) if wl == "train" else model((d["meta"], d["img"], d["tgt"]))
    return fn.soft_cross_entropy(out, d["tgt"])


for _ in range(2):
    ddp_train_step(opt.grads, loss_fn, opt)
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record()
t0 = time.perf_counter()
loss = ddp_train_step(opt.grads, loss_fn, opt)
cpu_ms = (time.perf_counter() - t0) * 1e3
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("step ms", e0.elapsed_time(e1), "cpu launch ms", cpu_ms, "loss", float(loss))
