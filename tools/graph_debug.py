"""Debug helper: capture + replay the training step of one module in a CUDA graph and report (run one module per process)."""
import sys

import torch

sys.path.insert(0, ".")
import applecider_b200 as ab  # noqa: E402
from applecider_b200 import fn, synth  # noqa: E402
from applecider_b200.graph import GraphedTrainStep  # noqa: E402
from applecider_b200.optim import FusedAdam  # noqa: E402

which, B = sys.argv[1], int(sys.argv[2])
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
cfg = ab.default_config()
for k in cfg["model"]:
    cfg["model"][k]["compute_dtype"] = dtype
x, pad, lens = synth.photometry_batch(B, seed=1)
ntok = int(lens.sum()) + B
tgt5 = torch.nn.functional.one_hot(synth.labels(B, seed=1), 5).float().cuda()
if which == "photo":
    m = ab.HyraxBaselineCLS(cfg)
    inp = {"x": x.cuda(), "pad": pad.cuda(), "t": tgt5}
    f = lambda d: fn.soft_cross_entropy(m((d["x"], d["pad"], None), total_tokens=ntok), d["t"])  # noqa: E731
elif which == "spectra":
    m = ab.SpectraNet(cfg)
    t9 = torch.nn.functional.one_hot(torch.arange(B) % 9, 9).float().cuda()
    inp = {"s": synth.spectra(B, seed=1, L=4096).cuda(), "t": t9}
    f = lambda d: fn.soft_cross_entropy(m((d["s"], None, None)), d["t"])  # noqa: E731
elif which == "astrominn":
    m = ab.AstroMiNN(cfg)
    inp = {"m": synth.metadata(B, seed=1).cuda(), "i": synth.cutouts(B, seed=1).cuda(), "t": tgt5}
    f = lambda d: fn.soft_cross_entropy(m((d["m"], d["i"], None)), d["t"])  # noqa: E731
else:
    m = ab.AppleCider(cfg, hidden_dim=64, fusion="avg", compute_dtype=dtype)
    inp = {"x": x.cuda(), "pad": pad.cuda(), "m": synth.metadata(B, seed=1).cuda(), "i": synth.cutouts(B, seed=1).cuda(),
           "s": synth.spectra(B, seed=1, L=4096).cuda(), "t": tgt5}
    f = lambda d: fn.soft_cross_entropy(m(d["x"], d["pad"], d["m"], d["i"], d["s"], total_tokens=ntok), d["t"])  # noqa: E731
m.load_state_dict(synth.det_state_dict(m, 0))
m = m.cuda().train()
opt = FusedAdam([p for p in m.parameters()], lr=1e-3, bf16_shadow=(dtype == "bf16"))
g = GraphedTrainStep(opt.grads, f, opt, inp, warmup=2)
torch.cuda.synchronize()
print(which, "captured:", g.launches_per_replay, "launches", flush=True)
for i in range(5):
    l = g().item()
    torch.cuda.synchronize()
    print(which, "replay", i, "loss", l, flush=True)
print(which, "OK", flush=True)
