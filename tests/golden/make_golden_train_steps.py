"""Generate tests/golden/train_steps.npz and tests/golden/refinit.npz from the REAL reference (build container only).

    python tests/golden/make_golden_train_steps.py

1. train_steps.npz -- two consecutive UNMODIFIED ``train_step`` calls of the reference's
   * ``AstroMiNN``  (models/astrominn.py:308-326: CrossEntropyLoss on soft targets, its own 11-group AdamW, :151-218),
   * ``SpectraNet`` (models/spectranet.py:172-184: optimizer / criterion injected the way Hyrax does -- here
     Adam(lr 1e-3, weight_decay 0.01) as in brew_cider.py:1211 and CrossEntropyLoss),
   in eval() mode (dropout off, so the step is deterministic), deterministic name-keyed weights.  Recorded per parameter
   tensor: the first 48 elements of (weights after two steps - weights before), the L2 norm of that update, and the first
   48 elements of the first-step gradient (to tell solid updates from sign-like ones: Adam's first update is
   lr * g / (|g| + eps)); plus the returned losses.  28 M-parameter checkpoints never need to be stored.
2. refinit.npz -- the reference modules constructed under ``torch.manual_seed(INIT_SEED)`` (the reference's OWN random
   init: randn Time2Vec, trunc-normal 0.02 ConvNeXt with layer scale 1e-6, default nn.Linear / Conv1d init):
   fingerprints of the initial state_dict (checked here to be bit-identical to oracle/models.py constructed under the same
   seed, which is how the GPU test regenerates the weights) and the reference's fp32 logits on a small batch.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from applecider_b200 import synth  # noqa: E402
from oracle import models as om  # noqa: E402
from oracle import ref_loader  # noqa: E402

HEAD = 48
INIT_SEED = 0


def _record(prefix, model, before, grads1, arrs):
    for n, p in model.named_parameters():
        d = (p.detach() - before[n]).flatten()
        arrs[f"{prefix}/delta/{n}"] = d[:HEAD].clone()
        arrs[f"{prefix}/dnorm/{n}"] = d.double().norm().float()
        g = grads1.get(n)
        arrs[f"{prefix}/g1/{n}"] = (g.flatten()[:HEAD].clone() if g is not None else torch.zeros(min(HEAD, d.numel())))


def astrominn_steps(R, cfg, arrs):
    ref = R.astrominn.AstroMiNN(cfg).eval()
    ref.load_state_dict(synth.det_state_dict(ref, 0))
    B = 6
    meta, img = synth.metadata(B, seed=301, missing_frac=0.0), synth.cutouts(B, seed=301)
    tgt = torch.nn.functional.one_hot(synth.labels(B, seed=301), 5).float()
    before = {n: p.detach().clone() for n, p in ref.named_parameters()}
    out1 = ref.train_step((meta, img, tgt))
    grads1 = {n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.grad is not None}
    out2 = ref.train_step((meta, img, tgt))
    arrs.update({"astrominn/meta": meta, "astrominn/img": img, "astrominn/tgt": tgt,
                 "astrominn/loss1": np.float32(out1["loss"]), "astrominn/loss2": np.float32(out2["loss"])})
    _record("astrominn", ref, before, grads1, arrs)
    print(f"AstroMiNN: running-mean losses {out1['loss']:.6f} {out2['loss']:.6f}")


def spectranet_steps(R, cfg, arrs):
    ref = R.spectra.SpectraNet(cfg).eval()
    ref.load_state_dict(synth.det_state_dict(ref, 0))
    ref.optimizer = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=0.01)
    ref.criterion = torch.nn.CrossEntropyLoss()
    B, L = 3, 1024
    x = synth.spectra(B, seed=302, L=L)
    labels = torch.tensor([1, 7, 4])
    before = {n: p.detach().clone() for n, p in ref.named_parameters()}
    out1 = ref.train_step((x, labels, None))
    grads1 = {n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.grad is not None}
    out2 = ref.train_step((x, labels, None))
    arrs.update({"spectranet/x": x, "spectranet/labels": labels, "spectranet/loss1": np.float32(out1["loss"]),
                 "spectranet/loss2": np.float32(out2["loss"])})
    _record("spectranet", ref, before, grads1, arrs)
    print(f"SpectraNet: losses {out1['loss']:.6f} {out2['loss']:.6f}")


def _fingerprint(sd):
    return np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in sd.values() if v.dtype.is_floating_point], np.float64)


def refinit(R, cfg, arrs):
    B = 4
    x, pad, _ = synth.photometry_batch(B, seed=303, L=64)
    meta, img, sp = synth.metadata(B, seed=303, missing_frac=0.0), synth.cutouts(B, seed=303), synth.spectra(B, seed=303, L=1024)
    arrs.update({"in/x": x, "in/pad": pad, "in/meta": meta, "in/img": img, "in/spec": sp})
    for name, ref_cls, batch in [("HyraxBaselineCLS", R.photo.HyraxBaselineCLS, (x, pad, None)),
                                 ("SpectraNet", R.spectra.SpectraNet, (sp, None, None)),
                                 ("AstroMiNN", R.astrominn.AstroMiNN, (meta, img, None))]:
        torch.manual_seed(INIT_SEED)
        ref = ref_cls(cfg).eval()
        torch.manual_seed(INIT_SEED)
        port = getattr(om, name)(om.default_config()).eval()
        rs, ps = ref.state_dict(), port.state_dict()
        assert list(rs) == list(ps)
        for k in rs:
            assert torch.equal(rs[k], ps[k]), f"{name}.{k}: the oracle's seeded init differs from the reference's"
        with torch.no_grad():
            out = ref(batch)
        arrs[f"fp/{name}"] = _fingerprint(rs)
        arrs[f"logits/{name}"] = out
        print(f"{name}: seeded init identical to the oracle port ({len(rs)} tensors); |logits|max {out.abs().max():.4f}")


def main():
    torch.set_num_threads(8)
    cfg = ref_loader.default_config()
    R = ref_loader.ref_models()
    arrs = {}
    astrominn_steps(R, cfg, arrs)
    spectranet_steps(R, cfg, arrs)
    path = os.path.join(HERE, "train_steps.npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path)/1024:.1f} KiB)")
    arrs = {}
    refinit(R, cfg, arrs)
    path = os.path.join(HERE, "refinit.npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path)/1024:.1f} KiB)")


if __name__ == "__main__":
    main()
