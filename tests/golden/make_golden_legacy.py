"""Generate tests/golden/legacy_spectra.npz from the REAL reference's archived "variant B" spectra encoder
(_archive/notebooks/brew_cider.py:585-708, `build_spec_model`).  Run in the build container only:

    python tests/golden/make_golden_legacy.py

brew_cider.py is a notebook export with top-level training code, so only the source of `build_spec_model` is executed
(its text is taken by line range at generation time, nothing is copied into this repository).  Deterministic name-keyed weights (incl. BatchNorm running
statistics), eval mode, CPU fp32; inputs and outputs are stored, weights are regenerated from the key names by the tests.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from applecider_b200 import synth  # noqa: E402

REF = "/root/reference/_archive/notebooks/brew_cider.py"


def load_build_spec_model():
    """The file contains IPython magics (not parseable as a module): take the text of the one top-level function."""
    lines = open(REF).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("def build_spec_model("))
    end = next(i for i in range(start + 1, len(lines)) if lines[i] and not lines[i][0].isspace() and not lines[i].startswith("#"))
    ns = {"torch": torch, "nn": nn, "F": F}
    exec(compile("\n".join(lines[start:end]), REF, "exec"), ns)
    return ns["build_spec_model"]


def main():
    torch.set_num_threads(8)
    build = load_build_spec_model()
    out = {}
    for mode in ("all", "spectra"):
        cfg = {"mode": mode, "classes": list(range(5))}
        ref = build(cfg).eval()
        ref.load_state_dict(synth.det_state_dict(ref, 0))
        x = synth.spectra(3, seed=31, L=4096)
        with torch.no_grad():
            y = ref(x)
        out[f"x_{mode}"] = x.numpy()
        out[f"y_{mode}"] = y.numpy()
        out[f"keys_{mode}"] = np.array(sorted(ref.state_dict().keys()))
        print(mode, tuple(y.shape), float(y.abs().max()))
    path = os.path.join(HERE, "legacy_spectra.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path)/1024:.1f} KiB)")
    train_record(build)


def train_record(build):
    """legacy_train.npz: ONE training-mode forward + backward of the real `build_spec_model` (BatchNorm on batch statistics,
    running statistics updated in place; the two Dropout layers set to p = 0 so that the step is deterministic):
    loss, a selection of gradients, and the BatchNorm running statistics after the step."""
    cfg = {"mode": "spectra", "classes": list(range(5))}
    ref = build(cfg).train()
    ref.load_state_dict(synth.det_state_dict(ref, 0))
    for m in ref.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    x = synth.spectra(4, seed=33, L=4096)
    y = torch.tensor([0, 3, 1, 4])
    logits = ref(x)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    out = {"x": x.numpy(), "y": y.numpy(), "logits": logits.detach().numpy(), "loss": np.float32(loss.item())}
    sd = ref.state_dict()
    for i in range(1, 5):
        out[f"rm{i}"] = sd[f"stage{i}.0.norm.running_mean"].numpy()
        out[f"rv{i}"] = sd[f"stage{i}.0.norm.running_var"].numpy()
    grads = {n: p.grad for n, p in ref.named_parameters()}
    for n in ["stage1.0.convs.2.weight", "stage1.0.norm.weight", "stage1.0.norm.bias", "stage1.0.proj.weight", "stage2.0.convs.1.weight",
              "stage3.0.convs.0.bias", "stage3.0.norm.weight", "stage4.0.proj.weight", "stage5.0.convs.2.weight", "stage5.0.norm.weight",
              "class_model.4.weight", "fc.weight"]:
        g = grads[n]
        out["g_" + n] = g.reshape(g.shape[0], -1)[:, :256].numpy() if g.dim() > 1 else g.numpy()
    w0 = grads["class_model.0.weight"]
    out["g_class_model.0.weight_rows"] = w0[:4].numpy()
    path = os.path.join(HERE, "legacy_train.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path)/1024:.1f} KiB) loss={loss.item():.6f}")


if __name__ == "__main__":
    main()
