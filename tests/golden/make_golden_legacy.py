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


if __name__ == "__main__":
    main()
