"""Generate tests/golden/mpt.npz from the REAL reference MPTModel (run in the build container only).

    python tests/golden/make_golden_mpt.py

MPTModel.train_step (HyraxBaselineCLS.py:241-281) is run unmodified with dropout 0 (F.dropout on the time
embedding and the encoder dropouts are otherwise random), deterministic name-keyed weights and an SGD(lr=0)
optimizer so that the step leaves the weights alone; recorded: the mask its _mask_batch drew under
torch.manual_seed(MASK_SEED), the masked data, the loss and the (clipped) gradients left in .grad.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from applecider_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

WEIGHT_SEED = 0
MASK_SEED = 123


def main():
    torch.set_num_threads(8)
    cfg = ref_loader.default_config()
    cfg["model"]["HyraxBaselineCLS"]["dropout"] = 0.0
    R = ref_loader.ref_models()
    ref = R.photo.MPTModel(cfg).train()
    ref.load_state_dict(synth.det_state_dict(ref, WEIGHT_SEED))
    ref.optimizer = torch.optim.SGD(ref.parameters(), lr=0.0)
    x, pad, lens = synth.photometry_batch(12, seed=21)
    # the mask the reference draws (and the data after its in-place zeroing)
    torch.manual_seed(MASK_SEED)
    xm = x.clone()
    masked = ref._mask_batch(xm, pad)
    # the unmodified train_step on a fresh copy, same seed -> same draws
    torch.manual_seed(MASK_SEED)
    xs = x.clone()
    out = ref.train_step((xs, pad, None))
    assert torch.equal(xs, xm)
    grads = {n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.grad is not None}
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    keep = ["head_flux.weight", "head_band.weight", "head_dt.bias", "in_proj.weight", "time2vec.w", "cls_tok",
            "encoder.layers.0.self_attn.in_proj_weight", "encoder.layers.3.linear2.weight", "encoder.layers.1.norm1.weight"]
    arrs = {"x": x, "pad": pad, "x_masked": xm, "masked": masked, "loss": np.float32(out["loss"]), "clipped_grad_norm": total.float()}
    for k in keep:
        arrs["g_" + k.replace(".", "_")] = grads[k]
    path = os.path.join(HERE, "mpt.npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path)/1024:.1f} KiB) loss={out['loss']:.6f} masked={int(masked.sum())} |g|={float(total):.6f}")


if __name__ == "__main__":
    main()
