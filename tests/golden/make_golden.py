"""Generate tests/golden/*.npz from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py

The reference modules are imported unmodified from /root/reference/src through
oracle/ref_loader.py (hyrax/astropy stubbed, timm ConvNeXt-T restated), loaded with
name-keyed deterministic weights (applecider_b200.synth.det_state_dict — the GPU
tests regenerate the same weights from the key names) and run on CPU in fp32.
Inputs and outputs are stored; weights are not (28 M parameters).
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

WEIGHT_SEED = 0


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path)/1024:.1f} KiB)")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    cfg = ref_loader.default_config()
    R = ref_loader.ref_models()

    # ---- photometry (HyraxBaselineCLS.py:49-86) ------------------------------------
    ref = R.photo.HyraxBaselineCLS(cfg).eval()
    ref.load_state_dict(synth.det_state_dict(ref, WEIGHT_SEED))
    x, pad, lens = synth.photometry_batch(8, seed=11)
    with torch.no_grad():
        logits_fast = ref((x, pad, None))  # nested-tensor fast path
    logits_slow = ref((x, pad, None)).detach()  # grad-enabled path (−inf masking)
    labels = synth.labels(8, seed=11)
    loss = R.photo.FocalLoss()(ref((x, pad, None)), labels)
    ref.zero_grad()
    loss.backward()
    g = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    # short padded layout as produced by the fusion collate (Time2Vec.py:18-45)
    Ls = int(lens.max())
    xs, ps = x[:, :Ls].contiguous(), pad[:, :Ls].contiguous()
    with torch.no_grad():
        logits_short = ref((xs, ps, None))
    save(
        "photo", x=x, pad=pad, labels=labels, logits=logits_fast, logits_slowpath=logits_slow, logits_short=logits_short,
        loss=loss.detach(), g_fc_weight=g["fc.weight"], g_in_proj_weight=g["in_proj.weight"],
        g_l0_in_proj_weight=g["encoder.layers.0.self_attn.in_proj_weight"], g_l3_linear2_weight=g["encoder.layers.3.linear2.weight"],
        g_time2vec_w=g["time2vec.w"], g_cls_tok=g["cls_tok"], g_l1_norm1_weight=g["encoder.layers.1.norm1.weight"],
    )

    # ---- spectra (spectranet.py:156-170) -------------------------------------------
    ref = R.spectra.SpectraNet(cfg).eval()
    ref.load_state_dict(synth.det_state_dict(ref, WEIGHT_SEED))
    s4096 = synth.spectra(2, seed=12, L=4096)
    s3481 = synth.spectra(2, seed=13, L=3481)
    with torch.no_grad():
        o4096 = ref((s4096, None, None))
        o3481 = ref((s3481, None, None))
        st0 = ref.all_stages[0](s4096)  # (2,64,1024)
    save("spectra", s4096=s4096, s3481=s3481, logits4096=o4096, logits3481=o3481, stage0_b0_c0=st0[0, 0], stage0_b1_c63=st0[1, 63])

    # tiny SpectraNet config incl. full weights (independent of det_state_dict)
    tcfg = ref_loader.cfg_copy(cfg)
    tcfg["model"]["SpectraNet"].update(
        channels=[8, 16, 16, 32, 32], kernel_sizes_per_stage=[[3, 9, 33], [3, 7, 17], [3, 5, 9], [3, 5, 7], [3, 5, 7]], flat_dim=96, class_order=4
    )
    torch.manual_seed(5)
    ref = R.spectra.SpectraNet(tcfg).eval()
    st = torch.randn(3, 1, 777)
    with torch.no_grad():
        ot = ref((st, None, None))
    save("spectra_tiny", x=st, logits=ot, **{"w::" + k: v for k, v in ref.state_dict().items()})

    # ---- image + metadata (astrominn.py:220-300) -----------------------------------
    ref = R.astrominn.AstroMiNN(cfg).eval()
    ref.load_state_dict(synth.det_state_dict(ref, WEIGHT_SEED))
    meta = synth.metadata(6, seed=14)
    img = synth.cutouts(6, seed=14)
    with torch.no_grad():
        feat = ref.image_tower.backbone(img)
        im32 = ref.image_tower(img)
        logits = ref((meta, img, None))
    tgt = torch.nn.functional.one_hot(synth.labels(6, seed=14), 5).float()
    out = ref((meta, img, tgt))
    loss = torch.nn.CrossEntropyLoss()(out, tgt)
    ref.zero_grad()
    loss.backward()
    g = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
    save(
        "astrominn", metadata=meta, image=img, target=tgt, backbone=feat, image_feats=im32, logits=logits, loss=loss.detach(),
        g_stem0_weight=g["image_tower.backbone.stem.0.weight"], g_router0_weight=g["fusion_router.0.weight"],
        g_s2b4_gamma=g["image_tower.backbone.stages.2.blocks.4.gamma"], g_mega_skip_weight=g["mega_tower.skip_path.weight"],
        g_s3b0_fc1_weight_row0=g["image_tower.backbone.stages.3.blocks.0.mlp.fc1.weight"][0],
    )

    # ---- fusion (brew_cider.py:834-860 head over the src encoders; the real encoders are
    #      the reference modules, the 20-line head is oracle.models.AppleCider) ---------
    for fusion in ("avg", "concat"):
        fm = om.AppleCider(om.default_config(), hidden_dim=64, fusion=fusion).eval()
        # swap in the REAL reference encoders
        pc = ref_loader.cfg_copy(cfg)
        pc["model"]["HyraxBaselineCLS"]["mode"] = "all"
        fm.photometry_encoder = R.photo.HyraxBaselineCLS(pc).eval()
        fm.spectra_encoder = R.spectra.SpectraNet(cfg).eval()
        fm.img_metadata_encoder = R.astrominn.AstroMiNN(cfg).eval()
        fm.load_state_dict(synth.det_state_dict(fm, WEIGHT_SEED))
        x, pad, lens = synth.photometry_batch(3, seed=15)
        Ls = int(lens.max())
        x, pad = x[:, :Ls].contiguous(), pad[:, :Ls].contiguous()
        meta, img, sp = synth.metadata(3, seed=15, missing_frac=0.0), synth.cutouts(3, seed=15), synth.spectra(3, seed=15)
        with torch.no_grad():
            lg = fm(x, pad, meta, img, sp)
            p, im, s = fm.get_embeddings(x, pad, meta, img, sp)
        save(f"fusion_{fusion}", x=x, pad=pad, metadata=meta, image=img, spectra=sp, logits=lg, p_emb=p, im_emb=im, s_emb=s)


if __name__ == "__main__":
    main()
