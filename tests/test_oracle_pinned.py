"""CPU suite: pin the oracle.

1. oracle/models.py (the torch restatement) reproduces the committed golden outputs that
   tests/golden/make_golden.py generated from the REAL reference modules.
2. When /root/reference is mounted (build container), the restatement is compared live against the
   real reference on fresh inputs, including state_dict key/shape identity.
3. The restated timm ConvNeXt-T is cross-checked against torchvision.models.convnext_tiny.
"""
import os

import numpy as np
import pytest
import torch

from util import assert_close, load_golden


def _oracle(name, **kw):
    from applecider_b200 import synth
    from oracle import models as om

    m = getattr(om, name)(om.default_config(), **kw).eval()
    m.load_state_dict(synth.det_state_dict(m, 0))
    return m


def test_oracle_photo_vs_golden(golden_dir):
    g = load_golden(golden_dir, "photo")
    m = _oracle("HyraxBaselineCLS")
    with torch.no_grad():
        out = m((g["x"], g["pad"], None))
    assert_close(out, g["logits"], 1e-5, "photo logits (fast path)")
    assert_close(out, g["logits_slowpath"], 1e-5, "photo logits (slow path)")
    # gradients of the focal loss through the restated encoder (eval mode, autograd on)
    from oracle.models import focal_loss

    m.zero_grad()
    loss = focal_loss(m((g["x"], g["pad"], None)), g["labels"])
    loss.backward()
    assert_close(loss, g["loss"], 1e-5, "focal loss")
    grads = dict(m.named_parameters())
    assert_close(grads["fc.weight"].grad, g["g_fc_weight"], 1e-4, "d fc.weight")
    assert_close(grads["in_proj.weight"].grad, g["g_in_proj_weight"], 1e-4, "d in_proj.weight", atol=1e-6)
    assert_close(grads["time2vec.w"].grad, g["g_time2vec_w"], 1e-4, "d time2vec.w", atol=1e-6)
    assert_close(grads["encoder.layers.0.self_attn.in_proj_weight"].grad, g["g_l0_in_proj_weight"], 1e-4, "d l0 in_proj", atol=1e-6)


def test_oracle_spectra_vs_golden(golden_dir):
    g = load_golden(golden_dir, "spectra")
    m = _oracle("SpectraNet")
    with torch.no_grad():
        assert_close(m((g["s4096"], None, None)), g["logits4096"], 1e-5, "spectra 4096")
        assert_close(m((g["s3481"], None, None)), g["logits3481"], 1e-5, "spectra 3481")


def test_oracle_spectra_tiny_vs_golden(golden_dir):
    from oracle import models as om

    g = load_golden(golden_dir, "spectra_tiny")
    cfg = om.default_config()
    cfg["model"]["SpectraNet"].update(
        channels=[8, 16, 16, 32, 32], kernel_sizes_per_stage=[[3, 9, 33], [3, 7, 17], [3, 5, 9], [3, 5, 7], [3, 5, 7]], flat_dim=96, class_order=4
    )
    m = om.SpectraNet(cfg).eval()
    m.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("w::")})
    with torch.no_grad():
        assert_close(m((g["x"], None, None)), g["logits"], 1e-5, "tiny spectra")


def test_oracle_astrominn_vs_golden(golden_dir):
    g = load_golden(golden_dir, "astrominn")
    m = _oracle("AstroMiNN")
    with torch.no_grad():
        assert_close(m.image_tower.backbone(g["image"]), g["backbone"], 1e-5, "convnext features")
        assert_close(m.image_tower(g["image"]), g["image_feats"], 1e-5, "split head")
        assert_close(m((g["metadata"], g["image"], None)), g["logits"], 1e-5, "astrominn logits")


@pytest.mark.parametrize("fusion", ["avg", "concat"])
def test_oracle_fusion_vs_golden(golden_dir, fusion):
    g = load_golden(golden_dir, f"fusion_{fusion}")
    m = _oracle("AppleCider", hidden_dim=64, fusion=fusion)
    with torch.no_grad():
        assert_close(m(g["x"], g["pad"], g["metadata"], g["image"], g["spectra"]), g["logits"], 1e-5, "fusion logits")


def test_convnext_restatement_vs_torchvision():
    tv = pytest.importorskip("torchvision")
    from applecider_b200 import synth
    from oracle import models as om

    mine = om.ConvNeXtTiny(in_chans=3).eval()
    sd = synth.det_state_dict(mine, 3)
    mine.load_state_dict(sd)
    ref = tv.models.convnext_tiny(weights=None).eval()
    remap = {}
    for k, v in sd.items():
        p = k.split(".")
        if p[0] == "stem":
            nk = f"features.0.{p[1]}.{p[2]}"
        elif p[0] == "stages" and p[2] == "downsample":
            nk = f"features.{2 * int(p[1])}.{p[3]}.{p[4]}"
        elif p[0] == "stages" and p[2] == "blocks":
            s, b = int(p[1]), int(p[3])
            if p[4] == "gamma":
                nk, v = f"features.{2 * s + 1}.{b}.layer_scale", v.view(-1, 1, 1)
            else:
                sub = {"conv_dw": "0", "norm": "2", "mlp": None}[p[4]]
                if sub is None:
                    sub = {"fc1": "3", "fc2": "5"}[p[5]]
                nk = f"features.{2 * s + 1}.{b}.block.{sub}.{p[-1]}"
        elif p[0] == "head":
            nk = f"classifier.0.{p[2]}"
        remap[nk] = v
    missing = ref.load_state_dict(remap, strict=False)
    assert set(missing.missing_keys) == {"classifier.2.weight", "classifier.2.bias"} and not missing.unexpected_keys
    x = synth.cutouts(3, seed=9)
    with torch.no_grad():
        a = mine(x)
        b = ref.classifier[0](ref.avgpool(ref.features(x))).flatten(1)
    assert_close(a, b, 1e-5, "ConvNeXt-T restatement vs torchvision")


def test_oracle_vs_real_reference_live():
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("reference tree not mounted (GPU box)")
    from applecider_b200 import synth
    from oracle import models as om

    cfg = ref_loader.default_config()
    R = ref_loader.ref_models()
    for name, ref_cls in [("HyraxBaselineCLS", R.photo.HyraxBaselineCLS), ("SpectraNet", R.spectra.SpectraNet), ("AstroMiNN", R.astrominn.AstroMiNN)]:
        ref = ref_cls(cfg).eval()
        port = getattr(om, name)(om.default_config()).eval()
        rs, ps = ref.state_dict(), port.state_dict()
        assert list(rs) == list(ps) and all(rs[k].shape == ps[k].shape for k in rs), name
        sd = synth.det_state_dict(ref, 7)
        ref.load_state_dict(sd)
        port.load_state_dict(sd)
        if name == "HyraxBaselineCLS":
            x, pad, _ = synth.photometry_batch(5, seed=71)
            batch = (x, pad, None)
        elif name == "SpectraNet":
            batch = (synth.spectra(1, seed=72, L=2000), None, None)
        else:
            batch = (synth.metadata(3, seed=73), synth.cutouts(3, seed=73), None)
        with torch.no_grad():
            assert_close(port(batch), ref(batch), 1e-5, f"{name}: port vs real reference")


def test_product_state_dict_schema_matches_oracle():
    import applecider_b200 as ab
    from oracle import models as om

    for name in ("HyraxBaselineCLS", "SpectraNet", "AstroMiNN"):
        a, b = getattr(ab, name)(ab.default_config()).state_dict(), getattr(om, name)(om.default_config()).state_dict()
        assert list(a) == list(b) and all(a[k].shape == b[k].shape for k in a), name
    a, b = ab.AppleCider(ab.default_config()).state_dict(), om.AppleCider(om.default_config()).state_dict()
    assert list(a) == list(b)


def _clip_grads(named_grads, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_ arithmetic on a dict of gradients."""
    total = torch.sqrt(sum((g.float() ** 2).sum() for g in named_grads.values()))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef for k, g in named_grads.items()}, total * coef


def test_oracle_mpt_vs_golden(golden_dir):
    """MPTModel restatement vs the unmodified reference train_step (tests/golden/make_golden_mpt.py)."""
    from applecider_b200 import synth
    from oracle import models as om

    g = load_golden(golden_dir, "mpt")
    cfg = om.default_config()
    cfg["model"]["HyraxBaselineCLS"]["dropout"] = 0.0
    m = om.MPTModel(cfg).train()
    m.load_state_dict(synth.det_state_dict(m, 0))
    # same seed -> same torch.randperm draws -> the reference's mask, bit for bit
    torch.manual_seed(123)
    x = g["x"].clone()
    masked = m.mask_batch(x, g["pad"])
    assert torch.equal(masked, g["masked"])
    assert torch.equal(x, g["x_masked"])
    loss, lf, lb, ldt = m.losses(x, g["pad"], masked)
    assert_close(loss, g["loss"], 1e-5, "mpt loss")
    mc = cfg["model"]["HyraxBaselineCLS"]
    assert_close(mc["lambda_f"] * lf * mc["lambda_b"] * lb * mc["lambda_dt"] * ldt, g["loss"], 1e-5, "mpt loss product")
    loss.backward()
    grads, norm = _clip_grads({n: p.grad for n, p in m.named_parameters() if p.grad is not None})
    assert_close(norm, g["clipped_grad_norm"], 1e-5, "clipped grad norm")
    for k in g:
        if k.startswith("g_"):
            name = [n for n in grads if n.replace(".", "_") == k[2:]][0]
            assert_close(grads[name], g[k], 2e-4, "d " + name, atol=1e-6)


def test_oracle_legacy_spectra_vs_golden(golden_dir):
    """Variant-B spectra encoder restatement vs the real `build_spec_model` (tests/golden/make_golden_legacy.py), incl. key identity."""
    from applecider_b200 import synth
    from oracle import models as om

    g = np.load(os.path.join(golden_dir, "legacy_spectra.npz"))
    for mode in ("all", "spectra"):
        m = om.SpectraClassificationB({"mode": mode, "classes": list(range(5))}).eval()
        assert sorted(m.state_dict().keys()) == list(g[f"keys_{mode}"])
        m.load_state_dict(synth.det_state_dict(m, 0))
        with torch.no_grad():
            assert_close(m(torch.from_numpy(g[f"x_{mode}"])), torch.from_numpy(g[f"y_{mode}"]), 1e-5, f"legacy spectra ({mode})")
