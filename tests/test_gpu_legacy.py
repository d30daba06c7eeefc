"""GPU parity of the legacy "variant B" spectra encoder (_archive/notebooks/brew_cider.py:585-708) against the
golden outputs of the REAL reference and the CPU oracle; fp32 path, |err| <= 1e-4 * max(1, |ref|_inf).  Also: strict
state_dict compatibility with the reference key list, and the fusion model composed with this encoder (spectra_proj 256 -> H)."""
import os

import numpy as np
import pytest
import torch

from util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("mode", ["all", "spectra"])
def test_legacy_spectra_matches_reference_golden(golden_dir, mode):
    from applecider_b200 import synth
    from applecider_b200.legacy import SpectraClassificationB

    g = np.load(os.path.join(golden_dir, "legacy_spectra.npz"))
    m = SpectraClassificationB({"mode": mode, "classes": list(range(5))})
    assert sorted(m.state_dict().keys()) == list(g[f"keys_{mode}"])
    m.load_state_dict(synth.det_state_dict(m, 0), strict=True)
    m = m.to(DEV).eval()
    y = m(torch.from_numpy(g[f"x_{mode}"]).to(DEV))
    assert_close(y, torch.from_numpy(g[f"y_{mode}"]), 1e-4, f"legacy spectra ({mode})")


def test_legacy_spectra_vs_oracle_fresh_batch():
    from applecider_b200 import synth
    from applecider_b200.legacy import SpectraClassificationB
    from oracle import models as om

    o = om.SpectraClassificationB().eval()
    sd = synth.det_state_dict(o, 3)
    o.load_state_dict(sd)
    m = SpectraClassificationB()
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    x = synth.spectra(5, seed=99, L=4096)
    with torch.no_grad():
        ref = o(x)
    assert_close(m(x.to(DEV)), ref, 1e-4, "legacy spectra vs oracle")
    with pytest.raises(RuntimeError):
        m(x)  # CPU tensor: no fallback


def test_fusion_with_legacy_spectra_encoder():
    """AppleCider(spectra_variant='B'): spectra_proj takes the 256-d embedding (brew_cider.py:826), forward == manual composition."""
    import applecider_b200 as ab
    from applecider_b200 import synth

    model = ab.AppleCider(ab.default_config(), hidden_dim=64, fusion="concat", spectra_variant="B")
    assert model.spectra_proj.in_features == 256
    model.load_state_dict(synth.det_state_dict(model, 0), strict=True)
    model = model.to(DEV).eval()
    B = 4
    x, pad, _ = synth.photometry_batch(B, seed=8)
    args = [t.to(DEV) for t in (x, pad, synth.metadata(B, seed=8), synth.cutouts(B, seed=8), synth.spectra(B, seed=8, L=4096))]
    with torch.no_grad():
        logits = model(*args)
        p, im, s = model.get_embeddings(*args)
        emb = torch.cat([p, im, s], 1)
        ref = emb @ model.fc.weight.t() + model.fc.bias
    assert logits.shape == (B, 5) and torch.isfinite(logits).all()
    assert_close(logits, ref, 1e-5, "fusion head over the legacy spectra embedding")
    assert_close(s.norm(dim=1), torch.ones(B), 1e-5, "unit-norm spectra embedding")


def test_legacy_xastrominn_four_channel_cutouts():
    """XastroMiNN (in_chans=4, forward(metadata, image)) against the CPU oracle with the same 4-channel stem."""
    import copy

    from applecider_b200 import synth
    from applecider_b200.legacy import XastroMiNN
    from oracle import models as om

    ocfg = copy.deepcopy(om.default_config())
    ocfg["model"]["AstroMiNN"]["in_chans"] = 4
    o = om.AstroMiNN(ocfg).eval()
    sd = synth.det_state_dict(o, 0)
    o.load_state_dict(sd)
    assert sd["image_tower.backbone.stem.0.weight"].shape == (96, 4, 4, 4)
    m = XastroMiNN()
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    B = 6
    meta = synth.metadata(B, seed=12)
    img3 = synth.cutouts(B, seed=12)
    img = torch.cat([img3, img3[:, :1] * 0.5 - img3[:, 1:2]], 1).contiguous()  # a 4th plane
    with torch.no_grad():
        ref = o((meta, img, None))
        got = m(meta.to(DEV), img.to(DEV))
    assert_close(got, ref, 1e-4, "XastroMiNN logits")


def test_redshift_head_softplus_forward_and_gradient():
    """SpectraNet(redshift=true, redshift_softplus=true): positive outputs = softplus of the plain regressor, gradients flow
    (archived SpectraNetRedshift.py:93-113)."""
    import copy

    import applecider_b200 as ab
    from applecider_b200 import synth

    cfg = copy.deepcopy(ab.default_config())
    cfg["model"]["SpectraNet"].update(redshift=True, compute_dtype="fp32",
                                      channels=[8, 16, 16, 32, 32], kernel_sizes_per_stage=[[3, 9, 33], [3, 7, 17], [3, 5, 9], [3, 5, 7], [3, 5, 7]], flat_dim=96)
    plain = ab.SpectraNet(cfg)
    sd = synth.det_state_dict(plain, 0)
    plain.load_state_dict(sd)
    cfg2 = copy.deepcopy(cfg)
    cfg2["model"]["SpectraNet"]["redshift_softplus"] = True
    soft = ab.SpectraNet(cfg2)
    soft.load_state_dict(sd)
    plain, soft = plain.to(DEV).eval(), soft.to(DEV).eval()
    x = synth.spectra(6, seed=5, L=1024).to(DEV)
    with torch.no_grad():
        z = plain((x, None, None))
        y = soft((x, None, None))
    assert y.shape == (6,) and (y > 0).all()
    assert_close(y, torch.nn.functional.softplus(z), 1e-6, "softplus(regressor)")
    yt = soft((x, None, None))  # autograd path
    assert_close(yt, y, 1e-5, "train-path forward")
    yt.sum().backward()
    g = soft.regressor[4].weight.grad
    assert g is not None and torch.isfinite(g).all() and g.abs().max() > 0
