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


def _ZERO_GRAD(name):
    """conv biases of the BatchNorm stages (1-4) in training mode"""
    return name.startswith(("stage1.", "stage2.", "stage3.", "stage4.")) and ".convs." in name and name.endswith(".bias")


def _legacy_train_pair(golden_dir, dtype):
    from applecider_b200 import synth
    from applecider_b200.legacy import SpectraClassificationB

    g = np.load(os.path.join(golden_dir, "legacy_train.npz"))
    m = SpectraClassificationB({"mode": "spectra", "classes": list(range(5))}, compute_dtype=dtype)
    m.load_state_dict(synth.det_state_dict(m, 0), strict=True)
    m = m.to(DEV).train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return g, m


def test_legacy_variant_b_training_matches_reference(golden_dir):
    """ONE training-mode forward + backward (BatchNorm on batch statistics, running statistics updated in place) against the
    record of the REAL `build_spec_model` in train() mode (brew_cider.py:585-708; tests/golden/make_golden_legacy.py)."""
    from applecider_b200 import fn

    g, m = _legacy_train_pair(golden_dir, "fp32")
    logits = m(torch.from_numpy(g["x"]).to(DEV))
    assert_close(logits, torch.from_numpy(g["logits"]), 1e-4, "variant-B train-mode logits")
    tgt = torch.nn.functional.one_hot(torch.from_numpy(g["y"]), 5).float().to(DEV)
    loss = fn.soft_cross_entropy(logits, tgt)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4
    sd = m.state_dict()
    for i in range(1, 5):
        assert_close(sd[f"stage{i}.0.norm.running_mean"], torch.from_numpy(g[f"rm{i}"]), 1e-5, f"running_mean stage {i}", atol=1e-6)
        assert_close(sd[f"stage{i}.0.norm.running_var"], torch.from_numpy(g[f"rv{i}"]), 1e-5, f"running_var stage {i}", atol=1e-6)
        assert int(sd[f"stage{i}.0.norm.num_batches_tracked"]) == 1
    grads = {n: p.grad for n, p in m.named_parameters()}
    for k in g.files:
        if k.startswith("g_") and not k.endswith("_rows"):
            gr = grads[k[2:]].detach().cpu()
            got = gr.reshape(gr.shape[0], -1)[:, :256] if gr.dim() > 1 else gr
            ref = torch.from_numpy(g[k])
            if _ZERO_GRAD(k[2:]):
                # a bias in front of a train-mode BatchNorm has an exactly zero gradient (the batch mean removes it): the reference
                # holds rounding noise there; ours must be noise too, not a value
                wmax = grads[k[2:].replace(".bias", ".weight")].abs().max().item()
                assert got.abs().max().item() <= 1e-4 * wmax, f"{k[2:]}: gradient of a pre-BatchNorm bias must vanish"
                continue
            s = ref.abs().max().clamp_min(1e-12)
            assert_close(got / s, ref / s, 1e-3, f"variant-B gradient {k[2:]}")
    ref = torch.from_numpy(g["g_class_model.0.weight_rows"])
    s = ref.abs().max().clamp_min(1e-12)
    assert_close(grads["class_model.0.weight"][:4].cpu() / s, ref / s, 1e-3, "class_model.0.weight gradient (C-major flatten re-layout)")


def test_legacy_variant_b_bf16_path(golden_dir):
    """bf16 path (stages 2-5 on the tcgen05 implicit GEMM, stage 1 fp32): logits within 1.5e-2, gradients aligned with the reference."""
    from applecider_b200 import fn

    g, m = _legacy_train_pair(golden_dir, "bf16")
    logits = m(torch.from_numpy(g["x"]).to(DEV))
    assert_close(logits, torch.from_numpy(g["logits"]), 1.5e-2, "variant-B bf16 train-mode logits")
    tgt = torch.nn.functional.one_hot(torch.from_numpy(g["y"]), 5).float().to(DEV)
    fn.soft_cross_entropy(logits, tgt).backward()
    grads = {n: p.grad for n, p in m.named_parameters()}
    for k in g.files:
        if k.startswith("g_") and not k.endswith("_rows") and not _ZERO_GRAD(k[2:]):
            gr = grads[k[2:]].detach().float().cpu()
            got = (gr.reshape(gr.shape[0], -1)[:, :256] if gr.dim() > 1 else gr).flatten()
            ref = torch.from_numpy(g[k]).flatten()
            cos = torch.nn.functional.cosine_similarity(got, ref, dim=0).item()
            assert cos > 0.97, f"{k[2:]}: cosine {cos:.4f}"
    m.eval()
    with torch.no_grad():
        e = m(torch.from_numpy(g["x"]).to(DEV))
    assert torch.isfinite(e).all()


def test_fusion_with_legacy_encoder_trains():
    """AppleCider(spectra_variant='B') through the whole-step training path: loss decreases over a few fused-Adam steps."""
    import applecider_b200 as ab
    from applecider_b200 import fn, synth
    from applecider_b200.ddp import ddp_train_step
    from applecider_b200.optim import FusedAdam

    model = ab.AppleCider(ab.default_config(), hidden_dim=64, fusion="avg", spectra_variant="B", compute_dtype="fp32")
    model.load_state_dict(synth.det_state_dict(model, 0), strict=True)
    model = model.to(DEV).train()
    B = 8
    x, pad, lens = synth.photometry_batch(B, seed=9, L=64)
    args = [t.to(DEV) for t in (x, pad, synth.metadata(B, seed=9, missing_frac=0.0), synth.cutouts(B, seed=9), synth.spectra(B, seed=9, L=4096))]
    tgt = torch.nn.functional.one_hot(synth.labels(B, seed=9), 5).float().to(DEV)
    opt = FusedAdam([p for p in model.parameters() if p.requires_grad], lr=2e-4)
    losses = [ddp_train_step(opt.grads, lambda: fn.soft_cross_entropy(model(*args), tgt), opt).item() for _ in range(8)]
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
