"""GPU parity of the masked-event pre-training path (MPTModel, HyraxBaselineCLS.py:194-319).

* loss + clipped gradients against the golden produced by the UNMODIFIED reference train_step
  (tests/golden/make_golden_mpt.py), feeding the reference's own mask;
* every parameter gradient against the CPU oracle's autograd;
* the device mask sampler against the rules of _mask_batch (:283-319): only valid tokens, k = max(int(n*p),3),
  k//3 per band capped by the band population, extras from the rest, channels 2:7 zeroed exactly there.
fp32: loss 1e-4 relative, gradients 2e-4 * max|g|; bf16: loss 3e-2, cosine >= 0.99."""
import numpy as np
import pytest
import torch

from test_oracle_pinned import _clip_grads
from util import assert_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pair(dtype="fp32", dropout=0.0):
    import applecider_b200 as ab
    from applecider_b200 import synth
    from oracle import models as om

    cfg, ocfg = ab.default_config(), om.default_config()
    for c in (cfg, ocfg):
        c["model"]["HyraxBaselineCLS"]["dropout"] = dropout
    cfg["model"]["HyraxBaselineCLS"]["compute_dtype"] = dtype
    oracle = om.MPTModel(ocfg).train()
    sd = synth.det_state_dict(oracle, 0)
    oracle.load_state_dict(sd)
    prod = ab.MPTModel(cfg)
    prod.load_state_dict(sd, strict=True)
    return prod.to(DEV).train(), oracle


def test_mpt_loss_and_gradients_vs_reference_golden(golden_dir):
    from applecider_b200.train import mpt_losses

    g = load_golden(golden_dir, "mpt")
    prod, oracle = _pair()
    xm, pad, masked = g["x_masked"].to(DEV), g["pad"].to(DEV), g["masked"].to(DEV)
    loss, parts = mpt_losses(prod, xm, pad, masked)
    assert_close(loss, g["loss"], 1e-4, "mpt loss")
    ol, olf, olb, oldt = oracle.losses(g["x_masked"].clone(), g["pad"], g["masked"])
    assert_close(parts, torch.stack([ol, olf, olb, oldt]).detach(), 1e-4, "loss parts (loss, L_f, L_b, L_dt)")
    loss.backward()
    grads, norm = _clip_grads({n: p.grad for n, p in prod.named_parameters() if p.grad is not None})
    assert_close(norm, g["clipped_grad_norm"], 1e-4, "clipped grad norm")
    for k in g:
        if k.startswith("g_"):
            name = [n for n in grads if n.replace(".", "_") == k[2:]][0]
            ref = g[k]
            assert_close(grads[name], ref, 2e-4 * ref.abs().max().item(), "d " + name, atol=1e-7)
    ol.backward()
    for n, p in prod.named_parameters():
        ref = dict(oracle.named_parameters())[n].grad
        assert p.grad is not None and ref is not None, n
        assert_close(p.grad, ref, 2e-4 * max(ref.abs().max().item(), 1e-6) / max(1.0, ref.abs().max().item()), "oracle grad " + n, atol=1e-6)


def test_mpt_bf16_close_to_oracle(golden_dir):
    from applecider_b200.train import mpt_losses

    g = load_golden(golden_dir, "mpt")
    prod, oracle = _pair("bf16")
    loss, parts = mpt_losses(prod, g["x_masked"].to(DEV), g["pad"].to(DEV), g["masked"].to(DEV))
    assert_close(loss, g["loss"], 3e-2, "mpt loss bf16")
    loss.backward()
    ol = oracle.losses(g["x_masked"].clone(), g["pad"], g["masked"])[0]
    ol.backward()
    og = dict(oracle.named_parameters())
    for n, p in prod.named_parameters():
        ref = og[n].grad
        if ref.abs().max() > 1e-9:
            cos = torch.nn.functional.cosine_similarity(p.grad.float().cpu().flatten(), ref.flatten(), dim=0).item()
            assert cos >= 0.99, f"{n}: cosine {cos:.4f}"


@pytest.mark.parametrize("B,seed", [(64, 5), (257, 6)])
def test_mpt_mask_sampler_rules(B, seed):
    from applecider_b200 import synth

    prod, _ = _pair()
    x, pad, lens = synth.photometry_batch(B, seed=seed)
    if B == 257:  # edge cases: empty, 1-token and 2-token light curves, and a single-band one
        pad[0, :] = True
        pad[1, :] = True; pad[1, 0] = False
        pad[2, :] = True; pad[2, :2] = False
        x[3, :, 4:7] = torch.tensor([0.0, 1.0, 0.0])
    x0 = x.clone()
    xd = x.to(DEV).contiguous()
    masked = prod.mask_batch(xd, pad.to(DEV), seed=77).cpu()
    xd = xd.cpu()
    p = prod.config["model"]["HyraxBaselineCLS"]["mask_p"]
    assert not (masked & pad).any(), "padded tokens must never be masked"
    assert torch.equal(xd[~masked], x0[~masked]), "unmasked rows must be untouched"
    assert (xd[masked][:, 2:7] == 0).all() and torch.equal(xd[masked][:, :2], x0[masked][:, :2])
    bands = x0[..., 4:7].argmax(-1)
    for b in range(B):
        valid = ~pad[b]
        n = int(valid.sum())
        k = max(int(n * p), 3)
        each, extras = k // 3, k - 3 * (k // 3)
        per_band = [int((valid & (bands[b] == c)).sum()) for c in range(3)]
        base = sum(min(c, each) for c in per_band)
        want = base + min(extras, n - base)
        assert int(masked[b].sum()) == want, (b, n, per_band, int(masked[b].sum()), want)
        for c in range(3):  # at least the balanced share of every band
            assert int((masked[b] & (bands[b] == c)).sum()) >= min(per_band[c], each)
    # a different seed draws a different mask; the same seed the same one
    x2 = x0.to(DEV).contiguous()
    assert torch.equal(prod.mask_batch(x2, pad.to(DEV), seed=77).cpu(), masked)
    x3 = x0.to(DEV).contiguous()
    assert not torch.equal(prod.mask_batch(x3, pad.to(DEV), seed=78).cpu(), masked)


def test_mpt_mask_is_uniform_within_band():
    """Each valid token of a band is drawn with the same probability (chi-square-ish bound over 4000 draws)."""
    from applecider_b200 import synth

    prod, _ = _pair()
    x, pad, _ = synth.photometry_batch(1, seed=9)
    L = x.shape[1]
    pad[:] = True
    pad[0, :30] = False
    x[0, :, 4:7] = 0
    x[0, torch.arange(L), 4 + (torch.arange(L) % 3)] = 1.0  # 10 tokens per band
    R = 4000
    xs = x.repeat(R, 1, 1).to(DEV).contiguous()
    m = prod.mask_batch(xs, pad.repeat(R, 1).to(DEV), seed=3).float().mean(0).cpu()  # k = 9 -> 3 per band -> p = 0.3
    assert (m[30:] == 0).all()
    assert (m[:30] - 0.3).abs().max().item() < 0.04, m[:30]


def test_mpt_train_step_runs_and_learns():
    from applecider_b200 import synth

    prod, _ = _pair(dropout=0.4)
    x, pad, _ = synth.photometry_batch(64, seed=3)
    losses = []
    for it in range(8):
        out = prod.train_step((x.clone().to(DEV), pad.to(DEV), None))
        assert np.isfinite(out["loss"])
        losses.append(out["loss"])
    assert min(losses[4:]) < losses[0], losses


def test_mpt_heads_forward():
    prod, oracle = _pair()
    z = torch.randn(3, 7, 128)
    f, b, d = prod(z.to(DEV))
    of, ob, od = oracle(z)
    assert_close(f, of, 1e-5, "head_flux"); assert_close(b, ob, 1e-5, "head_band"); assert_close(d, od, 1e-5, "head_dt")
