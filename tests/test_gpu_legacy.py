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
