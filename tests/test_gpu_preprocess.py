"""GPU parity of the preprocessing kernels (P1-P5) against the numpy oracle and the golden outputs of the real
reference.  Bit-exact: masks, lengths, one-hot, band ids, segmentation (event counts), crop geometry, median
selection.  Float columns: <= 2 float32 ulp (libm log1p/log10/pow differ between hosts and the GPU)."""
import numpy as np
import pytest
import torch

from util import assert_close  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def g(golden_dir):
    import os

    return np.load(os.path.join(golden_dir, "preprocess.npz"))


def _ulp_close(got, ref, ulps=2, name=""):
    got, ref = np.asarray(got, np.float32), np.asarray(ref, np.float32)
    assert got.shape == ref.shape, name
    assert np.array_equal(np.isnan(got), np.isnan(ref)), name
    tol = ulps * np.spacing(np.maximum(np.abs(ref), np.float32(1e-30)))
    bad = np.abs(got - ref) > tol
    assert not np.nansum(bad), f"{name}: {int(np.nansum(bad))} elements differ by more than {ulps} ulp; max abs {np.nanmax(np.abs(got-ref))}"
    return float(np.mean(got == ref))


def test_p1_lightcurve(g):
    from applecider_b200 import preprocess as pp

    raw, off = torch.from_numpy(g["p1_raw"]).to(DEV), torch.from_numpy(g["p1_offsets"]).to(DEV)
    x, mask, lens = pp.prep_lightcurves(raw, off, float(g["p1_horizon"]), torch.from_numpy(g["p1_mean"]).to(DEV), torch.from_numpy(g["p1_std"]).to(DEV))
    assert np.array_equal(mask.cpu().numpy(), g["p1_mask"])
    assert np.array_equal(lens.cpu().numpy(), (~g["p1_mask"]).sum(1))
    xg = x.cpu().numpy()
    assert np.array_equal(xg[..., 4:], g["p1_x"][..., 4:]), "one-hot band must be bit-exact"
    assert np.array_equal(xg[..., 2:4], g["p1_x"][..., 2:4]), "normalised logf/logfe are pure IEEE sub/div: bit-exact"
    # log1p differs by <= 1 ulp between libm and the GPU; the normalisation (x - mu) / sd then divides that by sd
    np.testing.assert_allclose(xg[..., :2], g["p1_x"][..., :2], rtol=0, atol=2e-6, err_msg="log1p columns")


def test_p1_large_batch_properties():
    from applecider_b200 import preprocess as pp, synth
    from oracle import preprocess as op

    raws = synth.raw_light_curves(3000, seed=77, max_len=600)
    raw, off = pp.ragged(raws)
    mean, std = torch.tensor([2.9, 0.9, 1.5, 0.08]), torch.tensor([1.1, 0.8, 0.5, 0.04])
    x, mask, lens = pp.prep_lightcurves(raw, off, 100.0, mean.to(DEV), std.to(DEV))
    seqs = [op.lightcurve_features(r, 100.0) for r in raws]
    xo, mo = op.collate_photometry(seqs, mean.numpy(), std.numpy())
    assert np.array_equal(mask.cpu().numpy(), mo)
    assert np.array_equal(x.cpu().numpy()[..., 2:], xo[..., 2:].astype(np.float32))
    np.testing.assert_allclose(x.cpu().numpy()[..., :2], xo[..., :2], rtol=0, atol=2e-6, err_msg="log1p columns (3000 objects)")


def test_p2_events(g):
    from applecider_b200 import preprocess as pp

    off = torch.from_numpy(g["p2_offsets"]).to(DEV)
    out = pp.prep_events(*[torch.from_numpy(g[k]).to(DEV) for k in ("p2_mjd", "p2_mag", "p2_magerr", "p2_fid")], off)
    n_ev = out["n_events"].cpu().numpy()
    oo = g["p2_out_offsets"]
    assert np.array_equal(n_ev, np.diff(oo)), "merged-event counts (window segmentation) must be bit-exact"
    o = g["p2_offsets"]
    for i in range(len(o) - 1):
        sl = slice(o[i], o[i] + n_ev[i])
        so = slice(oo[i], oo[i + 1])
        assert np.array_equal(out["band_id"][sl].cpu().numpy(), g["p2_band_id"][so]), "band ids must be bit-exact"
        for k in ("dt", "dt_prev"):
            _ulp_close(out[k][sl].cpu().numpy(), g["p2_" + k][so], 2, f"obj {i} {k}")
        for k in ("logflux", "logflux_err"):
            np.testing.assert_allclose(out[k][sl].cpu().numpy(), g["p2_" + k][so], rtol=2e-6, atol=1e-7, err_msg=f"obj {i} {k}")


def test_p3_spectra(g):
    from applecider_b200 import preprocess as pp

    off = torch.from_numpy(g["p3_offsets"]).to(DEV)
    grid = pp.wave_grid()
    out = pp.resample_spectra(torch.from_numpy(g["p3_wl"]).to(DEV), torch.from_numpy(g["p3_fx"]).to(DEV), off, grid)
    exact = _ulp_close(out.cpu().numpy(), g["p3_out"], 1, "resampled spectra")
    assert exact > 0.999, f"only {exact*100:.3f}% of the resampled samples are bit-identical"


def test_p3_interval_index_is_exact():
    """The searchsorted interval index of EVERY grid point (the integer/indexing part of P3, SURVEY §8a: bit-exact) against
    numpy: sorted and unsorted inputs, duplicated wavelengths (side='left' must pick the first), non-finite samples dropped,
    grid points outside the sampled range (extrapolation: index 0 / n), a grid point exactly on a sample."""
    from applecider_b200 import preprocess as pp, synth
    from oracle import preprocess as op

    rng = np.random.default_rng(7)
    specs = synth.raw_spectra(60, seed=79)
    grid_np = pp.wave_grid().cpu().numpy()
    # edge cases
    s = specs[0].copy(); s[::7, 0] = s[1::7, 0][: len(s[::7])]  # duplicated wavelengths
    specs.append(s[np.argsort(s[:, 0])])
    s = specs[1].copy(); rng.shuffle(s)                            # unsorted input
    specs.append(s)
    s = specs[2].copy(); s[5, 1] = np.nan; s[9, 0] = np.inf; s[11, 1] = -np.inf   # non-finite samples
    specs.append(s)
    s = specs[3].copy(); s = s[(s[:, 0] > 5000) & (s[:, 0] < 7000)]               # grid extends past both ends
    specs.append(s)
    s = specs[4].copy(); s[40:60, 0] = grid_np[1000:1020].astype(np.float64)       # samples exactly on grid points
    specs.append(s[np.argsort(s[:, 0])])
    specs.append(np.array([[5000.0, 1.0], [np.nan, 2.0]]))                         # < 2 finite samples
    wl, off = pp.ragged([s[:, 0] for s in specs])
    fx, _ = pp.ragged([s[:, 1] for s in specs])
    out, idx = pp.resample_spectra(wl, fx, off, pp.wave_grid(), return_index=True)
    idx = idx.cpu().numpy()
    for i, s in enumerate(specs):
        ref = op.searchsorted_index(s[:, 0], s[:, 1], grid_np)
        assert np.array_equal(idx[i], ref), f"spectrum {i}: {(idx[i] != ref).sum()} interval indices differ from numpy searchsorted"
    keep = [i for i in range(len(specs) - 1) if i != 60]  # 60 = duplicated wavelengths: the order of ties (and so 0/0 slopes) is not defined
    ref_vals = np.stack([op.resample_spectrum(specs[i][:, 0], specs[i][:, 1], grid_np) for i in keep])
    _ulp_close(out[keep].cpu().numpy(), ref_vals, 1, "resampled values of the edge-case spectra")
    assert torch.isnan(out[-1]).all()


def test_p3_properties_large():
    from applecider_b200 import preprocess as pp, synth
    from oracle import preprocess as op

    specs = synth.raw_spectra(300, seed=78)
    wl, off = pp.ragged([s[:, 0] for s in specs])
    fx, _ = pp.ragged([s[:, 1] for s in specs])
    grid = pp.wave_grid()
    out = pp.resample_spectra(wl, fx, off, grid).cpu().numpy()
    ref = np.stack([op.resample_spectrum(s[:, 0], s[:, 1], grid.cpu().numpy()) for s in specs[:40]])
    _ulp_close(out[:40], ref, 1, "resample vs oracle")
    # size-independent properties: zero mean, unit MAD
    assert np.abs(out.mean(1)).max() < 1e-4
    med = np.median(out, 1, keepdims=True)
    assert np.abs(np.median(np.abs(out - med), 1) - 1.0).max() < 1e-5


def test_p4_cutouts(g):
    from applecider_b200 import preprocess as pp
    from oracle import preprocess as op

    img = torch.from_numpy(g["p4_img"]).to(DEV)
    m63 = pp.normalize_cutouts(img, "median").cpu().numpy()
    _ulp_close(m63, g["p4_median63"], 2, "median/std 63")
    m49 = pp.normalize_cutouts(img, "median", 49).cpu().numpy()
    assert m49.shape == (4, 3, 49, 49)
    _ulp_close(m49, g["p4_median49"], 2, "median/std crop 49")
    l2 = pp.normalize_cutouts(img, "L2").cpu().numpy()
    np.testing.assert_allclose(l2, g["p4_l2_63"], rtol=5e-6, atol=0, err_msg="L2")  # the reference's torch.norm sums in float32
    nb = pp.normalize_cutouts(img, "median_notebook").cpu().numpy()
    ref = np.stack([op.normalize_cutout(i, "median", 63, variant="notebook") for i in g["p4_img"]])
    _ulp_close(nb, ref, 2, "notebook median/std")
    # even crop geometry follows the reference's int((63-cs)/2) rule: 32 -> 33x33
    assert pp.normalize_cutouts(img, "L2", 32).shape == (4, 3, 33, 33)
    # other plane sizes: 64x64 (even pixel count: the notebook median averages the two middle values; exactly fills the
    # register-resident kernel) and 80x80 (the shared-memory kernel), constant planes, planes with ties
    rng = np.random.default_rng(5)
    for hw in (64, 80, 9):
        big = (rng.standard_normal((3, 3, hw, hw)) * 20 + 300).astype(np.float32)
        big[1, 0] = np.round(big[1, 0] / 8) * 8     # many ties
        big[2, 1] = 7.0                             # constant plane: std 0
        for name, variant in (("median", "dataset"), ("median_notebook", "notebook")):
            got = pp.normalize_cutouts(torch.from_numpy(big).to(DEV), name).cpu().numpy()
            want = np.stack([op.normalize_cutout(i, "median", hw, variant=variant) for i in big])
            _ulp_close(got, want, 2, f"{name} {hw}x{hw}")


def test_p5_feature_stats(g):
    from applecider_b200 import preprocess as pp

    m, s = pp.feature_stats(torch.from_numpy(g["p5_data"]).to(DEV))
    np.testing.assert_allclose(m.cpu().numpy(), g["p5_mean"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(s.cpu().numpy(), g["p5_std"], rtol=1e-5, atol=2e-6)
    big = torch.randn(200_000, 46, device=DEV) * 3 + 1
    m, s = pp.feature_stats(big)
    np.testing.assert_allclose(m.cpu().numpy(), big.double().mean(0).float().cpu().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(s.cpu().numpy(), big.double().std(0, unbiased=False).float().cpu().numpy(), rtol=1e-4)


@pytest.mark.parametrize("log1p_dt", [False, True])
def test_pad_collate_dict_to_model_inputs(log1p_dt):
    """acb_collate_events vs the numpy restatement: gather of the 7 model channels, normalisation, mask polarity (bit-exact mask)."""
    import numpy as np
    from applecider_b200.fusion import from_pad_collate
    from oracle import preprocess as op

    rng = np.random.default_rng(3)
    samples = []
    for T in [1, 17, 108, 40, 0 + 5]:
        ev = rng.normal(size=(T, 14)).astype(np.float32)
        ev[:, 0] = np.sort(rng.uniform(0, 100, T)); ev[:, 1] = np.abs(rng.normal(size=T))
        samples.append({"events": ev, "image": rng.normal(size=(3, 63, 63)).astype(np.float32), "metadata": rng.normal(size=46).astype(np.float32), "label": T % 5})
    batch = op.pad_collate(samples)
    mean, std = rng.normal(size=4).astype(np.float32), np.abs(rng.normal(size=4)).astype(np.float32) + 0.1
    rx, rpad, rmeta, rimg = op.pad_collate_to_model_inputs(batch, mean, std, log1p_dt=log1p_dt)
    tb = {k: torch.from_numpy(v) for k, v in batch.items()}
    x, pad, meta, img, label = from_pad_collate(tb, mean, std, log1p_dt=log1p_dt)
    assert torch.equal(pad.cpu(), torch.from_numpy(rpad))
    assert torch.equal(x[..., 4:].cpu(), torch.from_numpy(rx[..., 4:]))  # one-hot columns pass through untouched
    assert_close(x, torch.from_numpy(rx), 2e-6, "normalised channels")
    assert torch.equal(meta.cpu(), torch.from_numpy(rmeta)) and torch.equal(img.cpu(), torch.from_numpy(rimg))
    assert meta.shape == (5, 24) and label.tolist() == [s["label"] for s in samples]
