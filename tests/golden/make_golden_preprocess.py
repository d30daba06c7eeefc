"""Generate tests/golden/preprocess.npz from the REAL reference preprocessing functions (build container only).

    python tests/golden/make_golden_preprocess.py

Calls, unmodified: PhotoEventsDataset.get_photometry/collate + HyraxBaselineCLS.to_tensor (P1),
merge_by_filter + build_event_features (P2), preprocess_spectra_df (P3),
ImageAndMetadataDataset.get_image (P4), and the sum/sum-of-squares maths of
compute_feature_stats_safe (P5, restated inline because the function only works on files).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from applecider_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402


def ref_p1(raws, horizon, mean, std):
    pd_mod = ref_loader.ref_photo_dataset()
    R = ref_loader.ref_models()
    ds = object.__new__(pd_mod.PhotoEventsDataset)
    ds.use_oversampling = False
    ds.horizon = horizon
    items = []
    with tempfile.TemporaryDirectory() as td:
        ds.filenames = []
        for i, r in enumerate(raws):
            p = os.path.join(td, f"{i}.npz")
            np.savez(p, data=r)
            ds.filenames.append(p)
        for i in range(len(raws)):
            items.append({"data": {"photometry": ds.get_photometry(i), "label": 0, "mean": mean, "std": std}})
    batch = pd_mod.PhotoEventsDataset.collate(items)
    x, mask, _ = R.photo.HyraxBaselineCLS.to_tensor(batch)
    return np.asarray(x, np.float32), np.asarray(mask)


def ref_p2(mjd, mag, magerr, fid):
    pm = ref_loader.ref_preprocess()
    flux, ferr = pm.mag_to_flux(mag, magerr)
    df = pd.DataFrame({"obj_id": "o", "jd": mjd + 2458000.5, "mjd": mjd, "mag": mag, "magerr": magerr, "flux": flux,
                       "flux_error": ferr, "fid": fid, "filter": [pm.FID2BAND[int(f)] for f in fid]})
    ev = pm.build_event_features(pm.merge_by_filter(df, 12.0))
    return {k: ev[k].to_numpy() for k in ("dt", "dt_prev", "band_id", "logflux", "logflux_err")}


def ref_p3(wl, fx):
    pm = ref_loader.ref_preprocess()
    cfg = pm.Config(data_dir=".", spec_csv=".", output_root=".")
    return pm.preprocess_spectra_df(pd.DataFrame({"wavelength": wl, "flux": fx}), cfg.wave_grid())


def ref_p4(img, mode, cutout_size):
    ref_loader.install_stubs()
    import importlib

    mod = importlib.import_module("applecider.datasets.image_and_metadata_dataset")
    ds = object.__new__(mod.ImageAndMetadataDataset)
    ds.use_oversampling = False
    ds.enable_cache = False
    ds.dataset_config = {"tags": [], "patch_size": [32, 32], "cutout_size": cutout_size, "image_norm": mode}
    ds.raw_files = [{"image": torch.from_numpy(img.copy())}]
    return ds.get_image(0).numpy()


def synth_detections(n_obj, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_obj):
        n = int(rng.integers(3, 120))
        # clustered nights: several exposures within hours, nights days apart
        nights = np.sort(rng.uniform(0, 90, size=max(1, n // 3)))
        mjd = np.sort(rng.choice(nights, size=n) + rng.uniform(0, 0.6, size=n))
        mjd = mjd - mjd.min()
        fid = rng.choice([1, 2, 3], size=n, p=[0.45, 0.45, 0.1])
        mag = rng.normal(19.0, 0.8, size=n)
        magerr = np.abs(rng.normal(0.08, 0.04, size=n)) + 0.005
        out.append((mjd, mag, magerr, fid))
    return out


def main():
    out = {}
    # ---- P1 ----
    raws = synth.raw_light_curves(12, seed=51, max_len=400)
    raws[3] = raws[3][:1]
    mean = np.array([2.9, 0.9, 1.5, 0.08], np.float32)
    std = np.array([1.1, 0.8, 0.5, 0.04], np.float32)
    x, mask = ref_p1(raws, 60.0, mean, std)
    out["p1_offsets"] = np.cumsum([0] + [len(r) for r in raws]).astype(np.int64)
    out["p1_raw"] = np.concatenate(raws, 0)
    out["p1_mean"], out["p1_std"], out["p1_horizon"] = mean, std, np.float32(60.0)
    out["p1_x"], out["p1_mask"] = x, mask
    # ---- P2 ----
    dets = synth_detections(10, seed=52)
    out["p2_offsets"] = np.cumsum([0] + [len(d[0]) for d in dets]).astype(np.int64)
    out["p2_mjd"] = np.concatenate([d[0] for d in dets])
    out["p2_mag"] = np.concatenate([d[1] for d in dets])
    out["p2_magerr"] = np.concatenate([d[2] for d in dets])
    out["p2_fid"] = np.concatenate([d[3] for d in dets]).astype(np.int32)
    evs = [ref_p2(*d) for d in dets]
    out["p2_out_offsets"] = np.cumsum([0] + [len(e["dt"]) for e in evs]).astype(np.int64)
    for k in ("dt", "dt_prev", "band_id", "logflux", "logflux_err"):
        out[f"p2_{k}"] = np.concatenate([e[k] for e in evs])
    # ---- P3 ----
    specs = synth.raw_spectra(6, seed=53)
    specs[1] = specs[1][::-1].copy()  # unsorted input
    specs[2] = np.stack([np.linspace(5000, 7000, 40), np.sin(np.linspace(0, 9, 40))], 1)  # needs extrapolation both sides
    specs[3][5, 1] = np.nan  # non-finite sample is dropped
    out["p3_offsets"] = np.cumsum([0] + [len(s) for s in specs]).astype(np.int64)
    out["p3_wl"] = np.concatenate([s[:, 0] for s in specs])
    out["p3_fx"] = np.concatenate([s[:, 1] for s in specs])
    out["p3_out"] = np.stack([ref_p3(s[:, 0], s[:, 1]) for s in specs])
    # ---- P4 ----
    img = synth.cutouts(4, seed=54, normalise=False).numpy()
    out["p4_img"] = img
    out["p4_median63"] = np.stack([ref_p4(i, "median", 63) for i in img])
    out["p4_l2_63"] = np.stack([ref_p4(i, "L2", 63) for i in img])
    out["p4_median49"] = np.stack([ref_p4(i, "median", 49) for i in img])
    # ---- P5 ---- (maths of compute_feature_stats_safe:863-895 on in-memory chunks)
    rng = np.random.default_rng(55)
    chunks = [rng.normal(1.0, 2.0, size=(int(rng.integers(1, 300)), 14)).astype(np.float32) for _ in range(9)]
    s = sq = None
    tot = 0
    for d in chunks:
        s = d.sum(axis=0) if s is None else s + d.sum(axis=0)
        sq = (d**2).sum(axis=0) if sq is None else sq + (d**2).sum(axis=0)
        tot += d.shape[0]
    m = s / tot
    out["p5_data"] = np.concatenate(chunks, 0)
    out["p5_mean"] = m.astype(np.float32)
    out["p5_std"] = np.sqrt(np.clip(sq / tot - m**2, 0, None)).astype(np.float32)
    path = os.path.join(HERE, "preprocess.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path)/1024:.1f} KiB)")


if __name__ == "__main__":
    main()
