"""numpy restatement of the array-level preprocessing (TEST INFRASTRUCTURE — see oracle/__init__.py).

Reference anchors (relative to /root/reference):
  P1  datasets/photo_dataset.py:85-101 (features), :117-152 (collate), models/HyraxBaselineCLS.py:157 (normalise)
  P2  preprocessing_utils/preprocess_multimodal.py:84-111 (_merge_jit), :176-180 (mag_to_flux),
      :291-312 (merge_by_filter), :315-336 (build_event_features numeric core)
  P3  preprocess_multimodal.py:146-170 (_interp_with_extrap), :135-143 (_mad), :598-609 (scaling)
  P4  datasets/image_and_metadata_dataset.py:78-99 (crop + median/std | L2);
      docs/pre_executed/Fusion_Dataset.ipynb cell 0 (_center_crop_chw_np, _normalize_image)
  P5  preprocess_multimodal.py:863-895 (streaming sum / sum-of-squares -> mean, std)
Pinned against the real functions by tests/test_oracle_preprocess.py (build container) and
tests/golden/preprocess.npz.
"""
from __future__ import annotations

import numpy as np

LOG_CONST = 1.0 / np.log(10)
BAND_ORDER = (1, 3, 2)  # groupby("filter") iterates 'ztfg','ztfi','ztfr' -> fid 1,3,2 (only matters for exact time ties)


# ---- P1 --------------------------------------------------------------------------------------------
def lightcurve_features(raw: np.ndarray, horizon: float) -> np.ndarray:
    """raw (L,5) [dt, dt_prev, band, logf, logfe] -> (L',7) after the horizon cut."""
    d = raw[raw[:, 0] <= horizon]
    vec4 = np.stack([np.log1p(d[:, 0]), np.log1p(d[:, 1]), d[:, 3], d[:, 4]], 1)
    one_hot = np.eye(3, dtype=np.float32)[d[:, 2].astype(np.int64)]
    return np.concatenate([vec4, one_hot], 1)


def collate_photometry(seqs, mean, std, max_len: int = 257):
    """Pad to >= max_len, truncate to exactly max_len, mask True = padding, normalise channels 0..3."""
    lengths = [s.shape[0] for s in seqs]
    L = max([max_len, max(lengths)])
    pad = np.stack([np.pad(s, ((0, L - s.shape[0]), (0, 0)), mode="constant", constant_values=0.0) for s in seqs], 0)
    mask = np.stack([np.concatenate([np.zeros(n), np.ones(L - n)]) for n in lengths]).astype(bool)
    pad, mask = pad[:, :max_len, :], mask[:, :max_len]
    pad[..., :4] = (pad[..., :4] - mean) / (std + 1e-8)
    return pad, mask


# ---- P2 --------------------------------------------------------------------------------------------
def mag_to_flux(mag, magerr):
    flux = 10 ** (-0.4 * (mag - 23.9))
    return flux, (magerr / (2.5 / np.log(10))) * flux


def merge_window(time, flux, err, dt_days, eps=1e-8):
    """Greedy window merge: anchor t0, absorb points while t - t0 <= dt_days; weights 1/(err+eps)."""
    t_out, f_out, e_out = [], [], []
    i, n = 0, len(time)
    while i < n:
        t0, j = time[i], i
        while j + 1 < n and time[j + 1] - t0 <= dt_days:
            j += 1
        totw = 0.0
        for k in range(i, j + 1):
            totw += 1.0 / (err[k] + eps)
        tw = fw = ew = 0.0
        for k in range(i, j + 1):
            w = (1.0 / (err[k] + eps)) / totw
            tw += w * time[k]
            fw += w * flux[k]
            ew += w * err[k]
        t_out.append(tw)
        f_out.append(fw)
        e_out.append(ew)
        i = j + 1
    return np.asarray(t_out, np.float64), np.asarray(f_out, np.float64), np.asarray(e_out, np.float64)


def event_features(mjd, mag, magerr, fid, delta_t_hours: float = 12.0):
    """mjd (relative days, f64), mag/magerr f64, fid in {1,2,3} -> merged event table sorted by time.

    Returns dict(dt, dt_prev, band_id, logflux, logflux_err) with the reference dtypes (f32 / int8)."""
    flux, ferr = mag_to_flux(np.asarray(mag, np.float64), np.asarray(magerr, np.float64))
    ts, fs, es, bs = [], [], [], []
    for band in BAND_ORDER:
        m = np.asarray(fid) == band
        if not m.any():
            continue
        order = np.argsort(np.asarray(mjd)[m], kind="stable")
        t, f, e = merge_window(np.asarray(mjd, np.float64)[m][order], flux[m][order], ferr[m][order], delta_t_hours / 24.0)
        ts.append(t)
        fs.append(f)
        es.append(e)
        bs.append(np.full(len(t), band - 1, np.int8))
    t, f, e, b = np.concatenate(ts), np.concatenate(fs), np.concatenate(es), np.concatenate(bs)
    order = np.argsort(t, kind="stable")
    t, f, e, b = t[order], f[order], e[order], b[order]
    dt = t - t[0]
    dt_prev = np.diff(np.r_[t[0], t])
    f32 = np.clip(f.astype(np.float32), 1e-6, None)
    logf = np.log10(f32)
    sig = e.astype(np.float32) * LOG_CONST / f32  # float64 maths (LOG_CONST is a float64 scalar), rounded below
    return {
        "dt": dt.astype(np.float32), "dt_prev": dt_prev.astype(np.float32), "band_id": b,
        "logflux": logf.astype(np.float32), "logflux_err": sig.astype(np.float32),
    }


# ---- P3 --------------------------------------------------------------------------------------------
def wave_grid(lo=4500.0, hi=7980.0, step=1.0):
    n = int(round((hi - lo) / step)) + 1
    return np.linspace(lo, hi, n, dtype=np.float32)


def interp_with_extrap(x, y, xnew):
    """Linear interpolation with linear extrapolation (scipy interp1d(kind='linear', fill_value='extrapolate'))."""
    x, y, xnew = np.asarray(x, np.float64), np.asarray(y, np.float64), np.asarray(xnew, np.float64)
    order = np.argsort(x)
    x, y = x[order], y[order]
    m = np.isfinite(x) & np.isfinite(y)
    x, y = x[m], y[m]
    if len(x) < 2:
        return np.full_like(xnew, np.nan)
    hi = np.clip(np.searchsorted(x, xnew), 1, len(x) - 1)  # side='left'
    lo = hi - 1
    slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
    return slope * (xnew - x[lo]) + y[lo]


def searchsorted_index(x, y, xnew):
    """The integer part of interp_with_extrap: np.searchsorted(side='left') of every new point among the finite samples sorted
    by wavelength (scipy interp1d then uses the interval [idx-1, idx] clipped to [1, n-1]; preprocess_multimodal.py:146-170).
    -1 everywhere when fewer than two finite samples remain."""
    x, y, xnew = np.asarray(x, np.float64), np.asarray(y, np.float64), np.asarray(xnew, np.float64)
    order = np.argsort(x)
    x, y = x[order], y[order]
    x = x[np.isfinite(x) & np.isfinite(y)]
    if len(x) < 2:
        return np.full(xnew.shape, -1, np.int32)
    return np.searchsorted(x, xnew).astype(np.int32)


def resample_spectrum(wl, fx, grid):
    yg = interp_with_extrap(wl, fx, np.asarray(grid, np.float64))
    mean = float(np.nanmean(yg))
    med = np.nanmedian(yg)
    mad = float(np.nanmedian(np.abs(yg - med)))
    if not np.isfinite(mad) or mad == 0.0:
        std = float(np.nanstd(yg))
        scale = std if (np.isfinite(std) and std > 0) else 1.0
    else:
        scale = mad
    return ((yg - mean) / scale).astype(np.float32)


# ---- P4 --------------------------------------------------------------------------------------------
def crop_indices(cutout_size: int, full: int = 63):
    if cutout_size == full:
        return 0, full
    i1 = int((full - cutout_size) / 2)
    return i1, int(full - i1)


def normalize_cutout(img: np.ndarray, mode: str, cutout_size: int = 63, variant: str = "dataset") -> np.ndarray:
    """img (3,63,63) f32.  variant 'dataset' = ImageAndMetadataDataset.get_image (torch.median = lower median,
    unbiased std, +1e-8); variant 'notebook' = Fusion_Dataset._normalize_image (np.median, population std, <=1e-8 -> 1)."""
    i1, i2 = crop_indices(cutout_size, img.shape[-1])
    x = np.array(img[:, i1:i2, i1:i2], dtype=np.float32, copy=True)
    if mode == "median":
        for c in range(x.shape[0]):
            flat = x[c].reshape(-1)
            if variant == "dataset":
                med = np.sort(flat)[(flat.size - 1) // 2]
                p = x[c] - med
                x[c] = p / (np.float32(p.astype(np.float64).std(ddof=1)) + np.float32(1e-8))
            else:
                p = x[c] - np.median(flat)
                s = float(p.std())
                x[c] = p / (s if (np.isfinite(s) and s > 1e-8) else 1.0)
    elif mode in ("L2", "l2"):
        n = np.float32(np.sqrt((x.astype(np.float64) ** 2).sum()))
        if variant == "notebook" and not (np.isfinite(n) and n > 1e-8):
            n = np.float32(1.0)
        x = x / n
    return x.astype(np.float32)


# ---- P5 --------------------------------------------------------------------------------------------
def feature_stats(chunks):
    """Streaming column sums over a list of (T_i, F) arrays -> (mean, std) f32 (population std, clipped at 0)."""
    s = sq = None
    total = 0
    for d in chunks:
        if d.size == 0:
            continue
        s = d.sum(axis=0) if s is None else s + d.sum(axis=0)
        sq = (d**2).sum(axis=0) if sq is None else sq + (d**2).sum(axis=0)
        total += d.shape[0]
    mean = s / total
    std = np.sqrt(np.clip(sq / total - mean**2, 0, None))
    return mean.astype(np.float32), std.astype(np.float32)


# ---- pad_collate dict format -> model inputs (Fusion_Dataset.ipynb cell 0) ---------------------------
EVENT_COLUMNS = ["dt", "dt_prev", "band_id", "logflux", "logflux_err", "band_ztfg", "band_ztfr", "band_ztfi",
                 "g_r", "g_r_err", "r_i", "r_i_err", "has_g_r", "has_r_i"]  # preprocess_multimodal.py:324-365 minus :680's drops


def pad_collate(samples, pad_value=0.0):
    """MultiModalDataset.pad_collate restated with numpy: list of dicts(events[T,Fe], image, metadata, label)."""
    B = len(samples)
    Tmax = max(s["events"].shape[0] for s in samples)
    Fe = samples[0]["events"].shape[1]
    ev = np.full((B, Tmax, Fe), pad_value, dtype=np.float32)
    mask = np.zeros((B, Tmax), dtype=bool)
    for i, s in enumerate(samples):
        T = s["events"].shape[0]
        ev[i, :T] = s["events"]
        mask[i, :T] = True
    return {"events": ev, "events_mask": mask, "image": np.stack([s["image"] for s in samples]),
            "metadata": np.stack([s["metadata"] for s in samples]), "label": np.asarray([s["label"] for s in samples])}


def pad_collate_to_model_inputs(batch, mean, std, log1p_dt=False, n_meta=24):
    """The 7 model channels of the 14 event columns, normalised like HyraxBaselineCLS.to_tensor (:157), and the
    key-padding mask in the encoder's polarity (True = padding)."""
    idx = [EVENT_COLUMNS.index(c) for c in ("dt", "dt_prev", "logflux", "logflux_err", "band_ztfg", "band_ztfr", "band_ztfi")]
    x = batch["events"][..., idx].astype(np.float32).copy()
    if log1p_dt:
        x[..., :2] = np.log1p(x[..., :2])
    x[..., :4] = (x[..., :4] - np.asarray(mean, np.float32)) / (np.asarray(std, np.float32) + np.float32(1e-8))
    return x, ~batch["events_mask"], batch["metadata"][:, :n_meta].astype(np.float32), batch["image"].astype(np.float32)
