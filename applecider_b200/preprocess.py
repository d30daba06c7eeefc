"""Array-level preprocessing on the GPU: thin wrappers over the C-ABI kernels.

All functions take CUDA tensors (ragged inputs as concatenated arrays + int64 offsets) and return CUDA
tensors; file I/O, pandas joins and manifests of the reference's preprocessing stay out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import call


def _offsets(lengths, device):
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    return torch.from_numpy(off).to(device)


def ragged(arrays, device="cuda", dtype=None):
    """list of (n_i, ...) numpy arrays -> (concatenated CUDA tensor, offsets int64 CUDA tensor)."""
    cat = np.concatenate(arrays, 0)
    t = torch.from_numpy(np.ascontiguousarray(cat))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(device), _offsets([len(a) for a in arrays], device)


def prep_lightcurves(raw, offsets, horizon, mean, std, max_len=257):
    """P1 (photo_dataset.py:85-152 + HyraxBaselineCLS.py:157) -> x[B,max_len,7] f32, pad mask bool, lengths."""
    B = offsets.numel() - 1
    x = torch.empty((B, max_len, 7), dtype=torch.float32, device=raw.device)
    mask = torch.empty((B, max_len), dtype=torch.uint8, device=raw.device)
    lengths = torch.empty(B, dtype=torch.int32, device=raw.device)
    call("acb_prep_lightcurve", raw.contiguous().float(), offsets, B, float(horizon), mean.float().contiguous(), std.float().contiguous(),
         max_len, x, mask, lengths)
    return x, mask.view(torch.bool), lengths


def prep_events(mjd, mag, magerr, fid, offsets, delta_t_hours=12.0):
    """P2 (preprocess_multimodal.py:84-111,176-180,291-336) -> dict of per-detection-slot arrays + n_events[B]."""
    B = offsets.numel() - 1
    total = mjd.numel()
    dev = mjd.device
    tmp = torch.empty(3 * total, dtype=torch.float64, device=dev)
    tmp_b = torch.empty(total, dtype=torch.int8, device=dev)
    out = {k: torch.empty(total, dtype=torch.float32, device=dev) for k in ("dt", "dt_prev", "logflux", "logflux_err")}
    out["band_id"] = torch.empty(total, dtype=torch.int8, device=dev)
    n_events = torch.empty(B, dtype=torch.int32, device=dev)
    call("acb_prep_events", mjd.double().contiguous(), mag.double().contiguous(), magerr.double().contiguous(), fid.int().contiguous(),
         offsets, B, total, float(delta_t_hours) / 24.0, tmp, tmp_b, out["dt"], out["dt_prev"], out["band_id"], out["logflux"],
         out["logflux_err"], n_events)
    out["n_events"] = n_events
    return out


def wave_grid(lo=4500.0, hi=7980.0, step=1.0, device="cuda"):
    """Config.wave_grid() of the reference (preprocess_multimodal.py:66-68): float32 linspace."""
    n = int(round((hi - lo) / step)) + 1
    return torch.from_numpy(np.linspace(lo, hi, n, dtype=np.float32)).to(device)


def resample_spectra(wl, fx, offsets, grid, max_n=None, return_index=False):
    """P3 (preprocess_multimodal.py:146-170,598-609) -> out[B, n_grid] f32 (mean 0, MAD 1)
    [, idx[B, n_grid] int32 = searchsorted(side='left') position of every grid point among the finite, sorted wavelengths]."""
    B = offsets.numel() - 1
    if max_n is None:
        max_n = int((offsets[1:] - offsets[:-1]).max().item())
    out = torch.empty((B, grid.numel()), dtype=torch.float32, device=wl.device)
    idx = torch.empty((B, grid.numel()), dtype=torch.int32, device=wl.device) if return_index else None
    call("acb_prep_spectrum_resample_idx", wl.double().contiguous(), fx.double().contiguous(), offsets, B, max_n, grid.float().contiguous(),
         grid.numel(), out, idx)
    return (out, idx) if return_index else out


_MODES = {"median": 0, "L2": 1, "l2": 1, "median_notebook": 2}


def normalize_cutouts(img, mode="median", cutout_size=None):
    """P4 (image_and_metadata_dataset.py:78-99) -> [B,C,S,S] f32."""
    B, C, H, W = img.shape
    cs = H if cutout_size is None else int(cutout_size)
    i1 = 0 if cs == H else int((H - cs) / 2)
    S = H - 2 * i1
    out = torch.empty((B, C, S, S), dtype=torch.float32, device=img.device)
    call("acb_prep_cutout_norm", img.contiguous().float(), B, C, H, W, cs, _MODES[mode], out)
    return out


def feature_stats(data):
    """P5 (preprocess_multimodal.py:863-895) -> (mean[F], std[F]) f32."""
    rows, F = data.shape
    work = torch.empty(2 * F, dtype=torch.float64, device=data.device)
    mean = torch.empty(F, dtype=torch.float32, device=data.device)
    std = torch.empty(F, dtype=torch.float32, device=data.device)
    call("acb_feature_stats", data.contiguous().float(), rows, F, work, mean, std)
    return mean, std
