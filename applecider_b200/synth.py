"""Deterministic synthetic weights and ZTF-shaped inputs.

Pure host-side helpers (torch CPU / numpy) shared by tests, bench.py and the
golden-vector generator.  No compute of the hot path happens here.

``det_state_dict`` builds a state_dict whose every tensor depends only on
(seed, key name, shape) so that 28 M-parameter checkpoints never have to be
committed: the golden generator (run against the real reference) and the GPU
parity tests regenerate byte-identical weights from the key names.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def det_tensor(seed: int, key: str, shape, dtype=torch.float32) -> torch.Tensor:
    """Deterministic 'trained-looking' values for a parameter called ``key``."""
    shape = tuple(shape)
    g = _gen(seed, key)
    leaf = key.rsplit(".", 1)[-1]
    if len(shape) >= 2 and leaf != "cls_tok":
        fan_in = int(np.prod(shape[1:]))
        t = torch.randn(shape, generator=g) * (1.0 / np.sqrt(max(fan_in, 1)))
    elif leaf == "gamma":  # ConvNeXt layer-scale: make the residual branch matter
        t = 0.2 + 0.05 * torch.randn(shape, generator=g)
    elif leaf == "cls_tok":
        t = 0.5 * torch.randn(shape, generator=g)
    elif leaf in ("w0", "w"):  # Time2Vec frequencies
        t = torch.randn(shape, generator=g)
    elif leaf in ("b0", "b"):
        t = 0.3 * torch.randn(shape, generator=g)
    elif leaf == "running_var":  # BatchNorm statistics of the legacy spectra encoder: strictly positive
        t = 0.5 + torch.rand(shape, generator=g)
    elif leaf == "weight":  # 1-D weight = a norm scale
        t = 1.0 + 0.1 * torch.randn(shape, generator=g)
    else:  # biases
        t = 0.05 * torch.randn(shape, generator=g)
    return t.to(dtype)


def det_state_dict(module: torch.nn.Module, seed: int = 0) -> dict:
    out = {}
    for k, v in module.state_dict().items():
        if v.dtype.is_floating_point:
            out[k] = det_tensor(seed, k, v.shape, v.dtype)
        else:
            out[k] = v.clone()
    return out


# ----------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------
CLASS_PRIORS = np.array([6426, 1629, 693, 321, 47], dtype=np.float64)  # _archive/files/cider_BTS.csv grouped


def light_curve_lengths(n: int, rng: np.random.Generator, max_len: int = 257) -> np.ndarray:
    ln = rng.lognormal(mean=np.log(40.0), sigma=0.9, size=n)
    return np.clip(np.round(ln), 1, max_len).astype(np.int64)


def raw_light_curves(n: int, seed: int = 1337, max_len: int = 257):
    """Ragged raw event arrays (L_i,5) = [dt, dt_prev, band, logf, logfe] (photo_dataset.py:85-101)."""
    rng = np.random.default_rng(seed)
    lens = light_curve_lengths(n, rng, max_len)
    out = []
    for L in lens:
        dt = np.sort(rng.uniform(0.0, 100.0, size=L))
        dt[0] = 0.0
        dt_prev = np.diff(np.r_[0.0, dt])
        band = rng.choice(3, size=L, p=[0.45, 0.45, 0.10]).astype(np.float64)
        logf = rng.normal(1.5, 0.5, size=L)
        logfe = np.abs(rng.normal(0.08, 0.04, size=L))
        out.append(np.stack([dt, dt_prev, band, logf, logfe], 1).astype(np.float32))
    return out


def photometry_batch(n: int, seed: int = 1337, L: int = 257, max_len: int = 257):
    """Hyrax-collated, normalised batch: x (n,L,7) f32, pad (n,L) bool True=pad, lengths."""
    raws = raw_light_curves(n, seed, max_len)
    x = np.zeros((n, L, 7), np.float32)
    pad = np.ones((n, L), bool)
    for i, r in enumerate(raws):
        k = min(len(r), L)
        x[i, :k, 0] = np.log1p(r[:k, 0])
        x[i, :k, 1] = np.log1p(r[:k, 1])
        x[i, :k, 2] = r[:k, 3]
        x[i, :k, 3] = r[:k, 4]
        x[i, np.arange(k), 4 + r[:k, 2].astype(np.int64)] = 1.0
        pad[i, :k] = False
    valid = ~pad
    mean = x[valid][:, :4].mean(0)
    std = x[valid][:, :4].std(0)
    x[..., :4] = (x[..., :4] - mean) / (std + 1e-8)  # applied to padded rows too (HyraxBaselineCLS.py:157)
    lens = valid.sum(1)
    return torch.from_numpy(x), torch.from_numpy(pad), torch.from_numpy(lens)


def cutouts(n: int, seed: int = 1337, normalise: bool = True) -> torch.Tensor:
    """(n,3,63,63) science/reference/difference cutouts: PSF + sky, optional median/std norm."""
    rng = np.random.default_rng(seed + 1)
    yy, xx = np.mgrid[0:63, 0:63].astype(np.float32)
    r2 = (yy - 31.0) ** 2 + (xx - 31.0) ** 2
    psf = np.exp(-r2 / (2 * 1.5**2)).astype(np.float32)
    amp = np.exp(rng.uniform(np.log(50.0), np.log(5000.0), size=(n, 1, 1))).astype(np.float32)
    sky = rng.uniform(100.0, 400.0, size=(n, 1, 1)).astype(np.float32)
    img = np.empty((n, 3, 63, 63), np.float32)
    for c in range(2):
        img[:, c] = amp * psf + sky + rng.standard_normal((n, 63, 63)).astype(np.float32) * np.sqrt(sky)
    img[:, 2] = 0.3 * amp * psf + rng.standard_normal((n, 63, 63)).astype(np.float32) * 30.0
    if normalise:
        flat = img.reshape(n, 3, -1)
        med = np.median(flat, axis=2, keepdims=True)
        flat = flat - med
        flat = flat / (flat.std(axis=2, ddof=1, keepdims=True) + 1e-8)
        img = flat.reshape(n, 3, 63, 63)
    return torch.from_numpy(img.astype(np.float32))


def spectra(n: int, seed: int = 1337, L: int = 4096) -> torch.Tensor:
    """(n,1,L) mean/MAD-scaled spectra: smooth continuum + 5 Gaussian lines + noise."""
    rng = np.random.default_rng(seed + 2)
    u = np.linspace(0.0, 1.0, L, dtype=np.float32)[None, :]
    a = rng.normal(0, 1, size=(n, 3)).astype(np.float32)
    y = 1.0 + a[:, :1] * u + a[:, 1:2] * u**2 + 0.3 * a[:, 2:3] * np.sin(6.28 * u)
    for _ in range(5):
        c = rng.uniform(0.05, 0.95, size=(n, 1)).astype(np.float32)
        w = rng.uniform(0.002, 0.02, size=(n, 1)).astype(np.float32)
        h = rng.normal(0, 1.0, size=(n, 1)).astype(np.float32)
        y = y + h * np.exp(-0.5 * ((u - c) / w) ** 2)
    y = y + rng.normal(0, 0.1, size=(n, L)).astype(np.float32)
    mean = y.mean(1, keepdims=True)
    med = np.median(y, axis=1, keepdims=True)
    mad = np.median(np.abs(y - med), axis=1, keepdims=True)
    y = (y - mean) / np.where(mad > 0, mad, 1.0)
    return torch.from_numpy(y.astype(np.float32))[:, None, :]


def metadata(n: int, seed: int = 1337, cols: int = 24, missing_frac: float = 0.05) -> torch.Tensor:
    rng = np.random.default_rng(seed + 3)
    m = rng.standard_normal((n, cols)).astype(np.float32)
    m[rng.uniform(size=(n, cols)) < missing_frac] = -999.0  # missing sentinel (preprocess_multimodal.py:720,743)
    return torch.from_numpy(m)


def labels(n: int, seed: int = 1337) -> torch.Tensor:
    rng = np.random.default_rng(seed + 4)
    return torch.from_numpy(rng.choice(5, size=n, p=CLASS_PRIORS / CLASS_PRIORS.sum()).astype(np.int64))


def raw_spectra(n: int, seed: int = 1337):
    """Ragged raw spectra for the resampling kernel: list of (n_i,2) f64 [wavelength, flux]."""
    rng = np.random.default_rng(seed + 5)
    out = []
    for _ in range(n):
        k = int(rng.integers(180, 2500))
        wl = np.sort(rng.uniform(3700.0, 9300.0, size=k))
        fx = 1.0 + 0.5 * np.sin(wl / 300.0) + rng.normal(0, 0.1, size=k)
        out.append(np.stack([wl, fx], 1))
    return out
