"""Shared helpers for the parity tests."""
import numpy as np
import torch


def load_golden(golden_dir, name):
    import os

    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def assert_close(got, ref, rtol_max, name="", atol=0.0):
    """|got-ref| <= rtol_max * max(1, |ref|_inf) + atol, with a diagnostic on failure."""
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{name}: non-finite values in result ({(~torch.isfinite(got)).sum().item()} of {got.numel()})"
    scale = max(1.0, ref.abs().max().item())
    err = (got - ref).abs()
    tol = rtol_max * scale + atol
    if err.max().item() > tol:
        bad = err > tol
        idx = bad.nonzero()[0].tolist()
        frac = bad.float().mean().item()
        raise AssertionError(
            f"{name}: max|err|={err.max().item():.4e} > tol={tol:.4e} (scale {scale:.3e}); {frac*100:.2f}% elements bad; "
            f"first bad index {idx} got={got[tuple(idx)].item():.6g} ref={ref[tuple(idx)].item():.6g}; "
            f"bad rows={sorted(set(bad.nonzero()[:, 0].tolist()))[:16] if bad.dim() > 1 else ''}"
        )
    return err.max().item() / scale
