"""Shared checker for the two-step train_step fixtures (tests/golden/train_steps.npz)."""
import os

import numpy as np
import torch


def load_steps(golden_dir, prefix):
    z = np.load(os.path.join(golden_dir, "train_steps.npz"))
    out = {"delta": {}, "dnorm": {}, "g1": {}}
    for k in z.files:
        if not k.startswith(prefix + "/"):
            continue
        parts = k.split("/", 2)
        if parts[1] in out and len(parts) == 3:
            out[parts[1]][parts[2]] = torch.from_numpy(np.asarray(z[k]))
        else:
            out[parts[1]] = torch.from_numpy(np.asarray(z[k]))
    return out


def check_two_steps(g, params, before, losses, loss_tol, solid_tol, grad_floor=1e-6):
    """losses: the two values train_step returned; params/before: name -> tensor after / before the two steps.

    Adam's first update is lr * g / (|g| + eps): it divides every element by its own magnitude, so an absolute gradient error e
    (the kernels meet 5e-4 * max|g| per tensor, tests/test_gpu_train.py) becomes a relative update error e / |g_i| -- sign-like
    at the noise floor.  Element-wise comparison of the recorded update head is therefore restricted to 'solid' elements
    (|g1| > grad_floor * max|g1| of that tensor; grad_floor 1e-6 for the bit-faithful CPU oracle, 0.05 on the GPU so that
    5e-4 / 0.05 = 1 % stays inside solid_tol); the L2 norm of the whole update is compared for every tensor (sign flips of
    noise-floor elements leave it unchanged)."""
    for got, key in zip(losses, ("loss1", "loss2")):
        ref = float(g[key])
        assert abs(got - ref) <= loss_tol * max(1.0, abs(ref)), f"{key}: {got} vs reference {ref}"
    worst = ("", 0.0)
    for n, ref_d in g["delta"].items():
        d = (params[n].detach().float().cpu() - before[n].float().cpu()).flatten()
        assert torch.isfinite(d).all(), n
        rn = float(g["dnorm"][n])
        assert abs(float(d.double().norm()) - rn) <= solid_tol * max(rn, 1e-12), f"{n}: update norm {float(d.norm()):.4e} vs reference {rn:.4e}"
        g1 = g["g1"][n]
        solid = g1.abs() > grad_floor * max(float(g1.abs().max()), 1e-30)
        if solid.any():
            h = d[: ref_d.numel()]
            scale = float(ref_d.abs().max())
            err = float((h[solid] - ref_d[solid]).abs().max())
            if scale > 0 and err / scale > worst[1]:
                worst = (n, err / scale)
            assert err <= solid_tol * scale + 1e-9, f"{n}: update differs from the reference train_step by {err:.3e} (scale {scale:.3e})"
    return worst
