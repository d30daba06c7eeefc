"""Fused optimiser step: gradient-norm clip + Adam/AdamW over flat buffers, no host sync.

Parameters are re-seated as views of ONE flat fp32 buffer (ordered by parameter group), their ``.grad`` as views of
a matching flat gradient buffer (the buffer ``ddp.FlatGradSync`` all-reduces), and the Adam moments live in two
more flat buffers.  ``step()`` is two kernels: ``acb_sumsq`` (only when clipping) and ``acb_adam_step``; the latter
can also refresh a flat bf16 shadow of the weights so the tcgen05 GEMMs need no per-parameter casts.

Same update rule as torch.optim.Adam / AdamW (amsgrad/maximize off) and torch.nn.utils.clip_grad_norm_; reference
call sites: HyraxBaselineCLS.py:15,108-120 (Adam 1e-4, clip 1.0), :228 (AdamW), astrominn.py:151-218 (11-group AdamW,
eps 5e-10), brew_cider.py:1211 (Adam, weight_decay 0.01).
"""
from __future__ import annotations

import ctypes

import torch

from ._lib import call
from .ddp import FlatGradSync

_ALIGN = 8  # elements: every tensor starts 16-byte aligned in the bf16 shadow too (TMA needs it), 32-byte in fp32


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, max_grad_norm=None,
                 bf16_shadow=False, process_group=None):
        groups = list(params)
        if not groups:
            raise ValueError("optimizer got an empty parameter list")
        if not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, decoupled=decoupled)
        self.param_groups = []
        for g in groups:
            pg = dict(self.defaults)
            pg.update({k: v for k, v in g.items() if k != "params"})
            pg["params"] = [p for p in g["params"] if p.requires_grad]
            if pg["params"]:
                self.param_groups.append(pg)
        if len(self.param_groups) > 16:
            raise ValueError("FusedAdam supports at most 16 parameter groups")
        self.max_grad_norm = max_grad_norm
        self.params = [p for g in self.param_groups for p in g["params"]]
        if len({id(p) for p in self.params}) != len(self.params):
            raise ValueError("some parameters appear in more than one parameter group")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("applecider_b200: FusedAdam needs CUDA parameters (no CPU fallback)")
        if any(p.dtype != torch.float32 for p in self.params):
            raise TypeError("FusedAdam: parameters must be fp32 (bf16 is a derived shadow copy)")
        # ---- flat layout ------------------------------------------------------------------------------
        self.offsets, off, self._group_end = [], 0, []
        for g in self.param_groups:
            for p in g["params"]:
                self.offsets.append(off)
                off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
            self._group_end.append(off)
        self.numel = off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                self.flat_p[o: o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o: o + p.numel()].view(p.shape)
        self.grads = FlatGradSync(self.params, process_group, offsets=self.offsets, numel=off)
        self.grads.broadcast_parameters(0, flat_params=self.flat_p)  # replicas start identical (no-op for one process)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        self._gnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.step_dev = None  # device int32 step counter (graph.GraphedTrainStep): overrides step_count inside the kernel
        self.flat_p16 = None
        if bf16_shadow:
            self.flat_p16 = torch.empty(off, dtype=torch.bfloat16, device=dev)
            self.refresh_shadow()

    # ---- bf16 shadow -------------------------------------------------------------------------------------
    def refresh_shadow(self):
        """(Re)build the bf16 copy, e.g. after load_state_dict changed the weights behind the optimiser's back."""
        if self.flat_p16 is None:
            return
        call("acb_cast", self.flat_p, 0, self.flat_p16, 1, self.numel)
        self._tag_shadow()

    def _tag_shadow(self):
        for p, o in zip(self.params, self.offsets):
            p._acb_bf16 = (self.flat_p16[o: o + p.numel()].view(p.shape), p._version)

    # ---- torch.optim-like surface -----------------------------------------------------------------------
    def zero_grad(self, set_to_none=False):
        self.grads.zero()

    def grad_norm(self):
        """Device scalar: the (pre-clip) global gradient norm of the last step()."""
        return self._gnorm.sqrt()

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        self.step_count += 1
        g = self.grads.flat
        gn = None
        if self.max_grad_norm is not None:
            call("acb_sumsq", g, self.numel, self._gnorm, 0)
            gn = self._gnorm
        n = len(self.param_groups)
        ends = (ctypes.c_longlong * n)(*self._group_end)
        hyper = (ctypes.c_float * (6 * n))()
        for i, pg in enumerate(self.param_groups):
            hyper[6 * i: 6 * i + 6] = [pg["lr"], pg["betas"][0], pg["betas"][1], pg["eps"], pg["weight_decay"], 1.0 if pg["decoupled"] else 0.0]
        call("acb_adam_step", self.flat_p, g, self.flat_m, self.flat_v, self.flat_p16, self.numel, n, ends, hyper, self.step_count, self.step_dev, gn,
             float(self.max_grad_norm or 0.0), float(grad_scale))
        torch.autograd.graph.increment_version(self.params)  # derived-weight caches key on the version counter
        if self.flat_p16 is not None:
            self._tag_shadow()

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.flat_m.clone(), "exp_avg_sq": self.flat_v.clone(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.flat_m.copy_(sd["exp_avg"])
        self.flat_v.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)


def fused_from_torch(opt: torch.optim.Optimizer, max_grad_norm=None, bf16_shadow=False, process_group=None) -> FusedAdam:
    """Build the fused equivalent of a torch.optim.Adam / AdamW instance (same groups and hyper-parameters)."""
    if not isinstance(opt, (torch.optim.Adam, torch.optim.AdamW)):
        raise TypeError(f"no fused equivalent for {type(opt).__name__}")
    decoupled = isinstance(opt, torch.optim.AdamW) or bool(opt.defaults.get("decoupled_weight_decay", False))
    groups = []
    for g in opt.param_groups:
        if g.get("amsgrad") or g.get("maximize"):
            raise NotImplementedError("amsgrad / maximize are not implemented")
        groups.append(dict(params=g["params"], lr=float(g["lr"]), betas=tuple(g["betas"]), eps=float(g["eps"]),
                           weight_decay=float(g["weight_decay"]), decoupled=decoupled))
    return FusedAdam(groups, max_grad_norm=max_grad_norm, bf16_shadow=bf16_shadow, process_group=process_group)
