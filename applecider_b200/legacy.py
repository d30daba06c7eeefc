"""Legacy fusion-checkpoint compatibility: the "variant B" spectra encoder of
_archive/notebooks/brew_cider.py:585-708 (BatchNorm stages, 1x1 skip projection, max/avg/min tri-pool, flatten
12288 -> 2048 -> 256) on the same kernel families as the src SpectraNet.  Three paths:
  * eval + no_grad + fp32: BatchNorm folded into the conv GEMM epilogue (CUDA-core implicit GEMM; the round-1 parity path);
  * differentiable (autograd on, or train()): convs through train.SpectraConvs (forward AND backward on the C-ABI kernels),
    BatchNorm1d with batch statistics + running-statistics update (csrc/legacy_train.cu) fused with the skip add and the GELU,
    tri-pool with its backward -- so the archived fusion model can be fine-tuned;
  * bf16: stages 2-5 (144 .. 1152 input channels) on the tcgen05 implicit GEMM, stage 1 (one input channel, 16 outputs) on the
    fp32 conv, activations bf16, statistics fp32.  State-dict keys are the reference's
(`stage{1..5}.0.convs.{j}`, `.norm` incl. BatchNorm running statistics, `.proj`, `class_model.{0,1,4,5}`, `fc`), so the
archived `cider_weights/*.pth` spectra sub-dicts load with strict=True.  Activations are channels-last [B, L, C]."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .config import resolve_dtype
from .spectra import SpectraNetBlock

KERNEL_SIZES = [[3, 61, 1021], [3, 31, 251], [3, 15, 61], [3, 11, 31], [3, 7, 13]]
USE_LN = [False, False, False, False, True]
CHANNELS = [1, 16, 32, 64, 128, 256]


class SpectraNetBlockB(nn.Module):
    """convs x3 -> cat -> LayerNorm | BatchNorm1d(eval) -> + proj(x) -> GELU -> [max|avg|min pool 4]  (brew_cider.py:586-636)."""

    def __init__(self, in_channels, out_channels, kernel_sizes, use_skip=True, use_ln=True, do_pool=False):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_sizes = in_channels, out_channels, list(kernel_sizes)
        self.use_skip, self.use_ln, self.do_pool, self.k = use_skip, use_ln, do_pool, len(kernel_sizes)
        nc = out_channels * self.k
        self.convs = nn.ModuleList([nn.Conv1d(in_channels, out_channels, kernel_size=k, padding=k // 2) for k in kernel_sizes])
        self.norm = nn.LayerNorm(nc) if use_ln else nn.BatchNorm1d(nc)
        if use_skip:
            self.proj = nn.Conv1d(in_channels, nc, kernel_size=1)
        self._derived = ops.DerivedCache()

    # the multi-kernel conv machinery of the src block (packed tap-major weights, fp32 / tcgen05 implicit GEMMs)
    _kmax = SpectraNetBlock._kmax
    _packed_dt = SpectraNetBlock._packed
    _convs_f32 = SpectraNetBlock._convs_f32
    _convs_bf16 = SpectraNetBlock._convs_bf16

    def _packed(self, dtype=None):
        if dtype is not None:
            return self._packed_dt(dtype)

        def build():
            kmax, cin, cout = max(self.kernel_sizes), self.in_channels, self.out_channels
            w = torch.zeros((self.k * cout, kmax * cin), dtype=torch.float32, device=self.convs[0].weight.device)
            for j, c in enumerate(self.convs):
                kj = self.kernel_sizes[j]
                ops.call("acb_pack_conv_weight", c.weight, ops._offset_ptr(w, j * cout * kmax * cin), 0, cout, cin, kj, kmax * cin, kmax // 2 - kj // 2)
            return w, torch.cat([c.bias.detach() for c in self.convs]).contiguous()

        return self._derived.get("packed", [c.weight for c in self.convs] + [c.bias for c in self.convs], build)

    def _bn_affine(self):
        """Eval-mode BatchNorm as a per-channel affine: scale = w / sqrt(var + eps), shift = b - mean * scale."""
        n = self.norm

        def build():
            scale = n.weight.detach() / torch.sqrt(n.running_var + n.eps)
            return scale.contiguous(), (n.bias.detach() - n.running_mean * scale).contiguous()

        return self._derived.get("bn", [n.weight, n.bias, n.running_mean, n.running_var], build)

    def forward_diff(self, h, B, L, dtype, training):
        """Differentiable path.  h: [B, L, Cin] channels-last (fp32 for the 1-channel first stage, else `dtype`).
        Returns (activation [B, L', C'], L', C')."""
        from . import fn
        from .train import SpectraConvs

        cin, nc = self.in_channels, self.out_channels * self.k
        params = [c.weight for c in self.convs] + [c.bias for c in self.convs]
        y = SpectraConvs.apply(h, None, self, B, L, dtype, *params)  # [B*L, nc] in `dtype`
        res = None
        if self.use_skip:
            hp = h.reshape(B * L, cin)
            wp = self.proj.weight.view(nc, cin)
            if hp.dtype != torch.float32 and cin % 8 == 0:
                res = fn.linear(hp, wp, self.proj.bias)
            else:  # first stage: K = 1
                res = fn.linear(hp.float(), wp, self.proj.bias)
        if self.use_ln:
            z = fn.layernorm(y, self.norm.weight, self.norm.bias, self.norm.eps)
            if res is not None:
                z = fn.add(fn.cast(res, z.dtype), z)
            g = fn.act(z, ops.ACT_GELU)
        else:
            g = fn.BatchNormResAct.apply(y, res, self.norm.weight, self.norm.bias, self.norm, bool(training), ops.ACT_GELU, dtype)
        if not self.do_pool:
            return g.view(B, L, nc), L, nc
        return fn.TriPool.apply(g.view(B, L, nc), B, L, nc), L // 4, 3 * nc

    def forward_cl(self, x, B, L):
        from .fn import ew  # thin acb_ew wrapper (no autograd involved here)

        cin, cout, kmax, nc = self.in_channels, self.out_channels, max(self.kernel_sizes), self.out_channels * self.k
        w, bias = self._packed()
        y = torch.empty((B * L, nc), dtype=torch.float32, device=x.device)
        res, scale = None, None
        if not self.use_ln:
            # eval BatchNorm folded into the conv GEMM epilogue: y = (proj(x) + shift) + scale[col] * (conv(x) + bias)
            scale, shift = self._bn_affine()
            if self.use_skip:
                res = ops.gemm(x.view(B * L, cin), self.proj.weight.view(nc, cin), (self.proj.bias.detach() + shift).contiguous())
            else:
                res = shift.expand(B * L, nc).contiguous()
        for j in range(self.k):  # conv j only touches its own taps (same call pattern as SpectraNetBlock._convs_f32)
            kj = self.kernel_sizes[j]
            off = (kmax // 2 - kj // 2) * cin
            ops.call("acb_gemm_f32", x, ops._offset_ptr(w, j * cout * kmax * cin + off), ops._offset_ptr(y, j * cout), B * L, cout, kj * cin, cin,
                     kmax * cin, nc, L, cin, kj // 2, ops._offset_ptr(bias, j * cout), ops.ACT_NONE,
                     (ops._offset_ptr(res, j * cout) if res is not None else None), nc,
                     (ops._offset_ptr(scale, j * cout) if scale is not None else None), (ops.RES_ADD if res is not None else ops.RES_NONE))
        if self.use_ln:
            y = ops.layernorm(y, self.norm.weight, self.norm.bias, self.norm.eps)
            if self.use_skip:
                y = ew(ops.gemm(x.view(B * L, cin), self.proj.weight.view(nc, cin), self.proj.bias), y, 0)
        g = torch.empty_like(y)
        ops.call("acb_act_fwd", y, 0, g, 0, ops.ACT_GELU, y.numel())
        if not self.do_pool:
            return g.view(B, L, nc), L, nc
        out = torch.empty((B, L // 4, 3 * nc), dtype=torch.float32, device=x.device)
        ops.call("acb_tripool4_cl", g, out, B, L, nc)
        return out, L // 4, 3 * nc


class SpectraClassificationB(nn.Module):
    """forward(x[B,1,4096]) -> (B,256) embedding, or (B,num_classes) when config['mode'] == 'spectra' (brew_cider.py:638-705)."""

    def __init__(self, config=None, depths=(1, 1, 1, 1, 1), length=4096, compute_dtype=None):
        super().__init__()
        config = config or {"mode": "all", "classes": list(range(5))}
        self.compute_dtype = resolve_dtype(compute_dtype if compute_dtype is not None else config.get("compute_dtype"))
        if list(depths) != [1, 1, 1, 1, 1]:
            raise NotImplementedError("applecider_b200: the archived checkpoints use depths [1,1,1,1,1]")
        self.classification = config["mode"] == "spectra"
        cin = CHANNELS[0]
        for i in range(5):
            blk = SpectraNetBlockB(cin, CHANNELS[i + 1], KERNEL_SIZES[i], use_skip=True, use_ln=USE_LN[i], do_pool=(i != 4))
            setattr(self, f"stage{i + 1}", nn.Sequential(blk))
            cin = CHANNELS[i + 1] * 3 * (3 if i != 4 else 1)
        self.length = length // 256
        self.flat_dim = CHANNELS[5] * 3 * self.length
        self.class_model = nn.Sequential(nn.Linear(self.flat_dim, 2048), nn.LayerNorm(2048), nn.GELU(), nn.Dropout(0.5),
                                         nn.Linear(2048, 256), nn.LayerNorm(256), nn.GELU(), nn.Dropout(0.3))
        if self.classification:
            self.fc = nn.Linear(256, len(config["classes"]))
        self._derived = ops.DerivedCache()

    def _w0_channels_last(self):
        """class_model.0 consumes x.reshape(B, C*L) (C-major); our activations are [B, L, C]: permute the weight columns once."""
        lin = self.class_model[0]
        C, L = CHANNELS[5] * 3, self.length
        return self._derived.get("w0", [lin.weight], lambda: lin.weight.detach().view(-1, C, L).permute(0, 2, 1).reshape(-1, L * C).contiguous())

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("applecider_b200: inputs must be CUDA tensors (no CPU fallback)")
        B, c, L = x.shape
        assert c == 1 and L // 256 == self.length, "the legacy encoder expects (B, 1, 4096) spectra"
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if self.training or grad or self.compute_dtype != torch.float32:
            return self._forward_diff(x)
        with torch.no_grad():
            return self._forward_eval_f32(x)

    def _forward_diff(self, x):
        """BatchNorm with batch statistics in train(), running statistics in eval(); every op has a C-ABI backward."""
        from . import fn

        dtype = self.compute_dtype
        B, _, L = x.shape
        h = x.contiguous().float().view(B, L, 1)
        for i in range(5):
            h, L, C = getattr(self, f"stage{i + 1}")[0].forward_diff(h, B, L, dtype, self.training)
        z = h.reshape(B, L * C)
        cm = self.class_model
        Cc, Ll = CHANNELS[5] * 3, self.length
        # class_model.0 consumes x.reshape(B, C*L) (C-major); our rows are [L, C]: a differentiable re-layout of the weight columns
        w0 = cm[0].weight.view(-1, Cc, Ll).permute(0, 2, 1).reshape(-1, Ll * Cc)
        if dtype != torch.float32 and B >= 64:
            z = fn.linear(z, w0, cm[0].bias, out_dtype=torch.float32)
        else:
            z = fn.linear(fn.cast(z, torch.float32), w0, cm[0].bias)
        z = fn.layernorm(z, cm[1].weight, cm[1].bias, cm[1].eps, post_act=ops.ACT_GELU)
        z = fn.dropout(z, cm[3].p, self.training)
        z = fn.linear(z, cm[4].weight, cm[4].bias)
        z = fn.layernorm(z, cm[5].weight, cm[5].bias, cm[5].eps, post_act=ops.ACT_GELU)
        z = fn.dropout(z, cm[7].p, self.training)
        if self.classification:
            z = fn.linear(z, self.fc.weight, self.fc.bias)
        return z

    def _forward_eval_f32(self, x):
        B, c, L = x.shape
        h = x.contiguous().float().view(B, L, 1)
        for i in range(5):
            h, L, C = getattr(self, f"stage{i + 1}")[0].forward_cl(h, B, L)
        z = h.reshape(B, L * C)
        cm = self.class_model
        z = ops.gemm(z, self._w0_channels_last(), cm[0].bias)
        z = ops.layernorm(z, cm[1].weight, cm[1].bias, cm[1].eps, post_act=ops.ACT_GELU)
        z = ops.gemm(z, cm[4].weight, cm[4].bias)
        z = ops.layernorm(z, cm[5].weight, cm[5].bias, cm[5].eps, post_act=ops.ACT_GELU)
        if self.classification:
            z = ops.gemm(z, self.fc.weight, self.fc.bias)
        return z


def XastroMiNN(config=None, **kw):
    """Legacy 4-channel image+metadata classifier (_archive/notebooks/brew_cider.py:438-582): the src AstroMiNN architecture with
    `in_chans=4` cutouts and the call signature forward(metadata, image).  Returns an AstroMiNN whose `forward` also accepts the
    two-argument legacy form; state_dict keys are identical to the archive (`image_tower.backbone.stem.0.weight` is (96,4,4,4))."""
    import copy

    from .astrominn import AstroMiNN
    from .config import default_config

    cfg = copy.deepcopy(config) if config is not None else default_config()
    cfg["model"]["AstroMiNN"]["in_chans"] = 4
    cfg["model"]["AstroMiNN"].update(kw)
    model = AstroMiNN(cfg)
    tuple_forward = model.forward

    def forward(metadata, image=None):
        if image is None and isinstance(metadata, (tuple, list)):
            return tuple_forward(metadata)
        return tuple_forward((metadata, image, None))

    model.forward = forward
    return model
