"""Drop-in SpectraNet (reference: src/applecider/models/spectranet.py:7-206).

Activations are channels-last [B, L, C].  Each stage is
    implicit-GEMM multi-kernel Conv1d  ->  LayerNorm(3C)+GELU  ->  1x1 conv (+MaxPool4 in the epilogue)
fp32: CUDA-core implicit GEMM (any length).  bf16: tcgen05 implicit GEMM fed by a 3-D TMA map
(tap = row shift, zero fill outside the sample); stage 0 (1 input channel, k up to 1021) uses the
8-phase polyphase view of the zero-padded signal so that K is a contiguous window of samples.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .config import resolve_dtype

FUSE_CONV_LN = True  # conv + bias + LayerNorm + GELU in one tcgen05 kernel where 3*C_out fits TMEM
FUSE_STAGE1 = False
FUSE_STAGE0_DOWN = True  # stage 0: also fuse the 1x1 downsample + pair max into the persistent kernel
FUSE_LN_DOWN = True  # stages 1-3: LayerNorm + GELU applied to the A operand of the 1x1 downsample GEMM (statistics from the conv epilogue)
_PHASES = 8
_HALO = 512  # zeros in front of every padded sample (>= max pad 510, multiple of 8)


class SpectraNetBlock(nn.Module):
    """Parameter container with the reference's keys: convs.{j}, norm, downsample."""

    def __init__(self, in_channels, out_channels, kernel_sizes, use_ln=True, do_pool=False):
        super().__init__()
        if not use_ln:
            raise NotImplementedError("applecider_b200: BatchNorm stages (use_ln=false) are not implemented")
        self.in_channels, self.out_channels, self.kernel_sizes = in_channels, out_channels, list(kernel_sizes)
        self.do_pool, self.use_ln, self.k = do_pool, use_ln, len(kernel_sizes)
        nc = out_channels * self.k
        self.convs = nn.ModuleList([nn.Conv1d(in_channels, out_channels, kernel_size=k, padding=k // 2) for k in kernel_sizes])
        self.norm = nn.LayerNorm(nc)
        if do_pool:
            self.total_pooled_channels = nc
            self.downsample = nn.Conv1d(nc, out_channels, kernel_size=1)
        self._derived = ops.DerivedCache()

    # ---- derived weights ---------------------------------------------------------------------
    def _kmax(self):
        return max(self.kernel_sizes)

    def _packed(self, dtype):
        """[k*Cout, kmax*Cin] tap-major weights, conv j centred in the kmax tap window; concatenated bias."""
        def build():
            kmax, cin, cout = self._kmax(), self.in_channels, self.out_channels
            w = torch.zeros((self.k * cout, kmax * cin), dtype=dtype, device=self.convs[0].weight.device)
            for j, c in enumerate(self.convs):
                kj = self.kernel_sizes[j]
                ops.call("acb_pack_conv_weight", c.weight, ops._offset_ptr(w, j * cout * kmax * cin), ops.dtype_tag(w), cout, cin, kj,
                         kmax * cin, kmax // 2 - kj // 2)
            bias = torch.cat([c.bias.detach() for c in self.convs]).contiguous()
            return w, bias

        return self._derived.get(("packed", dtype), [c.weight for c in self.convs] + [c.bias for c in self.convs], build)

    def _packed_polyphase(self, dtype):
        """Stage-0 (Cin == 1) weights for the 8-phase view: rows (conv j, phase r, co), columns = sample offset."""
        def build():
            cout = self.out_channels
            kp = ((_HALO + self._kmax() // 2 + _PHASES + 63) // 64) * 64
            w = torch.empty((self.k * _PHASES * cout, kp), dtype=dtype, device=self.convs[0].weight.device)
            for j, c in enumerate(self.convs):
                ops.call("acb_pack_polyphase_weight", c.weight, w, ops.dtype_tag(w), cout, self.kernel_sizes[j], _PHASES, _HALO, cout,
                         j * _PHASES * cout, kp, kp)
            bias = torch.cat([c.bias.detach().repeat(_PHASES) for c in self.convs]).contiguous()
            return w, bias, kp

        return self._derived.get(("poly", dtype), [c.weight for c in self.convs] + [c.bias for c in self.convs], build)

    def _down(self, dtype):
        def build():
            w = self.downsample.weight.detach().reshape(self.out_channels, -1).contiguous()
            return ops.cast(w, dtype)

        return self._derived.get(("down", dtype), [self.downsample.weight], build)

    # ---- forward -----------------------------------------------------------------------------
    def _convs_f32(self, x, B, L):
        cin, cout, kmax = self.in_channels, self.out_channels, self._kmax()
        w, bias = self._packed(torch.float32)
        y = torch.empty((B * L, self.k * cout), dtype=torch.float32, device=x.device)
        for j in range(self.k):
            kj = self.kernel_sizes[j]
            # conv j only touches its own taps: point at its first tap inside the packed row
            off = (kmax // 2 - kj // 2) * cin
            wj = ops._offset_ptr(w, j * cout * kmax * cin + off)
            ops.call("acb_gemm_f32", x, wj, ops._offset_ptr(y, j * cout), B * L, cout, kj * cin, cin, kmax * cin, self.k * cout, L, cin,
                     kj // 2, ops._offset_ptr(bias, j * cout), ops.ACT_NONE, None, 0, None, ops.RES_NONE)
        return y

    def _convs_bf16(self, x, B, L, want_stats=False):
        cin, cout, kmax = self.in_channels, self.out_channels, self._kmax()
        w, bias = self._packed(torch.bfloat16)
        N = self.k * cout
        bn = 256 if cout % 256 == 0 else (128 if cout % 128 == 0 else 64)
        if cin % 8 or N % 8:
            raise RuntimeError("applecider_b200: the bf16 conv path needs channel counts that are multiples of 8")
        cpt = (cin + 63) // 64
        ranges = []
        for nt in range((N + bn - 1) // bn):
            # an N tile normally holds channels of ONE conv; narrow convs (legacy variant B: 16 / 32 channels) share a tile, which
            # then runs the union of their tap ranges (the packed weights are zero outside a conv's own taps)
            js = range((nt * bn) // cout, min(self.k - 1, (min(N, (nt + 1) * bn) - 1) // cout) + 1)
            kj = max(self.kernel_sizes[j] for j in js)
            t_lo, t_hi = kmax // 2 - kj // 2, kmax // 2 + kj // 2 + 1
            ranges += [t_lo * cpt, t_hi * cpt]
        with ops.region(f"spectra.conv.cin{cin}"):
            if want_stats:  # per-row (sum, sum^2) of every N tile from the epilogue's fp32 accumulators: LayerNorm statistics for free
                parts = (N + bn - 1) // bn
                stats = torch.empty((B * L, parts, 2), dtype=torch.float32, device=x.device)
                return ops.gemm(x.view(B, L, cin), w, bias, conv=(kmax, kmax // 2), bn=bn, tile_kb=ranges, row_stats=stats), stats, parts
            return ops.gemm(x.view(B, L, cin), w, bias, conv=(kmax, kmax // 2), bn=bn, tile_kb=ranges)

    def _convs_bf16_polyphase(self, x_f32, B, L):
        """x_f32: [B, L] fp32 raw signal. Returns ([B*L8, 3C] bf16, L8)."""
        cout = self.out_channels
        w, bias, kp = self._packed_polyphase(torch.bfloat16)
        L8 = ((L + _PHASES - 1) // _PHASES) * _PHASES
        stride = L8 + kp
        xp = torch.zeros((B, stride), dtype=torch.bfloat16, device=x_f32.device)
        ops.call("acb_pad_signal", x_f32, xp, ops.BF16, B, L, stride, _HALO)
        N = self.k * _PHASES * cout
        bn = 256 if (_PHASES * cout) % 256 == 0 else 64
        ranges, coloff = [], []
        for nt in range(N // bn):
            j = (nt * bn) // (_PHASES * cout)
            pad = self.kernel_sizes[j] // 2
            lo = (_HALO - pad) // 64
            hi = min((_HALO + pad + _PHASES + 63) // 64, kp // 64)
            ranges += [lo, hi]
        for blk in range(N // 64):
            row = blk * 64
            j, rem = divmod(row, _PHASES * cout)
            r, co = divmod(rem, cout)
            coloff.append(r * self.k * cout + j * cout + co)
        y = torch.empty((B * L8, self.k * cout), dtype=torch.bfloat16, device=x_f32.device)
        with ops.region("spectra.conv.cin1"):
            ops.gemm(xp, w, bias, out=y.view(B * L8 // _PHASES, _PHASES * self.k * cout), bn=bn, tile_kb=ranges, colblk_off=coloff,
                     a_view=(B, L8 // _PHASES, kp, stride, _PHASES))
        return y, L8

    def _fusable(self):
        ks = self.kernel_sizes
        if not (self.k == 3 and ks == sorted(ks)):
            return False
        if self.in_channels == 1 and self.out_channels == 64:
            return True
        # the 128-channel case is functional (tests force it): with the persistent rotating-TMEM kernel it is 0.6 ms faster than
        # conv GEMM + streaming LayerNorm at B=4096 (19.6 vs 17.9 + 1.4 ms), but its single CTA per SM runs the main loop ~10 %
        # slower than two co-resident GEMM CTAs, so it stays opt-in
        return FUSE_STAGE1 and self.in_channels >= 64 and self.in_channels % 64 == 0 and self.out_channels == 128

    def _conv_ln_fused_bf16(self, x, B, L, raw_signal, fuse_down=False):
        """conv x3 + bias + LayerNorm + GELU in one tcgen05 kernel. Returns ([B*Lr, 3C] bf16, Lr)."""
        cin, cout, kmax = self.in_channels, self.out_channels, self._kmax()
        dev = self.norm.weight.device
        if cin == 1:
            w, bias, kp = self._packed_polyphase(torch.bfloat16)
            L8 = ((L + _PHASES - 1) // _PHASES) * _PHASES
            stride = L8 + kp
            xp = torch.zeros((B, stride), dtype=torch.bfloat16, device=dev)
            ops.call("acb_pad_signal", raw_signal, xp, ops.BF16, B, L, stride, _HALO)
            ranges = []
            for j in range(3):
                pad = self.kernel_sizes[j] // 2
                ranges += [(_HALO - pad) // 64, min((_HALO + pad + _PHASES + 63) // 64, kp // 64)]
            if fuse_down:
                # conv + LN + GELU + 1x1 downsample + max over position pairs in ONE kernel; the [B*L8, 192] activation never exists
                wd = self._down(torch.bfloat16)
                pm = torch.empty((B * L8 // 2, cout), dtype=torch.bfloat16, device=dev)
                with ops.region("spectra.conv.cin1"):
                    ops.call("acb_spectra_conv_ln_bf16", xp, w, None, B, L8 // _PHASES, kp, 1, 0, stride, _PHASES, kp, w.shape[0], ops._int_array(ranges),
                             ops._int_array([j * _PHASES * cout for j in range(3)]), 2 * cout, _PHASES // 2, 2, _PHASES, 2, B * L8, bias,
                             self.norm.weight, self.norm.bias, self.norm.eps, wd, self.downsample.bias, pm)
                    z = torch.empty((B * L8 // 4, cout), dtype=torch.bfloat16, device=dev)
                    ops.call("acb_pairmax_bf16", pm, z, B * L8 // 4, cout)
                return z, L8
            y = torch.empty((B * L8, 3 * cout), dtype=torch.bfloat16, device=dev)
            with ops.region("spectra.conv.cin1"):
                ops.call("acb_spectra_conv_ln_bf16", xp, w, y, B, L8 // _PHASES, kp, 1, 0, stride, _PHASES, kp, w.shape[0], ops._int_array(ranges),
                         ops._int_array([j * _PHASES * cout for j in range(3)]), 2 * cout, _PHASES // 2, 2, _PHASES, 2, B * L8, bias,
                         self.norm.weight, self.norm.bias, self.norm.eps, None, None, None)
            return y, L8
        w, bias = self._packed(torch.bfloat16)
        cpt = cin // 64
        ranges = []
        for j in range(3):
            kj = self.kernel_sizes[j]
            ranges += [(kmax // 2 - kj // 2) * cpt, (kmax // 2 + kj // 2 + 1) * cpt]
        y = torch.empty((B * L, 3 * cout), dtype=torch.bfloat16, device=dev)
        with ops.region(f"spectra.conv.cin{cin}"):
            ops.call("acb_spectra_conv_ln_bf16", x, w, y, B, L, cin, kmax, kmax // 2, L * cin, cin, kmax * cin, w.shape[0], ops._int_array(ranges),
                     ops._int_array([0, cout, 2 * cout]), 0, 1, 1, 1, 0, B * L, bias, self.norm.weight, self.norm.bias, self.norm.eps, None, None, None)
        return y, L

    def forward_cl(self, x, B, L, dtype, raw_signal=None):
        """x: channels-last [B, L, Cin] activations (dtype).  Returns (y, L_out)."""
        cout, nc = self.out_channels, self.out_channels * self.k
        Lr = L
        fused = False
        if dtype == torch.float32:
            y = self._convs_f32(x, B, L)
        elif FUSE_CONV_LN and self._fusable():
            L8 = ((L + _PHASES - 1) // _PHASES) * _PHASES
            # stage 0 with >= 148 signal windows (the persistent kernel): the 1x1 downsample and half of MaxPool(4) join the kernel
            if (FUSE_STAGE0_DOWN and self.in_channels == 1 and self.do_pool and self.out_channels == 64 and L8 % 128 == 0
                    and B * (L8 // _PHASES) // 128 >= 148 and L8 == L):
                z, _ = self._conv_ln_fused_bf16(x, B, L, raw_signal, fuse_down=True)
                return z.view(B, L // 4, cout), L // 4
            y, Lr = self._conv_ln_fused_bf16(x, B, L, raw_signal)
            fused = True
        elif self.in_channels == 1:
            if cout % 64:
                raise RuntimeError("applecider_b200: bf16 SpectraNet needs out_channels to be a multiple of 64")
            y, Lr = self._convs_bf16_polyphase(raw_signal, B, L)
        elif (FUSE_LN_DOWN and self.do_pool and nc % 64 == 0 and cout % 128 == 0 and L % 4 == 0 and B * L >= 128):
            # conv (+ LayerNorm statistics from its epilogue) -> ONE GEMM that normalises + GELUs its A tiles in shared memory,
            # multiplies by the 1x1 downsample weights and max-pools: the normalised [B*L, 3C] activation never exists in HBM
            y, stats, parts = self._convs_bf16(x, B, L, want_stats=True)
            with ops.region(f"spectra.ln_down.cin{self.in_channels}"):
                z = ops.gemm_ln(y, self._down(dtype), self.downsample.bias, stats, parts, self.norm.weight, self.norm.bias, self.norm.eps, pool4=True)
            return z.view(B, L // 4, cout), L // 4
        else:
            y = self._convs_bf16(x, B, L)
        if not fused:
            y = ops.layernorm(y, self.norm.weight, self.norm.bias, self.norm.eps, post_act=ops.ACT_GELU)
        if not self.do_pool:
            if Lr != L:
                y = y.view(B, Lr, nc)[:, :L].contiguous()
            return y.view(B, L, nc), L
        wd = self._down(dtype)
        fuse_pool = dtype == torch.bfloat16 and (B * Lr) % 4 == 0 and Lr % 4 == 0
        if fuse_pool:
            z = ops.gemm(y, wd, self.downsample.bias, pool4=True).view(B, Lr // 4, cout)
        else:
            z = ops.gemm(y, wd, self.downsample.bias)
            z = ops.maxpool4(z, B, Lr, cout)
        Lo = L // 4
        if z.shape[1] != Lo:
            z = z[:, :Lo].contiguous()
        return z, Lo


def make_stage(in_channel, out_channel, depth, kernel_sizes, use_ln=True, do_pool=True):
    k = len(kernel_sizes)
    blocks = []
    for i in range(depth):
        blocks.append(
            SpectraNetBlock(
                in_channels=in_channel if i == 0 else out_channel * k, out_channels=out_channel, kernel_sizes=kernel_sizes,
                use_ln=use_ln, do_pool=(do_pool if i == depth - 1 else False),
            )
        )
    return nn.Sequential(*blocks), k


class SpectraNet(nn.Module):
    """forward((flux[B,1,L] f32, labels, redshifts)) -> (B, class_order) logits or (B,) redshift."""

    def __init__(self, config=None, data_sample=None):
        super().__init__()
        self.config = config
        sc = config["model"]["SpectraNet"]
        self.redshift = sc["redshift"]
        ks, depths, lns, ch = sc["kernel_sizes_per_stage"], sc["depths"], sc["use_ln_stages"], sc["channels"]
        if not (len(depths) == len(lns) == len(ch) == len(ks)):
            raise ValueError("depths, use_ln_stages, channels, and kernel_sizes_per_stage must be the same length.")
        self.stages, self.ks = [], []
        for i in range(len(depths)):
            stage, k = make_stage(1 if i == 0 else ch[i - 1], ch[i], depths[i], ks[i], use_ln=lns[i], do_pool=(i != len(depths) - 1))
            self.stages.append(stage)
            self.ks.append(k)
        self.all_stages = nn.Sequential(*self.stages)
        out_dim = 1 if self.redshift else sc["class_order"]
        head = nn.Sequential(nn.Linear(sc["flat_dim"], 384), nn.LayerNorm(384), nn.GELU(), nn.Dropout(0.5), nn.Linear(384, out_dim))
        if self.redshift:
            self.regressor = head
        else:
            self.classifier = head
        self.compute_dtype = resolve_dtype(sc.get("compute_dtype"))
        self._derived = ops.DerivedCache()

    def features(self, x):
        """x: [B,1,L] -> global-max feature [B, 3*C_last] fp32."""
        if not x.is_cuda:
            raise RuntimeError("applecider_b200: inputs must be CUDA tensors (no CPU fallback)")
        dtype = self.compute_dtype
        B, cin0, L = x.shape
        assert cin0 == 1
        sig = x.contiguous().float().view(B, L)
        h = sig.view(B, L, 1) if dtype == torch.float32 else None
        first = True
        for stage in self.all_stages:
            for blk in stage:
                if first and dtype != torch.float32 and blk.in_channels == 1:
                    h, L = blk.forward_cl(None, B, L, dtype, raw_signal=sig)
                else:
                    if h is None:
                        h = ops.cast(sig.view(B, L, 1), dtype)
                    h, L = blk.forward_cl(h, B, L, dtype)
                first = False
        return ops.globalmax(h, B, L, h.shape[-1])

    def forward(self, batch):
        x, _, _ = batch
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train import spectra_forward_train

            return spectra_forward_train(self, x)
        feat = self.features(x)
        head = self.regressor if self.redshift else self.classifier
        dtype = self.compute_dtype
        if dtype == torch.float32:
            z = ops.gemm(feat, head[0].weight, head[0].bias)
        else:
            w0 = self._derived.get("h0", (head[0].weight,), lambda: ops.cast(head[0].weight.detach(), dtype))
            z = ops.gemm(ops.cast(feat, dtype), w0, head[0].bias, out_dtype=torch.float32)
        z = ops.layernorm(z, head[1].weight, head[1].bias, head[1].eps, post_act=ops.ACT_GELU)
        out = ops.gemm(z, head[4].weight, head[4].bias)
        if self.redshift and self.config["model"]["SpectraNet"].get("redshift_softplus", False):
            # archived redshift regressor ends in F.softplus (_archive/AppleCider/models/SpectraNetRedshift.py:112)
            sp = torch.empty_like(out)
            ops.call("acb_act_fwd", out, 0, sp, 0, ops.ACT_SOFTPLUS, out.numel())
            out = sp
        return out.squeeze(1) if self.redshift else out

    def train_step(self, batch):
        from .train import spectra_train_step

        return spectra_train_step(self, batch)

    @staticmethod
    def to_tensor(data_dict):
        """Same contract as the reference (spectranet.py:186-206)."""
        import numpy as np

        if "data" not in data_dict:
            raise ValueError("Data dictionary must have a 'data' key.")
        data = data_dict["data"]
        return (
            np.asarray(data.get("flux", []), dtype=np.float32),
            np.asarray(data.get("label", []), dtype=np.int16),
            np.asarray(data.get("redshift", []), dtype=np.float32),
        )
