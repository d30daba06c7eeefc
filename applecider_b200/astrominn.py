"""Drop-in AstroMiNN: ConvNeXt-T cutout CNN + gated metadata towers + top-2 MoE
(reference: src/applecider/models/astrominn.py:8-348; backbone = timm convnext_tiny, key names per
SURVEY.md Appendix A.3).  Channels-last activations; dense layers are GEMMs with fused epilogues,
depthwise 7x7 + LayerNorm is one kernel, towers/experts are single fused kernels, the MoE combine
runs on the device without the reference's per-expert host syncs."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .config import resolve_dtype

LN_EPS_CNX = 1e-6


# ---- parameter containers with timm's key names -----------------------------------------------------
class _LN(nn.Module):
    def __init__(self, c, eps=LN_EPS_CNX):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.eps = eps


class _Mlp(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.fc1 = nn.Linear(c, 4 * c)
        self.fc2 = nn.Linear(4 * c, c)


class _Block(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv_dw = nn.Conv2d(c, c, 7, padding=3, groups=c)
        self.norm = _LN(c)
        self.mlp = _Mlp(c)
        self.gamma = nn.Parameter(1e-6 * torch.ones(c))


class _Stage(nn.Module):
    def __init__(self, cin, cout, depth, first):
        super().__init__()
        self.downsample = nn.Identity() if first else nn.Sequential(_LN(cin), nn.Conv2d(cin, cout, 2, stride=2))
        self.blocks = nn.Sequential(*[_Block(cout) for _ in range(depth)])


class _Head(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = _LN(c)


class ConvNeXtTiny(nn.Module):
    """(B,C,H,W) f32 -> (B,768) f32; timm `convnext_tiny(in_chans, num_classes=0)` semantics."""

    num_features = 768

    def __init__(self, in_chans=3, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768)):
        super().__init__()
        self.dims = dims
        self.stem = nn.Sequential(nn.Conv2d(in_chans, dims[0], 4, stride=4), _LN(dims[0]))
        self.stages = nn.Sequential(*[_Stage(dims[max(i - 1, 0)], dims[i], depths[i], first=(i == 0)) for i in range(4)])
        self.head = _Head(dims[-1])
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
        self._derived = ops.DerivedCache()

    def _w(self, p, dtype, shape=None):
        def build():
            w = p.detach()
            if shape is not None:
                w = w.reshape(shape)
            return ops.cast(w.contiguous(), dtype)

        if dtype == torch.float32 and shape is None:
            return p
        return self._derived.get(("w", id(p), dtype), (p,), build)

    def _wds(self, conv, dtype):
        return self._derived.get(("ds", id(conv.weight), dtype), (conv.weight,), lambda: ops.pack_conv2d_weight(conv.weight.detach(), dtype))

    def forward_features(self, img, dtype):
        B, Cin, H, W = img.shape
        img = img.contiguous().float()
        c0 = self.dims[0]
        a = ops.patchify(img, 4, dtype)
        x = ops.gemm(a, self._w(self.stem[0].weight, dtype, (c0, Cin * 16)), self.stem[0].bias)
        x = ops.layernorm(x, self.stem[1].weight, self.stem[1].bias, self.stem[1].eps)
        h, w = H // 4, W // 4
        for si, st in enumerate(self.stages):
            C = self.dims[si]
            if si > 0:
                cprev = self.dims[si - 1]
                a = ops.ln_patch2(x, B, h, w, cprev, st.downsample[0].weight, st.downsample[0].bias, st.downsample[0].eps)
                h, w = h // 2, w // 2
                x = ops.gemm(a, self._wds(st.downsample[1], dtype), st.downsample[1].bias)
            for blk in st.blocks:
                y = ops.dwconv7_ln(x, B, h, w, C, blk.conv_dw.weight, blk.conv_dw.bias, blk.norm.weight, blk.norm.bias, blk.norm.eps)
                if ops.FUSE_CONVNEXT_MLP and dtype == torch.bfloat16 and C in (96, 192) and x.shape[0] >= 128:
                    # one kernel for fc1 + GELU + fc2 + layer scale + residual: the [rows, 4C] hidden activation never reaches HBM
                    x = ops.convnext_mlp(y, x, self._w(blk.mlp.fc1.weight, dtype), blk.mlp.fc1.bias, self._w(blk.mlp.fc2.weight, dtype),
                                         blk.mlp.fc2.bias, blk.gamma)
                    continue
                hid = ops.gemm(y, self._w(blk.mlp.fc1.weight, dtype), blk.mlp.fc1.bias, act=ops.ACT_GELU)
                x = ops.gemm(hid, self._w(blk.mlp.fc2.weight, dtype), blk.mlp.fc2.bias, res=x, gamma=blk.gamma, res_mode=ops.RES_ADD)
        return ops.gap_ln(x, B, h * w, self.dims[-1], self.head.norm.weight, self.head.norm.bias, self.head.norm.eps)


class SplitHeadConvNeXt(nn.Module):
    def __init__(self, pretrained=False, in_chans=4, outdims=4):
        super().__init__()
        if pretrained:
            raise NotImplementedError("pretrained timm weights are not available offline; load a state_dict instead")
        self.backbone = ConvNeXtTiny(in_chans=in_chans)
        f = self.backbone.num_features
        self.head_main = nn.Sequential(nn.GELU(), nn.LayerNorm(f), nn.Linear(f, f // 2), nn.ReLU(), nn.Dropout(0.4), nn.Linear(f // 2, f), nn.Linear(f, outdims))
        self.head_aux = nn.Sequential(nn.LayerNorm(f), nn.Linear(f, outdims), nn.Tanh())
        self._derived = ops.DerivedCache()

    def _w(self, p, dtype):
        if dtype == torch.float32:
            return p
        return self._derived.get(("w", id(p)), (p,), lambda: ops.cast(p.detach().contiguous(), dtype))

    def forward_into(self, img, dtype, out, out_col):
        """Writes head_main(f) * head_aux(f) into out[:, out_col:out_col+outdims] (fp32)."""
        feat = self.backbone.forward_features(img, dtype)
        hm, ha = self.head_main, self.head_aux
        a = ops.layernorm(feat, hm[1].weight, hm[1].bias, hm[1].eps, pre_gelu=True, out_dtype=dtype)
        a = ops.gemm(a, self._w(hm[2].weight, dtype), hm[2].bias, act=ops.ACT_RELU)
        a = ops.gemm(a, self._w(hm[5].weight, dtype), hm[5].bias)
        main = ops.gemm(a, self._w(hm[6].weight, dtype), hm[6].bias, out_dtype=torch.float32)
        x = ops.layernorm(feat, ha[0].weight, ha[0].bias, ha[0].eps, out_dtype=dtype)
        ops.gemm(x, self._w(ha[1].weight, dtype), ha[1].bias, act=ops.ACT_TANH, res=main, res_mode=ops.RES_MUL, out=out, out_col=out_col)
        return feat

    def forward(self, x):
        out = torch.empty((x.shape[0], self.head_main[6].out_features), dtype=torch.float32, device=x.device)
        self.forward_into(x, resolve_dtype(None), out, 0)
        return out


class ResidualTowerBlock(nn.Module):
    """Parameter container (astrominn.py:44-64); evaluated by acb_tower_fwd."""

    def __init__(self, input_dim, hidden_dim, output_dim):
        super().__init__()
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.start_path = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.GELU())
        self.main_path = nn.Sequential(nn.LayerNorm(hidden_dim), nn.Dropout(0.25), nn.Linear(hidden_dim, output_dim))
        self.activation = nn.Sequential(nn.LayerNorm(hidden_dim), nn.Dropout(0.25), nn.Linear(hidden_dim, output_dim), nn.Sigmoid())
        self.skip_path = nn.Linear(input_dim, output_dim) if input_dim != output_dim else nn.Identity()

    def run(self, X, cols, Y, y_off, s_pre=None, s_off=0):
        skip = self.skip_path if isinstance(self.skip_path, nn.Linear) else None
        ops.call(
            "acb_tower_fwd", X, X.shape[1], cols, self.input_dim, self.hidden_dim, self.output_dim,
            self.start_path[0].weight, self.start_path[0].bias, self.main_path[0].weight, self.main_path[0].bias,
            self.main_path[2].weight, self.main_path[2].bias, self.activation[0].weight, self.activation[0].bias,
            self.activation[2].weight, self.activation[2].bias, (skip.weight if skip is not None else None),
            (skip.bias if skip is not None else None), Y, Y.shape[1], y_off, X.shape[0],
            s_pre, (s_pre.shape[1] if s_pre is not None else 0), s_off,
        )

    def forward(self, x):
        x = x.contiguous().float()
        y = torch.empty((x.shape[0], self.output_dim), dtype=torch.float32, device=x.device)
        self.run(x, None, y, 0)
        return y


TOWER_COLS = {
    "nst1": [0, 2], "nst2": [1, 3], "spatial": [2, 3, 4], "psf": [5, 14],
    "mag": [6, 9, 10, 13, 15, 17, 18], "coord": [7, 8], "mega": list(range(19)),
    "lc": [6, 9, 10, 13, 15, 17, 18, 19, 20, 21, 22, 23],
}
CONCAT_ORDER = ["nst1", "nst2", "spatial", "psf", "mag", "coord", "mega", "image", "lc"]


class AstroMiNN(nn.Module):
    """forward((metadata[B,24] f32, image[B,3,63,63] f32, target)) -> (B,5) logits | probabilities."""

    def __init__(self, config=None, data_sample=None):
        super().__init__()
        self.config = config
        ac = config["model"]["AstroMiNN"]
        self.has_image = True
        self.num_classes = ac["num_classes"]
        self.num_mlp_experts = ac["num_mlp_experts"]
        self.towers_hidden_dims, self.towers_outdims = ac["towers_hidden_dims"], ac["towers_outdims"]
        self.fusion_hidden_dims, self.fusion_router_dims, self.fusion_outdims = ac["fusion_hidden_dims"], ac["fusion_router_dims"], ac["fusion_outdims"]
        th, to, fo = self.towers_hidden_dims, self.towers_outdims, self.fusion_outdims
        self.psf_tower = ResidualTowerBlock(2, th, to)
        self.mag_tower = ResidualTowerBlock(7, th * 2, to)
        self.lc_tower = ResidualTowerBlock(12, th * 3, to)
        self.spatial_tower = ResidualTowerBlock(3, th, to)
        self.nst1_tower = ResidualTowerBlock(2, th, fo)
        self.nst2_tower = ResidualTowerBlock(2, th, fo)
        self.coord_tower = ResidualTowerBlock(2, th, fo)
        self.mega_tower = ResidualTowerBlock(19, 128, to)
        self.image_tower = SplitHeadConvNeXt(pretrained=False, in_chans=int(ac.get("in_chans", 3)), outdims=to)  # 4 = legacy XastroMiNN
        fusion_dims = 6 * to + 3 * fo
        self.fusion_dims = fusion_dims
        self.fusion_experts = nn.ModuleList([ResidualTowerBlock(fusion_dims, self.fusion_hidden_dims, 5) for _ in range(self.num_mlp_experts)])
        self.fusion_router = nn.Sequential(nn.Linear(fusion_dims, fusion_dims // 2), nn.Tanh(), nn.Dropout(0.3), nn.Linear(fusion_dims // 2, self.num_mlp_experts), nn.Sigmoid())
        self.total_loss, self.total_correct_predictions, self.total_predictions = [], 0, 0
        self.this_criterion = nn.CrossEntropyLoss()
        self.compute_dtype = resolve_dtype(ac.get("compute_dtype"))
        self._derived = ops.DerivedCache()
        for n, c in TOWER_COLS.items():
            self.register_buffer(f"_cols_{n}", torch.tensor(c, dtype=torch.int32), persistent=False)
        self.this_optimizer = self._make_optimizer(ac)

    def _make_optimizer(self, c):
        """11 AdamW groups exactly as astrominn.py:151-218 (base LR 1.6e-4)."""
        LR = 1.6e-4
        g = lambda mod, wd, lr, **kw: dict(params=mod.parameters(), weight_decay=c[wd], lr=LR * c[lr], **kw)  # noqa: E731
        groups = [
            g(self.image_tower, "cnn_decay", "cnn_lr"), g(self.psf_tower, "psf_decay", "psf_lr"), g(self.lc_tower, "lc_decay", "lc_lr"),
            g(self.mag_tower, "mag_decay", "mag_lr"), g(self.spatial_tower, "spatial_decay", "spatial_lr"),
            g(self.coord_tower, "nst1_decay", "nst1_lr"), g(self.nst1_tower, "nst1_decay", "nst1_lr"), g(self.nst2_tower, "nst2_decay", "nst2_lr"),
            g(self.mega_tower, "lc_decay", "lc_lr"),
            g(self.fusion_experts, "fusion_decay", "fusion_lr", betas=(c["fusion_beta1"], c["fusion_beta2"])),
            g(self.fusion_router, "router_decay", "router_lr", betas=(c["router_beta1"], c["router_beta2"])),
        ]
        return torch.optim.AdamW(groups, lr=LR, betas=(c["beta1"], c["beta2"]), eps=c["eps"])

    def features(self, metadata, image):
        """All tower outputs concatenated in the reference order -> [B, 288] fp32."""
        if not metadata.is_cuda:
            raise RuntimeError("applecider_b200: inputs must be CUDA tensors (no CPU fallback)")
        B = metadata.shape[0]
        metadata = metadata.contiguous().float()
        feats = torch.empty((B, self.fusion_dims), dtype=torch.float32, device=metadata.device)
        off = 0
        for n in CONCAT_ORDER:
            if n == "image":
                if image is not None:
                    self.image_tower.forward_into(image, self.compute_dtype, feats, off)
                else:
                    feats[:, off:off + self.towers_outdims].zero_()
                off += self.towers_outdims
            else:
                tw = getattr(self, f"{n}_tower")
                tw.run(metadata, getattr(self, f"_cols_{n}"), feats, off)
                off += tw.output_dim
        return feats

    def forward(self, batch):
        metadata, image, _ = batch
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train import astrominn_forward_train

            return astrominn_forward_train(self, metadata, image)
        feats = self.features(metadata, image)
        B, E = feats.shape[0], self.num_mlp_experts
        r = self.fusion_router
        g1 = ops.gemm(feats, r[0].weight, r[0].bias, act=ops.ACT_TANH)
        gate = ops.gemm(g1, r[3].weight, r[3].bias, act=ops.ACT_SIGMOID)
        eo = torch.empty((B, E * 5), dtype=torch.float32, device=feats.device)
        # all expert start paths (288 -> 128, GELU) as ONE GEMM; the per-expert kernel then only does LN + heads + skip
        hid = self.fusion_hidden_dims
        w0, b0 = self._derived.get("experts_w0", [ex.start_path[0].weight for ex in self.fusion_experts] + [ex.start_path[0].bias for ex in self.fusion_experts],
                                   lambda: (torch.cat([ex.start_path[0].weight.detach() for ex in self.fusion_experts], 0).contiguous(),
                                            torch.cat([ex.start_path[0].bias.detach() for ex in self.fusion_experts], 0).contiguous()))
        s_all = ops.gemm(feats, w0, b0, act=ops.ACT_GELU)
        for e, ex in enumerate(self.fusion_experts):
            ex.run(feats, None, eo, e * 5, s_pre=s_all, s_off=e * hid)
        out = torch.empty((B, 5), dtype=torch.float32, device=feats.device)
        ops.call("acb_moe_combine", gate, eo, out, None, B, E, 5)
        if self.config["model"]["AstroMiNN"]["use_probabilities"]:
            out = ops.softmax_rows(out)
        return out

    def _update_stats(self, loss):
        self.total_loss.append(float(loss))

    def _calculate_stats(self):
        return sum(self.total_loss) / len(self.total_loss)

    def train_step(self, batch):
        from .train import astrominn_train_step

        return astrominn_train_step(self, batch)

    @staticmethod
    def to_tensor(data_dict: dict) -> tuple:
        """Same contract as the reference (astrominn.py:328-348)."""
        import numpy as np

        if "data" not in data_dict:
            raise ValueError("Input data dictionary does not contain 'data' key.")
        data = data_dict["data"]
        metadata = np.asarray(data["metadata"], dtype=np.float32)
        images = np.asarray(data["image"], dtype=np.float32)
        labels = np.asarray(data.get("target", []), dtype=np.float32)
        return (metadata, images, labels)
