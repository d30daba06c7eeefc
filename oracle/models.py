"""Plain-PyTorch CPU restatement of the reference models (TEST INFRASTRUCTURE).

Every class keeps the reference's constructor arguments, forward signature and
state_dict key schema but spells the arithmetic out with
elementary torch ops so that it doubles as a readable specification for the CUDA
kernels.  It is pinned bit-for-bit/ULP-level against the real reference modules
in tests/test_oracle_pinned.py and against tests/golden/*.npz.

Reference anchors (paths relative to /root/reference):
  Time2Vec              src/applecider/models/Time2Vec.py:48-72
  HyraxBaselineCLS      src/applecider/models/HyraxBaselineCLS.py:10-86
  FocalLoss             src/applecider/models/HyraxBaselineCLS.py:169-191
  SpectraNet(+Block)    src/applecider/models/spectranet.py:7-170
  ResidualTowerBlock    src/applecider/models/astrominn.py:44-64
  SplitHeadConvNeXt     src/applecider/models/astrominn.py:8-41
  AstroMiNN             src/applecider/models/astrominn.py:67-300
  ConvNeXt-T            timm 1.0.x timm/models/convnext.py (third-party, absent;
                        pinned in _archive/requirement.txt:194 → timm 1.0.15);
                        restated here and cross-checked against torchvision.
  AppleCider (fusion)   _archive/notebooks/brew_cider.py:807-862 (DECISION-1)
"""
from __future__ import annotations

import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# default hyper-parameters (restated from src/applecider/default_config.toml:1-119)
# --------------------------------------------------------------------------------------
_DEFAULT = {
    "model": {
        "AstroMiNN": {
            "num_classes": 9, "num_mlp_experts": 4, "use_probabilities": False,
            "towers_hidden_dims": 16, "towers_outdims": 32,
            "fusion_hidden_dims": 128, "fusion_router_dims": 128, "fusion_outdims": 32,
            "cnn_lr": 2, "cnn_decay": 5e-2, "psf_lr": 0.5, "psf_decay": 5e-2,
            "mag_lr": 2, "mag_decay": 0.0, "lc_lr": 2, "lc_decay": 0.05,
            "spatial_lr": 2, "spatial_decay": 0.0, "coord_lr": 0.5, "coord_decay": 0.0,
            "nst1_lr": 2, "nst1_decay": 0.0, "nst2_lr": 2, "nst2_decay": 0.0,
            "fusion_lr": 1, "fusion_decay": 1e-2, "fusion_beta1": 0.9, "fusion_beta2": 0.999,
            "router_decay": 0.0, "router_lr": 1.5, "router_beta1": 0.9, "router_beta2": 0.999,
            "beta1": 0.9, "beta2": 0.999, "eps": 5e-10,
        },
        "HyraxBaselineCLS": {
            "num_classes": 5, "pad_mask": 1, "mode": "photo", "d_model": 128, "n_heads": 8,
            "n_layers": 4, "dropout": 0.40, "max_len": 257, "lr": 5e-6, "weight_decay": 1e-2,
            "focal_gamma": 2.0, "use_probabilities": False, "pretrained_weights_path_": False,
            "lambda_f": 5.0, "lambda_b": 3.0, "lambda_dt": 5.0, "mask_p": 0.30,
        },
        "SpectraNet": {
            "redshift": False, "use_ln_stages": [True] * 5, "depths": [1] * 5,
            "channels": [64, 128, 256, 512, 1024],
            "kernel_sizes_per_stage": [[3, 61, 1021], [3, 31, 251], [3, 15, 61], [3, 11, 31], [3, 7, 13]],
            "class_order": 9, "flat_dim": 3072,
        },
    }
}


def default_config() -> dict:
    return copy.deepcopy(_DEFAULT)


# --------------------------------------------------------------------------------------
# photometry transformer
# --------------------------------------------------------------------------------------
class Time2Vec(nn.Module):
    """t -> [w0*t+b0, sin(w*t+b)]  (Time2Vec.py:63-72)."""

    def __init__(self, d_model):
        super().__init__()
        self.w0 = nn.Parameter(torch.randn(1))
        self.b0 = nn.Parameter(torch.zeros(1))
        self.w = nn.Parameter(torch.randn(d_model - 1))
        self.b = nn.Parameter(torch.zeros(d_model - 1))

    def forward(self, t):
        lin = (self.w0 * t + self.b0).unsqueeze(-1)
        per = torch.sin(t.unsqueeze(-1) * self.w + self.b)
        return torch.cat([lin, per], dim=-1)


def _encoder_layer_math(x, key_pad, lyr, n_heads):
    """One post-LN nn.TransformerEncoderLayer (ReLU FFN, eval / dropout off), spelled out.

    x: (B, S, D); key_pad: (B, S) bool, True = key ignored.
    """
    B, S, D = x.shape
    dh = D // n_heads
    sa = lyr.self_attn
    qkv = F.linear(x, sa.in_proj_weight, sa.in_proj_bias)
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, S, n_heads, dh).transpose(1, 2)
    k = k.view(B, S, n_heads, dh).transpose(1, 2)
    v = v.view(B, S, n_heads, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(dh))
    s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, S, D)
    o = F.linear(o, sa.out_proj.weight, sa.out_proj.bias)
    x = F.layer_norm(x + o, (D,), lyr.norm1.weight, lyr.norm1.bias, lyr.norm1.eps)
    f = F.linear(torch.relu(F.linear(x, lyr.linear1.weight, lyr.linear1.bias)), lyr.linear2.weight, lyr.linear2.bias)
    x = F.layer_norm(x + f, (D,), lyr.norm2.weight, lyr.norm2.bias, lyr.norm2.eps)
    return x


class HyraxBaselineCLS(nn.Module):
    """forward((data[B,L,7], pad[B,L] bool True=pad, labels)) -> (B,5) | (B,128)."""

    def __init__(self, config, data_sample=None):
        super().__init__()
        self.config = config
        mc = config["model"]["HyraxBaselineCLS"]
        d = mc["d_model"]
        self.n_heads = mc["n_heads"]
        self.in_proj = nn.Linear(7, d)
        self.cls_tok = nn.Parameter(torch.zeros(1, 1, d))
        self.time2vec = Time2Vec(d)
        layer = nn.TransformerEncoderLayer(d, mc["n_heads"], d * 4, mc["dropout"], batch_first=True)
        self.encoder = nn.TransformerEncoder(layer, mc["n_layers"])  # parameter container
        self.norm = nn.LayerNorm(d)
        self.head = nn.Linear(d, mc["num_classes"])  # unused in forward (HyraxBaselineCLS.py:35)
        self.classification = mc["mode"] == "photo"
        if self.classification:
            self.fc = nn.Linear(d, mc["num_classes"])

    def encode(self, data, pad):
        B = data.shape[0]
        h = self.in_proj(data) + self.time2vec(data[..., 0])
        h = torch.cat([self.cls_tok.expand(B, -1, -1), h], dim=1)
        kp = F.pad(pad, (1, 0), value=False)
        for lyr in self.encoder.layers:
            h = _encoder_layer_math(h, kp, lyr, self.n_heads)
        return self.norm(h[:, 0])

    def forward(self, x):
        data, pad, _ = x
        out = self.encode(data, pad)
        if self.classification:
            out = self.fc(out)
        if self.config["model"]["HyraxBaselineCLS"]["use_probabilities"]:
            out = F.softmax(out, dim=1)
        return out


class MPTModel(nn.Module):
    """Masked-event pre-training restated (HyraxBaselineCLS.py:194-319), dropout off.

    mask_batch follows _mask_batch (:283-319) draw for draw (same torch.randperm calls, so the same global
    seed gives the reference's mask); losses() follows train_step (:241-278) up to the loss, including the
    reference's quirks: targets are read from the ALREADY masked data, and the loss is a product.
    """

    def __init__(self, config, data_sample=None):
        super().__init__()
        self.config = config
        mc = config["model"]["HyraxBaselineCLS"]
        d = mc["d_model"]
        self.n_heads = mc["n_heads"]
        layer = nn.TransformerEncoderLayer(d, mc["n_heads"], d * 4, mc["dropout"], batch_first=True)
        self.encoder = nn.TransformerEncoder(layer, mc["n_layers"])  # parameter container
        self.in_proj = nn.Linear(7, d)
        self.cls_tok = nn.Parameter(torch.zeros(1, 1, d))
        self.time2vec = Time2Vec(d)
        self.head_flux = nn.Linear(d, 1)
        self.head_band = nn.Linear(d, 3)
        self.head_dt = nn.Linear(d, 1)

    def forward(self, z):
        return self.head_flux(z), self.head_band(z), self.head_dt(z)

    def mask_batch(self, x, pad_mask):
        mask_p = self.config["model"]["HyraxBaselineCLS"]["mask_p"]
        masked = torch.zeros_like(pad_mask)
        for b in range(x.shape[0]):
            valid = (~pad_mask[b]).nonzero(as_tuple=True)[0]
            k = max(int(len(valid) * mask_p), 3)
            each, extras = k // 3, k - 3 * (k // 3)
            bands = x[b, :, 4:7].argmax(-1)
            chosen = []
            for band in (0, 1, 2):
                vb = valid[bands[valid] == band]
                if len(vb) > 0:
                    chosen.append(vb[torch.randperm(len(vb))[: min(len(vb), each)]])
            if extras > 0:
                taken = torch.cat(chosen) if chosen else valid[:0]
                pool = valid[~torch.isin(valid, taken)]
                if len(pool) > 0:
                    chosen.append(pool[torch.randperm(len(pool))[:extras]])
            if chosen:
                idx = torch.cat(chosen)
                if len(idx) > 0:
                    x[b, idx, 2:7] = 0.0
                    masked[b, idx] = True
        return masked

    def losses(self, data, pad, masked):
        """data already masked. -> (loss, loss_f, loss_b, loss_dt)."""
        mc = self.config["model"]["HyraxBaselineCLS"]
        B = data.shape[0]
        h = self.in_proj(data) + self.time2vec(data[..., 0])
        h = torch.cat([self.cls_tok.expand(B, -1, -1), h], dim=1)
        kp = F.pad(pad, (1, 0), value=False)
        for lyr in self.encoder.layers:
            h = _encoder_layer_math(h, kp, lyr, self.n_heads)
        z = h[:, 1:, :]
        f_hat, b_hat, dt_hat = self.forward(z)
        mf = masked.reshape(-1)
        loss_f = F.mse_loss(f_hat.reshape(-1)[mf], data[..., 2].reshape(-1)[mf])
        true_b = data[..., 4:7].argmax(-1).reshape(-1)
        loss_b = F.cross_entropy(b_hat.reshape(-1, 3)[mf], true_b[mf])
        dt_gt = torch.roll(data[..., 1], -1, dims=1).clone()
        dt_gt[:, -1] = 0.0
        loss_dt = F.mse_loss(dt_hat[..., 0].reshape(-1)[mf], dt_gt.reshape(-1)[mf])
        loss = mc["lambda_f"] * loss_f * mc["lambda_b"] * loss_b * mc["lambda_dt"] * loss_dt
        return loss, loss_f, loss_b, loss_dt


def focal_loss(logits, target, gamma=2.0):
    """HyraxBaselineCLS.py:177-191 with alpha=None, eps=0, reduction='mean'."""
    logp = F.log_softmax(logits, dim=1)
    p = logp.exp()
    y = F.one_hot(target, num_classes=logits.shape[1]).float()
    return (-(y * (1.0 - p).pow(gamma) * logp).sum(dim=1)).mean()


# --------------------------------------------------------------------------------------
# SpectraNet
# --------------------------------------------------------------------------------------
class SpectraNetBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_sizes, use_ln=True, do_pool=False):
        super().__init__()
        self.do_pool, self.use_ln = do_pool, use_ln
        nc = out_channels * len(kernel_sizes)
        self.convs = nn.ModuleList(nn.Conv1d(in_channels, out_channels, k, padding=k // 2) for k in kernel_sizes)
        self.norm = nn.LayerNorm(nc) if use_ln else nn.BatchNorm1d(nc)
        if do_pool:
            self.downsample = nn.Conv1d(nc, out_channels, 1)

    def forward(self, x):
        y = torch.cat([c(x) for c in self.convs], dim=1)  # (B, 3C, L)
        if self.use_ln:
            y = self.norm(y.transpose(1, 2)).transpose(1, 2)
        else:
            y = self.norm(y)
        y = F.gelu(y)
        if self.do_pool:
            y = F.max_pool1d(self.downsample(y), 4)
        return y


class SpectraNet(nn.Module):
    """forward((flux[B,1,L], labels, redshifts)) -> (B,class_order) | (B,)."""

    def __init__(self, config=None, data_sample=None):
        super().__init__()
        self.config = config
        sc = config["model"]["SpectraNet"]
        self.redshift = sc["redshift"]
        ch, ks, dep, ln = sc["channels"], sc["kernel_sizes_per_stage"], sc["depths"], sc["use_ln_stages"]
        stages = []
        for i in range(len(dep)):
            cin = 1 if i == 0 else ch[i - 1]
            blocks = []
            for j in range(dep[i]):
                blocks.append(
                    SpectraNetBlock(
                        cin if j == 0 else ch[i] * len(ks[i]), ch[i], ks[i], use_ln=ln[i],
                        do_pool=(i != len(dep) - 1) and (j == dep[i] - 1),
                    )
                )
            stages.append(nn.Sequential(*blocks))
        self.all_stages = nn.Sequential(*stages)
        head_out = 1 if self.redshift else sc["class_order"]
        head = nn.Sequential(nn.Linear(sc["flat_dim"], 384), nn.LayerNorm(384), nn.GELU(), nn.Dropout(0.5), nn.Linear(384, head_out))
        if self.redshift:
            self.regressor = head
        else:
            self.classifier = head

    def forward(self, batch):
        x, _, _ = batch
        x = self.all_stages(x)
        feat = x.amax(dim=-1)
        if self.redshift:
            return self.regressor(feat).squeeze(1)
        return self.classifier(feat)


# --------------------------------------------------------------------------------------
# ConvNeXt-T (timm semantics) + AstroMiNN
# --------------------------------------------------------------------------------------
class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over C of an NCHW tensor (timm LayerNorm2d), eps 1e-6."""

    def __init__(self, c, eps=1e-6):
        super().__init__(c, eps=eps)

    def forward(self, x):
        return F.layer_norm(x.permute(0, 2, 3, 1), self.normalized_shape, self.weight, self.bias, self.eps).permute(0, 3, 1, 2)


class _Mlp(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.fc1 = nn.Linear(c, 4 * c)
        self.fc2 = nn.Linear(4 * c, c)


class ConvNeXtBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv_dw = nn.Conv2d(c, c, 7, padding=3, groups=c)
        self.norm = nn.LayerNorm(c, eps=1e-6)
        self.mlp = _Mlp(c)
        self.gamma = nn.Parameter(1e-6 * torch.ones(c))

    def forward(self, x):
        y = self.conv_dw(x).permute(0, 2, 3, 1)
        y = self.norm(y)
        y = self.mlp.fc2(F.gelu(self.mlp.fc1(y)))
        return x + (y * self.gamma).permute(0, 3, 1, 2)


class _Stage(nn.Module):
    def __init__(self, cin, cout, depth, first):
        super().__init__()
        if first:
            self.downsample = nn.Identity()
        else:
            self.downsample = nn.Sequential(LayerNorm2d(cin), nn.Conv2d(cin, cout, 2, stride=2))
        self.blocks = nn.Sequential(*[ConvNeXtBlock(cout) for _ in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class _Head(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = LayerNorm2d(c)

    def forward(self, x):
        return self.norm(x.mean(dim=(2, 3), keepdim=True)).flatten(1)


class ConvNeXtTiny(nn.Module):
    """timm ``convnext_tiny(in_chans, num_classes=0)``: (B,C,H,W) -> (B,768)."""

    num_features = 768

    def __init__(self, in_chans=3, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768)):
        super().__init__()
        self.stem = nn.Sequential(nn.Conv2d(in_chans, dims[0], 4, stride=4), LayerNorm2d(dims[0]))
        self.stages = nn.Sequential(
            *[_Stage(dims[max(i - 1, 0)], dims[i], depths[i], first=(i == 0)) for i in range(4)]
        )
        self.head = _Head(dims[-1])
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        return self.head(self.stages(self.stem(x)))


class SplitHeadConvNeXt(nn.Module):
    def __init__(self, pretrained=False, in_chans=4, outdims=4):
        super().__init__()
        self.backbone = ConvNeXtTiny(in_chans=in_chans)
        f = self.backbone.num_features
        self.head_main = nn.Sequential(
            nn.GELU(), nn.LayerNorm(f), nn.Linear(f, f // 2), nn.ReLU(), nn.Dropout(0.4), nn.Linear(f // 2, f), nn.Linear(f, outdims)
        )
        self.head_aux = nn.Sequential(nn.LayerNorm(f), nn.Linear(f, outdims), nn.Tanh())

    def forward(self, x):
        feat = self.backbone(x)
        return self.head_main(feat) * self.head_aux(feat)


class ResidualTowerBlock(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim):
        super().__init__()
        self.start_path = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.GELU())
        self.main_path = nn.Sequential(nn.LayerNorm(hidden_dim), nn.Dropout(0.25), nn.Linear(hidden_dim, output_dim))
        self.activation = nn.Sequential(nn.LayerNorm(hidden_dim), nn.Dropout(0.25), nn.Linear(hidden_dim, output_dim), nn.Sigmoid())
        self.skip_path = nn.Linear(input_dim, output_dim) if input_dim != output_dim else nn.Identity()

    def forward(self, x):
        s = self.start_path(x)
        return self.main_path(s) * self.activation(s) + self.skip_path(x)


# metadata column subsets (astrominn.py:249-261) and concat order (:264-267)
TOWER_COLS = {
    "nst1": [0, 2], "nst2": [1, 3], "spatial": [2, 3, 4], "psf": [5, 14],
    "mag": [6, 9, 10, 13, 15, 17, 18], "coord": [7, 8], "mega": list(range(19)),
    "lc": [6, 9, 10, 13, 15, 17, 18, 19, 20, 21, 22, 23],
}
CONCAT_ORDER = ["nst1", "nst2", "spatial", "psf", "mag", "coord", "mega", "image", "lc"]


class AstroMiNN(nn.Module):
    """forward((metadata[B,24], image[B,3,63,63], target)) -> (B,5)."""

    def __init__(self, config=None, data_sample=None):
        super().__init__()
        self.config = config
        ac = config["model"]["AstroMiNN"]
        th, to, fo = ac["towers_hidden_dims"], ac["towers_outdims"], ac["fusion_outdims"]
        self.psf_tower = ResidualTowerBlock(2, th, to)
        self.mag_tower = ResidualTowerBlock(7, th * 2, to)
        self.lc_tower = ResidualTowerBlock(12, th * 3, to)
        self.spatial_tower = ResidualTowerBlock(3, th, to)
        self.nst1_tower = ResidualTowerBlock(2, th, fo)
        self.nst2_tower = ResidualTowerBlock(2, th, fo)
        self.coord_tower = ResidualTowerBlock(2, th, fo)
        self.mega_tower = ResidualTowerBlock(19, 128, to)
        self.image_tower = SplitHeadConvNeXt(pretrained=False, in_chans=int(ac.get("in_chans", 3)), outdims=to)  # 4 = legacy XastroMiNN
        fd = 6 * to + 3 * fo
        self.fusion_experts = nn.ModuleList([ResidualTowerBlock(fd, ac["fusion_hidden_dims"], 5) for _ in range(ac["num_mlp_experts"])])
        self.fusion_router = nn.Sequential(nn.Linear(fd, fd // 2), nn.Tanh(), nn.Dropout(0.3), nn.Linear(fd // 2, ac["num_mlp_experts"]), nn.Sigmoid())

    def features(self, metadata, image):
        parts = {n: getattr(self, f"{n}_tower")(metadata[:, c]) for n, c in TOWER_COLS.items()}
        parts["image"] = self.image_tower(image)
        return torch.cat([parts[n] for n in CONCAT_ORDER], dim=1)

    def forward(self, batch):
        metadata, image, _ = batch
        feats = self.features(metadata, image)
        gate = self.fusion_router(feats)
        top_w, top_i = torch.topk(gate, k=2, dim=-1)
        out = torch.zeros(metadata.shape[0], 5, dtype=feats.dtype, device=feats.device)
        for e, expert in enumerate(self.fusion_experts):  # summation order e = 0..3 (astrominn.py:282-295)
            sel = top_i == e  # (B,2) at most one True per row
            w = (top_w * sel).sum(dim=-1, keepdim=True)
            out = out + torch.where(sel.any(dim=-1, keepdim=True), w * expert(feats), torch.zeros_like(out))
        if self.config["model"]["AstroMiNN"]["use_probabilities"]:
            out = F.softmax(out, dim=-1)
        return out


# --------------------------------------------------------------------------------------
# late-fusion head (DECISION-1: src encoders + brew_cider head)
# --------------------------------------------------------------------------------------
class AppleCider(nn.Module):
    """forward(photometry, photometry_mask, metadata, images, spectra) -> (B,num_classes)."""

    def __init__(self, config, hidden_dim=64, fusion="avg", num_classes=5):
        super().__init__()
        cfg = copy.deepcopy(config)
        cfg["model"]["HyraxBaselineCLS"]["mode"] = "all"  # encoder returns norm(z[:,0]) (HyraxBaselineCLS.py:37,81)
        cfg["model"]["HyraxBaselineCLS"]["use_probabilities"] = False
        cfg["model"]["AstroMiNN"]["use_probabilities"] = False
        self.fusion = fusion
        self.photometry_encoder = HyraxBaselineCLS(cfg)
        self.spectra_encoder = SpectraNet(cfg)
        self.img_metadata_encoder = AstroMiNN(cfg)
        sc = cfg["model"]["SpectraNet"]
        spec_out = 1 if sc["redshift"] else sc["class_order"]
        self.photometry_proj = nn.Linear(cfg["model"]["HyraxBaselineCLS"]["d_model"], hidden_dim)
        self.spectra_proj = nn.Linear(spec_out, hidden_dim)
        self.img_metadata_proj = nn.Linear(5, hidden_dim)
        self.fc = nn.Linear(hidden_dim * 3 if fusion == "concat" else hidden_dim, num_classes)

    def get_embeddings(self, photometry, photometry_mask, metadata, images, spectra):
        p = self.photometry_proj(self.photometry_encoder((photometry, photometry_mask, None)))
        s = self.spectra_proj(self.spectra_encoder((spectra, None, None)))
        im = self.img_metadata_proj(self.img_metadata_encoder((metadata, images, None)))
        p = p / p.norm(dim=-1, keepdim=True)
        im = im / im.norm(dim=-1, keepdim=True)
        s = s / s.norm(dim=-1, keepdim=True)
        return p, im, s

    def forward(self, photometry, photometry_mask, metadata, images, spectra):
        p, im, s = self.get_embeddings(photometry, photometry_mask, metadata, images, spectra)
        if self.fusion == "concat":
            emb = torch.cat((p, im, s), dim=1)
        elif self.fusion == "avg":
            emb = (p + im + s) / 3
        else:
            raise NotImplementedError
        return self.fc(emb)


# --------------------------------------------------------------------------------------
# legacy "variant B" spectra encoder
# --------------------------------------------------------------------------------------
class SpectraNetBlockB(nn.Module):
    """_archive/notebooks/brew_cider.py:586-636 restated (BatchNorm in eval mode spelled out as an affine)."""

    def __init__(self, in_channels, out_channels, kernel_sizes, use_skip=True, use_ln=True, do_pool=False):
        super().__init__()
        self.use_skip, self.use_ln, self.do_pool, self.k = use_skip, use_ln, do_pool, len(kernel_sizes)
        self.convs = nn.ModuleList([nn.Conv1d(in_channels, out_channels, kernel_size=k, padding=k // 2) for k in kernel_sizes])
        self.norm = nn.LayerNorm(out_channels * self.k) if use_ln else nn.BatchNorm1d(out_channels * self.k)
        if use_skip:
            self.proj = nn.Conv1d(in_channels, out_channels * self.k, kernel_size=1)

    def forward(self, x):
        residual = self.proj(x) if self.use_skip else None
        y = torch.cat([c(x) for c in self.convs], dim=1)
        if self.use_ln:
            y = F.layer_norm(y.permute(0, 2, 1), (y.shape[1],), self.norm.weight, self.norm.bias, self.norm.eps).permute(0, 2, 1)
        elif self.training:
            y = self.norm(y)  # nn.BatchNorm1d in train mode: batch statistics, running statistics blended in place
        else:
            n = self.norm
            y = (y - n.running_mean[None, :, None]) / torch.sqrt(n.running_var[None, :, None] + n.eps) * n.weight[None, :, None] + n.bias[None, :, None]
        if self.use_skip:
            y = residual + y
        y = F.gelu(y)
        if self.do_pool:
            y = torch.cat([F.max_pool1d(y, 4), F.avg_pool1d(y, 4), -F.max_pool1d(-y, 4)], dim=1)
        return y


class SpectraClassificationB(nn.Module):
    """brew_cider.py:638-705: five stages (BatchNorm x4, LayerNorm x1), flatten (C-major), 12288 -> 2048 -> 256 [-> classes]."""

    def __init__(self, config=None):
        super().__init__()
        config = config or {"mode": "all", "classes": list(range(5))}
        ks = [[3, 61, 1021], [3, 31, 251], [3, 15, 61], [3, 11, 31], [3, 7, 13]]
        ch, ln = [1, 16, 32, 64, 128, 256], [False, False, False, False, True]
        cin = ch[0]
        for i in range(5):
            setattr(self, f"stage{i + 1}", nn.Sequential(SpectraNetBlockB(cin, ch[i + 1], ks[i], True, ln[i], do_pool=(i != 4))))
            cin = ch[i + 1] * 3 * (3 if i != 4 else 1)
        self.class_model = nn.Sequential(nn.Linear(ch[5] * 3 * 16, 2048), nn.LayerNorm(2048), nn.GELU(), nn.Dropout(0.5),
                                         nn.Linear(2048, 256), nn.LayerNorm(256), nn.GELU(), nn.Dropout(0.3))
        self.classification = config["mode"] == "spectra"
        if self.classification:
            self.fc = nn.Linear(256, len(config["classes"]))

    def forward(self, x):
        for i in range(5):
            x = getattr(self, f"stage{i + 1}")(x)
        out = self.class_model(x.reshape(x.size(0), -1))
        return self.fc(out) if self.classification else out
