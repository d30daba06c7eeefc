"""Drop-in photometry transformer (reference: src/applecider/models/HyraxBaselineCLS.py:10-166,
Time2Vec.py:48-124).  Same constructor, forward signature and state_dict keys; the arithmetic runs in
the sm_100a kernels behind the C-ABI (varlen-packed tokens, no padded FLOPs)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .config import resolve_dtype


class Time2Vec(nn.Module):
    """Parameter container for Time2Vec (Time2Vec.py:48-72); evaluated inside acb_photo_embed."""

    def __init__(self, d_model):
        super().__init__()
        self.w0 = nn.Parameter(torch.randn(1))
        self.b0 = nn.Parameter(torch.zeros(1))
        self.w = nn.Parameter(torch.randn(d_model - 1))
        self.b = nn.Parameter(torch.zeros(d_model - 1))


class _PhotoEncoderBase(nn.Module):
    """Shared encoder: embed -> n x (MHA + FFN, post-LN) -> LayerNorm(CLS)."""

    def _build(self, d_model, n_heads, n_layers, dropout):
        self.d_model, self.n_heads = d_model, n_heads
        if d_model // n_heads != 16 or d_model % n_heads:
            raise ValueError("applecider_b200: the attention kernel supports head_dim == 16 (d_model 128 / 8 heads)")
        self.in_proj = nn.Linear(7, d_model)
        self.cls_tok = nn.Parameter(torch.zeros(1, 1, d_model))
        self.time2vec = Time2Vec(d_model)
        layer = nn.TransformerEncoderLayer(d_model, n_heads, d_model * 4, dropout, batch_first=True)
        self.encoder = nn.TransformerEncoder(layer, n_layers)  # parameter container (keys encoder.layers.{i}.*)
        self._derived = ops.DerivedCache()

    def _w(self, p, dtype):
        if dtype == torch.float32:
            return p
        return self._derived.get(("cast", id(p)), (p,), lambda: ops.cast(p.detach().contiguous(), dtype))

    def pack(self, data, pad, total_tokens=None):
        """Varlen packing plan of a key-padding mask: (cu_seqlens, src_idx, T, rows_dev).  NOTHING is read back from the device:
        T is a row capacity -- ``total_tokens`` when the collate supplies the packed count (sum of lengths + B; any upper
        bound works) or the worst case B*(L+1).  ``rows_dev`` points at cu_seqlens[B] (the real count, on the device): the
        GEMM / LayerNorm kernels skip the 128-row tiles past it, so the unused capacity costs no arithmetic."""
        ops.check_photo_inputs(data, pad)
        pad = pad.contiguous()
        if pad.dtype != torch.bool:
            pad = pad != 0
        B, L = pad.shape
        T = ops.token_capacity(B, L, total_tokens)
        cu, src = ops.photo_compact(pad, T)
        return cu, src, T, ops._offset_ptr(cu, B)

    def encode_tokens(self, data, pad, dtype, total_tokens=None, packed=None):
        """Returns (h [T,D] packed tokens after the last layer (rows past cu[B] are undefined), cu_seqlens)."""
        if not data.is_cuda:
            raise RuntimeError("applecider_b200: inputs must be CUDA tensors (no CPU fallback)")
        B, L, F = data.shape
        data = data.contiguous().float()
        cu, src, T, nv = packed if packed is not None else self.pack(data, pad, total_tokens)
        D = self.d_model
        t2v = self.time2vec
        h = ops.photo_embed(data, src, T, D, self.in_proj.weight, self.in_proj.bias, t2v.w0, t2v.b0, t2v.w, t2v.b,
                            self.cls_tok, dtype)
        mv = nv if dtype != torch.float32 else None  # the fp32 parity GEMM has no device row count: it runs all capacity rows
        plan = ops.attention_plan(cu, B, T) if dtype == torch.bfloat16 and ops.USE_PACKED_ATTENTION else None
        for lyr in self.encoder.layers:
            sa = lyr.self_attn
            qkv = ops.gemm(h, self._w(sa.in_proj_weight, dtype), sa.in_proj_bias, m_valid=mv)
            att = ops.attention_varlen(qkv, cu, B, self.n_heads, D // self.n_heads, L + 1, plan=plan)
            o = ops.gemm(att, self._w(sa.out_proj.weight, dtype), sa.out_proj.bias, res=h, res_mode=ops.RES_ADD, m_valid=mv)
            h1 = ops.layernorm(o, lyr.norm1.weight, lyr.norm1.bias, lyr.norm1.eps, rows_dev=nv)
            if ops.FUSE_FFN and dtype == torch.bfloat16 and D == 128 and lyr.linear1.out_features == 4 * D and T >= 128:
                # linear1 + ReLU + linear2 + residual in one kernel: the [T, 512] hidden activation never reaches HBM
                g = ops.ffn_relu(h1, self._w(lyr.linear1.weight, dtype), lyr.linear1.bias, self._w(lyr.linear2.weight, dtype), lyr.linear2.bias, nv)
            else:
                f = ops.gemm(h1, self._w(lyr.linear1.weight, dtype), lyr.linear1.bias, act=ops.ACT_RELU, m_valid=mv)
                g = ops.gemm(f, self._w(lyr.linear2.weight, dtype), lyr.linear2.bias, res=h1, res_mode=ops.RES_ADD, m_valid=mv)
            h = ops.layernorm(g, lyr.norm2.weight, lyr.norm2.bias, lyr.norm2.eps, rows_dev=nv)
        return h, cu


class FocalLoss(nn.Module):
    """HyraxBaselineCLS.py:169-191 (gamma 2, no alpha, eps 0, mean)."""

    def __init__(self, gamma: float = 2.0, alpha=None, eps: float = 0, reduction: str = "mean"):
        super().__init__()
        self.gamma, self.alpha, self.eps, self.reduction = gamma, alpha, eps, reduction

    def forward(self, logits, target):
        from .train import focal_loss

        return focal_loss(logits, target, self.gamma, self.reduction)


class HyraxBaselineCLS(_PhotoEncoderBase):
    """forward((data[B,L,7] f32, pad[B,L] bool True=pad, labels)) -> (B,num_classes) logits|probs,
    or the (B,d_model) CLS embedding when config mode != "photo"."""

    def __init__(self, config, data_sample=None):
        super().__init__()
        self.config = config
        mc = config["model"]["HyraxBaselineCLS"]
        self.criterion = FocalLoss()
        self._build(mc["d_model"], mc["n_heads"], mc["n_layers"], mc["dropout"])
        self.norm = nn.LayerNorm(mc["d_model"])
        self.head = nn.Linear(mc["d_model"], mc["num_classes"])  # unused in forward, kept for checkpoints
        self.classification = mc["mode"] == "photo"
        if self.classification:
            self.fc = nn.Linear(mc["d_model"], mc["num_classes"])
        self.compute_dtype = resolve_dtype(mc.get("compute_dtype"))
        self.optimizer = torch.optim.Adam(self.parameters(), lr=1e-4)
        path = mc.get("pretrained_weights_path_")
        if path:
            self.load_state_dict(torch.load(path), strict=False)
            print(f"Loaded pretrained weights from {path}")

    def encode(self, data, pad, total_tokens=None, packed=None):
        """total_tokens: optional packed token count from the collate (sum of lengths + B, or any upper bound)."""
        h, cu = self.encode_tokens(data, pad, self.compute_dtype, total_tokens, packed)
        cls = ops.gather_cls(h, cu, data.shape[0])
        return ops.layernorm(cls, self.norm.weight, self.norm.bias, self.norm.eps)

    def forward(self, x, total_tokens=None):
        data, pad = x[0], x[1]
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train import photo_forward_train

            return photo_forward_train(self, data, pad, total_tokens)
        out = self.encode(data, pad, total_tokens)
        if self.classification:
            out = ops.gemm(out, self.fc.weight, self.fc.bias)
        if self.config["model"]["HyraxBaselineCLS"]["use_probabilities"]:
            out = ops.softmax_rows(out)
        return out

    def train_step(self, batch):
        from .train import photo_train_step

        return photo_train_step(self, batch)

    @staticmethod
    def to_tensor(data_dict):
        """Same contract as the reference (HyraxBaselineCLS.py:122-166)."""
        import numpy as np

        if "data" not in data_dict:
            raise ValueError("Data dictionary must contain 'data' key.")
        data = data_dict["data"]
        photo_tensor = data["photometry"]
        label_tensor = np.asarray(data.get("label", []), dtype=np.int64)
        photo_tensor[..., :4] = (photo_tensor[..., :4] - data["mean"]) / (data["std"] + 1e-8)
        if "pad_mask" in data.keys():
            return (photo_tensor, data["pad_mask"], label_tensor)
        false_mask = np.zeros((photo_tensor.shape[0], photo_tensor.shape[1] + 1), dtype=bool)
        return (photo_tensor, false_mask, label_tensor)


class MPTModel(_PhotoEncoderBase):
    """Masked-event pre-training of the same encoder (HyraxBaselineCLS.py:194-319): heads flux(1)/band(3)/dt(1),
    loss = lambda_f*L_f * lambda_b*L_b * lambda_dt*L_dt (a product, as in the reference)."""

    def __init__(self, config, data_sample=None):
        super().__init__()
        self.config = config
        mc = config["model"]["HyraxBaselineCLS"]
        self._build(mc["d_model"], mc["n_heads"], mc["n_layers"], mc["dropout"])
        d = mc["d_model"]
        self.head_flux = nn.Linear(d, 1)
        self.head_band = nn.Linear(d, 3)
        self.head_dt = nn.Linear(d, 1)
        self.compute_dtype = resolve_dtype(mc.get("compute_dtype"))
        self.optimizer = torch.optim.AdamW(self.parameters(), lr=1e-4)

    def forward(self, z):
        """z [..., d_model] -> (flux, band logits, dt) like the reference's forward (:238-239)."""
        if not z.is_cuda:
            raise RuntimeError("applecider_b200: inputs must be CUDA tensors (no CPU fallback)")
        lead = z.shape[:-1]
        z2 = z.reshape(-1, z.shape[-1]).float().contiguous()
        w = torch.cat([self.head_flux.weight, self.head_band.weight, self.head_dt.weight]).detach().contiguous()
        b = torch.cat([self.head_flux.bias, self.head_band.bias, self.head_dt.bias]).detach().contiguous()
        o = ops.gemm(z2, w, b).view(*lead, 5)
        return o[..., 0:1], o[..., 1:4], o[..., 4:5]

    def mask_batch(self, data, pad, seed=None):
        """Device version of _mask_batch (:283-319): zeroes channels 2:7 of the drawn tokens IN PLACE, returns the mask."""
        from . import fn

        B, L, _ = data.shape
        masked = torch.empty((B, L), dtype=torch.bool, device=data.device)
        mp = float(self.config["model"]["HyraxBaselineCLS"]["mask_p"])
        ops.call("acb_mpt_mask", data, pad.view(torch.uint8), B, L, mp, fn.next_seed() if seed is None else int(seed), masked.view(torch.uint8))
        return masked

    def train_step(self, batch, masked=None):
        from .train import mpt_train_step

        return mpt_train_step(self, batch, masked)

    to_tensor = staticmethod(HyraxBaselineCLS.to_tensor)


class BaselineCLS(_PhotoEncoderBase):
    """Legacy signature forward(x, pad_mask) -> head(norm(z[:,0])) (Time2Vec.py:80-124)."""

    def __init__(self, d_model, n_heads, n_layers, num_classes, dropout, max_len=None, compute_dtype=None):
        super().__init__()
        self._build(d_model, n_heads, n_layers, dropout)
        self.norm = nn.LayerNorm(d_model)
        self.head = nn.Linear(d_model, num_classes)
        self.compute_dtype = resolve_dtype(compute_dtype)

    def forward(self, x, pad_mask, total_tokens=None):
        h, cu = self.encode_tokens(x, pad_mask, self.compute_dtype, total_tokens)
        cls = ops.gather_cls(h, cu, x.shape[0])
        z = ops.layernorm(cls, self.norm.weight, self.norm.bias, self.norm.eps)
        return ops.gemm(z, self.head.weight, self.head.bias)
