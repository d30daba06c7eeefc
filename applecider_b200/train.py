"""Training path: differentiable forward graphs of the drop-in modules (built from fn.py ops whose forward and
backward are C-ABI kernels) and the reference's train_step semantics.

Reference anchors: HyraxBaselineCLS.train_step (models/HyraxBaselineCLS.py:88-120: focal loss, clip-norm 1.0,
Adam lr 1e-4), AstroMiNN.train_step (models/astrominn.py:308-326: soft-target CE, 11-group AdamW),
SpectraNet.train_step (models/spectranet.py:172-184: injected optimizer/criterion).
Dropout sites follow torch's nn.TransformerEncoderLayer / the reference modules and are active only in
train() mode; eval() + autograd gives deterministic gradients (parity mode).
"""
from __future__ import annotations

import numpy as np
import torch

from . import fn, ops
from .fn import focal_loss  # noqa: F401  (FocalLoss module in photo.py)

F32 = torch.float32


def _wc(cache, p, dtype):
    if dtype == F32:
        return None
    sh = getattr(p, "_acb_bf16", None)  # bf16 shadow kept current by optim.FusedAdam
    if sh is not None and dtype == torch.bfloat16 and sh[1] == p._version:
        return sh[0]
    return cache.get(("cast", id(p)), (p,), lambda: ops.cast(p.detach().contiguous(), dtype))


# ---- photometry -------------------------------------------------------------------------------------------
def photo_encode_train(model, data, pad, total_tokens=None, tokens=False, te_dropout=False):
    """-> (B, d_model) fp32 LayerNorm(CLS) with an autograd graph; tokens=True: (h [T,D] fp32, cu, src) instead."""
    dtype = model.compute_dtype
    tr = model.training
    mc = model.config["model"]["HyraxBaselineCLS"] if hasattr(model, "config") else {"dropout": 0.0}
    p_drop = float(mc.get("dropout", 0.0)) if tr else 0.0
    ops.check_photo_inputs(data, pad)
    B, L, _ = data.shape
    data = data.contiguous().float()
    pad = pad.contiguous()
    if pad.dtype != torch.bool:
        pad = pad != 0
    # T is a row capacity (the collate's packed count, or B*(L+1)); nothing is read back from the device.  Capacity rows past
    # cu[B] are zero after the embedding and in the attention output, every other op is row-local, and their gradient is zero
    # (CLS scatter / loss), so the row reductions of the backward (weight gradients, bias sums) see exact zeros from them.
    T = ops.token_capacity(B, L, total_tokens)
    cu, src = ops.photo_compact(pad, T)
    D, H = model.d_model, model.n_heads
    t2v = model.time2vec
    te_p = float(mc.get("dropout", 0.0)) if te_dropout else 0.0  # MPTModel: F.dropout(te) is always active (:248)
    h = fn.PhotoEmbed.apply(data, src, T, D, model.in_proj.weight, model.in_proj.bias, t2v.w0, t2v.b0, t2v.w, t2v.b, model.cls_tok, dtype,
                            te_p, fn.next_seed() if te_p > 0 else 0)
    dc = model._derived
    n_layers = len(model.encoder.layers)
    plan = ops.attention_plan(cu, B, T) if dtype == torch.bfloat16 and ops.USE_PACKED_ATTENTION else None
    for li, lyr in enumerate(model.encoder.layers):
        sa = lyr.self_attn
        qkv = fn.linear(h, sa.in_proj_weight, sa.in_proj_bias, _wc(dc, sa.in_proj_weight, dtype))
        att = fn.attention(qkv, cu, B, H, D // H, L + 1, p_drop, fn.next_seed() if p_drop > 0 else 0, plan)
        o = fn.linear(att, sa.out_proj.weight, sa.out_proj.bias, _wc(dc, sa.out_proj.weight, dtype))
        o = fn.dropout(o, p_drop, tr)
        h1 = fn.layernorm(fn.add(h, o), lyr.norm1.weight, lyr.norm1.bias, lyr.norm1.eps)
        f = fn.act(fn.linear(h1, lyr.linear1.weight, lyr.linear1.bias, _wc(dc, lyr.linear1.weight, dtype)), ops.ACT_RELU)
        f = fn.dropout(f, p_drop, tr)
        g = fn.linear(f, lyr.linear2.weight, lyr.linear2.bias, _wc(dc, lyr.linear2.weight, dtype))
        g = fn.dropout(g, p_drop, tr)
        h = fn.layernorm(fn.add(h1, g), lyr.norm2.weight, lyr.norm2.bias, lyr.norm2.eps,
                         out_dtype=(F32 if tokens and li == n_layers - 1 else None))
    if tokens:
        return h, cu, src
    cls = fn.GatherCls.apply(h, cu, B)
    return fn.layernorm(cls, model.norm.weight, model.norm.bias, model.norm.eps)


def photo_forward_train(model, data, pad, total_tokens=None):
    out = photo_encode_train(model, data, pad, total_tokens)
    if model.classification:
        out = fn.linear(out, model.fc.weight, model.fc.bias)
    if model.config["model"]["HyraxBaselineCLS"]["use_probabilities"]:
        raise NotImplementedError("training through use_probabilities=True is not implemented")
    return out


def photo_train_step(model, batch):
    """HyraxBaselineCLS.py:88-120."""
    _, _, labels = batch
    decoded = model.forward(batch)
    loss = model.criterion(decoded, labels)
    model.optimizer.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
    model.optimizer.step()
    return {"loss": loss.item(), "num_tdes": np.sum([labels.cpu().numpy() == 4])}


def mpt_losses(model, data, pad, masked, total_tokens=None):
    """data must already be masked (channels 2:7 of masked tokens zeroed).  -> (loss, [loss, L_f, L_b, L_dt])."""
    mc = model.config["model"]["HyraxBaselineCLS"]
    B, L, _ = data.shape
    data = data.contiguous().float()
    h, cu, src = photo_encode_train(model, data, pad, total_tokens, tokens=True, te_dropout=True)
    w = torch.cat([model.head_flux.weight, model.head_band.weight, model.head_dt.weight])
    b = torch.cat([model.head_flux.bias, model.head_band.bias, model.head_dt.bias])
    pred = fn.linear(h, w, b)  # [T,5] fp32; CLS rows are ignored by the loss
    m8 = masked.contiguous().view(torch.uint8) if masked.dtype == torch.bool else masked.contiguous()
    return fn.MptLoss.apply(pred, src, data, m8, L, (float(mc["lambda_f"]), float(mc["lambda_b"]), float(mc["lambda_dt"])))


def mpt_train_step(model, batch, masked=None):
    """HyraxBaselineCLS.py:241-281: mask -> encode -> three heads -> product loss -> clip 1.0 -> AdamW."""
    data, pad = batch[0], batch[1]
    if not data.is_cuda:
        raise RuntimeError("applecider_b200: inputs must be CUDA tensors (no CPU fallback)")
    if pad.dtype != torch.bool:
        pad = pad != 0
    pad = pad.contiguous()
    if masked is None:
        if data.dtype != F32 or not data.is_contiguous():
            raise RuntimeError("applecider_b200: MPTModel.train_step masks `data` in place and needs a contiguous fp32 tensor")
        masked = model.mask_batch(data, pad)
    loss, _ = mpt_losses(model, data, pad, masked)
    model.optimizer.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
    model.optimizer.step()
    return {"loss": loss.item()}


# ---- SpectraNet -------------------------------------------------------------------------------------------
class SpectraConvs(torch.autograd.Function):
    """The three same-padded Conv1d of a SpectraNetBlock on a channels-last signal -> [B*L, 3*Cout].

    forward: implicit-GEMM kernels of spectra.py; backward: dgrad = conv with flipped kernels (implicit GEMM),
    wgrad = transposed-im2col GEMM with split-K, bias = column sums."""

    @staticmethod
    def forward(ctx, x, sig, blk, B, L, dtype, *params):
        ctx.blk, ctx.dims, ctx.dtype = blk, (B, L), dtype
        if dtype == F32:
            y, Lr = blk._convs_f32(x, B, L), L
        elif blk.in_channels == 1 and blk.out_channels % 64:
            # one input channel and a narrow output (legacy variant B stage 1: 16 channels): the polyphase tcgen05 view needs 64-channel
            # column blocks, and 0.14 GFLOP per spectrum is nothing -- fp32 CUDA-core conv, result handed on in the compute dtype
            xf = x if x is not None else sig.view(B, L, 1)
            y, Lr = ops.cast(blk._convs_f32(xf.float().contiguous(), B, L), dtype), L
            x = xf
        elif blk.in_channels == 1:
            y, Lr = blk._convs_bf16_polyphase(sig, B, L)
            if Lr != L:
                y = y.view(B, Lr, -1)[:, :L].contiguous().view(B * L, -1)
        else:
            y, Lr = blk._convs_bf16(x, B, L), L
        ctx.save_for_backward(x if x is not None else sig.view(B, L, 1))
        return y

    @staticmethod
    def backward(ctx, dy):
        blk, (B, L), dtype = ctx.blk, ctx.dims, ctx.dtype
        (x,) = ctx.saved_tensors
        dy = fn._c(dy)
        cin, cout, nk = blk.in_channels, blk.out_channels, blk.k
        ldy = nk * cout
        dev = dy.device
        grads = []
        dx = None
        need_dx = ctx.needs_input_grad[0] and cin > 1
        tc = dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and cin % 64 == 0 and cout % 8 == 0
        if dy.dtype == torch.bfloat16 and cin == 1 and L % 8 == 0 and cout % 64 == 0:
            # stage 0 (1 input channel): polyphase weight gradient on tcgen05.  dY is viewed as [B, L/8, 8*3C] (a GEMM row =
            # 8 positions), X as the overlapping-row view of the zero-padded signal; the 8 phases accumulate (fp32 atomics)
            # into one buffer per kernel size because the output pointer is shifted by (8 - r) columns.
            from .spectra import _HALO, _PHASES

            kp = ((_HALO + blk._kmax() // 2 + _PHASES + 63) // 64) * 64
            stride = L + kp
            xp = torch.zeros((B, stride), dtype=torch.bfloat16, device=dev)
            fn.call("acb_pad_signal", x.view(B, L).float() if x.dtype != F32 else x.view(B, L), xp, 1, B, L, stride, _HALO)
            out = [None] * 6
            for j, conv in enumerate(blk.convs):
                kj = blk.kernel_sizes[j]
                padj = kj // 2
                n_start = ((_HALO - padj) // 8) * 8
                width = ((_HALO + padj + _PHASES - n_start + 63) // 64) * 64
                width = min(width, kp - n_start)
                G = torch.zeros((cout, width + 8), dtype=F32, device=dev)
                if cout == 64:  # all 8 phases in one launch (phase r accumulates (8 - r) columns to the right)
                    fn.call("acb_wgrad_phases_bf16", dy, _PHASES * ldy, j * cout, _PHASES, ldy, ops._offset_ptr(xp, n_start), B, L // _PHASES, width,
                            stride, _PHASES, ops._offset_ptr(G, _PHASES), width + 8, -1)
                else:
                    for r in range(_PHASES):
                        fn.call("acb_wgrad_bf16", dy, _PHASES * ldy, r * ldy + j * cout, cout, ops._offset_ptr(xp, n_start), B, L // _PHASES, width, 1, 0,
                                stride, _PHASES, ops._offset_ptr(G, _PHASES - r), width + 8, 1)
                dW = torch.empty(conv.weight.shape, dtype=F32, device=dev)
                fn.call("acb_copy2d", ops._offset_ptr(G, _HALO - padj - n_start + _PHASES), 0, width + 8, dW, 0, kj, cout, kj)
                out[j] = dW
                out[3 + j] = fn.colsum(ops._offset_ptr(dy, j * cout), None, M=B * L, N=cout, ld=ldy, a_dt=fn.dtype_tag(dy), dev=dev)
            return (None, None, None, None, None, None, *out)
        for j, conv in enumerate(blk.convs):
            kj = blk.kernel_sizes[j]
            dW = torch.empty(conv.weight.shape, dtype=F32, device=dev)
            if tc:
                # tcgen05 wgrad (MN-major operands): G[co, tap*Cin+ci]; unpack = pack with (Cin <-> k) exchanged
                G, db = fn.wgrad_tc(dy, ldy, j * cout, cout, x, B, L, cin, kj, kj // 2, L * cin, cin, dev, want_db=True)
                fn.call("acb_pack_conv_weight", G, dW, 0, cout, kj, cin, cin * kj, 0)
            else:
                # G[(tap,ci), co] = sum_(b,l) X[b, l+tap-pad, ci] * dY[(b,l), j*cout+co]
                G = fn.gemm_ex(x, fn.dtype_tag(x), ops._offset_ptr(dy, j * cout), fn.dtype_tag(dy), kj * cin, cout, B * L, 0, 0, 1, ldy, dev,
                               convT=(L, cin, kj // 2), splits=fn._splits(B * L))
                fn.call("acb_unpack_conv_wgrad", G, dW, cout, cin, kj)
                db = fn.colsum(ops._offset_ptr(dy, j * cout), None, M=B * L, N=cout, ld=ldy, a_dt=fn.dtype_tag(dy), dev=dev)
            grads.append((dW, db))
            if need_dx and tc:
                # tcgen05 dgrad: conv of the dY_j column slice with the flipped/transposed kernel, accumulated over j
                wd = torch.empty((cin, kj * cout), dtype=torch.bfloat16, device=dev)
                fn.call("acb_pack_conv_dgrad_weight", conv.weight, wd, 1, cout, cin, kj)
                first = dx is None
                if first:
                    dx = torch.empty((B * L, cin), dtype=torch.bfloat16, device=dev)
                fn.call("acb_gemm_bf16", ops._offset_ptr(dy, j * cout), wd, dx, 1, B, L, cout, kj, kj // 2, L * ldy, ldy, cin, kj * cout, cin,
                        ops.pick_bn(cin), None, None, None, ops.ACT_NONE, (None if first else dx), 1, cin, None,
                        (ops.RES_NONE if first else ops.RES_ADD), 0, None, None)
            elif need_dx:
                # dgrad: dX = sum_j conv(dY_j, flipped W_j^T)
                wd = torch.empty((cin, kj * cout), dtype=F32, device=dev)
                fn.call("acb_pack_conv_dgrad_weight", conv.weight, wd, 0, cout, cin, kj)
                dyj = torch.empty((B, L, cout), dtype=F32, device=dev)
                fn.call("acb_copy2d", ops._offset_ptr(dy, j * cout), fn.dtype_tag(dy), ldy, dyj, 0, cout, B * L, cout)
                if dx is None:
                    dx = ops.gemm(dyj, wd, None, conv=(kj, kj // 2))
                else:
                    ops.gemm(dyj, wd, None, conv=(kj, kj // 2), res=dx, res_mode=ops.RES_ADD, out=dx)
        if dx is not None:
            dx = fn.cast_to(dx.view(B, L, cin), x.dtype)
        out = [dx, None, None, None, None, None]
        for dW, _ in grads:
            out.append(dW)
        for _, db in grads:
            out.append(db)
        return tuple(out)


def spectra_features_train(model, x):
    dtype = model.compute_dtype
    B, _, L = x.shape
    sig = x.contiguous().float().view(B, L)
    h = sig.view(B, L, 1) if dtype == F32 else None
    for stage in model.all_stages:
        for blk in stage:
            params = [c.weight for c in blk.convs] + [c.bias for c in blk.convs]
            y = SpectraConvs.apply(h, sig if h is None else None, blk, B, L, dtype, *params)
            y = fn.layernorm(y, blk.norm.weight, blk.norm.bias, blk.norm.eps, post_act=ops.ACT_GELU)
            nc = blk.out_channels * blk.k
            if blk.do_pool:
                wd = blk.downsample.weight.view(blk.out_channels, nc)
                z = fn.linear(y, wd, blk.downsample.bias, _wc(blk._derived, blk.downsample.weight, dtype) if dtype != F32 else None)
                z = fn.MaxPool.apply(z.view(B, L, blk.out_channels), B, L, blk.out_channels, 4)
                L = L // 4
                h = z
            else:
                h = y.view(B, L, nc)
    return fn.MaxPool.apply(h, B, L, h.shape[-1], 0)


def spectra_forward_train(model, x):
    feat = spectra_features_train(model, x)
    head = model.regressor if model.redshift else model.classifier
    tr = model.training
    if model.compute_dtype != F32 and feat.shape[0] >= 64:  # Linear(3072 -> 384) on the tensor cores, fp32 result
        f16 = fn.cast(feat, model.compute_dtype)
        z = fn.linear(f16, head[0].weight, head[0].bias, _wc(model._derived, head[0].weight, model.compute_dtype), out_dtype=F32)
    else:
        z = fn.linear(feat, head[0].weight, head[0].bias)
    z = fn.layernorm(z, head[1].weight, head[1].bias, head[1].eps, post_act=ops.ACT_GELU)
    z = fn.dropout(z, head[3].p, tr)
    out = fn.linear(z, head[4].weight, head[4].bias)
    if model.redshift and model.config["model"]["SpectraNet"].get("redshift_softplus", False):
        out = fn.act(out, ops.ACT_SOFTPLUS)
    return out.squeeze(1) if model.redshift else out


def spectra_train_step(model, batch):
    """spectranet.py:172-184 (optimizer / criterion injected by the framework)."""
    _, labels, redshifts = batch
    model.optimizer.zero_grad()
    outputs = model(batch)
    loss = model.criterion(outputs, redshifts if model.redshift else labels)
    loss.backward()
    model.optimizer.step()
    return {"loss": loss.item()}


# ---- AstroMiNN ----------------------------------------------------------------------------------------------
def _tower_train(tw, x, training):
    """ResidualTowerBlock (astrominn.py:59-64) from differentiable ops; x fp32 [B, in]."""
    s = fn.act(fn.linear(x, tw.start_path[0].weight, tw.start_path[0].bias), ops.ACT_GELU)
    m = fn.layernorm(s, tw.main_path[0].weight, tw.main_path[0].bias, tw.main_path[0].eps)
    m = fn.linear(fn.dropout(m, tw.main_path[1].p, training), tw.main_path[2].weight, tw.main_path[2].bias)
    g = fn.layernorm(s, tw.activation[0].weight, tw.activation[0].bias, tw.activation[0].eps)
    g = fn.act(fn.linear(fn.dropout(g, tw.activation[1].p, training), tw.activation[2].weight, tw.activation[2].bias), ops.ACT_SIGMOID)
    skip = fn.linear(x, tw.skip_path.weight, tw.skip_path.bias) if isinstance(tw.skip_path, torch.nn.Linear) else x
    return fn.add(fn.mul(m, g), skip)


def convnext_features_train(bb, img, dtype):
    B, Cin, H, W = img.shape
    img = img.contiguous().float()
    c0 = bb.dims[0]
    a = ops.patchify(img, 4, dtype)
    w0 = bb.stem[0].weight.view(c0, Cin * 16)
    x = fn.linear(a, w0, bb.stem[0].bias, bb._w(bb.stem[0].weight, dtype, (c0, Cin * 16)) if dtype != F32 else None)
    x = fn.layernorm(x, bb.stem[1].weight, bb.stem[1].bias, bb.stem[1].eps)
    h, w = H // 4, W // 4
    for si, st in enumerate(bb.stages):
        C = bb.dims[si]
        if si > 0:
            cp = bb.dims[si - 1]
            xn = fn.layernorm(x, st.downsample[0].weight, st.downsample[0].bias, st.downsample[0].eps)
            p = fn.Patch2.apply(xn, (B, h, w, cp))
            h, w = h // 2, w // 2
            # conv weight (C, cp, 2, 2) viewed in the gather's (ky, kx, ci) column order
            wds = PackDown.apply(st.downsample[1].weight)
            x = fn.linear(p, wds, st.downsample[1].bias, ops.cast(wds.detach(), dtype) if dtype != F32 else None)
        for blk in st.blocks:
            y = fn.DwConv7.apply(x, blk.conv_dw.weight, blk.conv_dw.bias, (B, h, w, C))
            y = fn.layernorm(y, blk.norm.weight, blk.norm.bias, blk.norm.eps)
            if dtype == torch.bfloat16 and FUSE_MLP_BLOCK:
                # fc1 + GELU (pre-activation saved by the same epilogue), fc2 + layer scale + residual in the epilogue, and in
                # the backward the GELU derivative inside the fc2 dgrad GEMM: 2 launches forward instead of 4
                x = fn.MlpBlock.apply(x, y, blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.mlp.fc2.weight, blk.mlp.fc2.bias, blk.gamma,
                                      bb._w(blk.mlp.fc1.weight, dtype), bb._w(blk.mlp.fc2.weight, dtype))
                continue
            hid = fn.act(fn.linear(y, blk.mlp.fc1.weight, blk.mlp.fc1.bias, bb._w(blk.mlp.fc1.weight, dtype) if dtype != F32 else None), ops.ACT_GELU)
            v = fn.linear(hid, blk.mlp.fc2.weight, blk.mlp.fc2.bias, bb._w(blk.mlp.fc2.weight, dtype) if dtype != F32 else None)
            x = fn.scale_add(x, v, blk.gamma)
    g = fn.Gap.apply(x, B, h * w, bb.dims[-1])
    return fn.layernorm(g, bb.head.norm.weight, bb.head.norm.bias, bb.head.norm.eps)


class PackDown(torch.autograd.Function):
    """(Cout, Cin, 2, 2) -> [Cout, (ky,kx,ci)] and the inverse re-layout for its gradient."""

    @staticmethod
    def forward(ctx, w):
        ctx.shape = w.shape
        return ops.pack_conv2d_weight(w.detach(), F32)

    @staticmethod
    def backward(ctx, g):
        cout, cin, kh, kw = ctx.shape
        # inverse of pack_conv2d_weight = the same kernel with the roles (cin <-> kh*kw) exchanged
        out = torch.empty((cout, cin * kh * kw), dtype=F32, device=g.device)
        fn.call("acb_pack_conv_weight", fn._c(g), out, 0, cout, kh * kw, cin, cin * kh * kw, 0)
        return out.view(cout, cin, kh, kw)


def _lin(x, lin, dtype, cache, out_dtype=None):
    """Linear on the tensor cores when the model runs in bf16 and the batch is large enough for a 128-row tile to pay."""
    if dtype != F32 and x.shape[0] >= 64 and lin.in_features % 8 == 0 and lin.out_features % 8 == 0:
        return fn.linear(fn.cast(x, dtype), lin.weight, lin.bias, _wc(cache, lin.weight, dtype), out_dtype=out_dtype)
    return fn.linear(x, lin.weight, lin.bias)


def image_tower_train(it, img, dtype, training):
    feat = convnext_features_train(it.backbone, img, dtype)
    hm, ha = it.head_main, it.head_aux
    dc = it.backbone._derived
    a = fn.layernorm(fn.act(feat, ops.ACT_GELU), hm[1].weight, hm[1].bias, hm[1].eps)
    a = fn.act(_lin(a, hm[2], dtype, dc), ops.ACT_RELU)
    a = fn.dropout(a, hm[4].p, training)
    a = _lin(_lin(a, hm[5], dtype, dc), hm[6], dtype, dc, out_dtype=F32)
    x = fn.act(_lin(fn.layernorm(feat, ha[0].weight, ha[0].bias, ha[0].eps), ha[1], dtype, dc, out_dtype=F32), ops.ACT_TANH)
    return fn.mul(a, x)


FUSE_MLP_BLOCK = True  # ConvNeXt MLP + layer scale + residual as one autograd node with fused epilogues (bf16)
FUSE_TOWERS = True  # tower groups as one forward + one backward launch (acb_tower_group_*)


def _dims(tw):
    return tw.start_path[0].in_features, tw.start_path[0].out_features, tw.main_path[2].out_features


def _towers_fused(model, metadata, image_feats, tr):
    """The 8 metadata towers + the image features -> the 288-wide concat (astrominn.py:249-267), one launch each way."""
    from .astrominn import CONCAT_ORDER

    towers, params, y_off, a_off, extra_off = [], [], 0, 0, 0
    p_drop = 0.0
    for n in CONCAT_ORDER:
        if n == "image":
            extra_off = y_off
            y_off += image_feats.shape[1]
            continue
        tw = getattr(model, f"{n}_tower")
        i, h, o = _dims(tw)
        towers.append(dict(cols=getattr(model, f"_cols_{n}"), in_dim=i, hid=h, out_dim=o, y_off=y_off, a_off=a_off))
        params += fn.tower_params(tw)
        y_off += o
        a_off += h
        p_drop = tw.main_path[1].p if tr else 0.0
    spec = dict(towers=towers, ldy=y_off, lda=a_off, drop_p=float(p_drop), seed=(fn.next_seed() if p_drop > 0 else 0), need_dx=False,
                extra_off=extra_off)
    return fn.TowerGroup.apply(spec, metadata, None, image_feats, *params)


def _experts_fused(model, feats, tr):
    """The 4 experts: start paths as ONE GEMM (pre-GELU), the rest as one fused launch -> [B, 4*5]."""
    exs = list(model.fusion_experts)
    w0 = torch.cat([ex.start_path[0].weight for ex in exs], 0)
    b0 = torch.cat([ex.start_path[0].bias for ex in exs], 0)
    if model.compute_dtype != F32 and feats.shape[0] >= 64:  # [B,288] x [512,288]^T on the tensor cores, fp32 result
        a_pre = fn.linear(fn.cast(feats, model.compute_dtype), w0, b0, None, out_dtype=F32)
    else:
        a_pre = fn.linear(feats, w0, b0)
    towers, params, y_off, a_off = [], [], 0, 0
    for ex in exs:
        i, h, o = _dims(ex)
        towers.append(dict(cols=None, in_dim=i, hid=h, out_dim=o, y_off=y_off, a_off=a_off))
        params += fn.tower_params(ex, with_start=False)
        y_off += o
        a_off += h
    p_drop = exs[0].main_path[1].p if tr else 0.0
    spec = dict(towers=towers, ldy=y_off, lda=a_off, drop_p=float(p_drop), seed=(fn.next_seed() if p_drop > 0 else 0), need_dx=True, extra_off=0)
    return fn.TowerGroup.apply(spec, feats, a_pre, None, *params)


def astrominn_forward_train(model, metadata, image):
    from .astrominn import CONCAT_ORDER

    tr = model.training
    metadata = metadata.contiguous().float()
    fuse = FUSE_TOWERS and len(model.fusion_experts) <= 8
    if fuse:
        feats = _towers_fused(model, metadata, image_tower_train(model.image_tower, image, model.compute_dtype, tr), tr)
    else:
        parts = []
        for n in CONCAT_ORDER:
            if n == "image":
                parts.append(image_tower_train(model.image_tower, image, model.compute_dtype, tr))
            else:
                tw = getattr(model, f"{n}_tower")
                parts.append(_tower_train(tw, fn.gather_cols(metadata, getattr(model, f"_cols_{n}")), tr))
        feats = fn.ConcatCols.apply(*parts)
    r = model.fusion_router
    g1 = fn.act(_lin(feats, r[0], model.compute_dtype, model._derived, out_dtype=F32), ops.ACT_TANH)
    gate = fn.act(fn.linear(fn.dropout(g1, r[2].p, tr), r[3].weight, r[3].bias), ops.ACT_SIGMOID)
    if fuse:
        eo = _experts_fused(model, feats, tr)
    else:
        eo = fn.ConcatCols.apply(*[_tower_train(ex, feats, tr) for ex in model.fusion_experts])
    out = fn.MoeCombine.apply(gate, eo, model.num_mlp_experts, 5)
    if model.config["model"]["AstroMiNN"]["use_probabilities"]:
        raise NotImplementedError("training through use_probabilities=True is not implemented")
    return out


class _SoftCE(torch.nn.Module):
    def forward(self, logits, target):
        return fn.soft_cross_entropy(logits, target)


def astrominn_train_step(model, batch):
    """astrominn.py:308-326."""
    _, _, labels = batch
    model.this_optimizer.zero_grad()
    logits = model.forward(batch)
    crit = model.this_criterion
    loss = fn.soft_cross_entropy(logits, labels) if isinstance(crit, torch.nn.CrossEntropyLoss) and labels.dtype.is_floating_point else crit(logits, labels)
    model._update_stats(loss.item())
    loss.backward()
    model.this_optimizer.step()
    return {"loss": model._calculate_stats()}


# ---- fusion ---------------------------------------------------------------------------------------------------
def fusion_forward_train(model, photometry, photometry_mask, metadata, images, spectra, total_tokens=None):
    p = photo_encode_train(model.photometry_encoder, photometry, photometry_mask, total_tokens)
    s = model.spectra_encoder(spectra) if getattr(model, "spectra_variant", "src") == "B" else spectra_forward_train(model.spectra_encoder, spectra)
    if s.dim() == 1:
        s = s[:, None]
    im = astrominn_forward_train(model.img_metadata_encoder, metadata, images)
    pe = fn.L2Norm.apply(fn.linear(p, model.photometry_proj.weight, model.photometry_proj.bias))
    ie = fn.L2Norm.apply(fn.linear(im, model.img_metadata_proj.weight, model.img_metadata_proj.bias))
    se = fn.L2Norm.apply(fn.linear(s, model.spectra_proj.weight, model.spectra_proj.bias))
    if model.fusion == "concat":
        emb = fn.ConcatCols.apply(pe, ie, se)
    else:
        emb = fn.ew_scaled_sum3(pe, ie, se)
    return fn.linear(emb, model.fc.weight, model.fc.bias)
