"""Differentiable ops: torch.autograd.Function wrappers whose forward AND backward call the C-ABI kernels.

torch autograd is only the tape; no torch arithmetic runs on activations here.  Conventions as in ops.py:
activations are [rows, C] CUDA tensors (fp32 or bf16), parameters fp32.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import ops
from ._lib import call, dtype_tag

F32 = torch.float32


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _zeros(shape, dev):
    return torch.zeros(shape, dtype=F32, device=dev)


def gemm_ex(A, a_dt, B, b_dt, M, N, K, sam, sak, sbn, sbk, dev, convT=None, splits=1, out=None, accumulate=False):
    """C[m,n] (+)= sum_k A(m,k) B(n,k) with element strides; returns fp32 [M,N]."""
    if out is None:
        out = torch.empty((M, N), dtype=F32, device=dev)
    cv = convT or (0, 0, 0)
    call("acb_gemm_ex", A, a_dt, B, b_dt, out, M, N, K, sam, sak, sbn, sbk, out.shape[-1], int(convT is not None), cv[0], cv[1], cv[2],
         splits, int(accumulate))
    return out


def _splits(k, tiles=None):
    """Split-K factor of the CUDA-core GEMM for a reduction of length k.  With few output tiles (64 x 64 each) a single CTA would
    walk the whole reduction serially (router 144 -> 4 at 1024 rows: 195 us): split until ~2 CTAs per SM are busy."""
    if tiles is not None and tiles < 64:
        return max(1, min(128, k // 128, 296 // max(1, tiles)))
    return max(1, min(128, k // 2048))


def colsum(a, b=None, M=None, N=None, ld=0, a_dt=None, dev=None):
    """out[n] = sum_m a[m*ld+n] (* b[m*ld+n]); a may be a raw device pointer when M, N, ld, a_dt, dev are given."""
    if M is None:
        M, N = a.shape
    out = torch.empty(N, dtype=F32, device=(dev if dev is not None else a.device))
    call("acb_colsum", a, (a_dt if a_dt is not None else dtype_tag(a)), b, (dtype_tag(b) if b is not None else 0), M, N, ld, out, 0)
    return out


def ew(a, b, op, g=None, C=1, s0=1.0, s1=1.0, out_dtype=None):
    y = torch.empty(a.shape, dtype=(out_dtype or a.dtype), device=a.device)
    call("acb_ew", a, dtype_tag(a), b, (dtype_tag(b) if b is not None else 0), g, y, dtype_tag(y), op, C, s0, s1, a.numel())
    return y


def cast_to(x, dtype):
    return ops.cast(_c(x), dtype)


class Cast(Function):
    """Differentiable dtype change (gradient is cast back to the input dtype)."""

    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        return ops.cast(_c(x), dtype)

    @staticmethod
    def backward(ctx, dy):
        return ops.cast(_c(dy), ctx.src), None


def cast(x, dtype):
    return x if x.dtype == dtype else Cast.apply(x, dtype)


def transpose(x, dtype):
    R, C = x.shape
    y = torch.empty((C, R), dtype=dtype, device=x.device)
    call("acb_transpose", _c(x), dtype_tag(x), y, dtype_tag(y), R, C)
    return y


class _TransposedWeights:
    """bf16 W^T copies of Linear weights for the input-gradient GEMMs (dX = dY @ W as a plain x @ B^T GEMM with B = W^T).

    Every step needs all of them again (the optimizer has changed every weight), so they are refreshed TOGETHER: the first
    request that finds its copy stale relaunches ONE `acb_transpose_batch` over every registered weight instead of one small
    transpose kernel per layer (66 launches per fusion training step).  Staleness = the version counter of the parameter (a
    view shares its base's counter; `optim.FusedAdam` bumps it, so does any in-place torch update).  An entry is keyed by
    (data pointer, shape) and holds only a weak reference to the parameter; buffers and the device job table are stable across
    steps, so a CUDA-graph capture records the one batched launch."""

    def __init__(self):
        self._by_dev = {}

    def get(self, W):
        import weakref

        if W.dtype != F32 or not W.is_contiguous() or W.dim() != 2:
            return transpose(W.detach(), BF16)
        base = W._base if W._is_view() and W._base is not None else W
        if not isinstance(base, torch.nn.Parameter):  # temporaries (concatenated / sliced copies of weights) die after the step
            return transpose(W.detach(), BF16)
        reg = self._by_dev.setdefault(W.device, {"entries": {}, "sig": None, "table": None})
        key = (W.data_ptr(), W.shape[0], W.shape[1])
        e = reg["entries"].get(key)
        if e is not None and e[0]() is base:
            if e[2] != base._version:
                self._refresh(reg)
            return e[1]
        out = transpose(W.detach(), BF16)
        reg["entries"][key] = [weakref.ref(base), out, base._version]
        reg["sig"] = None
        return out

    @staticmethod
    def _refresh(reg):
        ents = reg["entries"]
        for k in [k for k, e in ents.items() if e[0]() is None]:  # weights of models that are gone
            del ents[k]
            reg["sig"] = None
        if reg["sig"] is None:
            src, dst, meta, t0 = [], [], [], 0
            for (ptr, R, C), e in ents.items():
                src.append(ptr)
                dst.append(e[1].data_ptr())
                meta += [R, C, (C + 31) // 32, t0]
                t0 += ((R + 31) // 32) * ((C + 31) // 32)
            dev = next(iter(ents.values()))[1].device
            reg["table"] = (torch.tensor(src, dtype=torch.int64).to(dev), torch.tensor(dst, dtype=torch.int64).to(dev),
                            torch.tensor(meta, dtype=torch.int32).to(dev), len(src), t0)
            reg["sig"] = True
        src_t, dst_t, meta_t, n, tiles = reg["table"]
        call("acb_transpose_batch", src_t, dst_t, meta_t, n, tiles)
        for e in ents.values():
            w = e[0]()
            if w is not None:
                e[2] = w._version


_wT = _TransposedWeights()


def transposed_weight(W):
    """bf16 [K, N] copy of the fp32 parameter W [N, K], refreshed in one batched launch per optimizer step."""
    return _wT.get(W)


import os as _os

FOLD_BIAS_GRAD = _os.environ.get("ACB_FOLD_BIAS_GRAD", "1") != "0"  # bias gradient from the wgrad launch (ones tile) vs a colsum pass


def wgrad_tc(dy, ldy, a_col0, M_out, x, nb, L, Cin, taps, pad, x_bstride, x_rstride, dev, want_db=False):
    """tcgen05 weight gradient (bf16 operands, fp32 result [M_out, taps*Cin]); want_db: also the bias gradient (column sums of
    the dY slice) from the same launch -> (dW, db)."""
    out = torch.empty((M_out, taps * Cin), dtype=F32, device=dev)
    # the ones tile is one more N tile of the same K loop: next to >= 8 real N tiles (convolutions: taps * Cin columns) it is free and
    # saves a pass over dY; next to 1-2 (Linear layers) it would cost more than the HBM-speed column sum (measured: +0.25 ms per step)
    if want_db and not (FOLD_BIAS_GRAD and taps * Cin >= 2048):
        call("acb_wgrad_bf16", dy, ldy, a_col0, M_out, x, nb, L, Cin, taps, pad, x_bstride, x_rstride, out, taps * Cin, 0)
        return out, colsum(ops._offset_ptr(dy, a_col0), None, M=nb * L, N=M_out, ld=ldy, a_dt=dtype_tag(dy), dev=dev)
    if want_db:
        db = torch.empty(M_out, dtype=F32, device=dev)
        call("acb_wgrad_bias_bf16", dy, ldy, a_col0, M_out, x, nb, L, Cin, taps, pad, x_bstride, x_rstride, out, taps * Cin, 0, db)
        return out, db
    call("acb_wgrad_bf16", dy, ldy, a_col0, M_out, x, nb, L, Cin, taps, pad, x_bstride, x_rstride, out, taps * Cin, 0)
    return out


BF16 = torch.bfloat16


# ------------------------------------------------------------------------------------------------------
class Linear(Function):
    """y = x W^T + b  (x [M,K] fp32|bf16, W [N,K] fp32 parameter, b [N] or None); output dtype = x dtype
    unless out_dtype is given."""

    @staticmethod
    def forward(ctx, x, W, b, wcast, out_dtype):
        x = _c(x)
        w_use = W if x.dtype == F32 else wcast
        y = ops.gemm(x, w_use, b, out_dtype=out_dtype)
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = _c(dy)
        M, K = x.shape
        N = W.shape[0]
        dx = dW = db = None
        if x.dtype == BF16 and dy.dtype == F32 and N % 8 == 0 and K % 8 == 0 and M >= 64:
            dy = cast_to(dy, BF16)  # fp32 output of a bf16 GEMM: the gradient GEMMs run on the tensor cores too
        tc = x.dtype == BF16 and dy.dtype == BF16 and N % 8 == 0 and K % 8 == 0 and M >= 64
        if ctx.needs_input_grad[0]:
            if tc:  # tcgen05 dgrad: dX = dY @ W through a bf16 W^T copy
                dx = ops.gemm(dy, transposed_weight(W), None)
            else:
                dx = cast_to(gemm_ex(dy, dtype_tag(dy), W, 0, M, K, N, N, 1, 1, K, x.device), x.dtype)
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            if tc and want_db:  # tcgen05 wgrad with MN-major operands (no activation transposes); bias gradient from the same launch
                dW, db = wgrad_tc(dy, N, 0, N, x, 1, M, K, 1, 0, M * K, K, x.device, want_db=True)
            elif tc:
                dW = wgrad_tc(dy, N, 0, N, x, 1, M, K, 1, 0, M * K, K, x.device)
            else:
                dW = gemm_ex(dy, dtype_tag(dy), x, dtype_tag(x), N, K, M, 1, N, 1, K, x.device, splits=_splits(M, ((N + 63) // 64) * ((K + 63) // 64)))
        if want_db and db is None:
            db = colsum(dy)
        return dx, dW, db, None, None


def linear(x, W, b=None, wcast=None, out_dtype=None):
    if x.dtype != F32 and wcast is None:
        wcast = ops.cast(W.detach(), x.dtype)
    return Linear.apply(x, W, b, wcast, out_dtype)


class Act(Function):
    @staticmethod
    def forward(ctx, x, act):
        x = _c(x)
        y = torch.empty_like(x)
        call("acb_act_fwd", x, dtype_tag(x), y, dtype_tag(y), act, x.numel())
        ctx.save_for_backward(x)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(x)
        call("acb_act_bwd", dy, dtype_tag(dy), x, dtype_tag(x), dx, dtype_tag(dx), ctx.act, x.numel())
        return dx, None


def act(x, kind):
    return Act.apply(x, kind)


class LayerNorm(Function):
    """y = [gelu](LN(x)); the optional GELU is fused in both directions (its input is recomputed from x)."""

    @staticmethod
    def forward(ctx, x, w, b, eps, out_dtype, post_act=ops.ACT_NONE):
        x = _c(x)
        y = ops.layernorm(x, w, b, eps, out_dtype=out_dtype, post_act=post_act)
        ctx.save_for_backward(x, w, b)
        ctx.eps, ctx.post_act = eps, post_act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b = ctx.saved_tensors
        dy = _c(dy)
        C = x.shape[-1]
        rows = x.numel() // C
        dx = torch.empty_like(x)
        dwb = _zeros(2 * C, x.device)
        dw, db = dwb[:C], dwb[C:]
        call("acb_layernorm_bwd", x, dtype_tag(x), dy, dtype_tag(dy), w, b, ctx.post_act, dx, dtype_tag(dx), dw, db, rows, C, ctx.eps)
        return dx, dw, db, None, None, None


def layernorm(x, w, b, eps, out_dtype=None, post_act=ops.ACT_NONE):
    return LayerNorm.apply(x, w, b, eps, out_dtype, post_act)


class Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        return ew(_c(a), _c(b), 0)

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    return Add.apply(a, b)


class Mul(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _c(a), _c(b)
        ctx.save_for_backward(a, b)
        return ew(a, b, 1)

    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        dy = _c(dy)
        return ew(dy, b, 1, out_dtype=a.dtype), ew(dy, a, 1, out_dtype=b.dtype)


def mul(a, b):
    return Mul.apply(a, b)


class ScaleAdd(Function):
    """y = x + gamma[col] * v   (ConvNeXt layer scale + residual)."""

    @staticmethod
    def forward(ctx, x, v, gamma):
        x, v = _c(x), _c(v)
        ctx.save_for_backward(v, gamma)
        return ew(x, v, 2, g=gamma, C=gamma.numel())

    @staticmethod
    def backward(ctx, dy):
        v, gamma = ctx.saved_tensors
        dy = _c(dy)
        C = gamma.numel()
        dv = ew(dy, None, 3, g=gamma, C=C, out_dtype=v.dtype)
        dgamma = colsum(dy.view(-1, C), v.view(-1, C))
        return dy, dv, dgamma


def scale_add(x, v, gamma):
    return ScaleAdd.apply(x, v, gamma)


class MlpBlock(Function):
    """ConvNeXt block tail (bf16): out = xres + gamma * fc2(gelu(fc1(y))).

    forward : GEMM1 writes h = gelu(u) AND u (pre_out) from one epilogue; GEMM2 adds bias, scales by gamma, adds the
              residual in its epilogue and also writes v = fc2(h) (needed for d gamma).
    backward: dv = gamma * dy; d h -> d u inside the dgrad GEMM (ACB_RES_MUL_GELU_GRAD with res = u); tcgen05 wgrads."""

    @staticmethod
    def forward(ctx, xres, y, W1, b1, W2, b2, gamma, w1c, w2c):
        xres, y = _c(xres), _c(y)
        M = y.shape[0]
        H, C = W1.shape[0], W2.shape[0]
        u = torch.empty((M, H), dtype=BF16, device=y.device)
        h = ops.gemm(y, w1c, b1, act=ops.ACT_GELU, pre_out=u)
        v = torch.empty((M, C), dtype=BF16, device=y.device)
        out = ops.gemm(h, w2c, b2, res=xres, gamma=gamma, res_mode=ops.RES_ADD, pre_out=v)
        ctx.save_for_backward(y, u, h, v, W1, W2, gamma)
        return out

    @staticmethod
    def backward(ctx, dy):
        y, u, h, v, W1, W2, gamma = ctx.saved_tensors
        dy = _c(dy)
        M = y.shape[0]
        H, C = W1.shape[0], W2.shape[0]
        dgamma = colsum(dy.view(-1, C), v.view(-1, C))
        dv = ew(dy, None, 3, g=gamma, C=C, out_dtype=BF16)
        dW2, db2 = wgrad_tc(dv, C, 0, C, h, 1, M, H, 1, 0, M * H, H, y.device, want_db=True)
        du = ops.gemm(dv, transposed_weight(W2), None, res=u, res_mode=ops.RES_MUL_GELU_GRAD)  # [M,H] = (dv W2) * gelu'(u)
        dW1, db1 = wgrad_tc(du, H, 0, H, y, 1, M, C, 1, 0, M * C, C, y.device, want_db=True)
        dyin = ops.gemm(du, transposed_weight(W1), None)
        return dy, dyin, dW1, db1, dW2, db2, dgamma, None, None


class Attention(Function):
    @staticmethod
    def forward(ctx, qkv, cu, B, H, dh, maxlen, drop_p, seed, plan=None):
        qkv = _c(qkv)
        out = ops.attention_varlen(qkv, cu, B, H, dh, maxlen, drop_p, seed, zero_tail=True, plan=plan)
        ctx.save_for_backward(qkv, cu)
        ctx.cfg = (B, H, dh, maxlen, drop_p, seed)
        ctx.plan = plan
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, cu = ctx.saved_tensors
        B, H, dh, maxlen, drop_p, seed = ctx.cfg
        dout = _c(dout)
        dqkv = torch.zeros_like(qkv)  # capacity rows past the last sequence must carry zero gradient
        plan = ctx.plan
        if (plan is not None and ops.USE_PACKED_ATTENTION and ops.USE_TC_ATTENTION_BWD and qkv.dtype == BF16 and dout.dtype == BF16 and dh == 16
                and H % 4 == 0 and maxlen <= 1024):
            call("acb_attention_packed_bwd", qkv, dout, cu, plan[0], B, plan[1], qkv.shape[0], H, dh, maxlen, drop_p, seed, dqkv)
            return dqkv, None, None, None, None, None, None, None, None
        call("acb_attention_varlen_bwd", qkv, dtype_tag(qkv), dout, dtype_tag(dout), cu, B, H, dh, maxlen, drop_p, seed, dqkv, dtype_tag(dqkv))
        return dqkv, None, None, None, None, None, None, None, None


def attention(qkv, cu, B, H, dh, maxlen, drop_p=0.0, seed=0, plan=None):
    return Attention.apply(qkv, cu, B, H, dh, maxlen, drop_p, seed, plan)


class PhotoEmbed(Function):
    @staticmethod
    def forward(ctx, x, src, T, D, w_in, b_in, w0, b0, w, b, cls_tok, dtype, te_drop_p=0.0, te_seed=0):
        h = ops.photo_embed(x, src, T, D, w_in, b_in, w0, b0, w, b, cls_tok, dtype, te_drop_p, te_seed)
        ctx.save_for_backward(x, src, w, b)
        ctx.dims = (T, D, cls_tok.shape, te_drop_p, te_seed)
        return h

    @staticmethod
    def backward(ctx, dh):
        x, src, w, b = ctx.saved_tensors
        T, D, cls_shape, te_drop_p, te_seed = ctx.dims
        dh = _c(dh)
        g = torch.empty(11 * D, dtype=F32, device=x.device)
        call("acb_photo_embed_bwd", x, src, T, D, dh, dtype_tag(dh), w, b, te_drop_p, te_seed, g)
        return (None, None, None, None, g[: 7 * D].view(D, 7), g[7 * D: 8 * D], g[8 * D: 8 * D + 1], g[8 * D + 1: 8 * D + 2],
                g[8 * D + 2: 9 * D + 1], g[9 * D + 1: 10 * D], g[10 * D: 11 * D].view(cls_shape), None, None, None)


class GatherCls(Function):
    @staticmethod
    def forward(ctx, h, cu, B):
        h = _c(h)
        ctx.save_for_backward(cu)
        ctx.meta = (h.shape, h.dtype, B)
        return ops.gather_cls(h, cu, B)

    @staticmethod
    def backward(ctx, dcls):
        (cu,) = ctx.saved_tensors
        shape, dtype, B = ctx.meta
        dh = torch.empty(shape, dtype=dtype, device=dcls.device)
        call("acb_scatter_cls", _c(dcls), cu, B, shape[1], dh, dtype_tag(dh), shape[0])
        return dh, None, None


class DwConv7(Function):
    @staticmethod
    def forward(ctx, x, w, b, dims):
        x = _c(x)
        B, H, W, C = dims
        y = torch.empty_like(x)
        call("acb_dwconv7", x, dtype_tag(x), w, b, 0, y, dtype_tag(y), B, H, W, C)
        ctx.save_for_backward(x, w)
        ctx.dims = dims
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        B, H, W, C = ctx.dims
        dy = _c(dy)
        dx = torch.empty_like(x)
        call("acb_dwconv7", dy, dtype_tag(dy), w, None, 1, dx, dtype_tag(dx), B, H, W, C)
        dw = torch.empty(w.shape, dtype=F32, device=x.device)
        db = torch.empty(C, dtype=F32, device=x.device)
        call("acb_dwconv7_wgrad", x, dtype_tag(x), dy, dtype_tag(dy), B, H, W, C, dw, db, 0)
        return dx, dw, db, None


class Patch2(Function):
    @staticmethod
    def forward(ctx, x, dims):
        x = _c(x)
        B, H, W, C = dims
        p = torch.empty((B * (H // 2) * (W // 2), 4 * C), dtype=x.dtype, device=x.device)
        call("acb_patch2", x, dtype_tag(x), p, dtype_tag(p), B, H, W, C, 0)
        ctx.meta = (dims, x.shape)
        return p

    @staticmethod
    def backward(ctx, dp):
        (B, H, W, C), shape = ctx.meta
        dp = _c(dp)
        dx = torch.empty(shape, dtype=dp.dtype, device=dp.device)
        call("acb_patch2", dx, dtype_tag(dx), dp, dtype_tag(dp), B, H, W, C, 1)
        return dx, None


class Gap(Function):
    @staticmethod
    def forward(ctx, x, B, HW, C):
        x = _c(x)
        y = torch.empty((B, C), dtype=F32, device=x.device)
        call("acb_gap", x, dtype_tag(x), y, B, HW, C, 0, None, 0)
        ctx.meta = (B, HW, C, x.shape, x.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, HW, C, shape, dtype = ctx.meta
        dx = torch.empty(shape, dtype=dtype, device=dy.device)
        call("acb_gap", None, 0, _c(dy), B, HW, C, 1, dx, dtype_tag(dx))
        return dx, None, None, None


class MaxPool(Function):
    """window 4: [B,L,C] -> [B,L//4,C]; window 0: global max -> [B,C] fp32."""

    @staticmethod
    def forward(ctx, x, B, L, C, window):
        x = _c(x)
        y = ops.maxpool4(x, B, L, C) if window else ops.globalmax(x, B, L, C)
        ctx.save_for_backward(x)
        ctx.meta = (B, L, C, window)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        B, L, C, window = ctx.meta
        dy = _c(dy)
        dx = torch.empty_like(x)
        call("acb_maxpool_bwd", x, dtype_tag(x), dy, dtype_tag(dy), dx, dtype_tag(dx), B, L, C, window)
        return dx, None, None, None, None


class BatchNormResAct(Function):
    """out = act(res + BatchNorm1d(y)) over channels-last rows (legacy variant-B block, brew_cider.py:618-627).

    training: batch statistics, running statistics blended in place (momentum); eval: running statistics.  Backward
    recomputes the pre-activation, so only y and res are kept."""

    @staticmethod
    def forward(ctx, y, res, w, b, bn, training, act, out_dtype):
        y = _c(y)
        res = _c(res) if res is not None else None
        rows, C = y.shape
        dev = y.device
        if bn.momentum is None:
            raise NotImplementedError("BatchNorm with cumulative moving average (momentum=None) is not implemented")
        stats = torch.empty(6 * C, dtype=F32, device=dev)
        sums, scale, shift, mean, rstd = stats[:2 * C], stats[2 * C:3 * C], stats[3 * C:4 * C], stats[4 * C:5 * C], stats[5 * C:]
        if training:
            call("acb_bn_stats", y, dtype_tag(y), rows, C, sums)
            bn.num_batches_tracked.add_(1)
        call("acb_bn_finalize", sums if training else None, rows, C, w, b, bn.eps, float(bn.momentum), int(training), bn.running_mean,
             bn.running_var, scale, shift, mean, rstd)
        out = torch.empty((rows, C), dtype=out_dtype, device=dev)
        call("acb_affine_res_act", y, dtype_tag(y), scale, shift, res, (dtype_tag(res) if res is not None else 0), act, out, dtype_tag(out), rows, C)
        ctx.save_for_backward(y, res, scale, shift, mean, rstd)
        ctx.cfg = (training, act)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, res, scale, shift, mean, rstd = ctx.saved_tensors
        training, act = ctx.cfg
        dout = _c(dout)
        rows, C = y.shape
        dpre = torch.empty((rows, C), dtype=(res.dtype if res is not None else y.dtype), device=y.device)
        sums = torch.empty(2 * C, dtype=F32, device=y.device)
        call("acb_affine_res_act_bwd", y, dtype_tag(y), scale, shift, res, (dtype_tag(res) if res is not None else 0), act, dout, dtype_tag(dout),
             mean, rstd, dpre, dtype_tag(dpre), sums, rows, C)
        dy = torch.empty_like(y)
        call("acb_bn_bwd_apply", y, dtype_tag(y), dpre, dtype_tag(dpre), scale, mean, rstd, sums, int(training), dy, dtype_tag(dy), rows, C)
        return dy, (dpre if res is not None else None), sums[C:], sums[:C], None, None, None, None


class TriPool(Function):
    """[B, L, C] -> [B, L//4, 3C] = [max | mean | min] over windows of 4 (brew_cider.py:629-634)."""

    @staticmethod
    def forward(ctx, x, B, L, C):
        x = _c(x)
        y = torch.empty((B, L // 4, 3 * C), dtype=x.dtype, device=x.device)
        call("acb_tripool4", x, dtype_tag(x), y, dtype_tag(y), B, L, C)
        ctx.save_for_backward(x)
        ctx.dims = (B, L, C)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        B, L, C = ctx.dims
        dy = _c(dy)
        dx = torch.empty_like(x)
        call("acb_tripool4_bwd", x, dtype_tag(x), dy, dtype_tag(dy), dx, dtype_tag(dx), B, L, C)
        return dx, None, None, None


class ConcatCols(Function):
    @staticmethod
    def forward(ctx, *parts):
        rows = parts[0].shape[0]
        widths = [p.shape[1] for p in parts]
        out = torch.empty((rows, sum(widths)), dtype=F32, device=parts[0].device)
        off = 0
        for p, w in zip(parts, widths):
            call("acb_copy2d", _c(p), dtype_tag(p), w, ops._offset_ptr(out, off), 0, out.shape[1], rows, w)
            off += w
        ctx.widths = widths
        return out

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        rows, tot = dy.shape
        outs, off = [], 0
        for w in ctx.widths:
            g = torch.empty((rows, w), dtype=F32, device=dy.device)
            call("acb_copy2d", ops._offset_ptr(dy, off), 0, tot, g, 0, w, rows, w)
            outs.append(g)
            off += w
        return tuple(outs)


class TowerGroup(Function):
    """N ResidualTowerBlocks in one forward and one backward launch (acb_tower_group_fwd/bwd).

    spec = dict(towers=[dict(cols=int32 tensor|None, in_dim, hid, out_dim, y_off, a_off)], ldy, lda, drop_p, seed,
    need_dx, extra_off); X [rows, ldx] fp32; A_pre [rows, lda] fp32 or None (then every tower has its start path W0);
    extra: optional [rows, w] tensor copied into Y at column extra_off (the image features inside the 288-wide concat);
    params: 12 per tower (W0, b0, ln1w, ln1b, W1, b1, ln2w, ln2b, W2, b2, Ws, bs), None where absent."""

    @staticmethod
    def _tables(spec, params, grads):
        import ctypes

        n = len(spec["towers"])
        ptrs = (ctypes.c_longlong * (25 * n))()
        dims = (ctypes.c_int * (5 * n))()
        for t, tw in enumerate(spec["towers"]):
            ptrs[25 * t] = tw["cols"].data_ptr() if tw["cols"] is not None else 0
            for j in range(12):
                p = params[12 * t + j]
                ptrs[25 * t + 1 + j] = p.data_ptr() if p is not None else 0
                g = grads[12 * t + j] if grads is not None else None
                ptrs[25 * t + 13 + j] = g.data_ptr() if g is not None else 0
            dims[5 * t: 5 * t + 5] = [tw["in_dim"], tw["hid"], tw["out_dim"], tw["y_off"], tw["a_off"]]
        return n, ptrs, dims

    @staticmethod
    def forward(ctx, spec, X, A_pre, extra, *params):
        X = _c(X)
        rows = X.shape[0]
        Y = torch.empty((rows, spec["ldy"]), dtype=F32, device=X.device)
        A = _c(A_pre) if A_pre is not None else torch.empty((rows, spec["lda"]), dtype=F32, device=X.device)
        pd = [None if p is None else p.detach() for p in params]
        n, ptrs, dims = TowerGroup._tables(spec, pd, None)
        call("acb_tower_group_fwd", X, X.shape[1], rows, n, ptrs, dims, Y, spec["ldy"], A, spec["lda"], spec["drop_p"], spec["seed"])
        if extra is not None:
            call("acb_copy2d", _c(extra), dtype_tag(extra), extra.shape[1], ops._offset_ptr(Y, spec["extra_off"]), 0, spec["ldy"], rows, extra.shape[1])
            ctx.extra_w = extra.shape[1]
        ctx.spec, ctx.has_apre, ctx.has_extra = spec, A_pre is not None, extra is not None
        ctx.save_for_backward(X, A, *[p for p in pd if p is not None])
        ctx.present = [p is not None for p in pd]
        return Y

    @staticmethod
    def backward(ctx, dY):
        spec = ctx.spec
        X, A = ctx.saved_tensors[:2]
        it = iter(ctx.saved_tensors[2:])
        params = [next(it) if pres else None for pres in ctx.present]
        dY = _c(dY)
        rows = X.shape[0]
        sizes = [0 if p is None else (p.numel() + 3) // 4 * 4 for p in params]
        flat = torch.zeros(sum(sizes), dtype=F32, device=X.device)
        grads, off = [], 0
        for p, sz in zip(params, sizes):
            grads.append(None if p is None else flat[off: off + p.numel()].view(p.shape))
            off += sz
        n, ptrs, dims = TowerGroup._tables(spec, params, grads)
        dA = torch.empty_like(A) if ctx.has_apre else None
        dX = torch.empty_like(X) if spec["need_dx"] else None
        call("acb_tower_group_bwd", X, X.shape[1], rows, n, ptrs, dims, A, spec["lda"], dY, spec["ldy"], dA, dX, spec["drop_p"], spec["seed"])
        dextra = None
        if ctx.has_extra:
            dextra = torch.empty((rows, ctx.extra_w), dtype=F32, device=X.device)
            call("acb_copy2d", ops._offset_ptr(dY, spec["extra_off"]), 0, spec["ldy"], dextra, 0, ctx.extra_w, rows, ctx.extra_w)
        return (None, dX, dA, dextra, *grads)


def tower_params(tw, with_start=True):
    """The 12 parameter slots of a ResidualTowerBlock in acb_tower_group order."""
    skip = isinstance(tw.skip_path, torch.nn.Linear)
    return [tw.start_path[0].weight if with_start else None, tw.start_path[0].bias if with_start else None,
            tw.main_path[0].weight, tw.main_path[0].bias, tw.main_path[2].weight, tw.main_path[2].bias,
            tw.activation[0].weight, tw.activation[0].bias, tw.activation[2].weight, tw.activation[2].bias,
            tw.skip_path.weight if skip else None, tw.skip_path.bias if skip else None]


def gather_cols(X, cols):
    Y = torch.empty((X.shape[0], cols.numel()), dtype=F32, device=X.device)
    call("acb_gather_cols", X, X.shape[1], cols, cols.numel(), Y, X.shape[0])
    return Y


class MoeCombine(Function):
    @staticmethod
    def forward(ctx, gate, eo, E, C):
        gate, eo = _c(gate), _c(eo)
        B = gate.shape[0]
        out = torch.empty((B, C), dtype=F32, device=gate.device)
        call("acb_moe_combine", gate, eo, out, None, B, E, C)
        ctx.save_for_backward(gate, eo)
        ctx.meta = (B, E, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        gate, eo = ctx.saved_tensors
        B, E, C = ctx.meta
        dgate, deo = torch.empty_like(gate), torch.empty_like(eo)
        call("acb_moe_combine_bwd", gate, eo, _c(dout), dgate, deo, B, E, C)
        return dgate, deo, None, None


class L2Norm(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        call("acb_l2norm", x, None, y, x.shape[0], x.shape[1], 0)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        call("acb_l2norm", x, _c(dy), dx, x.shape[0], x.shape[1], 1)
        return dx


class Loss(Function):
    """Focal loss (int64 labels) or soft-target cross entropy (float targets), mean reduction."""

    @staticmethod
    def forward(ctx, logits, labels, soft, gamma):
        logits = _c(logits.float())
        B, C = logits.shape
        loss = torch.empty(1, dtype=F32, device=logits.device)
        dlogits = torch.empty_like(logits)
        call("acb_loss_fwd_bwd", logits, labels, soft, gamma, B, C, loss, dlogits)
        ctx.save_for_backward(dlogits)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return ew(dlogits, None, 5, g=_c(g.float().view(1))), None, None, None


class MptLoss(Function):
    """Masked-event pre-training loss (HyraxBaselineCLS.py:262-278): product of the three masked-token losses."""

    @staticmethod
    def forward(ctx, pred, src, x, masked, L, lam):
        pred = _c(pred)
        T = pred.shape[0]
        losses = torch.empty(4, dtype=F32, device=pred.device)
        dpred = torch.empty((T, 5), dtype=F32, device=pred.device)
        ws = torch.empty(4, dtype=F32, device=pred.device)
        call("acb_mpt_loss_fwd_bwd", pred, dtype_tag(pred), src, T, x, masked, L, lam[0], lam[1], lam[2], losses, dpred, ws)
        ctx.save_for_backward(dpred)
        ctx.pdtype = pred.dtype
        ctx.mark_non_differentiable(losses)
        return losses[0].view(()).clone(), losses

    @staticmethod
    def backward(ctx, g, _):
        (dpred,) = ctx.saved_tensors
        return ew(dpred, None, 5, g=_c(g.float().view(1)), out_dtype=ctx.pdtype), None, None, None, None, None


def focal_loss(logits, target, gamma=2.0, reduction="mean"):
    if reduction != "mean":
        raise NotImplementedError("only mean reduction is implemented")
    return Loss.apply(logits, _c(target.long()), None, float(gamma))


def soft_cross_entropy(logits, target):
    return Loss.apply(logits, None, _c(target.float()), 0.0)


class Dropout(Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        x = _c(x)
        y = torch.empty_like(x)
        call("acb_dropout", x, dtype_tag(x), y, dtype_tag(y), p, seed, x.numel())
        ctx.meta = (p, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, seed = ctx.meta
        dy = _c(dy)
        dx = torch.empty_like(dy)
        call("acb_dropout", dy, dtype_tag(dy), dx, dtype_tag(dx), p, seed, dy.numel())
        return dx, None, None


class Mean3(Function):
    """(a + b + c) / 3  (late-fusion 'avg', brew_cider.py:854)."""

    @staticmethod
    def forward(ctx, a, b, c):
        t = ew(_c(a), _c(b), 0)
        return ew(t, _c(c), 4, s0=1.0 / 3.0, s1=1.0 / 3.0)

    @staticmethod
    def backward(ctx, dy):
        g = ew(_c(dy), None, 4, s0=1.0 / 3.0, s1=0.0)
        return g, g, g


def ew_scaled_sum3(a, b, c):
    return Mean3.apply(a, b, c)


# ---- seeds of the counter-hash RNG streams (dropout, attention dropout, Time2Vec dropout, MPT masking) -----------------
# Every stochastic op draws its 62-bit seed from torch's CUDA generator of the current device: (initial_seed, philox offset)
# mixed with the data-parallel rank, and advances the offset -- the same bookkeeping torch's own CUDA dropout does.  So
# torch.manual_seed / torch.cuda.manual_seed reseed these kernels, replaying a seed replays the masks, and replicas that
# were seeded identically still draw independent masks.
_seed_state = {"rank": None}
_MASK62 = (1 << 62) - 1
_M64 = (1 << 64) - 1


def _mix64(x):
    x &= _M64
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & _M64
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & _M64
    x ^= x >> 33
    return x


def _default_rank():
    import os

    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank()
    return int(os.environ.get("RANK", "0"))


def set_seed(seed=None, rank=None):
    """Reseed the stochastic kernels: seed -> torch.cuda.manual_seed(seed) on the current device (None keeps torch's seed and
    only rewinds nothing); rank (None -> torch.distributed rank or $RANK) decorrelates data-parallel replicas."""
    if seed is not None:
        torch.cuda.manual_seed(int(seed))
    _seed_state["rank"] = _default_rank() if rank is None else int(rank)


def next_seed():
    if _seed_state["rank"] is None:
        _seed_state["rank"] = _default_rank()
    if torch.cuda.is_current_stream_capturing():
        # the generator may not be touched during a CUDA-graph capture: derive the seeds of the captured call sites from the
        # last eager draw and a counter; the graph's device epoch (acb_set_seed_epoch_ptr) varies them from replay to replay
        _seed_state["cap"] = _seed_state.get("cap", 0) + 1
        return _mix64(_seed_state.get("last", 0x5EED) + _seed_state["cap"] * 0x9E3779B97F4A7C15) & _MASK62
    g = torch.cuda.default_generators[torch.cuda.current_device()]
    off = g.get_offset()
    g.set_offset(off + 4)
    s = _mix64(_mix64(g.initial_seed() * 0x9E3779B97F4A7C15 + _seed_state["rank"] + 1) + off) & _MASK62
    _seed_state["last"] = s
    return s


def dropout(x, p, training):
    if not training or p <= 0.0:
        return x
    return Dropout.apply(x, float(p), next_seed())
