"""Thin Python wrappers over the C-ABI (allocation + argument marshalling only; no math here)."""
from __future__ import annotations

import ctypes

import torch

from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SOFTPLUS, ACT_TANH, BF16, F32, RES_ADD, RES_MUL, RES_MUL_GELU_GRAD, RES_NONE, call, dtype_tag  # noqa: F401


# ---- optional per-region CUDA-event timing (bench.py roofline; off by default) ----------------------
_PROF = None  # dict: tag -> list[(start_event, end_event)]


def profile_start():
    global _PROF
    _PROF = {}


def profile_stop():
    """Returns {tag: [ms, ...]} (synchronises)."""
    global _PROF
    torch.cuda.synchronize()
    out = {k: [a.elapsed_time(b) for a, b in v] for k, v in (_PROF or {}).items()}
    _PROF = None
    return out


class region:
    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if _PROF is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _PROF is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            _PROF.setdefault(self.tag, []).append((self.a, b))
        return False


def _elt(t: torch.Tensor) -> int:
    return t.element_size()


def _offset_ptr(t: torch.Tensor, elems: int) -> int:
    if not t.is_cuda or not t.is_contiguous():
        raise RuntimeError("applecider_b200: output views need a contiguous CUDA base tensor")
    return t.data_ptr() + elems * t.element_size()


def _int_array(vals):
    return (ctypes.c_int * len(vals))(*[int(v) for v in vals])


def pick_bn(n: int) -> int:
    if n % 256 == 0:
        return 256
    if n > 64:
        return 128
    return 64


def gemm(a, w, bias=None, act=ACT_NONE, res=None, gamma=None, res_mode=RES_NONE, out=None, out_dtype=None,
         out_col=0, conv=None, pool4=False, bn=None, tile_kb=None, colblk_off=None, a_view=None, pre_out=None, m_valid=None,
         row_stats=None):
    """C = epilogue(A @ W^T).  a: [M,K] (or [B,L,Cin] with conv=(taps,pad)); w: [N,K'] row-major.

    f32 operands run the CUDA-core kernel, bf16 operands the tcgen05 kernel.
    ``out``/``out_col`` let the result land in a column slice of a wider row-major buffer.
    ``a_view`` = (nbatch, L, Cin, batch_stride, row_stride) overrides the A geometry (bf16 only).
    ``m_valid`` = device pointer to an int32 row count (bf16 only): 128-row tiles past it exit at once, their output rows
    stay unwritten (capacity-sized token matrices, see photo.pack).
    """
    N = w.shape[0]
    ldb = w.shape[1]
    if conv is not None:
        taps, pad = conv
        nb, L, cin = a.shape
        M = nb * L
    else:
        taps, pad = 1, 0
        if a_view is not None:
            nb, L, cin = a_view[0], a_view[1], a_view[2]
            M = nb * L
        else:
            M, cin = a.shape[0], a.shape[-1]
            nb, L = 1, M
    K = taps * cin
    if out is None:
        odt = out_dtype if out_dtype is not None else a.dtype
        rows = M // 4 if pool4 else M
        out = torch.empty((rows, N), dtype=odt, device=a.device)
    ldc = out.shape[-1]
    c_ptr = _offset_ptr(out, out_col)
    ldr = res.shape[-1] if res is not None else 0
    if a.dtype == torch.float32:
        assert w.dtype == torch.float32 and out.dtype == torch.float32 and not pool4 and a_view is None and pre_out is None
        assert res is None or res.dtype == torch.float32
        call("acb_gemm_f32", a, w, c_ptr, M, N, K, cin, ldb, ldc, (L if conv is not None else 0), (cin if conv is not None else 0), pad,
             bias, act, res, ldr, gamma, res_mode)
    elif a.dtype == torch.bfloat16:
        assert w.dtype == torch.bfloat16
        if a_view is not None:
            bstride, rstride = a_view[3], a_view[4]
        else:
            bstride, rstride = L * cin, cin
        if bn is None:
            bn = pick_bn(N)
        kb = _int_array(tile_kb) if tile_kb is not None else None
        co = _int_array(colblk_off) if colblk_off is not None else None
        if row_stats is not None:  # conv / GEMM that also leaves per-row LayerNorm partial sums (one pair per N tile)
            assert co is None and res is None and act == ACT_NONE and not pool4 and pre_out is None and m_valid is None and out_col == 0
            call("acb_gemm_bf16_stats", a, w, c_ptr, dtype_tag(out), nb, L, cin, taps, pad, bstride, rstride, N, ldb, ldc, bn, kb, bias, row_stats)
            return out
        call("acb_gemm_bf16", a, w, c_ptr, dtype_tag(out), nb, L, cin, taps, pad, bstride, rstride, N, ldb, ldc, bn, kb, co,
             bias, act, res, (dtype_tag(res) if res is not None else 0), ldr, gamma, res_mode, int(pool4), m_valid, pre_out)
    else:
        raise TypeError(f"gemm: unsupported dtype {a.dtype}")
    return out


def gemm_ln(a, w, bias, row_stats, parts, ln_w, ln_b, ln_eps, pool4=False, out_dtype=None):
    """[maxpool4]( gelu(LayerNorm(a)) @ w^T + bias ) with the LayerNorm statistics taken from `row_stats` (acb_gemm_ln_bf16)."""
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M // 4 if pool4 else M, N), dtype=(out_dtype or a.dtype), device=a.device)
    call("acb_gemm_ln_bf16", a, w, out, dtype_tag(out), M, K, N, N, bias, int(pool4), row_stats, parts, ln_w, ln_b, float(ln_eps))
    return out


def layernorm(x, w, b, eps, res=None, pre_gelu=False, post_act=ACT_NONE, out_dtype=None, rows_dev=None):
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty(x.shape, dtype=(out_dtype or x.dtype), device=x.device)
    call("acb_layernorm_n", x, dtype_tag(x), res, (dtype_tag(res) if res is not None else 0), w, b, y, dtype_tag(y), rows, C, eps,
         int(pre_gelu), post_act, rows_dev)
    return y


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    call("acb_cast", x, dtype_tag(x), y, dtype_tag(y), x.numel())
    return y


def photo_compact(pad, capacity=0):
    """-> (cu_seqlens[B+1], src_idx[B*(L+1)]); src entries past cu[B] are dead markers (acb_photo_embed writes zero rows)."""
    B, L = pad.shape
    pad_u8 = pad.view(torch.uint8) if pad.dtype == torch.bool else pad
    cu = torch.empty(B + 1, dtype=torch.int32, device=pad.device)
    src = torch.empty(B * (L + 1), dtype=torch.int32, device=pad.device)
    call("acb_photo_compact", pad_u8, B, L, int(capacity), cu, src)
    return cu, src


def check_photo_inputs(data, pad):
    """The mask must cover exactly the (B, L) rows of data: the packing indices address data by b*L+l with no bounds
    checks on the device (the reference raises a mask-shape error in nn.MultiheadAttention for the same input)."""
    if not data.is_cuda or not pad.is_cuda:
        raise RuntimeError("applecider_b200: inputs must be CUDA tensors (no CPU fallback)")
    if data.dim() != 3 or data.shape[-1] != 7:
        raise ValueError(f"applecider_b200: photometry must be (B, L, 7), got {tuple(data.shape)}")
    if tuple(pad.shape) != tuple(data.shape[:2]):
        raise ValueError(f"applecider_b200: pad mask shape {tuple(pad.shape)} does not match the photometry batch {tuple(data.shape[:2])} "
                         "(the mask covers the L events; the CLS column is added internally)")


def token_capacity(B, L, total_tokens=None):
    """Rows to allocate for the packed token matrix: a caller-supplied upper bound of the packed count (exact from the
    collate is best) or the worst case B*(L+1).  Never read back from the device."""
    cap = B * (L + 1)
    if total_tokens is None:
        return cap
    T = int(total_tokens)
    if T < B or T > cap:
        raise ValueError(f"applecider_b200: total_tokens={T} outside [{B}, {cap}] for a ({B}, {L}) batch")
    return T


def photo_embed(x, src, total, D, w_in, b_in, w0, b0, w, b, cls_tok, dtype, te_drop_p=0.0, te_seed=0):
    out = torch.empty((total, D), dtype=dtype, device=x.device)
    call("acb_photo_embed", x, src, None, total, D, w_in, b_in, w0, b0, w, b, cls_tok, te_drop_p, te_seed, out, dtype_tag(out))
    return out


USE_TC_ATTENTION = True
USE_PACKED_ATTENTION = True  # several whole sequences per 128-row tile, 4 heads per CTA, TMA-fed (attention_packed.cu)
USE_TC_ATTENTION_BWD = True  # backward of the packed attention on tcgen05 (attention_packed_bwd_kernel)


def attention_plan(cu, B, total_rows):
    """Tile plan of the packed attention kernel for one batch (acb_attention_plan): (plan int32 tensor, max_tiles).
    Built once per forward and shared by every layer; sized from host-side bounds only."""
    max_tiles = max(1, min(B, 2 * (total_rows // 128) + total_rows // 129 + 2))
    plan = torch.empty(2 + 2 * max_tiles + B, dtype=torch.int32, device=cu.device)
    call("acb_attention_plan", cu, B, max_tiles, plan)
    return plan, max_tiles


def attention_varlen(qkv, cu, B, n_heads, dh, max_seqlen, drop_p=0.0, seed=0, zero_tail=False, plan=None):
    T = qkv.shape[0]
    out = (torch.zeros if zero_tail else torch.empty)((T, n_heads * dh), dtype=qkv.dtype, device=qkv.device)
    if (qkv.dtype == torch.bfloat16 and USE_TC_ATTENTION and USE_PACKED_ATTENTION and plan is not None and max_seqlen <= 1024
            and dh == 16 and n_heads % 4 == 0):
        call("acb_attention_packed", qkv, cu, plan[0], B, plan[1], T, n_heads, dh, max_seqlen, drop_p, seed, out)
        return out
    if qkv.dtype == torch.bfloat16 and USE_TC_ATTENTION and max_seqlen <= 1024 and dh == 16:
        call("acb_attention_varlen_tc", qkv, cu, B, n_heads, dh, max_seqlen, drop_p, seed, out)
        return out
    call("acb_attention_varlen", qkv, dtype_tag(qkv), cu, B, n_heads, dh, max_seqlen, drop_p, seed, out)
    return out


def gather_cls(x, cu, B):
    D = x.shape[-1]
    out = torch.empty((B, D), dtype=torch.float32, device=x.device)
    call("acb_gather_cls", x, dtype_tag(x), cu, B, D, out)
    return out


def patchify(img, p, dtype):
    B, C, H, W = img.shape
    out = torch.empty((B * (H // p) * (W // p), C * p * p), dtype=dtype, device=img.device)
    call("acb_patchify_nchw", img, B, C, H, W, p, out, dtype_tag(out))
    return out


FUSE_CONVNEXT_MLP = True  # fc1 -> GELU -> fc2 -> layer scale -> residual in one tcgen05 kernel (C = 96 / 192), hidden never in HBM


def convnext_mlp(y, res, w1, b1, w2, b2, gamma):
    """out = res + gamma * (fc2(gelu(fc1(y) + b1)) + b2); y, res [M, C] bf16; w1 [4C, C], w2 [C, 4C] bf16 (acb_convnext_mlp_bf16)."""
    M, C = y.shape
    out = torch.empty_like(res)
    call("acb_convnext_mlp_bf16", y, res, w1, b1, w2, b2, gamma, out, M, C)
    return out


FUSE_FFN = True  # transformer feed-forward block (linear1 + ReLU + linear2 + residual) in one tcgen05 kernel, hidden on chip

_ones_cache = {}


def ffn_relu(x, w1, b1, w2, b2, rows_dev=None):
    """out = x + linear2(relu(linear1(x) + b1)) + b2 for bf16 tokens x [T, 128], w1 [512, 128], w2 [128, 512] (acb_ffn_relu_bf16)."""
    T, C = x.shape
    ones = _ones_cache.get((x.device, C))
    if ones is None:
        ones = _ones_cache[(x.device, C)] = torch.ones(C, dtype=torch.float32, device=x.device)
    out = torch.empty_like(x)
    call("acb_ffn_relu_bf16", x, w1, b1, w2, b2, ones, out, T, C, rows_dev)
    return out


def dwconv7_ln(x, B, H, W, C, w, b, ln_w, ln_b, eps):
    y = torch.empty_like(x)
    call("acb_dwconv7_ln", x, dtype_tag(x), w, b, ln_w, ln_b, eps, y, B, H, W, C)
    return y


def ln_patch2(x, B, H, W, C, ln_w, ln_b, eps):
    out = torch.empty((B * (H // 2) * (W // 2), 4 * C), dtype=x.dtype, device=x.device)
    call("acb_ln_patch2", x, dtype_tag(x), ln_w, ln_b, eps, out, B, H, W, C)
    return out


def gap_ln(x, B, HW, C, ln_w, ln_b, eps):
    out = torch.empty((B, C), dtype=torch.float32, device=x.device)
    call("acb_gap_ln", x, dtype_tag(x), ln_w, ln_b, eps, out, B, HW, C)
    return out


def maxpool4(x, B, L, C):
    y = torch.empty((B, L // 4, C), dtype=x.dtype, device=x.device)
    call("acb_maxpool4_cl", x, dtype_tag(x), y, B, L, C)
    return y


def globalmax(x, B, L, C):
    y = torch.empty((B, C), dtype=torch.float32, device=x.device)
    call("acb_globalmax_cl", x, dtype_tag(x), y, B, L, C)
    return y


def softmax_rows(x):
    y = torch.empty_like(x)
    call("acb_softmax_rows", x, y, x.shape[0], x.shape[1])
    return y


def pack_conv_weight(w, out, row_stride, tap_off):
    cout, cin, k = w.shape
    call("acb_pack_conv_weight", w, out, dtype_tag(out), cout, cin, k, row_stride, tap_off)


def pack_conv2d_weight(w, dtype):
    cout, cin, kh, kw = w.shape
    out = torch.empty((cout, kh * kw * cin), dtype=dtype, device=w.device)
    call("acb_pack_conv2d_weight", w, out, dtype_tag(out), cout, cin, kh, kw)
    return out


class DerivedCache:
    """Derived (re-laid-out / down-cast) copies of parameters, refreshed when a source changes.

    The state_dict schema stays the reference's; these buffers are never persisted.
    """

    def __init__(self):
        self._store = {}

    def get(self, key, params, builder):
        sig = tuple((p.data_ptr(), p._version, p.device) for p in params)
        hit = self._store.get(key)
        if hit is not None and hit[0] == sig:
            return hit[1]
        with torch.no_grad():
            val = builder()
        self._store[key] = (sig, val)
        return val

    def clear(self):
        self._store.clear()
