"""GPU: every differentiable op in applecider_b200/fn.py (forward + backward kernels) against torch autograd
on a plain fp32 statement of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0, grad=False):
    g = torch.Generator().manual_seed(seed)
    t = (torch.randn(*shape, generator=g) * scale).to(DEV)
    return t.requires_grad_(grad)


def _check(got_out, ref_out, got_inputs, ref_inputs, tol=2e-5, seed=99):
    assert_close(got_out, ref_out, tol, "forward")
    go = _rand(*ref_out.shape, seed=seed)
    got_out.backward(go.to(got_out.dtype))
    ref_out.backward(go)
    for i, (a, b) in enumerate(zip(got_inputs, ref_inputs)):
        s = b.grad.abs().max().clamp_min(1e-12)
        assert_close(a.grad / s, b.grad / s, tol, f"grad of input {i}")


@pytest.mark.parametrize("M,N,K", [(2000, 64, 192), (2048, 64, 192), (130, 70, 50), (5000, 128, 128), (7, 5, 288)])
def test_linear(M, N, K):
    from applecider_b200 import fn

    x, W, b = _rand(M, K, seed=1, grad=True), _rand(N, K, seed=2, scale=K**-0.5, grad=True), _rand(N, seed=3, grad=True)
    x2, W2, b2 = [t.detach().clone().requires_grad_(True) for t in (x, W, b)]
    _check(fn.linear(x, W, b), F.linear(x2, W2, b2), (x, W, b), (x2, W2, b2))


@pytest.mark.parametrize("act,ref", [(1, torch.relu), (2, F.gelu), (3, torch.tanh), (4, torch.sigmoid)])
def test_act(act, ref):
    from applecider_b200 import fn

    x = _rand(300, 37, seed=4, grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    _check(fn.act(x, act), ref(x2), (x,), (x2,))


@pytest.mark.parametrize("rows,C", [(2000, 192), (33, 128), (5, 3072), (70000, 96)])
def test_layernorm(rows, C):
    from applecider_b200 import fn

    x, w, b = _rand(rows, C, seed=5, scale=2.0, grad=True), (1 + 0.1 * _rand(C, seed=6)).requires_grad_(True), _rand(C, seed=7, grad=True)
    x2, w2, b2 = [t.detach().clone().requires_grad_(True) for t in (x, w, b)]
    _check(fn.layernorm(x, w, b, 1e-5), F.layer_norm(x2, (C,), w2, b2, 1e-5), (x, w, b), (x2, w2, b2), tol=5e-5)


@pytest.mark.parametrize("rows,C", [(3000, 192), (2500, 1536), (1100, 3072), (1030, 1024), (70000, 192)])
@pytest.mark.parametrize("gelu", [False, True])
def test_layernorm_bf16(rows, C, gelu):
    """bf16 LayerNorm (+GELU) forward / backward: the register kernels (C <= 768), the wide-row kernel (C >= 1024) and the streaming
    forward (>= 65536 rows) against torch fp32 on the same bf16-rounded input."""
    from applecider_b200 import fn, ops

    x = _rand(rows, C, seed=5, scale=2.0).to(torch.bfloat16).requires_grad_(True)
    w, b = (1 + 0.1 * _rand(C, seed=6)).requires_grad_(True), _rand(C, seed=7, grad=True)
    x2 = x.detach().float().requires_grad_(True)
    w2, b2 = [t.detach().clone().requires_grad_(True) for t in (w, b)]
    y = fn.layernorm(x, w, b, 1e-5, post_act=ops.ACT_GELU if gelu else ops.ACT_NONE)
    ref = F.layer_norm(x2, (C,), w2, b2, 1e-5)
    ref = F.gelu(ref) if gelu else ref
    assert_close(y.float(), ref, 1.2e-2, "forward")
    go = _rand(rows, C, seed=99).to(torch.bfloat16)
    y.backward(go)
    ref.backward(go.float())
    for name, a, r in (("dx", x.grad.float(), x2.grad), ("dw", w.grad, w2.grad), ("db", b.grad, b2.grad)):
        s = r.abs().max().clamp_min(1e-12)
        assert_close(a / s, r / s, 1.2e-2, name)


@pytest.mark.parametrize("B,L,C", [(2, 1000, 64), (2, 250, 128), (3, 15, 16), (2, 1024, 64)])
def test_maxpool(B, L, C):
    from applecider_b200 import fn

    x = _rand(B, L, C, seed=8, grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    _check(fn.MaxPool.apply(x, B, L, C, 4), F.max_pool1d(x2.transpose(1, 2), 4).transpose(1, 2), (x,), (x2,), tol=1e-6)
    x.grad = None
    x2.grad = None
    _check(fn.MaxPool.apply(x, B, L, C, 0), x2.amax(1), (x,), (x2,), tol=1e-6)


def test_linear_dgrad_weight_copies_follow_weight_updates():
    """The bf16 W^T copies behind the tcgen05 input-gradient GEMMs are refreshed in ONE batched launch when a weight has changed
    (fn.transposed_weight): the gradients must follow in-place updates of the weights, for plain parameters and for views."""
    from applecider_b200 import _lib, fn

    bf = torch.bfloat16
    Ws = [torch.nn.Parameter(_rand(n, k, seed=60 + i, scale=k**-0.5)) for i, (n, k) in enumerate([(128, 96), (64, 128), (40, 72)])]
    Wv = torch.nn.Parameter(_rand(2 * 64 * 48, seed=70, scale=0.1))  # used through a reshaping view, like the 1x1 downsample conv weights
    xs = [_rand(256, W.shape[1], seed=80 + i).to(bf).requires_grad_(True) for i, W in enumerate(Ws)]
    xv = _rand(256, 96, seed=90).to(bf).requires_grad_(True)

    def run():
        outs = [fn.linear(x, W) for x, W in zip(xs, Ws)] + [fn.linear(xv, Wv.view(64, 96))]
        for x in xs + [xv]:
            x.grad = None
        sum(o.float().sum() for o in outs).backward()
        return [x.grad.float().clone() for x in xs + [xv]]

    def ref():
        return [W.detach().to(bf).float().sum(0)[None, :].expand(256, -1) for W in Ws] + [Wv.detach().view(64, 96).to(bf).float().sum(0)[None, :].expand(256, -1)]

    for step in range(3):
        before = _lib.lib().acb_launch_count() if hasattr(_lib.lib(), "acb_launch_count") else None
        got = run()
        for g, r in zip(got, ref()):
            assert_close(g, r, 1e-2, f"dX after {step} weight updates")
        with torch.no_grad():  # in-place updates bump the version counters
            for W in Ws + [Wv]:
                W.mul_(-1.5).add_(0.01)
        del before


@pytest.mark.parametrize("rows,C", [(1000, 96), (333, 20)])
def test_elementwise_bf16(rows, C):
    """bf16 elementwise ops and activation backward: 16-byte kernels when the size allows (C = 96), scalar kernels otherwise."""
    from applecider_b200 import fn, ops

    bf = torch.bfloat16
    a, b = _rand(rows, C, seed=21).to(bf), _rand(rows, C, seed=22).to(bf)
    g = _rand(C, seed=23)
    af, bfl = a.float(), b.float()
    for op, ref in ((0, af + bfl), (1, af * bfl), (2, af + g * bfl), (4, af * 0.25 + bfl * 0.5)):
        got = fn.ew(a, b, op, g=g if op == 2 else None, C=C, s0=0.25, s1=0.5)
        assert_close(got.float(), ref, 8e-3, f"ew op {op}")  # (fused multiply-adds: not bit-identical to torch's two roundings)
    assert_close(fn.ew(a, None, 3, g=g, C=C).float(), g * af, 8e-3, "ew op 3")
    for act, tf in ((ops.ACT_RELU, torch.relu), (ops.ACT_GELU, lambda t: F.gelu(t, approximate="tanh"))):
        x2 = af.clone().requires_grad_(True)
        tf(x2).backward(bfl)
        dx = torch.empty_like(a)
        fn.call("acb_act_bwd", b, 1, a, 1, dx, 1, act, a.numel())
        assert_close(dx.float(), x2.grad, 1e-2, f"act_bwd {act}")


@pytest.mark.parametrize("M,N", [(5000, 96), (777, 384), (2049, 2048), (4000, 24), (600, 20)])
def test_colsum_bf16(M, N):
    """Column sums of bf16 matrices (bias gradients): 16-byte kernel when N % 8 == 0, scalar kernel otherwise; optional elementwise factor."""
    from applecider_b200 import fn

    a, b = _rand(M, N, seed=11).to(torch.bfloat16), _rand(M, N, seed=12).to(torch.bfloat16)
    ref = a.float().sum(0)
    s = ref.abs().max().clamp_min(1.0)
    assert_close(fn.colsum(a) / s, ref / s, 2e-5, "colsum")
    ref2 = (a.float() * b.float()).sum(0)
    s2 = ref2.abs().max().clamp_min(1.0)
    assert_close(fn.colsum(a, b) / s2, ref2 / s2, 2e-5, "colsum of products")


@pytest.mark.parametrize("B,L,C", [(3, 1003, 64), (2, 64, 1024), (5, 17, 8), (2, 40, 12)])
def test_maxpool_bf16(B, L, C):
    """bf16 MaxPool1d(4) backward (16-byte vector kernel when C % 8 == 0, scalar otherwise): exact, ties go to the first maximum."""
    from applecider_b200 import fn

    x = _rand(B, L, C, seed=8).to(torch.bfloat16).requires_grad_(True)
    x2 = x.detach().float().requires_grad_(True)
    y = fn.MaxPool.apply(x, B, L, C, 4)
    ref = F.max_pool1d(x2.transpose(1, 2), 4).transpose(1, 2)
    assert torch.equal(y.float().view(ref.shape), ref)
    go = _rand(*ref.shape, seed=9).to(torch.bfloat16)
    y.backward(go.view(y.shape))
    ref.backward(go.float())
    assert torch.equal(x.grad.float(), x2.grad)


@pytest.mark.parametrize("B,L,Cin,Cout,ks", [(2, 250, 64, 128, [3, 31, 251]), (2, 1000, 1, 64, [3, 61, 1021]), (3, 62, 16, 32, [3, 7, 13]), (2, 256, 64, 128, [3, 31, 251])])
def test_spectra_convs(B, L, Cin, Cout, ks):
    from applecider_b200.spectra import SpectraNetBlock
    from applecider_b200.train import SpectraConvs

    torch.manual_seed(0)
    blk = SpectraNetBlock(Cin, Cout, ks, do_pool=True).to(DEV)
    x = _rand(B, L, Cin, seed=9, grad=Cin > 1)
    params = [c.weight for c in blk.convs] + [c.bias for c in blk.convs]
    y = SpectraConvs.apply(x, None, blk, B, L, torch.float32, *params)
    x2 = x.detach().clone().requires_grad_(Cin > 1)
    ref = torch.cat([c(x2.transpose(1, 2)) for c in blk.convs], 1).transpose(1, 2).reshape(B * L, -1)
    assert_close(y, ref, 2e-5, "convs forward")
    go = _rand(*ref.shape, seed=10)
    refs = torch.autograd.grad(ref, ([x2] if Cin > 1 else []) + params, go, retain_graph=True)
    y.backward(go)
    if Cin > 1:
        assert_close(x.grad, refs[0], 3e-5, "dX")
        refs = refs[1:]
    for p, r in zip(params, refs):
        s = r.abs().max().clamp_min(1e-12)
        assert_close(p.grad / s, r / s, 3e-5, f"param grad {tuple(p.shape)}")
        p.grad = None


def test_dwconv_patch_gap():
    from applecider_b200 import fn

    B = 3
    for (H, C) in [(15, 96), (7, 192), (3, 384), (1, 768)]:
        x, w, b = _rand(B, H, H, C, seed=11, grad=True), _rand(C, 1, 7, 7, seed=12, scale=0.15, grad=True), _rand(C, seed=13, grad=True)
        x2, w2, b2 = [t.detach().clone().requires_grad_(True) for t in (x, w, b)]
        got = fn.DwConv7.apply(x.view(B * H * H, C), w, b, (B, H, H, C))
        ref = F.conv2d(x2.permute(0, 3, 1, 2), w2, b2, padding=3, groups=C).permute(0, 2, 3, 1).reshape(B * H * H, C)
        _check(got, ref, (x, w, b), (x2, w2, b2), tol=3e-5)
        if H >= 2:
            x3 = _rand(B * H * H, C, seed=14, grad=True)
            x4 = x3.detach().clone().requires_grad_(True)
            got = fn.Patch2.apply(x3, (B, H, H, C))
            Ho = H // 2
            v = x4.view(B, H, H, C)[:, : 2 * Ho, : 2 * Ho].reshape(B, Ho, 2, Ho, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B * Ho * Ho, 4 * C)
            _check(got, v, (x3,), (x4,), tol=1e-6)
        x5 = _rand(B * H * H, C, seed=15, grad=True)
        x6 = x5.detach().clone().requires_grad_(True)
        _check(fn.Gap.apply(x5, B, H * H, C), x6.view(B, H * H, C).mean(1), (x5,), (x6,), tol=1e-6)


def test_attention_backward():
    from applecider_b200 import fn, ops, synth

    _, pad, _ = synth.photometry_batch(7, seed=16, L=70)
    cu, _ = ops.photo_compact(pad.to(DEV))
    T, D, H = int(cu[-1]), 128, 8
    qkv = _rand(T, 3 * D, seed=17, grad=True)
    q2 = qkv.detach().clone().requires_grad_(True)
    got = fn.attention(qkv, cu, 7, H, 16, 71)
    outs = []
    for bi in range(7):
        s, e = int(cu[bi]), int(cu[bi + 1])
        q, k, v = [z.view(e - s, H, 16).transpose(0, 1) for z in q2[s:e].split(D, 1)]
        p = torch.softmax(q @ k.transpose(1, 2) / 4.0, -1)
        outs.append((p @ v).transpose(0, 1).reshape(e - s, D))
    _check(got, torch.cat(outs, 0), (qkv,), (q2,), tol=2e-5)


def test_small_ops():
    from applecider_b200 import fn

    a, b, c = _rand(50, 32, seed=18, grad=True), _rand(50, 5, seed=19, grad=True), _rand(50, 64, seed=20, grad=True)
    a2, b2, c2 = [t.detach().clone().requires_grad_(True) for t in (a, b, c)]
    _check(fn.ConcatCols.apply(a, b, c), torch.cat([a2, b2, c2], 1), (a, b, c), (a2, b2, c2), tol=1e-6)
    x, y = _rand(40, 24, seed=21, grad=True), _rand(40, 24, seed=22, grad=True)
    x2, y2 = x.detach().clone().requires_grad_(True), y.detach().clone().requires_grad_(True)
    _check(fn.add(fn.mul(x, y), x), x2 * y2 + x2, (x, y), (x2, y2), tol=1e-6)
    v, g = _rand(90, 96, seed=23, grad=True), _rand(96, seed=24, grad=True)
    x = _rand(90, 96, seed=25, grad=True)
    v2, g2, x2 = [t.detach().clone().requires_grad_(True) for t in (v, g, x)]
    _check(fn.scale_add(x, v, g), x2 + g2 * v2, (x, v, g), (x2, v2, g2), tol=2e-6)
    z = _rand(33, 64, seed=26, grad=True)
    z2 = z.detach().clone().requires_grad_(True)
    _check(fn.L2Norm.apply(z), z2 / z2.norm(dim=-1, keepdim=True), (z,), (z2,), tol=2e-6)
    gate = torch.sigmoid(_rand(64, 4, seed=27)).requires_grad_(True)
    eo = _rand(64, 20, seed=28, grad=True)
    gate2, eo2 = gate.detach().clone().requires_grad_(True), eo.detach().clone().requires_grad_(True)
    tw, ti = torch.topk(gate2, 2, -1)
    ref = sum((tw * (ti == e)).sum(-1, keepdim=True) * eo2.view(64, 4, 5)[:, e] for e in range(4))
    _check(fn.MoeCombine.apply(gate, eo, 4, 5), ref, (gate, eo), (gate2, eo2), tol=2e-6)
    p, q, r = _rand(9, 64, seed=29, grad=True), _rand(9, 64, seed=30, grad=True), _rand(9, 64, seed=31, grad=True)
    p2, q2, r2 = [t.detach().clone().requires_grad_(True) for t in (p, q, r)]
    _check(fn.ew_scaled_sum3(p, q, r), (p2 + q2 + r2) / 3, (p, q, r), (p2, q2, r2), tol=2e-6)


def test_losses():
    from applecider_b200 import fn
    from oracle.models import focal_loss as oracle_focal

    z = _rand(37, 5, seed=32, scale=2.0, grad=True)
    z2 = z.detach().clone().requires_grad_(True)
    y = torch.randint(0, 5, (37,), generator=torch.Generator().manual_seed(1)).to(DEV)
    l1, l2 = fn.focal_loss(z, y), oracle_focal(z2, y)
    assert_close(l1, l2, 1e-6, "focal loss")
    (l1 * 3.0).backward()
    (l2 * 3.0).backward()
    assert_close(z.grad, z2.grad, 2e-6, "focal dlogits", atol=1e-8)
    t = torch.softmax(_rand(37, 5, seed=33), -1)
    z.grad = None
    z2.grad = None
    l1, l2 = fn.soft_cross_entropy(z, t), F.cross_entropy(z2, t)
    assert_close(l1, l2, 1e-6, "soft CE")
    l1.backward()
    l2.backward()
    assert_close(z.grad, z2.grad, 2e-6, "CE dlogits", atol=1e-8)


# ---- tcgen05 backward paths (bf16 operands, MN-major UMMA descriptors) ---------------------------------
@pytest.mark.parametrize("M,N,K", [(5000, 384, 96), (2048, 128, 512), (300, 96, 384), (70000, 512, 128), (64, 32, 768)])
def test_wgrad_tc_linear(M, N, K):
    from applecider_b200 import fn

    dy = _rand(M, N, seed=40).to(torch.bfloat16)
    x = _rand(M, K, seed=41).to(torch.bfloat16)
    got = fn.wgrad_tc(dy, N, 0, N, x, 1, M, K, 1, 0, M * K, K, DEV)
    ref = dy.float().T @ x.float()
    assert_close(got / ref.abs().max(), ref / ref.abs().max(), 2e-5, f"wgrad_tc {M}x{N}x{K}")
    t = fn.transpose(_rand(77, 130, seed=42), torch.bfloat16)
    assert torch.equal(t, _rand(77, 130, seed=42).T.contiguous().to(torch.bfloat16))


@pytest.mark.parametrize("M,N,K", [(5000, 384, 96), (1000, 128, 512)])
def test_linear_bf16_backward(M, N, K):
    from applecider_b200 import fn

    x = _rand(M, K, seed=43).to(torch.bfloat16).requires_grad_(True)
    W, b = _rand(N, K, seed=44, scale=K**-0.5, grad=True), _rand(N, seed=45, grad=True)
    y = fn.linear(x, W, b)
    go = _rand(M, N, seed=46).to(torch.bfloat16)
    y.backward(go)
    x2 = x.detach().float().requires_grad_(True)
    W2 = W.detach().to(torch.bfloat16).float().requires_grad_(True)
    F.linear(x2, W2, b.detach()).backward(go.float())
    assert_close(x.grad.float() / x2.grad.abs().max(), x2.grad / x2.grad.abs().max(), 1e-2, "bf16 dX")
    assert_close(W.grad / W2.grad.abs().max(), W2.grad / W2.grad.abs().max(), 2e-5, "bf16 dW (fp32 accumulate)")
    assert_close(b.grad, go.float().sum(0), 1e-5, "bf16 db")


@pytest.mark.parametrize("B,L,Cin,Cout,ks", [(3, 200, 64, 128, [3, 5, 9]), (2, 250, 64, 128, [3, 31, 251]), (5, 62, 128, 256, [3, 15, 61]), (9, 13, 512, 256, [3, 7, 13])])
def test_spectra_convs_bf16_backward(B, L, Cin, Cout, ks):
    from applecider_b200.spectra import SpectraNetBlock
    from applecider_b200.train import SpectraConvs

    torch.manual_seed(0)
    blk = SpectraNetBlock(Cin, Cout, ks, do_pool=True).to(DEV)
    x = _rand(B, L, Cin, seed=47).to(torch.bfloat16).requires_grad_(True)
    params = [c.weight for c in blk.convs] + [c.bias for c in blk.convs]
    y = SpectraConvs.apply(x, None, blk, B, L, torch.bfloat16, *params)
    go = _rand(B * L, 3 * Cout, seed=48).to(torch.bfloat16)
    y.backward(go)
    x2 = x.detach().float().requires_grad_(True)
    ws = [c.weight.detach().to(torch.bfloat16).float().requires_grad_(True) for c in blk.convs]
    ref = torch.cat([F.conv1d(x2.transpose(1, 2), w, c.bias.detach(), padding=k // 2) for w, c, k in zip(ws, blk.convs, ks)], 1).transpose(1, 2).reshape(B * L, -1)
    assert_close(y.float(), ref, 1e-2, "bf16 convs forward")
    ref.backward(go.float())
    assert_close(x.grad.float() / x2.grad.abs().max(), x2.grad / x2.grad.abs().max(), 2e-2, "bf16 conv dX")
    for c, w in zip(blk.convs, ws):
        s = w.grad.abs().max()
        assert_close(c.weight.grad / s, w.grad / s, 3e-5, f"bf16 conv dW k={c.kernel_size[0]}")


@pytest.mark.parametrize("L", [1024, 4096, 1000])
def test_spectra_stage0_bf16_wgrad(L):
    """Stage-0 (1 input channel, k up to 1021) weight gradients through the polyphase tcgen05 wgrad."""
    from applecider_b200.spectra import SpectraNetBlock
    from applecider_b200.train import SpectraConvs

    torch.manual_seed(0)
    B = 2
    blk = SpectraNetBlock(1, 64, [3, 61, 1021], do_pool=True).to(DEV)
    sig = _rand(B, L, seed=50)
    params = [c.weight for c in blk.convs] + [c.bias for c in blk.convs]
    y = SpectraConvs.apply(None, sig, blk, B, L, torch.bfloat16, *params)
    go = _rand(B * L, 192, seed=51).to(torch.bfloat16)
    y.backward(go)
    xb = sig.to(torch.bfloat16).float()[:, None, :]
    ws = [c.weight.detach().clone().requires_grad_(True) for c in blk.convs]
    bs = [c.bias.detach().clone().requires_grad_(True) for c in blk.convs]
    ref = torch.cat([F.conv1d(xb, w, b, padding=w.shape[-1] // 2) for w, b in zip(ws, bs)], 1).transpose(1, 2).reshape(B * L, -1)
    ref.backward(go.float())
    for c, w, b in zip(blk.convs, ws, bs):
        s = w.grad.abs().max()
        assert_close(c.weight.grad / s, w.grad / s, 5e-5, f"stage-0 dW k={c.kernel_size[0]} L={L}")
        assert_close(c.bias.grad, b.grad, 1e-5, "stage-0 db")
