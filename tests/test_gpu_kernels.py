"""GPU kernel parity tests (run on the B200 box): every C-ABI kernel against a plain torch fp32
statement of the same op.  Integer/index outputs must be bit-exact; fp32 kernels within 1e-5 of the
result scale; bf16 tensor-core kernels within the bf16 bound stated per test."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import assert_close

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from applecider_b200 import ops

    return ops


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


ACTS = {0: lambda v: v, 1: torch.relu, 2: F.gelu, 3: torch.tanh, 4: torch.sigmoid}


# ---------------------------------------------------------------------------------------------------
# fp32 GEMM
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(130, 70, 50), (64, 64, 16), (1, 5, 128), (257, 384, 128), (33, 9, 384)])
@pytest.mark.parametrize("act", [0, 1, 2, 3, 4])
def test_gemm_f32_plain(M, N, K, act):
    ops = _ops()
    a, w, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K**-0.5), _rand(N, seed=3)
    got = ops.gemm(a, w, b, act=act)
    ref = ACTS[act](a.double() @ w.double().T + b.double()).float()
    assert_close(got, ref, 1e-5, f"gemm_f32 {M}x{N}x{K} act{act}")


def test_gemm_f32_residual_modes_and_slices():
    ops = _ops()
    M, N, K = 100, 32, 48
    a, w, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=K**-0.5), _rand(N, seed=3)
    res, gamma = _rand(M, N, seed=4), _rand(N, seed=5)
    v = a @ w.T + b
    assert_close(ops.gemm(a, w, b, res=res, gamma=gamma, res_mode=ops.RES_ADD), res + gamma * v, 1e-5, "res_add")
    assert_close(ops.gemm(a, w, b, res=res, res_mode=ops.RES_ADD), res + v, 1e-5, "res_add nogamma")
    assert_close(ops.gemm(a, w, b, act=3, res=res, res_mode=ops.RES_MUL), res * torch.tanh(v), 1e-5, "res_mul")
    wide = torch.full((M, 100), 7.0, device=DEV)
    ops.gemm(a, w, b, out=wide, out_col=40)
    assert_close(wide[:, 40:72], v, 1e-5, "slice")
    assert (wide[:, :40] == 7).all() and (wide[:, 72:] == 7).all()


@pytest.mark.parametrize("B,L,Cin,Cout,k", [(2, 50, 1, 8, 9), (3, 37, 16, 24, 5), (2, 130, 8, 16, 33), (2, 13, 32, 8, 7)])
def test_gemm_f32_conv1d(B, L, Cin, Cout, k):
    ops = _ops()
    x = _rand(B, L, Cin, seed=1)
    w = _rand(Cout, Cin, k, seed=2, scale=(Cin * k) ** -0.5)
    b = _rand(Cout, seed=3)
    wp = torch.zeros(Cout, k * Cin, device=DEV)
    ops.pack_conv_weight(w, wp, k * Cin, 0)
    assert torch.equal(wp, w.permute(0, 2, 1).reshape(Cout, -1)), "pack_conv_weight mismatch"
    got = ops.gemm(x, wp, b, conv=(k, k // 2))
    ref = F.conv1d(x.transpose(1, 2), w, b, padding=k // 2).transpose(1, 2).reshape(B * L, Cout)
    assert_close(got, ref, 1e-5, "conv1d f32")


# ---------------------------------------------------------------------------------------------------
# tcgen05 bf16 GEMM
# ---------------------------------------------------------------------------------------------------
def _bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 128, 64), (256, 128, 128), (128, 256, 64), (1000, 384, 128),
                                   (300, 96, 48), (4096, 512, 128), (515, 128, 512), (77, 32, 768), (8, 384, 3072), (640, 192, 96)])
def test_gemm_bf16_plain(M, N, K):
    ops = _ops()
    a, w, b = _bf(_rand(M, K, seed=1)), _bf(_rand(N, K, seed=2, scale=K**-0.5)), _rand(N, seed=3)
    ref = a.float() @ w.float().T + b
    got = ops.gemm(a, w, b, out_dtype=torch.float32)
    assert_close(got, ref, 2e-5, f"gemm_bf16->f32 {M}x{N}x{K}", atol=1e-5)
    got16 = ops.gemm(a, w, b)
    assert got16.dtype == torch.bfloat16
    assert_close(got16, ref, 8e-3, f"gemm_bf16->bf16 {M}x{N}x{K}")


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_bf16_tile_widths(bn):
    ops = _ops()
    M, N, K = 384, 512, 192
    a, w = _bf(_rand(M, K, seed=1)), _bf(_rand(N, K, seed=2, scale=K**-0.5))
    got = ops.gemm(a, w, None, out_dtype=torch.float32, bn=bn)
    assert_close(got, a.float() @ w.float().T, 2e-5, f"bn{bn}", atol=1e-5)


def test_gemm_bf16_epilogues():
    ops = _ops()
    M, N, K = 260, 128, 128
    a, w, b = _bf(_rand(M, K, seed=1)), _bf(_rand(N, K, seed=2, scale=K**-0.5)), _rand(N, seed=3)
    v = a.float() @ w.float().T + b
    for act in (1, 2, 3):
        assert_close(ops.gemm(a, w, b, act=act, out_dtype=torch.float32), ACTS[act](v), 3e-5, f"tc act{act}", atol=1e-5)
    res32, gamma = _rand(M, N, seed=4), _rand(N, seed=5)
    assert_close(ops.gemm(a, w, b, res=res32, gamma=gamma, res_mode=ops.RES_ADD, out_dtype=torch.float32), res32 + gamma * v, 3e-5, "tc res_add f32", atol=1e-5)
    res16 = _bf(res32)
    assert_close(ops.gemm(a, w, b, res=res16, res_mode=ops.RES_ADD, out_dtype=torch.float32), res16.float() + v, 3e-5, "tc res_add bf16", atol=1e-5)
    assert_close(ops.gemm(a, w, b, act=3, res=res32, res_mode=ops.RES_MUL, out_dtype=torch.float32), res32 * torch.tanh(v), 3e-5, "tc res_mul", atol=1e-5)
    # column slice of a wider buffer
    wide = torch.full((M, 288), 7.0, device=DEV)
    ops.gemm(a, w[:32].contiguous(), b[:32].contiguous(), out=wide, out_col=224)
    assert_close(wide[:, 224:256], v[:, :32], 3e-5, "tc slice", atol=1e-5)
    assert (wide[:, :224] == 7).all() and (wide[:, 256:] == 7).all()
    # fused MaxPool(4) over rows
    got = ops.gemm(a, w, b, pool4=True, out_dtype=torch.float32)
    assert_close(got, v.view(M // 4, 4, N).amax(1), 3e-5, "tc pool4", atol=1e-5)


@pytest.mark.parametrize("B,L,Cin,Cout,k", [(3, 200, 64, 128, 5), (2, 128, 64, 128, 3), (5, 50, 128, 128, 7), (9, 13, 512, 256, 7),
                                             (2, 1024, 64, 128, 31), (3, 217, 128, 256, 15)])
def test_gemm_bf16_conv1d(B, L, Cin, Cout, k):
    ops = _ops()
    x = _bf(_rand(B, L, Cin, seed=1))
    w = _rand(Cout, Cin, k, seed=2, scale=(Cin * k) ** -0.5)
    b = _rand(Cout, seed=3)
    wp = torch.zeros(Cout, k * Cin, device=DEV, dtype=torch.bfloat16)
    ops.pack_conv_weight(w, wp, k * Cin, 0)
    got = ops.gemm(x, wp, b, conv=(k, k // 2), out_dtype=torch.float32)
    ref = F.conv1d(x.float().transpose(1, 2), _bf(w).float(), b, padding=k // 2).transpose(1, 2).reshape(B * L, Cout)
    assert_close(got, ref, 3e-5, f"conv1d bf16 L={L}", atol=1e-5)


def test_gemm_bf16_k_ranges_and_column_remap():
    ops = _ops()
    M, K = 256, 512
    a = _bf(_rand(M, K, seed=1))
    w = _bf(_rand(256, K, seed=2, scale=K**-0.5))
    # tile 0 only uses K blocks [2,5), tile 1 uses [0,8): emulate by zeroing the weights outside
    wz = w.clone()
    wz[:128, :128] = 0
    wz[:128, 320:] = 0
    ref = a.float() @ wz.float().T
    coloff = [64, 0, 192, 128]  # swap 64-column blocks pairwise
    got = ops.gemm(a, w, None, out_dtype=torch.float32, bn=128, tile_kb=[2, 5, 0, 8], colblk_off=coloff)
    ref_perm = torch.cat([ref[:, 64:128], ref[:, 0:64], ref[:, 192:256], ref[:, 128:192]], 1)
    assert_close(got, ref_perm, 3e-5, "k-ranges + column remap", atol=1e-5)


def test_spectra_stage0_polyphase_conv():
    """k up to 1021 on a 1-channel signal through the 8-phase TMA view == F.conv1d."""
    from applecider_b200.spectra import SpectraNetBlock

    torch.manual_seed(0)
    for L in (4096, 3481, 1000):
        blk = SpectraNetBlock(1, 64, [3, 61, 1021], do_pool=True).to(DEV)
        x = _rand(2, L, seed=L)
        y, L8 = blk._convs_bf16_polyphase(x, 2, L)
        y = y.view(2, L8, 192)[:, :L].float()
        xb = _bf(x).float()[:, None, :]
        ref = torch.cat([F.conv1d(xb, _bf(c.weight).float(), c.bias, padding=c.kernel_size[0] // 2) for c in blk.convs], 1).transpose(1, 2)
        assert_close(y, ref, 8e-3, f"polyphase L={L}")


# ---------------------------------------------------------------------------------------------------
# row kernels
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,C", [(7, 128), (33, 192), (5, 3072), (4, 10), (1000, 96)])
@pytest.mark.parametrize("dt_in,dt_out", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16)])
def test_layernorm(rows, C, dt_in, dt_out):
    ops = _ops()
    x = _rand(rows, C, seed=1, scale=2.0).to(dt_in)
    r = _rand(rows, C, seed=2).to(dt_in)
    w, b = 1 + 0.1 * _rand(C, seed=3), _rand(C, seed=4)
    tol = 1e-5 if dt_out == torch.float32 else 8e-3
    assert_close(ops.layernorm(x, w, b, 1e-5, out_dtype=dt_out), F.layer_norm(x.float(), (C,), w, b, 1e-5), tol, "ln")
    assert_close(ops.layernorm(x, w, b, 1e-6, res=r, post_act=ops.ACT_GELU, out_dtype=dt_out), F.gelu(F.layer_norm(x.float() + r.float(), (C,), w, b, 1e-6)), tol, "ln+res+gelu")
    assert_close(ops.layernorm(x, w, b, 1e-5, pre_gelu=True, out_dtype=dt_out), F.layer_norm(F.gelu(x.float()), (C,), w, b, 1e-5), tol, "gelu+ln")


@pytest.mark.parametrize("C", [128, 192, 384])
def test_layernorm_stream(C):
    """>= 65536 bf16 rows take the streaming kernels (C = 192: two rows per warp); odd row count exercises the guarded tail."""
    ops = _ops()
    rows = 65536 + 77
    x = _rand(rows, C, seed=5, scale=2.0).to(torch.bfloat16)
    w, b = 1 + 0.1 * _rand(C, seed=3), _rand(C, seed=4)
    ref = F.gelu(F.layer_norm(x.float(), (C,), w, b, 1e-5))
    assert_close(ops.layernorm(x, w, b, 1e-5, post_act=ops.ACT_GELU), ref, 8e-3, f"stream ln+gelu C={C}")
    assert_close(ops.layernorm(x, w, b, 1e-5), F.layer_norm(x.float(), (C,), w, b, 1e-5), 8e-3, f"stream ln C={C}")
    # small-batch (cached kernel) and large-batch (streaming kernel) results agree to bf16 rounding
    small = ops.layernorm(x[:1000].contiguous(), w, b, 1e-5).float()
    assert_close(ops.layernorm(x, w, b, 1e-5)[:1000].float(), small, 8e-3, "stream vs cached")


def test_cast_pool_softmax():
    ops = _ops()
    x = _rand(3, 43, 20, seed=1)
    assert torch.equal(ops.cast(x, torch.bfloat16), x.to(torch.bfloat16))
    assert torch.equal(ops.maxpool4(x, 3, 43, 20), F.max_pool1d(x.transpose(1, 2), 4).transpose(1, 2).contiguous())
    assert torch.equal(ops.globalmax(x, 3, 43, 20), x.amax(1))
    xb = x.to(torch.bfloat16)
    assert torch.equal(ops.maxpool4(xb, 3, 43, 20), F.max_pool1d(xb.float().transpose(1, 2), 4).transpose(1, 2).contiguous().to(torch.bfloat16))
    z = _rand(11, 5, seed=2)
    assert_close(ops.softmax_rows(z), torch.softmax(z, 1), 1e-6, "softmax")


# ---------------------------------------------------------------------------------------------------
# photometry pieces
# ---------------------------------------------------------------------------------------------------
def test_photo_compact_is_exact():
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    B, L = 37, 77
    pad = torch.rand(B, L, generator=g) < 0.4  # arbitrary (non-suffix) masks
    pad[3] = True  # fully padded row -> CLS only
    pad[4] = False
    cu, src = ops.photo_compact(pad.to(DEV))
    cu, src = cu.cpu(), src.cpu()
    lens = (~pad).sum(1) + 1
    assert torch.equal(cu, torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)]).int())
    exp = []
    for b in range(B):
        exp.append(-1 - b)
        exp += [b * L + l for l in range(L) if not pad[b, l]]
    assert src[: len(exp)].tolist() == exp


def test_photo_embed_and_attention():
    ops = _ops()
    from applecider_b200 import synth

    x, pad, lens = synth.photometry_batch(9, seed=3)
    x, pad = x.to(DEV), pad.to(DEV)
    B, L, D, H = 9, x.shape[1], 128, 8
    w_in, b_in = _rand(D, 7, seed=1, scale=0.4), _rand(D, seed=2)
    w0, b0, w, b, cls = _rand(1, seed=3), _rand(1, seed=4), _rand(D - 1, seed=5), _rand(D - 1, seed=6), _rand(1, 1, D, seed=7)
    cu, src = ops.photo_compact(pad)
    T = int(cu[-1])
    h = ops.photo_embed(x, src, T, D, w_in, b_in, w0, b0, w, b, cls, torch.float32)
    t = x[..., 0]
    full = x @ w_in.T + b_in + torch.cat([(w0 * t + b0)[..., None], torch.sin(t[..., None] * w + b)], -1)
    full = torch.cat([cls.expand(B, 1, D), full], 1)  # (B, L+1, D)
    keep = torch.cat([torch.ones(B, 1, dtype=torch.bool, device=DEV), ~pad], 1)
    assert_close(h, full[keep], 2e-6, "photo_embed")

    qkv = _rand(T, 3 * D, seed=8)
    att = ops.attention_varlen(qkv, cu, B, H, D // H, L + 1)
    ref = torch.empty_like(att)
    for bi in range(B):
        s, e = int(cu[bi]), int(cu[bi + 1])
        q, k, v = [z.view(e - s, H, D // H).transpose(0, 1) for z in qkv[s:e].split(D, 1)]
        p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(D // H), -1)
        ref[s:e] = (p @ v).transpose(0, 1).reshape(e - s, D)
    assert_close(att, ref, 2e-6, "attention f32")
    att16 = ops.attention_varlen(qkv.to(torch.bfloat16), cu, B, H, D // H, L + 1)
    assert_close(att16, ref, 1.5e-2, "attention bf16")
    cl = ops.gather_cls(h, cu, B)
    assert torch.equal(cl, h[cu[:-1].long()])


# ---------------------------------------------------------------------------------------------------
# ConvNeXt pieces
# ---------------------------------------------------------------------------------------------------
def test_convnext_pieces():
    ops = _ops()
    B = 3
    img = _rand(B, 3, 63, 63, seed=1)
    w = _rand(96, 3, 4, 4, seed=2, scale=0.15)
    a = ops.patchify(img, 4, torch.float32)
    ref = F.conv2d(img, w, stride=4).permute(0, 2, 3, 1).reshape(-1, 96)
    assert_close(a @ w.reshape(96, -1).T, ref, 1e-5, "patchify+gemm == conv k4s4")

    for (H, C) in [(15, 96), (7, 192), (3, 384), (1, 768)]:
        x = _rand(B, H, H, C, seed=H)
        dw, db = _rand(C, 1, 7, 7, seed=3, scale=0.15), _rand(C, seed=4)
        lw, lb = 1 + 0.1 * _rand(C, seed=5), _rand(C, seed=6)
        y = ops.dwconv7_ln(x, B, H, H, C, dw, db, lw, lb, 1e-6)
        ref = F.layer_norm(F.conv2d(x.permute(0, 3, 1, 2), dw, db, padding=3, groups=C).permute(0, 2, 3, 1), (C,), lw, lb, 1e-6)
        assert_close(y, ref, 2e-5, f"dwconv7_ln {H}x{H}x{C}")
        yb = ops.dwconv7_ln(x.to(torch.bfloat16), B, H, H, C, dw, db, lw, lb, 1e-6)
        assert_close(yb, ref, 2e-2, f"dwconv7_ln bf16 {H}")
        if H >= 2:
            cw, cb = _rand(2 * C, C, 2, 2, seed=7, scale=(4 * C) ** -0.5), _rand(2 * C, seed=8)
            pa = ops.ln_patch2(x, B, H, H, C, lw, lb, 1e-6)
            wp = ops.pack_conv2d_weight(cw, torch.float32)
            got = ops.gemm(pa, wp, cb)
            xn = F.layer_norm(x, (C,), lw, lb, 1e-6).permute(0, 3, 1, 2)
            ref = F.conv2d(xn, cw, cb, stride=2).permute(0, 2, 3, 1).reshape(-1, 2 * C)
            assert_close(got, ref, 2e-5, f"ln_patch2+gemm == LN2d+conv k2s2 ({H})")
        g = ops.gap_ln(x, B, H * H, C, lw, lb, 1e-6)
        assert_close(g, F.layer_norm(x.mean((1, 2)), (C,), lw, lb, 1e-6), 1e-5, "gap_ln")


# ---------------------------------------------------------------------------------------------------
# towers / MoE / fusion head
# ---------------------------------------------------------------------------------------------------
def test_tower_moe_fusion_head():
    ops = _ops()
    import applecider_b200 as ab
    from oracle import models as om

    torch.manual_seed(0)
    X = _rand(50, 24, seed=1)
    for (i, h, o, cols) in [(2, 16, 32, [5, 14]), (19, 128, 32, list(range(19))), (12, 48, 32, [6, 9, 10, 13, 15, 17, 18, 19, 20, 21, 22, 23])]:
        ref_m = om.ResidualTowerBlock(i, h, o).eval()
        m = ab.ResidualTowerBlock(i, h, o)
        m.load_state_dict(ref_m.state_dict())
        m = m.to(DEV)
        Y = torch.full((50, 40), 3.0, device=DEV)
        m.run(X, torch.tensor(cols, dtype=torch.int32, device=DEV), Y, 4)
        assert_close(Y[:, 4:4 + o], ref_m(X.cpu()[:, cols]), 1e-5, f"tower {i}->{h}->{o}")
        assert (Y[:, :4] == 3).all() and (Y[:, 4 + o:] == 3).all()
    # identity skip (in == out) and expert-sized block
    for (i, h, o) in [(32, 16, 32), (288, 128, 5)]:
        ref_m = om.ResidualTowerBlock(i, h, o).eval()
        m = ab.ResidualTowerBlock(i, h, o)
        m.load_state_dict(ref_m.state_dict())
        xx = _rand(21, i, seed=2)
        assert_close(m.to(DEV)(xx), ref_m(xx.cpu()), 1e-5, f"tower {i}->{h}->{o}")

    gate = torch.sigmoid(_rand(64, 4, seed=3))
    eo = _rand(64, 4 * 5, seed=4)
    out = torch.empty(64, 5, device=DEV)
    idx = torch.empty(64, 2, dtype=torch.int32, device=DEV)
    ops.call("acb_moe_combine", gate, eo, out, idx, 64, 4, 5)
    tw, ti = torch.topk(gate, 2, -1)
    assert torch.equal(idx.long(), ti), "top-2 indices must match torch.topk exactly"
    ref = torch.zeros(64, 5, device=DEV)
    for e in range(4):
        sel = (ti == e)
        ref += (tw * sel).sum(-1, keepdim=True) * eo.view(64, 4, 5)[:, e] * sel.any(-1, keepdim=True)
    assert_close(out, ref, 1e-6, "moe_combine")


@pytest.mark.parametrize("lens", [[1, 5, 16, 17, 33, 64, 100], [128, 129, 200, 257, 258], [300, 470, 3], [96, 97, 112, 192, 193, 288], [700, 31]])
def test_attention_tc_matches_reference(lens):
    """tcgen05 attention (bf16) vs fp32 torch on the same bf16-rounded q/k/v, incl. > 128 queries and > 256 keys."""
    ops = _ops()
    B, H, D = len(lens), 8, 128
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=DEV)
    T = int(cu[-1])
    qkv = _bf(_rand(T, 3 * D, seed=sum(lens)))
    got = ops.attention_varlen(qkv, cu, B, H, 16, max(lens))
    ref = torch.empty(T, D, device=DEV)
    f = qkv.float()
    for bi in range(B):
        s, e = int(cu[bi]), int(cu[bi + 1])
        q, k, v = [z.view(e - s, H, 16).transpose(0, 1) for z in f[s:e].split(D, 1)]
        p = torch.softmax(q @ k.transpose(1, 2) / 4.0, -1)
        ref[s:e] = (p @ v).transpose(0, 1).reshape(e - s, D)
    assert_close(got, ref, 1.2e-2, f"tc attention lens={lens}")
    ops.USE_TC_ATTENTION = False
    try:
        old = ops.attention_varlen(qkv, cu, B, H, 16, max(lens))
    finally:
        ops.USE_TC_ATTENTION = True
    assert_close(got, old.float(), 1.2e-2, "tc vs CUDA-core attention")


def test_gemm_bf16_pre_out_and_gelu_grad_epilogue():
    """Training epilogues of the tcgen05 GEMM: (a) pre_out = acc + bias next to the activated output, (b) ACB_RES_MUL_GELU_GRAD:
    out = (A W^T) * gelu'(res).  Reference: torch fp32 on the bf16-rounded operands; tolerance = bf16 rounding of the result."""
    from applecider_b200 import ops

    torch.manual_seed(0)
    M, K, N = 1000, 96, 384
    a = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device=DEV)
    u = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    h = ops.gemm(a, w, b, act=ops.ACT_GELU, pre_out=u)
    ref_u = a.float() @ w.float().t() + b
    assert_close(u, ref_u, 1e-2, "pre_out (acc + bias)")
    assert_close(h, torch.nn.functional.gelu(ref_u, approximate="tanh"), 1.5e-2, "gelu(acc + bias)")
    # residual + layer scale with the un-scaled value as second output
    g = torch.rand(N, device=DEV) + 0.5
    r = torch.randn(M, N, device=DEV).to(torch.bfloat16)
    v = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    o = ops.gemm(a, w, b, res=r, gamma=g, res_mode=ops.RES_ADD, pre_out=v)
    assert_close(v, ref_u, 1e-2, "pre_out with residual epilogue")
    assert_close(o, r.float() + g * ref_u, 1.5e-2, "res + gamma * v")
    # gradient GEMM fused with the GELU backward
    x = torch.randn(M, N, device=DEV).to(torch.bfloat16)  # plays d h
    wt = (torch.randn(K, N, device=DEV) * N ** -0.5).to(torch.bfloat16)  # [K_out, N]
    pre = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    du = ops.gemm(x, wt, None, res=pre, res_mode=ops.RES_MUL_GELU_GRAD)
    pf = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(pf, approximate="tanh").sum().backward()
    assert_close(du, (x.float() @ wt.float().t()) * pf.grad, 1.5e-2, "(A W^T) * gelu'(res)")


@pytest.mark.parametrize("M,K,N,bn", [(1024, 384, 128, 128), (640, 768, 256, 256), (520, 1536, 512, 256)])
def test_gemm_row_stats_and_ln_fused_downsample(M, K, N, bn):
    """(a) acb_gemm_bf16_stats: per-row (sum, sum of squares) of every N tile from the epilogue's fp32 accumulators;
    (b) acb_gemm_ln_bf16: maxpool4(gelu(LayerNorm(y)) Wd^T + b) with those statistics, the normalised activation staying in
    shared memory (SpectraNet block tail, spectranet.py:36-40).  Reference: torch fp32 on the bf16-rounded operands."""
    from applecider_b200 import ops

    torch.manual_seed(M + K)
    K0 = 192
    x = torch.randn(M, K0, device=DEV).to(torch.bfloat16)
    w0 = (torch.randn(K, K0, device=DEV) * K0 ** -0.5).to(torch.bfloat16)
    b0 = torch.randn(K, device=DEV) * 0.5
    parts = (K + bn - 1) // bn
    stats = torch.full((M, parts, 2), float("nan"), device=DEV)
    y = ops.gemm(x, w0, b0, bn=bn, row_stats=stats)
    y_ref = x.float() @ w0.float().t() + b0
    assert_close(y, y_ref, 1e-2, "producer GEMM output")
    assert_close(stats[:, :, 0].sum(1), y_ref.sum(1), 2e-3, "row sums from the epilogue")
    assert_close(stats[:, :, 1].sum(1), (y_ref * y_ref).sum(1), 2e-3, "row sums of squares from the epilogue")
    for j in range(parts):
        assert_close(stats[:, j, 0], y_ref[:, j * bn:(j + 1) * bn].sum(1), 2e-3, f"partial sums of N tile {j}")
    g, be = 1.0 + 0.1 * torch.randn(K, device=DEV), 0.1 * torch.randn(K, device=DEV)
    wd = (torch.randn(N, K, device=DEV) * K ** -0.5).to(torch.bfloat16)
    bd = torch.randn(N, device=DEV) * 0.1
    mean, var = y_ref.mean(1, keepdim=True), y_ref.var(1, unbiased=False, keepdim=True)
    yn = torch.nn.functional.gelu((y.float() - mean) * torch.rsqrt(var + 1e-5) * g + be).to(torch.bfloat16).float()
    full = yn @ wd.float().t() + bd
    for pool in (False, True):
        got = ops.gemm_ln(y, wd, bd, stats, parts, g, be, 1e-5, pool4=pool)
        ref = full.view(M // 4, 4, N).amax(1) if pool else full
        assert_close(got, ref, 2e-2, f"gelu(LN(y)) Wd^T (pool4={pool})")


@pytest.mark.parametrize("C,M", [(96, 1000), (96, 128 * 37 + 5), (192, 640), (192, 49 * 64)])
def test_convnext_mlp_fused(C, M):
    """acb_convnext_mlp_bf16 (fc1 + GELU + fc2 + layer scale + residual, hidden activation on chip) vs torch fp32 on the
    bf16-rounded operands and vs the two-GEMM path it replaces."""
    from applecider_b200 import ops

    torch.manual_seed(C + M)
    y = torch.randn(M, C, device=DEV).to(torch.bfloat16)
    res = torch.randn(M, C, device=DEV).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, device=DEV) * C ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device=DEV) * (4 * C) ** -0.5).to(torch.bfloat16)
    b1, b2 = torch.randn(4 * C, device=DEV) * 0.3, torch.randn(C, device=DEV) * 0.3
    gamma = 0.2 + 0.05 * torch.randn(C, device=DEV)
    got = ops.convnext_mlp(y, res, w1, b1, w2, b2, gamma)
    hid = torch.nn.functional.gelu(y.float() @ w1.float().t() + b1).to(torch.bfloat16).float()
    ref = res.float() + gamma * (hid @ w2.float().t() + b2)
    assert_close(got, ref, 1.5e-2, f"fused ConvNeXt MLP C={C} M={M}")
    two = ops.gemm(ops.gemm(y, w1, b1, act=ops.ACT_GELU), w2, b2, res=res, gamma=gamma, res_mode=ops.RES_ADD)
    assert_close(got, two.float(), 1.5e-2, "fused vs two-GEMM path")


@pytest.mark.parametrize("T,valid", [(1000, None), (128 * 9 + 17, 700)])
def test_ffn_relu_fused(T, valid):
    """acb_ffn_relu_bf16 (transformer feed-forward block in one kernel) vs the two-GEMM path; with a device row count the tiles
    past it are skipped (their rows stay untouched)."""
    from applecider_b200 import ops

    torch.manual_seed(T)
    C = 128
    x = torch.randn(T, C, device=DEV).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, device=DEV) * C ** -0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device=DEV) * (4 * C) ** -0.5).to(torch.bfloat16)
    b1, b2 = torch.randn(4 * C, device=DEV) * 0.3, torch.randn(C, device=DEV) * 0.3
    nv = torch.tensor([valid], dtype=torch.int32, device=DEV) if valid is not None else None
    got = ops.ffn_relu(x, w1, b1, w2, b2, nv.data_ptr() if nv is not None else None)
    hid = torch.relu(x.float() @ w1.float().t() + b1).to(torch.bfloat16).float()
    ref = x.float() + hid @ w2.float().t() + b2
    n = T if valid is None else valid
    assert_close(got[:n], ref[:n], 1.5e-2, "fused transformer FFN")
    two = ops.gemm(ops.gemm(x, w1, b1, act=ops.ACT_RELU), w2, b2, res=x, res_mode=ops.RES_ADD)
    assert_close(got[:n], two[:n].float(), 1.5e-2, "fused vs two-GEMM path")
