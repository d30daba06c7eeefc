"""GPU: packed multi-sequence tcgen05 attention (csrc/attention_packed.cu) vs fp32 torch on the same bf16-rounded q/k/v.

Covers: many short sequences per 128-row tile, sequences that exactly fill / just overflow a tile, 1-token sequences, long
sequences (> 128 tokens: per-(sequence, head) kernel over the plan's list) mixed in, capacity rows past the last sequence
(uninitialised, possibly NaN), the plan itself (bit-exact vs a Python greedy packing), and training-time dropout bit-for-bit
against the CUDA-core kernel (same counter hash)."""
import numpy as np
import pytest
import torch

from util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
H, D = 8, 128


def _ref(qkv, cu):
    T = qkv.shape[0]
    ref = torch.zeros(T, D, device=DEV)
    f = qkv.float()
    for bi in range(cu.numel() - 1):
        s, e = int(cu[bi]), int(cu[bi + 1])
        if e <= s:
            continue
        q, k, v = [z.view(e - s, H, 16).transpose(0, 1) for z in f[s:e].split(D, 1)]
        p = torch.softmax(q @ k.transpose(1, 2) / 4.0, -1)
        ref[s:e] = (p @ v).transpose(0, 1).reshape(e - s, D)
    return ref


def _greedy(lens):
    """A tile = maximal run of consecutive sequences with <= 128 rows in total, each <= 128 rows, at most 128 sequences."""
    tiles, longs, first, cnt, rows = [], [], 0, 0, 0
    for b, n in enumerate(lens):
        if (n > 128 or rows + n > 128 or cnt >= 128) and cnt:
            tiles.append((first, cnt))
            cnt = rows = 0
        if n > 128:
            longs.append(b)
        else:
            if cnt == 0:
                first = b
            cnt += 1
            rows += n
    if cnt:
        tiles.append((first, cnt))
    return tiles, longs


LENS = [
    [2, 5, 16, 17, 33, 64, 100, 2, 2, 3],
    [128, 128, 127, 1, 129, 64, 64, 1, 127, 2],
    [258, 40, 41, 200, 3, 130, 128, 58, 58, 58, 58, 12],
    [1] * 300,
    list(np.random.default_rng(0).integers(2, 259, size=200)),
    list(np.clip(np.round(np.random.default_rng(1).lognormal(np.log(40), 0.9, size=1500)), 1, 257).astype(int) + 1),
]


@pytest.mark.parametrize("case", range(len(LENS)))
@pytest.mark.parametrize("slack", [0, 777])
def test_packed_attention_matches_reference(case, slack):
    from applecider_b200 import ops

    lens = [int(x) for x in LENS[case]]
    B = len(lens)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    T = int(cu[-1])
    cap = T + slack  # capacity rows past the packed tokens hold garbage (incl. NaN): they must not leak into real rows
    torch.manual_seed(case)
    qkv = torch.full((cap, 3 * D), float("nan"), device=DEV, dtype=torch.bfloat16)
    qkv[:T] = torch.randn(T, 3 * D, device=DEV).to(torch.bfloat16)
    plan, max_tiles = ops.attention_plan(cu, B, cap)
    tiles, longs = _greedy(lens)
    p = plan.cpu().numpy()
    assert p[0] == len(tiles) and p[1] == len(longs), (p[:2], len(tiles), len(longs))
    assert [(int(p[2 + 2 * t]), int(p[3 + 2 * t])) for t in range(len(tiles))] == tiles
    assert list(p[2 + 2 * max_tiles: 2 + 2 * max_tiles + len(longs)]) == longs
    got = ops.attention_varlen(qkv, cu, B, H, 16, max(lens), plan=(plan, max_tiles), zero_tail=True)
    ref = _ref(qkv[:T], cu)
    assert_close(got[:T], ref, 1.2e-2, f"packed attention case {case}")
    assert (got[T:] == 0).all(), "capacity rows past the last sequence must stay zero"
    old = ops.attention_varlen(qkv, cu, B, H, 16, max(lens), zero_tail=True)  # per-(sequence, head) kernel
    assert_close(got[:T], old[:T].float(), 1.2e-2, "packed vs per-sequence tcgen05 kernel")


def test_packed_attention_dropout_matches_cuda_core_kernel():
    """Same counter hash (seed, sequence, head, i, j) as the fp32 CUDA-core kernel: with dropout on, both draw the same mask."""
    from applecider_b200 import ops

    lens = [30, 70, 5, 128, 200, 17, 99]
    B = len(lens)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    T = int(cu[-1])
    torch.manual_seed(3)
    qkv = torch.randn(T, 3 * D, device=DEV).to(torch.bfloat16)
    plan = ops.attention_plan(cu, B, T)
    got = ops.attention_varlen(qkv, cu, B, H, 16, max(lens), drop_p=0.4, seed=12345, plan=plan)
    ops.USE_TC_ATTENTION = False
    try:
        ref = ops.attention_varlen(qkv.float(), cu, B, H, 16, max(lens), drop_p=0.4, seed=12345)
    finally:
        ops.USE_TC_ATTENTION = True
    assert_close(got, ref, 2e-2, "packed attention with dropout vs CUDA-core fp32 kernel (same masks)")


@pytest.mark.parametrize("case", [0, 1, 2, 5])
@pytest.mark.parametrize("drop_p", [0.0, 0.4])
def test_packed_attention_backward_matches_cuda_core(case, drop_p):
    """tcgen05 backward (dQ, dK, dV from recomputed S / dP in TMEM) vs the fp32 CUDA-core backward on the same bf16 inputs, same
    dropout hash; long sequences (> 128) go through the list mode of the CUDA-core kernel inside the same call."""
    from applecider_b200 import ops

    lens = [int(x) for x in LENS[case]]
    B = len(lens)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    T = int(cu[-1])
    cap = T + 300
    torch.manual_seed(case + 10)
    qkv = torch.zeros(cap, 3 * D, device=DEV, dtype=torch.bfloat16)
    qkv[:T] = torch.randn(T, 3 * D, device=DEV).to(torch.bfloat16)
    dout = torch.zeros(cap, D, device=DEV, dtype=torch.bfloat16)
    dout[:T] = torch.randn(T, D, device=DEV).to(torch.bfloat16)
    plan, max_tiles = ops.attention_plan(cu, B, cap)
    got = torch.zeros_like(qkv)
    ops.call("acb_attention_packed_bwd", qkv, dout, cu, plan, B, max_tiles, cap, H, 16, max(lens), drop_p, 777, got)
    ref = torch.zeros(cap, 3 * D, device=DEV)
    ops.call("acb_attention_varlen_bwd", qkv.float(), 0, dout.float(), 0, cu, B, H, 16, max(lens), drop_p, 777, ref, 0)
    for name, sl in (("dQ", slice(0, D)), ("dK", slice(D, 2 * D)), ("dV", slice(2 * D, 3 * D))):
        assert_close(got[:T, sl], ref[:T, sl], 2e-2, f"{name} case {case} p={drop_p}")
    assert (got[T:] == 0).all()


def test_attention_autograd_uses_packed_backward():
    """fn.attention with a plan: forward and backward both on the packed tcgen05 kernels; gradient vs fp32 torch autograd."""
    from applecider_b200 import fn, ops

    lens = [40, 90, 17, 128, 3, 150, 64]
    B = len(lens)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    T = int(cu[-1])
    torch.manual_seed(4)
    qkv = torch.randn(T, 3 * D, device=DEV).to(torch.bfloat16).requires_grad_(True)
    w = torch.randn(T, D, device=DEV)
    plan = ops.attention_plan(cu, B, T)
    out = fn.attention(qkv, cu, B, H, 16, max(lens), 0.0, 0, plan)
    (out.float() * w).sum().backward()
    q32 = qkv.detach().float().requires_grad_(True)
    ref = torch.zeros(T, D, device=DEV)
    parts = []
    for bi in range(B):
        s, e = int(cu[bi]), int(cu[bi + 1])
        q, k, v = [z.view(e - s, H, 16).transpose(0, 1) for z in q32[s:e].split(D, 1)]
        p = torch.softmax(q @ k.transpose(1, 2) / 4.0, -1)
        parts.append((p @ v).transpose(0, 1).reshape(e - s, D))
    (torch.cat(parts) * w.to(torch.bfloat16).float()).sum().backward()
    assert_close(qkv.grad, q32.grad, 3e-2, "d qkv through the packed attention")


def test_plan_serial_fallback_for_huge_batches():
    """More than 8192 sequences: the planner's serial kernel must produce the same greedy packing, and the attention must match."""
    from applecider_b200 import ops

    rng = np.random.default_rng(9)
    lens = [int(x) for x in rng.integers(2, 40, size=9000)]
    lens[100], lens[5000] = 200, 131
    B = len(lens)
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32, device=DEV)
    T = int(cu[-1])
    plan, max_tiles = ops.attention_plan(cu, B, T)
    tiles, longs = _greedy(lens)
    p = plan.cpu().numpy()
    assert p[0] == len(tiles) and p[1] == len(longs)
    assert [(int(p[2 + 2 * t]), int(p[3 + 2 * t])) for t in range(len(tiles))] == tiles
    assert list(p[2 + 2 * max_tiles: 2 + 2 * max_tiles + len(longs)]) == longs
    torch.manual_seed(1)
    qkv = torch.randn(T, 3 * D, device=DEV).to(torch.bfloat16)
    got = ops.attention_varlen(qkv, cu, B, H, 16, max(lens), plan=(plan, max_tiles))
    old = ops.attention_varlen(qkv, cu, B, H, 16, max(lens))
    assert_close(got, old.float(), 1.2e-2, "packed attention over 9000 sequences vs the per-sequence kernel")
