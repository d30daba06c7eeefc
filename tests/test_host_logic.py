"""Host-side logic that needs no GPU: the registry behind `fn.transposed_weight` (which weight copies are rebuilt when, and that a
step costs ONE batched launch), the split-K heuristic of the CUDA-core gradient GEMM and the token-capacity rule of the varlen packing."""
import gc

import pytest
import torch


@pytest.fixture
def registry(monkeypatch):
    from applecider_b200 import fn

    log = {"single": 0, "batch": 0}

    def fake_transpose(x, dtype):
        log["single"] += 1
        return x.t().contiguous().to(dtype)

    def fake_call(name, src, dst, meta, n, tiles):
        assert name == "acb_transpose_batch"
        log["batch"] += 1
        log["jobs"], log["tiles"], log["meta"] = n, tiles, meta.view(-1, 4).tolist()

    monkeypatch.setattr(fn, "transpose", fake_transpose)
    monkeypatch.setattr(fn, "call", fake_call)
    return fn._TransposedWeights(), log


def test_transposed_weights_refresh_once_per_update(registry):
    reg, log = registry
    Ws = [torch.nn.Parameter(torch.randn(n, k)) for n, k in [(40, 33), (64, 128), (5, 7)]]
    for W in Ws:  # first use: one plain transpose each, nothing batched yet
        assert reg.get(W).shape == (W.shape[1], W.shape[0])
    assert (log["single"], log["batch"]) == (3, 0)
    for W in Ws:  # unchanged weights (gradient accumulation, eval): no launch at all
        reg.get(W)
    assert (log["single"], log["batch"]) == (3, 0)
    for step in range(1, 4):
        torch.autograd.graph.increment_version(Ws)  # what optim.FusedAdam does after its in-place kernel
        for W in Ws:
            reg.get(W)
        assert (log["single"], log["batch"]) == (3, step), "one batched launch per optimizer step"
    assert log["jobs"] == 3
    # job table: {R, C, tiles per row, first tile}; tiles are 32 x 32
    assert log["meta"] == [[40, 33, 2, 0], [64, 128, 4, 4], [5, 7, 1, 12]] and log["tiles"] == 13


def test_transposed_weights_views_temporaries_and_dead_models(registry):
    reg, log = registry
    flat = torch.nn.Parameter(torch.randn(64 * 96))
    reg.get(flat.view(64, 96))
    reg.get(flat.view(64, 96))  # a new view object of the same parameter: same entry
    assert (log["single"], log["batch"]) == (1, 0)
    with torch.no_grad():
        flat.mul_(2.0)  # in-place torch update bumps the shared version counter
    reg.get(flat.view(64, 96))
    assert (log["single"], log["batch"]) == (1, 1)
    # temporaries (concatenated weights) and non-fp32 inputs are never registered
    reg.get(torch.cat([torch.randn(4, 8), torch.randn(4, 8)]))
    reg.get(torch.randn(4, 8, dtype=torch.float64))
    assert (log["single"], log["batch"]) == (3, 1)
    # a parameter that is gone leaves the table at the next refresh
    tmp = torch.nn.Parameter(torch.randn(16, 16))
    reg.get(tmp)
    del tmp
    gc.collect()
    with torch.no_grad():
        flat.add_(1.0)
    reg.get(flat.view(64, 96))
    assert log["jobs"] == 1


def test_split_k_heuristic_and_token_capacity():
    from applecider_b200 import fn, ops

    assert fn._splits(100) == 1 and fn._splits(1 << 20) == 128  # long reductions split, capped
    assert fn._splits(4096, tiles=2) == 32 and fn._splits(4096, tiles=200) == 2  # few output tiles -> more splits
    B, L = 8, 257
    assert ops.token_capacity(B, L, None) == B * (L + 1)           # worst case: every event valid + CLS
    assert ops.token_capacity(B, L, 1000) == 1000                  # the collate's packed count
    for bad in (B - 1, B * (L + 1) + 1):                           # fewer than one CLS per object / more than the worst case
        with pytest.raises(ValueError):
            ops.token_capacity(B, L, bad)
