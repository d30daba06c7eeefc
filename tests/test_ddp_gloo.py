"""CPU suite: world_size-2 gloo test of the data-parallel host logic (flat gradient views + one averaged
all-reduce): a 2-rank step on half batches equals a 1-rank step on the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(7, 16), torch.nn.Tanh(), torch.nn.Linear(16, 5))


def _worker(rank, world, port, q, overlap=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from applecider_b200.ddp import FlatGradSync, ddp_train_step, shard_batch

    torch.manual_seed(1)
    x, y = torch.randn(12, 7), torch.randint(0, 5, (12,))
    model = _model()
    sync = FlatGradSync(model)
    if overlap:
        sync.enable_overlap(bucket_elems=64)  # 4 parameters -> several buckets, launched from the autograd hooks
        assert len(sync._buckets) >= 2
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    lo, hi = shard_batch(12, rank, world)
    for _ in range(3):
        ddp_train_step(sync, lambda: torch.nn.functional.cross_entropy(model(x[lo:hi]), y[lo:hi]), opt, clip_norm=1.0)
    assert all(p.grad.data_ptr() >= sync.flat.data_ptr() for p in sync.params), "grads must stay views of the flat buffer"
    q.put((rank, [p.detach().numpy().copy() for p in model.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
def test_two_rank_step_equals_single_rank_full_batch(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the full batch (mean loss over 12 == average of the two half-batch means)
    torch.manual_seed(1)
    x, y = torch.randn(12, 7), torch.randint(0, 5, (12,))
    model = _model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), y).backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
    import numpy as np

    for r in (0, 1):
        for a, b in zip(results[r], model.parameters()):
            assert np.allclose(a, b.detach().numpy(), atol=1e-6), f"rank {r} diverged from the single-rank step"
    for a, b in zip(results[0], results[1]):
        assert np.array_equal(a, b), "replicas must stay bit-identical"


def _worker_accum(rank, world, port, q):
    """Gradient accumulation (two micro-batches, all-reduce only after the second), replicas that start from DIFFERENT
    weights (broadcast_parameters must fix that), enable_overlap called twice (must not double-count)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from applecider_b200.ddp import FlatGradSync, shard_batch

    torch.manual_seed(1)
    x, y = torch.randn(16, 7), torch.randint(0, 5, (16,))
    torch.manual_seed(100 + rank)  # different init per rank
    model = torch.nn.Sequential(torch.nn.Linear(7, 16), torch.nn.Tanh(), torch.nn.Linear(16, 5))
    sync = FlatGradSync(model)
    sync.broadcast_parameters(0)
    sync.enable_overlap(bucket_elems=64)
    sync.enable_overlap(bucket_elems=64)
    assert sum(len(p._post_accumulate_grad_hooks or {}) for p in sync.params) == len(sync.params)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    lo, hi = shard_batch(16, rank, world)
    mid = (lo + hi) // 2
    loss = lambda a, b: torch.nn.functional.cross_entropy(model(x[a:b]), y[a:b], reduction="sum") / 16.0 * world  # noqa: E731
    for _ in range(2):
        sync.zero()
        with sync.no_sync():
            loss(lo, mid).backward()
        loss(mid, hi).backward()
        sync.sync()
        opt.step()
    q.put((rank, [p.detach().numpy().copy() for p in model.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_accumulation_broadcast_and_idempotent_overlap():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_accum, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(1)
    x, y = torch.randn(16, 7), torch.randint(0, 5, (16,))
    torch.manual_seed(100)  # rank 0's init is what every replica must have used
    model = torch.nn.Sequential(torch.nn.Linear(7, 16), torch.nn.Tanh(), torch.nn.Linear(16, 5))
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    for _ in range(2):
        opt.zero_grad()
        torch.nn.functional.cross_entropy(model(x), y).backward()
        opt.step()
    import numpy as np

    for r in (0, 1):
        for a, b in zip(results[r], model.parameters()):
            assert np.allclose(a, b.detach().numpy(), atol=1e-6), f"rank {r}: accumulated 2-rank step != full-batch step"
    for a, b in zip(results[0], results[1]):
        assert np.array_equal(a, b)


def test_shard_batch_covers_everything():
    from applecider_b200.ddp import shard_batch

    for n in (1, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            spans = [shard_batch(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
