"""Data-parallel training: identical replicas, one process per GPU, ONE collective per step — an all-reduce
(average) of a persistent flat gradient buffer over NCCL/NVLink (gloo on CPU for the host-logic tests).

Every parameter's ``.grad`` is a view into the flat buffer, so autograd accumulates straight into it: no
pack/unpack copies, one memset + one collective per step.  The reference has no distributed code of its own
(data parallelism comes from Hyrax/ignite); this is the B200-native equivalent for §8(e).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradSync:
    def __init__(self, module, process_group=None, offsets=None, numel=None):
        """module: an nn.Module or a list of parameters; offsets/numel: an explicit (aligned) flat layout."""
        self.module = module
        self.group = process_group
        plist = module.parameters() if isinstance(module, torch.nn.Module) else module
        self.params = [p for p in plist if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if offsets is None:
            offsets, off = [], 0
            for p in self.params:
                offsets.append(off)
                off += p.numel()
            numel = off
        self.offsets = list(offsets)
        self.flat = torch.zeros(numel, dtype=dt, device=dev)
        self.numel = numel
        self._rebind()
        self._buckets = None  # see enable_overlap()
        self._hooks = []
        self._accumulating = False
        self.allreduce_bytes = 0  # payload of the last sync() (diagnostics for the bench)
        self.time_sync = False    # bench: bracket the exposed wait of sync() with CUDA events (see exposed_ms())
        self._sync_events = []

    def broadcast_parameters(self, src: int = 0, flat_params=None) -> None:
        """Make every replica start from rank `src`'s weights (DDP does this at construction).  flat_params: the flat
        parameter buffer when the parameters are views of one (optim.FusedAdam) -- one collective instead of one per tensor."""
        if self.world_size == 1:
            return
        with torch.no_grad():
            if flat_params is not None:
                dist.broadcast(flat_params, src=src, group=self.group)
            else:
                for p in self.params:
                    dist.broadcast(p.data, src=src, group=self.group)

    def no_sync(self):
        """Context manager for gradient accumulation: backward passes inside it only accumulate into the flat buffer (the
        bucket hooks stay disarmed); the all-reduce happens in the first sync() after the last micro-batch's backward, which
        must run OUTSIDE the context -- call zero() once before the first micro-batch."""
        sync = self

        class _NoSync:
            def __enter__(self_):
                sync._accumulating = True
                sync._armed = False

            def __exit__(self_, *exc):
                sync._accumulating = False
                if sync._buckets is not None:  # re-arm for the final micro-batch; counters restart, buffers keep their sums
                    for bk in sync._buckets:
                        bk.update(seen=0, work=None, done=False)
                    sync._armed = True
                return False

        return _NoSync()

    # ---- bucketed all-reduce overlapped with the backward pass --------------------------------------------------
    def enable_overlap(self, bucket_elems: int = 16 * 1024 * 1024) -> "FlatGradSync":
        """Split the flat buffer into contiguous buckets of about `bucket_elems` gradients; a bucket's all-reduce starts from
        the post-accumulate hook of its last parameter, i.e. while autograd is still producing the other buckets
        (NCCL orders the collective after the gradient kernels enqueued so far).  sync() then only waits / finishes the
        buckets whose parameters received no gradient.  Buckets are contiguous slices: still no pack / unpack copies.
        Idempotent: a second call replaces the bucket plan and its hooks instead of stacking another set."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        bounds, start, acc = [], 0, 0
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            acc += p.numel()
            last = i == len(self.params) - 1
            if acc >= bucket_elems or last:
                end = self.numel if last else self.offsets[i + 1]
                bounds.append((start, end))
                start, acc = end, 0
        self._buckets = [dict(lo=lo, hi=hi, need=0, seen=0, work=None, done=False) for lo, hi in bounds]
        self._bucket_of = {}
        for p, off in zip(self.params, self.offsets):
            b = next(i for i, bk in enumerate(self._buckets) if bk["lo"] <= off < bk["hi"])
            self._bucket_of[id(p)] = b
            self._buckets[b]["need"] += 1
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        return self

    def _launch(self, bk) -> None:
        w = self.world_size
        if w == 1:
            bk["done"] = True
            return
        view = self.flat[bk["lo"]: bk["hi"]]
        if self.flat.is_cuda:
            bk["work"] = dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:  # gloo has no AVG
            bk["work"] = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        bk["done"] = True

    def _on_grad(self, p) -> None:
        if self._buckets is None or not self._armed or self._accumulating:
            return
        bk = self._buckets[self._bucket_of[id(p)]]
        bk["seen"] += 1
        if bk["seen"] == bk["need"] and not bk["done"]:
            self._launch(bk)

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    _armed = False

    def zero(self) -> None:
        """Replaces optimizer.zero_grad(): keeps the .grad views alive."""
        self.flat.zero_()
        if self._buckets is not None:
            for bk in self._buckets:
                bk.update(seen=0, work=None, done=False)
            self._armed = True
        for p in self.params:  # an optimizer.zero_grad(set_to_none=True) elsewhere would have dropped the views
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * self.flat.element_size():
                self._rebind()
                break

    def _rebind(self) -> None:
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off: off + p.numel()].view_as(p)

    def sync(self) -> None:
        """Average gradients over all ranks (the single exchange step of the data-parallel path)."""
        w = self.world_size
        if self._accumulating:
            raise RuntimeError("FlatGradSync.sync() called inside no_sync(): run the last micro-batch outside the context")
        self.allreduce_bytes = self.flat.numel() * self.flat.element_size() if w > 1 else 0
        timed = self.time_sync and self.flat.is_cuda and w > 1
        if timed:  # e0 completes when the backward's last kernel does, e1 when the last collective has landed
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        try:
            self._sync_impl(w)
        finally:
            if timed:
                e1.record()
                self._sync_events.append((e0, e1))

    def exposed_ms(self):
        """Mean time per step the compute stream spent waiting for the all-reduce after the backward had finished (the part
        of the collective that the bucketed overlap did not hide).  Synchronises; clears the samples."""
        if not self._sync_events:
            return 0.0
        torch.cuda.synchronize()
        v = [a.elapsed_time(b) for a, b in self._sync_events]
        self._sync_events = []
        return sum(v) / len(v)

    def _sync_impl(self, w) -> None:
        if self._buckets is not None:
            self._armed = False
            for bk in self._buckets:  # buckets with parameters that got no gradient this step were never launched
                if not bk["done"]:
                    self._launch(bk)
            for bk in self._buckets:
                if bk["work"] is not None:
                    bk["work"].wait()
                    if not self.flat.is_cuda:
                        self.flat[bk["lo"]: bk["hi"]].div_(w)
            return
        if w == 1:
            return
        if self.flat.is_cuda:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:  # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(w)


def ddp_train_step(sync: FlatGradSync, forward_loss, optimizer, clip_norm=None):
    """zero -> forward+loss -> backward -> all-reduce(avg) -> [clip] -> optimizer.step.  Returns the local loss tensor.

    With an optim.FusedAdam (whose ``grads`` is the FlatGradSync) clip + update are two kernels and nothing syncs."""
    sync.zero()
    loss = forward_loss()
    loss.backward()
    sync.sync()
    if hasattr(optimizer, "flat_p"):  # optim.FusedAdam: the clip is part of the fused step
        if clip_norm is not None and optimizer.max_grad_norm != clip_norm:
            optimizer.max_grad_norm = clip_norm
        optimizer.step()
        return loss.detach()
    if clip_norm is not None:
        torch.nn.utils.clip_grad_norm_(sync.params, max_norm=clip_norm)
    optimizer.step()
    return loss.detach()


def shard_batch(n: int, rank: int, world: int):
    """Contiguous alert shard [lo, hi) of rank `rank` (inference: independent shards, no collective)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
