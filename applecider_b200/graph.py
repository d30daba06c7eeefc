"""Whole-step CUDA graph for training: zero -> forward + loss -> backward -> gradient all-reduce -> clip/Adam, replayed as
ONE graph launch.

Why: a training step of the fusion model is ~800 kernel launches (AstroMiNN alone ~470), each issued through Python ->
autograd.Function -> ctypes; on a B200 the host needs about as long to enqueue them (37 ms measured) as the GPU needs to run
them, so the step is launch-bound.  The reference's loop (brew_cider.py:986-1014; Hyrax/ignite for the src models) has the
same structure -- one Python step per batch -- and no answer to it.

What makes the step capturable (all of it is in this package, nothing relies on torch.compile):
* no host read anywhere in the step (token packing works on a row capacity, photo.pack);
* every tensor of the step comes from torch's caching allocator, so the capture owns a private pool and addresses are
  stable across replays; TMA descriptors and argument structs are kernel parameters, baked at capture;
* the stochastic kernels take their seed as (host value + device epoch): the graph increments the epoch
  (acb_set_seed_epoch_ptr), so every replay draws new dropout masks and its backward regenerates the same ones;
* the optimizer's bias correction reads a device step counter the graph increments (FusedAdam.step_dev);
* derived weights (bf16 shadow, tap-major conv weights ...) are rebuilt by kernels inside the graph, because the capture
  follows a real optimizer step (parameter versions changed), exactly as in eager mode;
* the NCCL all-reduce buckets launched from autograd hooks are captured as forked streams and joined by sync().

Inputs are copied into static device buffers before each replay; the loss is a static device scalar.
"""
from __future__ import annotations

import torch

from . import _lib
from .ddp import FlatGradSync


class GraphedTrainStep:
    def __init__(self, sync: FlatGradSync, forward_loss, optimizer, example_inputs: dict, clip_norm=None, warmup: int = 2):
        """forward_loss(inputs: dict of CUDA tensors) -> scalar loss tensor; optimizer: optim.FusedAdam (its `grads` is `sync`).

        Runs `warmup` eager steps on a side stream (they are REAL training steps on the example inputs), then captures the
        step (capturing executes nothing)."""
        if not hasattr(optimizer, "flat_p"):
            raise TypeError("GraphedTrainStep needs optim.FusedAdam (torch optimizers read their step count on the host)")
        self.sync, self.optimizer, self.forward_loss = sync, optimizer, forward_loss
        if clip_norm is not None:
            optimizer.max_grad_norm = clip_norm
        dev = optimizer.flat_p.device
        self.static = {k: v.clone() for k, v in example_inputs.items()}
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        optimizer.step_dev = torch.full((1,), optimizer.step_count, dtype=torch.int32, device=dev)
        _lib.set_seed_epoch(self.epoch)
        self._timed = sync.time_sync
        sync.time_sync = False  # CUDA events with timing cannot be recorded inside a capture
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.loss = self._one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        count0 = optimizer.step_count
        with torch.cuda.graph(self.graph):
            self.loss = self._one_step()
        optimizer.step_count = count0  # a capture records the step, it does not run it
        self.launches_per_replay = _lib.launch_count() - n0
        self.replays = 0

    def _one_step(self):
        opt = self.optimizer
        self.epoch.add_(1)
        opt.step_dev.add_(1)
        self.sync.zero()
        loss = self.forward_loss(self.static)
        loss.backward()
        self.sync.sync()
        opt.step()
        return loss.detach()

    def __call__(self, inputs: dict | None = None):
        """Copies `inputs` (same shapes/dtypes as the example) into the static buffers and replays the step.  Returns the
        static loss tensor (overwritten by the next replay; `.item()` or clone it to keep a value)."""
        if inputs is not None:
            for k, v in inputs.items():
                if v is not self.static[k]:
                    self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        self.optimizer.step_count += 1
        torch.autograd.graph.increment_version(self.optimizer.params)  # derived-weight caches outside the graph key on it
        return self.loss

    def close(self):
        """Back to eager stepping: clears the device seed epoch and the device step counter."""
        if _lib._seed_epoch_tensor is self.epoch:
            _lib.set_seed_epoch(None)
        self.optimizer.step_dev = None
        self.sync.time_sync = self._timed

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
