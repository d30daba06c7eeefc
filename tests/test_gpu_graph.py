"""GPU: the whole-step CUDA graph (applecider_b200.graph.GraphedTrainStep) against the eager step it captures."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _setup(dropout, dtype="fp32", seed=0):
    import applecider_b200 as ab
    from applecider_b200 import fn, synth
    from applecider_b200.optim import FusedAdam

    cfg = ab.default_config()
    cfg["model"]["HyraxBaselineCLS"]["dropout"] = dropout
    cfg["model"]["HyraxBaselineCLS"]["compute_dtype"] = dtype
    m = ab.HyraxBaselineCLS(cfg)
    m.load_state_dict(synth.det_state_dict(m, seed))
    m = m.to(DEV).train()
    opt = FusedAdam([p for n, p in m.named_parameters() if not n.startswith("head.")], lr=1e-3, max_grad_norm=1.0, bf16_shadow=(dtype == "bf16"))
    x, pad, lens = synth.photometry_batch(24, seed=501, L=64)
    labels = synth.labels(24, seed=501)
    inputs = {"x": x.to(DEV), "pad": pad.to(DEV), "y": labels.to(DEV)}
    ntok = int(lens.clamp(max=64).sum()) + 24

    def fwd_loss(d):
        return fn.focal_loss(m((d["x"], d["pad"], None), total_tokens=ntok), d["y"])

    return m, opt, inputs, fwd_loss


def test_graph_replay_equals_eager_steps():
    """Dropout off: warm-up (2 eager steps; the capture itself executes nothing) + 3 replays == 5 eager steps; Adam's bias
    correction must follow the device step counter, and a new batch copied into the static buffers must be the one the replay
    trains on."""
    from applecider_b200.ddp import ddp_train_step
    from applecider_b200.graph import GraphedTrainStep

    m1, o1, inp, f1 = _setup(0.0)
    m2, o2, _, f2 = _setup(0.0)
    inp_b = {k: (v.flip(0).contiguous() if k != "pad" else v.flip(0).contiguous()) for k, v in inp.items()}
    seq = [inp, inp, inp_b, inp, inp_b]
    for d in seq:
        ddp_train_step(o1.grads, lambda d=d: f1(d), o1)
    g = GraphedTrainStep(o2.grads, f2, o2, inp, warmup=2)  # consumes seq[0:2]
    losses = [g(d).item() for d in seq[2:]]
    g.close()
    assert o2.step_count == o1.step_count == 5
    assert all(l == l for l in losses)
    for (n, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        # 5 Adam steps of lr 1e-3 move a weight by <= 5e-3; Adam divides by |g|, so the fp32-atomic ordering noise of the gradient
        # kernels shows up at the 1e-5 level -- a wrong bias correction or a stale batch would show at 1e-3
        assert torch.allclose(a, b, rtol=0, atol=1e-4), f"{n}: graph replay diverged from eager ({(a - b).abs().max().item():.3e})"


def test_graph_dropout_changes_every_replay_and_trains():
    """Dropout on, identical batch every replay: the device epoch must give each replay its own masks (different losses), and
    the backward must regenerate the forward's masks (the loss goes down over a few dozen steps)."""
    from applecider_b200.graph import GraphedTrainStep

    m, o, inp, f = _setup(0.4, dtype="bf16")
    g = GraphedTrainStep(o.grads, f, o, inp, warmup=2)
    losses = [g().item() for _ in range(40)]
    g.close()
    assert len({round(l, 6) for l in losses[:6]}) >= 5, f"replays repeat the same dropout masks: {losses[:6]}"
    assert sum(losses[-8:]) / 8 < sum(losses[:8]) / 8, f"not training: {losses[:8]} -> {losses[-8:]}"
