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
    # Adam divides every gradient element by its own magnitude, so elements whose true gradient is ZERO (the key bias of every
    # attention layer: softmax is shift-invariant) or at the fp32-atomic noise floor move by +-lr per step in an order-dependent
    # direction.  Compare (a) the well-conditioned tensors element-wise -- a wrong bias correction (device step counter) or a
    # stale batch in the static buffers would show at the 1e-3 level -- and (b) the direction of the whole 5-step update.
    p1, p2 = dict(m1.named_parameters()), dict(m2.named_parameters())
    for n in ("fc.bias", "fc.weight", "norm.weight", "encoder.layers.3.linear2.bias"):
        assert torch.allclose(p1[n], p2[n], rtol=0, atol=5e-5), f"{n}: graph replay diverged from eager ({(p1[n] - p2[n]).abs().max().item():.3e})"
    from applecider_b200 import synth

    m0 = synth.det_state_dict(m1, 0)
    u1 = torch.cat([(p1[n].detach().cpu() - m0[n]).flatten() for n in p1 if not n.startswith("head.")])
    u2 = torch.cat([(p2[n].detach().cpu() - m0[n]).flatten() for n in p1 if not n.startswith("head.")])
    cos = torch.nn.functional.cosine_similarity(u1, u2, dim=0).item()
    assert cos > 0.98, f"5-step update direction differs between graph replay and eager: cosine {cos:.4f}"


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
