"""CPU suite: the oracle's restated train_steps (oracle/train_steps.py) and its seeded random init against the fixtures
recorded from the UNMODIFIED reference (tests/golden/make_golden_train_steps.py)."""
import os

import numpy as np
import torch

from train_step_util import check_two_steps, load_steps


def test_oracle_astrominn_train_steps_vs_reference(golden_dir):
    from applecider_b200 import synth
    from oracle import models as om
    from oracle.train_steps import AstroMiNNTrainer

    g = load_steps(golden_dir, "astrominn")
    m = om.AstroMiNN(om.default_config()).eval()
    m.load_state_dict(synth.det_state_dict(m, 0))
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    tr = AstroMiNNTrainer(m)
    batch = (g["meta"], g["img"], g["tgt"])
    l1 = tr.train_step(batch)["loss"]
    l2 = tr.train_step(batch)["loss"]
    check_two_steps(g, dict(m.named_parameters()), before, (l1, l2), loss_tol=1e-5, solid_tol=2e-3)


def test_oracle_spectranet_train_steps_vs_reference(golden_dir):
    from applecider_b200 import synth
    from oracle import models as om
    from oracle.train_steps import spectranet_train_step

    g = load_steps(golden_dir, "spectranet")
    m = om.SpectraNet(om.default_config()).eval()
    m.load_state_dict(synth.det_state_dict(m, 0))
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=0.01)
    crit = torch.nn.CrossEntropyLoss()
    batch = (g["x"], g["labels"], None)
    l1 = spectranet_train_step(m, opt, crit, batch)["loss"]
    l2 = spectranet_train_step(m, opt, crit, batch)["loss"]
    check_two_steps(g, dict(m.named_parameters()), before, (l1, l2), loss_tol=1e-5, solid_tol=2e-3)


def test_oracle_seeded_init_matches_reference_fingerprint(golden_dir):
    """The GPU reference-init test regenerates the reference's own random init by constructing the oracle under
    torch.manual_seed(0); the fixture holds the fingerprint of the REAL reference's init and its logits."""
    from oracle import models as om

    z = np.load(os.path.join(golden_dir, "refinit.npz"))
    x, pad = torch.from_numpy(z["in/x"]), torch.from_numpy(z["in/pad"])
    meta, img, sp = torch.from_numpy(z["in/meta"]), torch.from_numpy(z["in/img"]), torch.from_numpy(z["in/spec"])
    for name, batch in [("HyraxBaselineCLS", (x, pad, None)), ("SpectraNet", (sp, None, None)), ("AstroMiNN", (meta, img, None))]:
        torch.manual_seed(0)
        m = getattr(om, name)(om.default_config()).eval()
        fp = np.array([[float(v.double().sum()), float(v.double().abs().sum())] for v in m.state_dict().values() if v.dtype.is_floating_point])
        assert np.allclose(fp, z[f"fp/{name}"], rtol=1e-12, atol=1e-12), f"{name}: seeded init differs from the recorded reference init"
        with torch.no_grad():
            out = m(batch)
        ref = torch.from_numpy(z[f"logits/{name}"])
        assert (out - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), name


def test_oracle_legacy_variant_b_training_vs_reference(golden_dir):
    """Train-mode forward + backward of the variant-B spectra encoder (BatchNorm on batch statistics) vs the record of the REAL
    `build_spec_model` (tests/golden/make_golden_legacy.py::train_record)."""
    from applecider_b200 import synth
    from oracle import models as om

    g = np.load(os.path.join(golden_dir, "legacy_train.npz"))
    m = om.SpectraClassificationB({"mode": "spectra", "classes": list(range(5))}).train()
    m.load_state_dict(synth.det_state_dict(m, 0))
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    logits = m(torch.from_numpy(g["x"]))
    loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(g["y"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5
    sd = m.state_dict()
    for i in range(1, 5):
        assert np.allclose(sd[f"stage{i}.0.norm.running_mean"].numpy(), g[f"rm{i}"], rtol=1e-5, atol=1e-6)
        assert np.allclose(sd[f"stage{i}.0.norm.running_var"].numpy(), g[f"rv{i}"], rtol=1e-5, atol=1e-6)
    grads = {n: p.grad for n, p in m.named_parameters()}
    for k in g.files:
        if k.startswith("g_") and not k.endswith("_rows"):
            gr = grads[k[2:]]
            got = gr.reshape(gr.shape[0], -1)[:, :256].numpy() if gr.dim() > 1 else gr.numpy()
            assert np.abs(got - g[k]).max() <= 1e-4 * max(np.abs(g[k]).max(), 1e-8) + 1e-8, k
