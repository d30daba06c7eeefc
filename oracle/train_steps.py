"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of the reference's ``train_step`` protocols on the
oracle modules of oracle/models.py.  Pinned by tests/test_oracle_train_steps.py against tests/golden/train_steps.npz,
which tests/golden/make_golden_train_steps.py recorded from the UNMODIFIED reference train_steps.

* AstroMiNN  -- models/astrominn.py:147-218 (CrossEntropyLoss, 11-group AdamW, base LR 1.6e-4) and :308-326
  (zero_grad -> forward -> loss -> running-mean bookkeeping -> backward -> step; returns the running mean).
* SpectraNet -- models/spectranet.py:172-184 (optimizer / criterion injected by the framework).
"""
from __future__ import annotations

import torch
import torch.nn as nn


def astrominn_optimizer(model, ac: dict) -> torch.optim.AdamW:
    """The eleven parameter groups in the reference's order (astrominn.py:151-218)."""
    LR = 1.6e-4

    def g(mod, wd, lr, **kw):
        return dict(params=mod.parameters(), weight_decay=ac[wd], lr=LR * ac[lr], **kw)

    groups = [
        g(model.image_tower, "cnn_decay", "cnn_lr"), g(model.psf_tower, "psf_decay", "psf_lr"), g(model.lc_tower, "lc_decay", "lc_lr"),
        g(model.mag_tower, "mag_decay", "mag_lr"), g(model.spatial_tower, "spatial_decay", "spatial_lr"),
        g(model.coord_tower, "nst1_decay", "nst1_lr"),  # sic: the coordinate tower uses the nst1 hyper-parameters (:181-185)
        g(model.nst1_tower, "nst1_decay", "nst1_lr"), g(model.nst2_tower, "nst2_decay", "nst2_lr"),
        g(model.mega_tower, "lc_decay", "lc_lr"),       # sic: the mega tower uses the lc hyper-parameters (:197-201)
        g(model.fusion_experts, "fusion_decay", "fusion_lr", betas=(ac["fusion_beta1"], ac["fusion_beta2"])),
        g(model.fusion_router, "router_decay", "router_lr", betas=(ac["router_beta1"], ac["router_beta2"])),
    ]
    return torch.optim.AdamW(groups, lr=LR, betas=(ac["beta1"], ac["beta2"]), eps=ac["eps"])


class AstroMiNNTrainer:
    def __init__(self, model):
        self.model = model
        self.optimizer = astrominn_optimizer(model, model.config["model"]["AstroMiNN"])
        self.criterion = nn.CrossEntropyLoss()
        self.total_loss = []

    def train_step(self, batch):
        _, _, labels = batch
        self.optimizer.zero_grad()
        loss = self.criterion(self.model(batch), labels)
        self.total_loss.append(loss.item())
        loss.backward()
        self.optimizer.step()
        return {"loss": sum(self.total_loss) / len(self.total_loss)}


def spectranet_train_step(model, optimizer, criterion, batch):
    _, labels, redshifts = batch
    optimizer.zero_grad()
    outputs = model(batch)
    loss = criterion(outputs, redshifts if getattr(model, "redshift", False) else labels)
    loss.backward()
    optimizer.step()
    return {"loss": loss.item()}
