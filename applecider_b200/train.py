"""Training path (backward kernels + optimizer glue).  Filled in incrementally; see DESIGN.md."""
from __future__ import annotations


def _todo(*a, **k):
    raise NotImplementedError("applecider_b200: the training path of this module is not implemented yet")


focal_loss = photo_forward_train = photo_train_step = _todo
spectra_forward_train = spectra_train_step = _todo
astrominn_forward_train = astrominn_train_step = _todo
fusion_forward_train = _todo
