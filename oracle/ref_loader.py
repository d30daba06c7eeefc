"""Import the REAL reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE — see oracle/__init__.py.  The reference is never copied:
this file only arranges ``sys.modules`` so that the reference's own files import
(recipe probed in SURVEY.md Appendix B):

* ``hyrax`` is an un-vendored dependency (reference pyproject.toml:18); its
  ``@hyrax_model`` decorator is replaced by the identity and ``HyraxDataset`` by
  an empty base class.
* ``astropy`` is only used for FITS/time I/O (preprocess_multimodal.py:433-456,
  563-570) which is outside the hot path; a stub lets the numeric functions import.
* ``timm`` (models/astrominn.py:12-17) is absent; ``timm.create_model`` is routed
  to the restated ConvNeXt-T in oracle/models.py (cross-checked against
  torchvision in tests/test_oracle_pinned.py).
* ``applecider/__init__.py`` imports a generated ``_version.py`` that is not in
  the tree, so the package is registered as a namespace shim.
"""
from __future__ import annotations

import copy
import os
import sys
import tomllib
import types

REF_ROOT = os.environ.get("APPLECIDER_REF_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_SRC, "applecider", "models"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install_stubs() -> None:
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")

    # --- hyrax ---------------------------------------------------------------
    def hyrax_model(cls):
        return cls

    class HyraxDataset:  # noqa: D401 - trivial base
        def __init__(self, config=None, metadata_table=None):
            pass

    hy = _mod("hyrax")
    hy.__path__ = []
    _mod("hyrax.models", hyrax_model=hyrax_model)
    ds = _mod("hyrax.data_sets", HyraxDataset=HyraxDataset)
    ds.__path__ = []
    _mod("hyrax.data_sets.data_set_registry", HyraxDataset=HyraxDataset)

    # --- astropy (I/O only) ----------------------------------------------------
    if "astropy" not in sys.modules:
        ap = _mod("astropy")
        ap.__path__ = []
        io = _mod("astropy.io", fits=None)
        io.__path__ = []
        _mod("astropy.io.fits")
        _mod("astropy.time", Time=type("Time", (), {}))
        _mod("astropy.table", Table=type("Table", (), {}))
        ut = _mod("astropy.utils")
        ut.__path__ = []
        _mod("astropy.utils.exceptions", AstropyWarning=type("AstropyWarning", (Warning,), {}))

    # --- timm → restated ConvNeXt-T -------------------------------------------
    if "timm" not in sys.modules:
        from . import models as _om

        def create_model(name, pretrained=False, in_chans=3, num_classes=0, **kw):
            assert name == "convnext_tiny" and not pretrained and num_classes == 0
            return _om.ConvNeXtTiny(in_chans=in_chans)

        _mod("timm", create_model=create_model)

    # --- applecider namespace shim ---------------------------------------------
    pkg = types.ModuleType("applecider")
    pkg.__path__ = [os.path.join(REF_SRC, "applecider")]
    sys.modules["applecider"] = pkg
    _installed = True


def default_config() -> dict:
    with open(os.path.join(REF_SRC, "applecider", "default_config.toml"), "rb") as f:
        cfg = tomllib.load(f)
    cfg["model"]["HyraxBaselineCLS"]["pretrained_weights_path_"] = False
    return cfg


def ref_models():
    """Return the reference model module namespace (real reference code)."""
    install_stubs()
    import importlib

    ns = types.SimpleNamespace()
    ns.photo = importlib.import_module("applecider.models.HyraxBaselineCLS")
    ns.time2vec = importlib.import_module("applecider.models.Time2Vec")
    ns.spectra = importlib.import_module("applecider.models.spectranet")
    ns.astrominn = importlib.import_module("applecider.models.astrominn")
    return ns


def ref_preprocess():
    install_stubs()
    import importlib

    return importlib.import_module("applecider.preprocessing_utils.preprocess_multimodal")


def ref_photo_dataset():
    install_stubs()
    import importlib

    return importlib.import_module("applecider.datasets.photo_dataset")


def cfg_copy(cfg: dict) -> dict:
    return copy.deepcopy(cfg)
