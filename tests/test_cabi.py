"""CPU suite: the C-ABI library builds, loads and exports every symbol include/applecider_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "applecider_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from applecider_b200 import _lib, build

    build.build(verbose=False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"symbol {n} declared in the header but not exported"
    assert _lib.lib().acb_version() >= 100


def test_signatures_cover_header():
    from applecider_b200 import _lib

    declared = set(_declared()) - {"acb_last_error", "acb_version", "acb_launch_count", "acb_reset_launch_count"} - set(_lib.NO_STREAM)
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_cpu_tensors_fail_loudly():
    import pytest
    import torch

    from applecider_b200 import ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8), 1e-5)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "applecider_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
