"""applecider_b200 — B200-native (sm_100a) hot path of the AppleCiDEr multimodal classifier.

Drop-in torch.nn.Modules with the reference's constructor/forward signatures and state_dict keys;
all arithmetic runs in hand-written CUDA kernels behind the C-ABI in include/applecider_b200.h.
There is no CPU fallback: the first op call raises if the library is not built.
"""
__version__ = "0.1.0"

from .config import default_config, resolve_dtype  # noqa: F401


def __getattr__(name):
    # lazy: keeps `import applecider_b200` cheap and side-effect free (synth / build helpers need no GPU)
    import importlib

    table = {
        "HyraxBaselineCLS": "photo", "MPTModel": "photo", "BaselineCLS": "photo", "Time2Vec": "photo", "FocalLoss": "photo",
        "SpectraNet": "spectra", "SpectraNetBlock": "spectra",
        "AstroMiNN": "astrominn", "SplitHeadConvNeXt": "astrominn", "ResidualTowerBlock": "astrominn", "ConvNeXtTiny": "astrominn",
        "AppleCider": "fusion", "fusion_collate": "fusion", "SpectraClassificationB": "legacy",
    }
    if name in table:
        return getattr(importlib.import_module(f".{table[name]}", __name__), name)
    raise AttributeError(name)
