"""Hyper-parameters of the hot path (restated from the reference's src/applecider/default_config.toml:1-119)
and the compute-dtype switch (fp32 parity path vs bf16 tcgen05 path)."""
from __future__ import annotations

import copy
import os

import torch

_DEFAULT = {
    "model": {
        "AstroMiNN": {
            "num_classes": 9, "num_mlp_experts": 4, "use_probabilities": False,
            "towers_hidden_dims": 16, "towers_outdims": 32,
            "fusion_hidden_dims": 128, "fusion_router_dims": 128, "fusion_outdims": 32,
            "cnn_lr": 2, "cnn_decay": 5e-2, "psf_lr": 0.5, "psf_decay": 5e-2,
            "mag_lr": 2, "mag_decay": 0.0, "lc_lr": 2, "lc_decay": 0.05,
            "spatial_lr": 2, "spatial_decay": 0.0, "coord_lr": 0.5, "coord_decay": 0.0,
            "nst1_lr": 2, "nst1_decay": 0.0, "nst2_lr": 2, "nst2_decay": 0.0,
            "fusion_lr": 1, "fusion_decay": 1e-2, "fusion_beta1": 0.9, "fusion_beta2": 0.999,
            "router_decay": 0.0, "router_lr": 1.5, "router_beta1": 0.9, "router_beta2": 0.999,
            "beta1": 0.9, "beta2": 0.999, "eps": 5e-10,
        },
        "HyraxBaselineCLS": {
            "num_classes": 5, "pad_mask": 1, "mode": "photo", "d_model": 128, "n_heads": 8,
            "n_layers": 4, "dropout": 0.40, "max_len": 257, "lr": 5e-6, "weight_decay": 1e-2,
            "focal_gamma": 2.0, "use_probabilities": False, "pretrained_weights_path_": False,
            "lambda_f": 5.0, "lambda_b": 3.0, "lambda_dt": 5.0, "mask_p": 0.30,
        },
        "SpectraNet": {
            "redshift": False, "use_ln_stages": [True] * 5, "depths": [1] * 5,
            "channels": [64, 128, 256, 512, 1024],
            "kernel_sizes_per_stage": [[3, 61, 1021], [3, 31, 251], [3, 15, 61], [3, 11, 31], [3, 7, 13]],
            "class_order": 9, "flat_dim": 3072,
        },
    }
}


def default_config() -> dict:
    return copy.deepcopy(_DEFAULT)


def resolve_dtype(v=None) -> torch.dtype:
    """"fp32" (CUDA-core parity path) or "bf16" (tcgen05 path).  Default: env APPLECIDER_B200_DTYPE or fp32."""
    if isinstance(v, torch.dtype):
        return v
    if v is None:
        v = os.environ.get("APPLECIDER_B200_DTYPE", "fp32")
    v = str(v).lower()
    if v in ("fp32", "float32", "f32"):
        return torch.float32
    if v in ("bf16", "bfloat16"):
        return torch.bfloat16
    raise ValueError(f"unknown compute dtype {v!r}")
