"""ctypes binding of libapplecider_b200.so (the C-ABI declared in include/applecider_b200.h).

There is no CPU fallback: if the library is missing or a tensor is not on a CUDA
device, the call raises.  Signature strings: p = pointer, i = int, l = long long,
f = float; the trailing ``stream`` pointer is appended automatically.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libapplecider_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_TANH, ACT_SIGMOID, ACT_SOFTPLUS = 0, 1, 2, 3, 4, 5
RES_NONE, RES_ADD, RES_MUL, RES_MUL_GELU_GRAD = 0, 1, 2, 3

HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "applecider_b200.h")


def _parse_header(path: str) -> dict:
    """Derive ctypes signatures from the C header: p = pointer, i = int, l = long long, f = float."""
    import re

    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    sigs = {}
    for m in re.finditer(r"\bint\s+(acb_\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(1), m.group(2)
        if name in NO_STREAM:
            continue
        kinds = []
        for a in args.split(","):
            a = a.strip()
            if not a or a == "void":
                continue
            if "*" in a:
                kinds.append("p")
            elif "long long" in a:
                kinds.append("l")
            elif re.search(r"\bfloat\b", a):
                kinds.append("f")
            elif re.search(r"\bdouble\b", a):
                kinds.append("d")
            else:
                kinds.append("i")
        if kinds and kinds[-1] == "p" and "stream" in args.split(",")[-1]:
            kinds = kinds[:-1]  # the stream is appended by call()
            sigs[name] = "".join(kinds)
    return sigs


NO_STREAM = ("acb_version", "acb_set_seed_epoch_ptr")  # management calls without a trailing stream argument
SIGNATURES = _parse_header(HEADER_PATH)

_KIND = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_longlong, "f": ctypes.c_float, "d": ctypes.c_double}

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"applecider_b200: CUDA library not built ({LIB_PATH} missing). "
                "Run `python -m applecider_b200.build` (nvcc, sm_100a); there is no CPU fallback."
            )
        l = ctypes.CDLL(LIB_PATH)
        l.acb_last_error.restype = ctypes.c_char_p
        l.acb_launch_count.restype = ctypes.c_longlong
        l.acb_set_seed_epoch_ptr.restype = ctypes.c_int
        l.acb_set_seed_epoch_ptr.argtypes = [ctypes.c_void_p]
        for name, sig in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [_KIND[k] for k in sig] + [ctypes.c_void_p]
        _lib = l
    return _lib


def dtype_tag(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"applecider_b200: unsupported dtype {t.dtype}")


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise RuntimeError("applecider_b200: tensors must live on a CUDA device (no CPU fallback)")
        if not a.is_contiguous():
            raise RuntimeError("applecider_b200: tensors must be contiguous")
        return a.data_ptr()
    if isinstance(a, ctypes.Array):
        return ctypes.cast(a, ctypes.c_void_p)
    if isinstance(a, int):
        return a
    raise TypeError(f"cannot pass {type(a)} as a pointer")


_FN_CACHE = {}
_raw_stream = None


def _current_stream() -> int:
    global _raw_stream
    if _raw_stream is None:
        _raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or False
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    ent = _FN_CACHE.get(name)
    if ent is None:
        l = lib()
        ent = _FN_CACHE[name] = (getattr(l, name), SIGNATURES[name], l.acb_last_error)
    fn, sig, last_error = ent
    if len(args) != len(sig):
        raise TypeError(f"{name}: expected {len(sig)} arguments, got {len(args)}")
    conv = []
    dev = None
    for k, a in zip(sig, args):
        if k == "p":
            if a is None or type(a) is int:
                conv.append(a)
            elif isinstance(a, torch.Tensor):
                if not a.is_cuda:
                    raise RuntimeError("applecider_b200: tensors must live on a CUDA device (no CPU fallback)")
                if not a.is_contiguous():
                    raise RuntimeError("applecider_b200: tensors must be contiguous")
                if dev is None:
                    dev = a.device.index
                elif a.device.index != dev:
                    raise RuntimeError(f"applecider_b200: {name}: tensors live on different devices (cuda:{dev} and cuda:{a.device.index})")
                conv.append(a.data_ptr())
            else:
                conv.append(_ptr(a))
        elif k == "i" or k == "l":
            conv.append(int(a))
        else:
            conv.append(float(a))
    if dev is not None and dev != torch.cuda.current_device():
        # the launch goes to the CURRENT device's stream: run under torch.cuda.device(tensor.device) (one process per GPU is the norm)
        raise RuntimeError(f"applecider_b200: {name}: tensors live on cuda:{dev} but the current device is cuda:{torch.cuda.current_device()}")
    rc = fn(*conv, _current_stream())
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {last_error().decode()}")


def launch_count() -> int:
    return int(lib().acb_launch_count())


def reset_launch_count() -> None:
    lib().acb_reset_launch_count()


_seed_epoch_tensor = None


def set_seed_epoch(tensor) -> None:
    """tensor: a 1-element CUDA int64 tensor the stochastic kernels add to their seeds (see acb_set_seed_epoch_ptr), or None to
    clear.  The module keeps a reference, so the device pointer the library holds can never dangle."""
    global _seed_epoch_tensor
    if tensor is None:
        lib().acb_set_seed_epoch_ptr(None)
        _seed_epoch_tensor = None
        return
    if not (tensor.is_cuda and tensor.dtype == torch.int64 and tensor.numel() == 1):
        raise TypeError("seed epoch must be a 1-element CUDA int64 tensor")
    lib().acb_set_seed_epoch_ptr(tensor.data_ptr())
    _seed_epoch_tensor = tensor


def exported_symbols():
    return sorted(SIGNATURES) + ["acb_last_error", "acb_version", "acb_launch_count", "acb_reset_launch_count", "acb_set_seed_epoch_ptr"]
