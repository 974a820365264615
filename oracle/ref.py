"""ctypes front-end of oracle/allreduce_oracle.c plus a numpy-only twin used to cross-check it."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence

import numpy as np

B2O_F32_WIRE_BF16 = 0
B2O_F32 = 1
B2O_BF16 = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/allreduce_oracle.c with gcc (seconds). Returns the .so path."""
    src = os.path.join(_HERE, "allreduce_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "_build/liboracle.so"], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.b2o_allreduce.restype = ctypes.c_int
        lib.b2o_allreduce.argtypes = [
            ctypes.c_int,
            ctypes.c_int,
            ctypes.POINTER(ctypes.c_void_p),
            ctypes.c_size_t,
            ctypes.c_float,
            ctypes.c_void_p,
        ]
        lib.b2o_compress.restype = None
        lib.b2o_compress.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float, ctypes.c_void_p]
        lib.b2o_bf16_rne.restype = ctypes.c_uint16
        lib.b2o_bf16_rne.argtypes = [ctypes.c_float]
        _lib = lib
    return _lib


def _elem_dtype(mode: int):
    return np.uint16 if mode == B2O_BF16 else np.float32


def allreduce(mode: int, inputs: Sequence[np.ndarray], scale: float) -> np.ndarray:
    """inputs[r]: rank r's bucket (float32, or uint16 bf16 bit patterns for B2O_BF16). Returns the
    value every rank holds afterwards, same dtype."""
    lib = _load()
    dt = _elem_dtype(mode)
    arrs = [np.ascontiguousarray(a, dtype=dt) for a in inputs]
    n = arrs[0].size
    assert all(a.size == n for a in arrs)
    out = np.empty(n, dtype=dt)
    ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    rc = lib.b2o_allreduce(mode, len(arrs), ptrs, n, ctypes.c_float(scale), out.ctypes.data)
    if rc != 0:
        raise ValueError(f"b2o_allreduce rc={rc}")
    return out


def compress(mode: int, x: np.ndarray, scale: float) -> np.ndarray:
    lib = _load()
    x = np.ascontiguousarray(x, dtype=_elem_dtype(mode))
    out = np.empty(x.size, dtype=np.float32)
    lib.b2o_compress(mode, x.ctypes.data, x.size, ctypes.c_float(scale), out.ctypes.data)
    return out


# ---- numpy twins (independent restatement used to cross-check the C file) ------------------------
def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = (u & 0x7FFFFFFF) > 0x7F800000
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = 0x7FFF
    return r


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def allreduce_numpy(mode: int, inputs: Sequence[np.ndarray], scale: float) -> np.ndarray:
    sc = np.float32(scale)
    acc = None
    with np.errstate(all="ignore"):
        for a in inputs:
            if mode == B2O_F32:
                c = np.asarray(a, np.float32) * sc
            else:
                v = bf16_bits_to_f32(a) if mode == B2O_BF16 else bf16_bits_to_f32(f32_to_bf16_bits(a))
                c = bf16_bits_to_f32(f32_to_bf16_bits(v * sc))
            acc = c.astype(np.float32) if acc is None else (acc + c).astype(np.float32)
    if mode == B2O_F32:
        return acc
    bits = f32_to_bf16_bits(acc)
    return bits if mode == B2O_BF16 else bf16_bits_to_f32(bits)


def torch_hook_restatement(inputs: List["torch.Tensor"], hook: str):  # noqa: F821
    """The reference's own op sequence on CPU torch tensors, reduced in rank order with the wire dtype's
    rounding after every add (what a ring would do).  hook in {"none", "allreduce", "bf16_compress"}.
    Cites: default_hooks.py:18-33 (`_allreduce_fut`), :57-93 (`_compress_hook`)."""
    import torch

    w = len(inputs)
    if hook in ("none", "allreduce"):
        acc = None
        for g in inputs:
            c = g.clone().float()
            c = c * (1.0 / w) if hook == "none" else c.div_(w)
            acc = c if acc is None else acc + c
        return acc
    if hook == "bf16_compress":
        acc = None
        for g in inputs:
            c = g.to(torch.bfloat16).div_(w)
            acc = c if acc is None else acc + c  # bf16 + bf16 -> rounds to bf16 each step
        return acc.float()
    raise ValueError(hook)
