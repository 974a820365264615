"""ctypes binding of libb200ddp.so (include/b200ddp.h).  No torch types cross this boundary: tensors are
passed as ``data_ptr()`` + element counts and streams as raw ``cudaStream_t`` values.

There is deliberately NO fallback: if the shared library is missing the import of the data plane raises,
so a GPU box can never silently run a CPU or library path in its place.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "lib", "libb200ddp.so")
SRC_DIR = os.path.join(_PKG, "csrc")
SRC_PATH = os.path.join(SRC_DIR, "b200ddp.cu")  # the one translation unit; it includes the other files of csrc/
INCLUDE_DIR = os.path.join(os.path.dirname(_PKG), "include")

B2_ABI_VERSION = 2
B2_MAX_WORLD = 8

B2_OK = 0
B2_EINVAL = -1
B2_ECUDA = -2
B2_ESYS = -3
B2_ETIMEOUT = -4
B2_ENOPEER = -5
B2_ESTATE = -6
B2_ENOTSUP = -7

B2_F32_WIRE_BF16 = 0
B2_F32 = 1
B2_BF16 = 2

B2_ALGO_AUTO = 0
B2_ALGO_ONESHOT = 1
B2_ALGO_TWOSHOT = 2
B2_ALGO_TWOSHOT_PIPE = 3
B2_ALGO_NVLS = 4
B2_ALGO_TWOSHOT_LL = 5

B2_CAP_VMM = 1
B2_CAP_MULTICAST = 2

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC",
    "-shared",
]

# every symbol include/b200ddp.h declares (tests/test_abi.py checks the header and this list agree)
SYMBOLS = [
    "b2_version",
    "b2_last_error",
    "b2_comm_create",
    "b2_comm_create_local",
    "b2_comm_destroy",
    "b2_comm_rank",
    "b2_comm_world",
    "b2_comm_device",
    "b2_comm_caps",
    "b2_comm_set_timeout_ms",
    "b2_comm_set_max_ctas",
    "b2_comm_set_param",
    "b2_comm_status",
    "b2_comm_launch_count",
    "b2_comm_last_algo",
    "b2_auto_algo",
    "b2_comm_trace",
    "b2_allreduce",
    "b2_allreduce_gather",
    "b2_broadcast",
    "b2_barrier",
    "b2_local_pass",
]


B2_MAX_SEGMENTS = 128


class B2Segment(ctypes.Structure):
    """b2_segment_t: bucket elements [begin, end) live at device pointer src."""

    _fields_ = [("src", ctypes.c_void_p), ("begin", ctypes.c_uint64), ("end", ctypes.c_uint64)]


class B2Error(RuntimeError):
    def __init__(self, code: int, msg: str) -> None:
        super().__init__(f"libb200ddp error {code}: {msg}")
        self.code = code


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile torchx_b200/csrc/b200ddp.cu for sm_100a into torchx_b200/lib/ (in-tree, so the .so travels
    with the repo snapshot).  nvcc cross-compiles without a GPU."""
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    hdr = os.path.join(INCLUDE_DIR, "b200ddp.h")
    sources = [hdr] + [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR) if f.endswith((".cu", ".cuh", ".h"))]
    newest = max(os.path.getmtime(f) for f in sources)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, SRC_PATH, "-o", LIB_PATH, "-lrt"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA data plane has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, i, sz, u64, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_float
    L.b2_version.restype = i
    L.b2_version.argtypes = []
    L.b2_last_error.restype = ctypes.c_char_p
    L.b2_last_error.argtypes = []
    L.b2_comm_create.restype = i
    L.b2_comm_create.argtypes = [ctypes.POINTER(vp), i, i, i, ctypes.c_char_p, u64, sz, i]
    L.b2_comm_create_local.restype = i
    L.b2_comm_create_local.argtypes = [ctypes.POINTER(vp), i, ctypes.POINTER(i), sz]
    L.b2_comm_destroy.restype = i
    L.b2_comm_destroy.argtypes = [vp]
    for name in ("b2_comm_rank", "b2_comm_world", "b2_comm_device", "b2_comm_status", "b2_comm_caps", "b2_comm_last_algo"):
        getattr(L, name).restype = i
        getattr(L, name).argtypes = [vp]
    L.b2_comm_set_timeout_ms.restype = i
    L.b2_comm_set_timeout_ms.argtypes = [vp, i]
    L.b2_comm_set_max_ctas.restype = i
    L.b2_comm_set_max_ctas.argtypes = [vp, i]
    L.b2_comm_set_param.restype = i
    L.b2_comm_set_param.argtypes = [vp, ctypes.c_char_p, ctypes.c_longlong]
    L.b2_auto_algo.restype = i
    L.b2_auto_algo.argtypes = [i, i, sz, i]
    L.b2_comm_launch_count.restype = u64
    L.b2_comm_launch_count.argtypes = [vp]
    L.b2_comm_trace.restype = i
    L.b2_comm_trace.argtypes = [vp, i, ctypes.POINTER(u64), i]
    L.b2_allreduce.restype = i
    L.b2_allreduce.argtypes = [vp, vp, sz, i, f, i, vp]
    L.b2_allreduce_gather.restype = i
    L.b2_allreduce_gather.argtypes = [vp, vp, sz, ctypes.POINTER(B2Segment), i, i, f, i, vp]
    L.b2_broadcast.restype = i
    L.b2_broadcast.argtypes = [vp, vp, sz, i, vp]
    L.b2_barrier.restype = i
    L.b2_barrier.argtypes = [vp, vp]
    L.b2_local_pass.restype = i
    L.b2_local_pass.argtypes = [vp, sz, i, f, i, vp]
    if L.b2_version() != B2_ABI_VERSION:
        raise ImportError(f"libb200ddp ABI {L.b2_version()} != binding {B2_ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != B2_OK:
        raise B2Error(rc, lib().b2_last_error().decode("utf-8", "replace"))
