"""Communicator: the worker-side handle on the NVSwitch peer-buffer fabric.

Replaces, for workers launched by the ``local_cuda`` scheduler, what ``torchx.distributed.init_pg`` obtains
from ``torch.distributed.init_process_group("nccl")`` (reference torchx/distributed/__init__.py:164-225):
rank/world discovery from the torchrun env contract, device pinning, and the collectives the DDP path needs.
torch is used for tensors and streams only.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import torch

from . import _native as N

_MODE_FOR = {
    ("f32", "bf16"): N.B2_F32_WIRE_BF16,
    ("f32", "f32"): N.B2_F32,
    ("bf16", "bf16"): N.B2_BF16,
}
ALGOS = {"auto": N.B2_ALGO_AUTO, "oneshot": N.B2_ALGO_ONESHOT, "twoshot": N.B2_ALGO_TWOSHOT, "twoshot_pipe": N.B2_ALGO_TWOSHOT_PIPE,
         "nvls": N.B2_ALGO_NVLS, "twoshot_ll": N.B2_ALGO_TWOSHOT_LL}


def mode_for(tensor: torch.Tensor, wire: str = "bf16") -> int:
    if tensor.dtype == torch.float32:
        key = ("f32", wire)
    elif tensor.dtype == torch.bfloat16:
        key = ("bf16", "bf16")
    else:
        raise TypeError(f"unsupported gradient dtype {tensor.dtype}; expected float32 or bfloat16")
    if key not in _MODE_FOR:
        raise ValueError(f"unsupported wire format {wire!r} for dtype {tensor.dtype}")
    return _MODE_FOR[key]


def _stream_ptr(stream: Optional[torch.cuda.Stream], device: int) -> int:
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return int(s.cuda_stream)


def default_shm_name() -> str:
    """Name of the rendezvous control block.  The ``local_cuda`` scheduler exports B2_SHM_NAME; under a plain
    ``torchrun`` (the bench driver) all workers share a parent agent and a MASTER_PORT, which is unique per job
    on one box."""
    name = os.environ.get("B2_SHM_NAME")
    if name:
        return name
    run_id = os.environ.get("TORCHELASTIC_RUN_ID", "none")
    port = os.environ.get("MASTER_PORT", "0")
    return f"/b2_{os.getppid()}_{port}_{''.join(ch for ch in run_id if ch.isalnum())[:32]}"


class Communicator:
    """One rank's endpoint. Collectives are asynchronous on the given (or current) CUDA stream and must be
    issued in the same order on every rank."""

    def __init__(self, handle: int, owner: bool = True) -> None:
        self._h = ctypes.c_void_p(handle)
        self._owner = owner
        L = N.lib()
        self.rank = L.b2_comm_rank(self._h)
        self.world = L.b2_comm_world(self._h)
        self.device = L.b2_comm_device(self._h)

    # ---- construction --------------------------------------------------------------------------
    @classmethod
    def from_env(cls, stage_mb: int = 0, timeout_s: float = 120.0) -> "Communicator":
        """Bootstrap from the env contract torchrun / local_cuda give every worker
        (torch/distributed/elastic/agent/server/local_elastic_agent.py:309-323)."""
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
        device = int(os.environ.get("B2_DEVICE", str(local_rank)))
        epoch = int(os.environ.get("B2_EPOCH", os.environ.get("TORCHELASTIC_RESTART_COUNT", "0")))
        return cls.create(rank, world, device, default_shm_name(), epoch, stage_mb, timeout_s)

    @classmethod
    def create(cls, rank: int, world: int, device: int, shm_name: str, epoch: int = 0, stage_mb: int = 0,
               timeout_s: float = 120.0) -> "Communicator":
        torch.cuda.set_device(device)
        torch.cuda.init()
        out = ctypes.c_void_p()
        N.check(N.lib().b2_comm_create(ctypes.byref(out), rank, world, device, shm_name.encode(), epoch,
                                       stage_mb << 20, int(timeout_s * 1000)))
        return cls(out.value)

    @classmethod
    def create_local(cls, devices: Sequence[int], stage_mb: int = 0) -> List["Communicator"]:
        """All ranks inside this process (tests / single-GPU parity topology)."""
        torch.cuda.init()
        w = len(devices)
        outs = (ctypes.c_void_p * w)()
        devs = (ctypes.c_int * w)(*devices)
        N.check(N.lib().b2_comm_create_local(outs, w, devs, stage_mb << 20))
        return [cls(outs[i]) for i in range(w)]

    def close(self) -> None:
        if self._h and self._owner:
            N.lib().b2_comm_destroy(self._h)
        self._h = ctypes.c_void_p()

    def __enter__(self) -> "Communicator":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

    # ---- tuning / health -----------------------------------------------------------------------
    def set_timeout(self, seconds: float) -> None:
        N.check(N.lib().b2_comm_set_timeout_ms(self._h, int(seconds * 1000)))

    def set_max_ctas(self, n: int) -> None:
        N.check(N.lib().b2_comm_set_max_ctas(self._h, n))

    def set_param(self, name: str, value: int) -> None:
        """AUTO thresholds / pipeline chunking (include/b200ddp.h: b2_comm_set_param); same value on every rank."""
        N.check(N.lib().b2_comm_set_param(self._h, name.encode(), int(value)))

    @property
    def caps(self) -> int:
        return int(N.lib().b2_comm_caps(self._h))

    @property
    def has_multicast(self) -> bool:
        """True when every rank's arena is bound into one NVSwitch multicast object (the NVLS algorithm is available)."""
        return bool(self.caps & N.B2_CAP_MULTICAST)

    def check(self) -> None:
        """Raise if any kernel of this communicator timed out waiting for a peer."""
        N.check(N.lib().b2_comm_status(self._h))

    def trace(self, enable: bool, read_ctas: int = 0):
        """Phase-boundary timestamps (ns, %globaltimer) of the last collective: list of 8-tuples per CTA."""
        buf = (ctypes.c_uint64 * (8 * read_ctas))() if read_ctas else None
        N.check(N.lib().b2_comm_trace(self._h, int(enable), buf, read_ctas))
        return [tuple(buf[8 * i: 8 * i + 8]) for i in range(read_ctas)] if buf is not None else []

    @property
    def last_algo(self) -> str:
        """Name of the algorithm the most recent multi-rank allreduce ran (what "auto" resolved to)."""
        k = int(N.lib().b2_comm_last_algo(self._h))
        return {v: n for n, v in ALGOS.items()}.get(k, "none") if k else "none"

    @property
    def launches(self) -> int:
        return int(N.lib().b2_comm_launch_count(self._h))

    # ---- collectives ---------------------------------------------------------------------------
    def allreduce_(self, t: torch.Tensor, scale: Optional[float] = None, wire: str = "bf16", algo: str = "auto",
                   stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """In-place ``t <- round(sum_r wire(scale * t_r))``; scale defaults to 1/world (gradient averaging)."""
        self._check_tensor(t)
        if scale is None:
            scale = 1.0 / self.world
        N.check(N.lib().b2_allreduce(self._h, ctypes.c_void_p(t.data_ptr()), t.numel(), mode_for(t, wire),
                                     ctypes.c_float(scale), ALGOS[algo], ctypes.c_void_p(_stream_ptr(stream, self.device))))
        return t

    def allreduce_gather_(self, out: torch.Tensor, segments, n_segments: int, scale: Optional[float] = None, wire: str = "bf16",
                          algo: str = "auto", stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        """``out <- round(sum_r wire(scale * concat(segments_r)))``: the bucket is only written; the input is read
        straight from the tensors the segment table points at (include/b200ddp.h: b2_allreduce_gather).  ``segments`` is a
        ctypes array of ``_native.B2Segment`` (device pointer, begin, end) covering the bucket in order; it is copied into
        the kernel parameters by the call."""
        self._check_tensor(out)
        if scale is None:
            scale = 1.0 / self.world
        N.check(N.lib().b2_allreduce_gather(self._h, ctypes.c_void_p(out.data_ptr()), out.numel(), segments, n_segments, mode_for(out, wire),
                                            ctypes.c_float(scale), ALGOS[algo], ctypes.c_void_p(_stream_ptr(stream, self.device))))
        return out

    def broadcast_(self, t: torch.Tensor, root: int = 0, stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
        self._check_tensor(t)
        N.check(N.lib().b2_broadcast(self._h, ctypes.c_void_p(t.data_ptr()), t.numel() * t.element_size(), root,
                                     ctypes.c_void_p(_stream_ptr(stream, self.device))))
        return t

    def barrier(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        N.check(N.lib().b2_barrier(self._h, ctypes.c_void_p(_stream_ptr(stream, self.device))))

    def _check_tensor(self, t: torch.Tensor) -> None:
        if not t.is_cuda or t.device.index != self.device:
            raise ValueError(f"tensor on {t.device}, communicator on cuda:{self.device}")
        if not t.is_contiguous():
            raise ValueError("collectives need a contiguous tensor")


def local_pass_(t: torch.Tensor, scale: float = 1.0, wire: str = "bf16", stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """The W == 1 fused cast/scale pass (b2_local_pass) on its own."""
    dev = t.device.index
    N.check(N.lib().b2_local_pass(ctypes.c_void_p(t.data_ptr()), t.numel(), mode_for(t, wire), ctypes.c_float(scale), dev,
                                  ctypes.c_void_p(_stream_ptr(stream, dev))))
    return t
