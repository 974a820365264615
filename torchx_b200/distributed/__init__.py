"""Worker-side helpers, importable from the training script the launcher starts.

Mirror of reference torchx/distributed/__init__.py (local_rank:26, local_cuda_device:57, local_device:70, rank:91,
world_size:109, init_pg:164, on_rank0_first:231, on_local_rank0_first:278): rank/device discovery from the torchrun
environment contract and barrier-ordered critical sections.  Two backends behind the same functions:

  * a ``torch.distributed`` process group, exactly as the reference does (``init_pg("auto")`` -> nccl on GPU hosts, gloo
    otherwise) - this is what config #1 (CPU/gloo) and unmodified TorchX scripts use;
  * ``init_pg("b200")`` - the peer-buffer communicator from ``libb200ddp.so``: no TCP store, no NCCL.  ``barrier`` /
    ``on_rank0_first`` then run on that fabric.
"""
from __future__ import annotations

import os
import warnings
from contextlib import contextmanager
from typing import Any, Iterator, Optional

import torch
import torch.distributed as dist

from torchx_b200.util.cuda import has_cuda_devices

_COMM: Optional[Any] = None  # torchx_b200.ddp.Communicator when init_pg("b200") was used


def local_rank() -> int:
    """``LOCAL_RANK``; 0 when the launcher did not set it - with a warning if a process group is nevertheless up, because
    then every process of the node would claim device 0 (reference torchx/distributed/__init__.py:26-54)."""
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"])
    if dist.is_available() and dist.is_initialized():
        warnings.warn("the default torch.distributed process group is initialized but the `LOCAL_RANK` environment variable is "
                      "not set: local_rank() trivially returns 0.  Launch the script with torchx (local_cuda / local_cwd) or "
                      "torchrun, or set `LOCAL_RANK` yourself.")
    return 0


def rank() -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank()
    return int(os.environ.get("RANK", "0")) if _COMM is None else _COMM.rank


def world_size() -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return int(os.environ.get("WORLD_SIZE", "1")) if _COMM is None else _COMM.world


def local_cuda_device() -> torch.device:
    """``cuda:$B2_DEVICE`` under ``local_cuda`` (the scheduler pins the device), else ``cuda:$LOCAL_RANK``."""
    return torch.device("cuda", int(os.environ.get("B2_DEVICE", str(local_rank()))))


def local_device() -> torch.device:
    """The device this process should place its model on: ``cuda:<pinned ordinal>`` when the B200 communicator or an
    nccl process group is active, ``cpu`` under any other process group; with nothing initialised, ``cuda`` if the host
    has GPUs else ``cpu`` (reference torchx/distributed/__init__.py:70-88)."""
    if _COMM is not None:
        return torch.device("cuda", _COMM.device)
    if dist.is_available() and dist.is_initialized():
        return local_cuda_device() if dist.get_backend() == "nccl" else torch.device("cpu")
    return torch.device("cuda") if has_cuda_devices() else torch.device("cpu")


def is_rank0() -> bool:
    return rank() == 0


def is_local_rank0() -> bool:
    return local_rank() == 0


def communicator() -> Any:
    """The process-wide B200 communicator (``init_pg("b200")`` must have been called)."""
    if _COMM is None:
        raise RuntimeError('no B200 communicator: call torchx_b200.distributed.init_pg("b200") first')
    return _COMM


def is_torchelastic_launched() -> bool:
    return "TORCHELASTIC_RUN_ID" in os.environ


def init_pg(backend: str = "auto", **kwargs: Any) -> torch.device:
    """Initialise this worker's collective backend and return the device it should use.

    ``auto``: nccl when CUDA devices exist, gloo otherwise (reference behaviour).  When the process was not launched
    by a torchrun-compatible launcher a trivial single-rank group is created so scripts also run with plain ``python``.
    As in the reference the caller selects the returned device (``torch.cuda.set_device`` / ``.to(device)``).
    ``b200``: rendezvous through the ``local_cuda`` scheduler's shm control block and CUDA IPC; no process group.
    """
    global _COMM
    if backend == "b200":
        if _COMM is None:
            from torchx_b200.ddp import Communicator

            _COMM = Communicator.from_env(**kwargs)
        return torch.device("cuda", _COMM.device)
    if backend == "auto":
        # a CUDA build of torch says is_available() on some CPU hosts: ask for actual devices and for NCCL itself
        has_gpu = torch.cuda.is_available() and torch.cuda.device_count() > 0 and dist.is_nccl_available()
        backend = "nccl" if has_gpu else "gloo"
    if is_torchelastic_launched():
        dist.init_process_group(backend=backend, **kwargs)
    else:  # plain `python script.py`: a one-rank group on a free port, so the same script runs either way
        os.environ["MASTER_ADDR"] = "localhost"
        os.environ["MASTER_PORT"] = "0"
        dist.init_process_group(backend=backend, rank=0, world_size=1, **kwargs)
    return local_cuda_device() if backend == "nccl" else torch.device("cpu")


def barrier() -> None:
    if _COMM is not None and not dist.is_initialized():
        _COMM.barrier()
        torch.cuda.current_stream(_COMM.device).synchronize()
        _COMM.check()  # a barrier kernel that gave up on a stalled peer must not let this rank into the critical section
    elif dist.is_initialized():
        dist.barrier()


@contextmanager
def on_rank0_first() -> Iterator[None]:
    """Run the block on rank 0 first, then on everyone else (download-once patterns)."""
    if rank() != 0:
        barrier()
    try:
        yield
    finally:
        if rank() == 0:
            barrier()


@contextmanager
def on_local_rank0_first() -> Iterator[None]:
    """Single-box topology: local rank 0 == rank 0 of its node; with one node this equals on_rank0_first()."""
    first = local_rank() == 0
    if not first:
        barrier()
    try:
        yield
    finally:
        if first:
            barrier()
