"""DistributedDataParallel on the B200 peer-buffer fabric: no torch.distributed, no NCCL.

Public surface mirrors ``torch.nn.parallel.DistributedDataParallel`` where the reference path relies on it
(torch/nn/parallel/distributed.py:664-890: rank-0 parameter/buffer broadcast at construction, 25 MiB buckets with
a 1 MiB first bucket, per-forward buffer broadcast, ``no_sync``, ``module.``-prefixed state dict), so a
``dist.ddp``-launched training script swaps one constructor.  Differences that matter for speed:

  * each bucket is reduced by ONE fused kernel (fp32->bf16 cast + 1/W scale + NVSwitch reduction + bf16->fp32)
    on a side stream while backward continues, instead of bf16_compress_hook's 4 launches;
  * zero-copy bucket fill: the kernel reads the gradients straight from the per-parameter tensors autograd produced
    (a device pointer table per bucket) and writes the averaged values into the persistent flat bucket, which
    ``param.grad`` aliases afterwards (gradient_as_bucket_view semantics) - the Reducer's copy-in pass
    (reducer.cpp mark_variable_ready_dense) and its 8 bytes per element are gone;
  * the layout is the steady-state one from iteration 0 (no rebuild pass).
"""
from __future__ import annotations

import contextlib
from typing import Dict, Iterator, List, Optional

import torch
from torch import nn

from .bucketing import MIB, BucketSpec, plan_buckets
from . import _native as N
from .comm import Communicator


def _dense_non_overlapping(t: torch.Tensor) -> bool:
    """True if the tensor's elements tile a contiguous block exactly once in SOME dimension order
    (contiguous, channels_last, any permutation)."""
    expected = 1
    for stride, size in sorted((st, sz) for sz, st in zip(t.shape, t.stride()) if sz != 1):
        if stride != expected:
            return False
        expected *= size
    return True


def _view_like(flat_slice: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    if p.is_contiguous() or not _dense_non_overlapping(p):
        return flat_slice.view(p.shape)
    return flat_slice.as_strided(p.shape, p.stride())


class _Bucket:
    def __init__(self, spec: BucketSpec, params: List[nn.Parameter], device: torch.device) -> None:
        self.spec = spec
        self.params = params
        self.flat = torch.zeros(spec.numel, dtype=params[0].dtype, device=device)
        # like the Reducer, give each view the parameter's own (dense) strides so a channels_last weight gets a
        # channels_last gradient view and the bucket's element order is the gradient's memory order
        self.views = [_view_like(self.flat[o : o + n], p) for o, n, p in zip(spec.offsets, spec.numels, params)]
        # strides of the dims that matter (size > 1): two tensors with these equal have the same memory order
        self.view_strides = [tuple(st for st, sz in zip(v.stride(), v.shape) if sz != 1) for v in self.views]
        self.pending = len(params)
        self.ready = False
        self.launched = False
        self.done = torch.cuda.Event()
        # ---- zero-copy fill: the segment table (begin / end fixed, pointers refreshed per iteration: autograd hands out
        # new tensors); it travels in the kernel parameters, so nothing is copied to the device or has to stay alive
        P = len(params)
        self.segments = None
        if P <= N.B2_MAX_SEGMENTS:
            self.segments = (N.B2Segment * P)()
            for i, (o, n) in enumerate(zip(spec.offsets, spec.numels)):
                self.segments[i].begin = o
                self.segments[i].end = o + n


class DistributedDataParallel(nn.Module):
    def __init__(
        self,
        module: nn.Module,
        comm: Optional[Communicator] = None,
        bucket_cap_mb: float = 25.0,
        first_bucket_mb: float = 1.0,
        wire: str = "bf16",
        broadcast_buffers: bool = True,
        algo: str = "auto",
        zero_copy: bool = True,
    ) -> None:
        super().__init__()
        self.module = module
        if comm is None:
            # one communicator per process: reuse the one init_pg("b200") created (each one owns a 1 GiB arena)
            from torchx_b200 import distributed as _dist

            comm = _dist._COMM if _dist._COMM is not None else Communicator.from_env()
        self.comm = comm
        self.world_size = self.comm.world
        self.wire = wire
        self.algo = algo
        self.zero_copy = zero_copy
        self.broadcast_buffers = broadcast_buffers
        self.require_backward_grad_sync = True
        self.device = torch.device("cuda", self.comm.device)
        self._params = [p for p in module.parameters() if p.requires_grad]
        if not self._params:
            raise RuntimeError("DistributedDataParallel is not needed when a module doesn't have any parameter that requires a gradient.")
        for p in self._params:
            if p.device != self.device:
                raise ValueError(f"parameter on {p.device}, communicator on {self.device}")
        specs = plan_buckets(
            [p.numel() for p in self._params],
            [p.element_size() for p in self._params],
            [str(p.dtype) for p in self._params],
            int(first_bucket_mb * MIB),
            int(bucket_cap_mb * MIB),
        )
        self.buckets = [_Bucket(s, [self._params[i] for i in s.param_indices], self.device) for s in specs]
        self._bucket_of: Dict[int, _Bucket] = {}
        for b in self.buckets:
            for p in b.params:
                self._bucket_of[id(p)] = b
        self._comm_stream = torch.cuda.Stream(device=self.device)
        self._ready_event = torch.cuda.Event()
        self._next_bucket = 0
        self._callback_queued = False
        self._profile: Optional[list] = None
        self.copied_in_buckets = 0   # buckets that needed the multi-tensor copy-in (zero-copy not applicable), for tests / bench
        self.gathered_buckets = 0    # buckets whose gradients were read in place by the kernel
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad_ready) for p in self._params]
        self._sync_module_states()

    # ---- construction-time / per-forward state sync (reference: distributed.py:881-890, 2176-2243) ----
    def _broadcast_coalesced(self, tensors: List[torch.Tensor], chunk_bytes: int = 250 * MIB) -> None:
        """Rank 0's values -> every rank, coalesced per dtype into flat chunks (logical element order, so
        channels_last and contiguous replicas agree)."""
        if self.world_size == 1 or not tensors:
            return
        by_dtype: Dict[torch.dtype, List[torch.Tensor]] = {}
        for t in tensors:
            by_dtype.setdefault(t.dtype, []).append(t)
        for group in by_dtype.values():
            chunks: List[List[torch.Tensor]] = [[]]
            size = 0
            for t in group:
                nbytes = t.numel() * t.element_size()
                if chunks[-1] and size + nbytes > chunk_bytes:
                    chunks.append([])
                    size = 0
                chunks[-1].append(t)
                size += nbytes
            for chunk in chunks:
                flat = torch.cat([t.detach().reshape(-1) for t in chunk])
                self.comm.broadcast_(flat, root=0)
                if self.comm.rank != 0:
                    outs = torch.split(flat, [t.numel() for t in chunk])
                    torch._foreach_copy_([t.detach() for t in chunk], [o.view_as(t) for o, t in zip(outs, chunk)])

    def _sync_module_states(self) -> None:
        self._broadcast_coalesced([p.data for p in self.module.parameters()] + [b.data for b in self.module.buffers()])

    def _sync_buffers(self) -> None:
        if self.broadcast_buffers and self.world_size > 1:
            bufs = [b.data for b in self.module.buffers()]
            if bufs:
                self._broadcast_coalesced(bufs)

    # ---- training step -----------------------------------------------------------------------------
    def _reset_reducer_state(self) -> None:
        """What Reducer::prepare_for_backward does: a backward that raised midway (caught OOM, skipped batch) must
        not leave stale counters behind - the engine drops its queued callbacks in that case."""
        self._callback_queued = False
        self._next_bucket = 0
        for b in self.buckets:
            b.pending, b.ready, b.launched = len(b.params), False, False

    def forward(self, *args, **kwargs):
        if torch.is_grad_enabled() and self.require_backward_grad_sync:
            # a kernel that gave up on a stalled peer leaves undefined bucket contents: never train on them
            self.comm.check()
            self._reset_reducer_state()
            self._sync_buffers()
        return self.module(*args, **kwargs)

    @contextlib.contextmanager
    def no_sync(self) -> Iterator[None]:
        """Gradient accumulation: skip the allreduce inside this context (distributed.py:1507-1531)."""
        old, self.require_backward_grad_sync = self.require_backward_grad_sync, False
        try:
            yield
        finally:
            self.require_backward_grad_sync = old

    def _on_grad_ready(self, p: nn.Parameter) -> None:
        if not self.require_backward_grad_sync:
            return
        if not self._callback_queued:
            self._callback_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize_backward)
        b = self._bucket_of[id(p)]
        b.pending -= 1
        if b.pending == 0:
            b.ready = True
            # launch strictly in bucket order so every rank issues the same collective sequence
            while self._next_bucket < len(self.buckets) and self.buckets[self._next_bucket].ready:
                self._launch(self.buckets[self._next_bucket])
                self._next_bucket += 1

    def _gatherable(self, b: _Bucket, grads: List[torch.Tensor]) -> bool:
        """The kernel can read a gradient in place when its memory order IS the bucket's element order: same dtype,
        same (dense) strides as the bucket view the parameter's layout produced."""
        dt = b.flat.dtype
        for g, st in zip(grads, b.view_strides):
            if g.dtype != dt or not g.is_cuda or tuple(s_ for s_, z in zip(g.stride(), g.shape) if z != 1) != st:
                return False
        return True

    def _launch(self, b: _Bucket) -> None:
        grads = []
        for p in b.params:
            g = p.grad
            if g is None:
                raise RuntimeError("a parameter finished backward without a gradient (unused parameters are not supported)")
            grads.append(g)
        gather = self.zero_copy and b.segments is not None and self._gatherable(b, grads)
        if not gather:
            src, dst = [], []
            for g, v in zip(grads, b.views):
                if g.data_ptr() != v.data_ptr():
                    src.append(g)
                    dst.append(v)
            if src:
                torch._foreach_copy_(dst, src)
            self.copied_in_buckets += 1
        else:
            for seg, g in zip(b.segments, grads):
                seg.src = g.data_ptr()
            self.gathered_buckets += 1
        cur = torch.cuda.current_stream(self.device)
        self._ready_event.record(cur)
        self._comm_stream.wait_event(self._ready_event)
        if self._profile is not None:
            t0 = torch.cuda.Event(enable_timing=True)
            t0.record(self._comm_stream)
        if gather:
            # the gradient tensors stay referenced by p.grad until _finalize_backward has made the compute stream wait for
            # b.done, so the caching allocator cannot hand their memory out before the kernel has read it
            self.comm.allreduce_gather_(b.flat, b.segments, len(b.params), scale=1.0 / self.world_size, wire=self.wire, algo=self.algo,
                                        stream=self._comm_stream)
        else:
            self.comm.allreduce_(b.flat, scale=1.0 / self.world_size, wire=self.wire, algo=self.algo, stream=self._comm_stream)
        b.done.record(self._comm_stream)
        if self._profile is not None:
            t1 = torch.cuda.Event(enable_timing=True)
            t1.record(self._comm_stream)
            self._profile.append((t0, t1, b.spec.numel * 2 * b.flat.element_size()))
        b.launched = True

    def _finalize_backward(self) -> None:
        self._callback_queued = False
        try:
            if self._next_bucket != len(self.buckets):
                missing = [b.spec.index for b in self.buckets if not b.launched]
                raise RuntimeError(
                    f"backward finished but buckets {missing} never became ready: some parameters received no gradient "
                    "(find_unused_parameters is not supported on this path)")
            cur = torch.cuda.current_stream(self.device)
            for b in self.buckets:
                cur.wait_event(b.done)
                for p, v in zip(b.params, b.views):
                    p.grad = v  # gradient_as_bucket_view: the optimizer reads the averaged bucket in place
            self.comm.check()
        finally:
            self._next_bucket = 0
            for b in self.buckets:
                b.pending, b.ready, b.launched = len(b.params), False, False

    # ---- measurement hooks used by bench.py (CUDA events on the comm stream, around every bucket kernel) ----
    def start_profile(self) -> None:
        self._profile = []

    def stop_profile(self) -> dict:
        """Sum of device time and algorithmic bytes (read the gradients once + write the bucket once) over the bucket
        kernels launched since start_profile(), plus the per-bucket-index mean device time."""
        torch.cuda.synchronize(self.device)
        prof, self._profile = self._profile or [], None
        times = [a.elapsed_time(b) * 1e-3 for a, b, _ in prof]
        seconds = sum(times)
        nb = len(self.buckets)
        per_bucket = []
        if prof and len(prof) % nb == 0:
            for i in range(nb):
                ts = times[i::nb]
                per_bucket.append(round(sum(ts) / len(ts) * 1e6, 2))
        name = "k_local_pass (W=1 fused cast/scale pass)" if self.world_size == 1 else "k_pipe / k_oneshot / k_twoshot (fused bucket allreduce)"
        return {"name": name, "launches": len(prof), "seconds": seconds, "alg_bytes": sum(n for _, _, n in prof), "per_bucket_us": per_bucket}

    def bucket_sizes_mib(self) -> List[float]:
        return [round(b.spec.nbytes / MIB, 2) for b in self.buckets]
