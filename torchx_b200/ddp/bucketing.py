"""Gradient-bucket layout, index-for-index identical to stock DistributedDataParallel's steady state.

Reference behaviour being matched (third-party torch on the reference path):
torch/nn/parallel/distributed.py:1219-1257 feeds ``dist._compute_bucket_assignment_by_size`` with limits
``[1 MiB, bucket_cap_mb MiB]`` and, after the first backward, ``Reducer::rebuild_buckets`` re-buckets the
parameters in gradient-ready order - which for a feed-forward model is reverse registration order.  The
layout is therefore static here: walk parameters last-to-first, close a bucket when it reaches its byte limit,
the first bucket (the first gradients to arrive) using the small limit so communication starts early.
tests/test_ddp_layout.py checks this module against torch's function and tests/golden/bucket_layouts.json.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Hashable, List, Sequence

MIB = 1 << 20
DEFAULT_FIRST_BUCKET_BYTES = 1 * MIB
DEFAULT_BUCKET_CAP_BYTES = 25 * MIB


@dataclass
class BucketSpec:
    """One flat bucket: parameter indices in fill order, element offsets into the flat buffer."""

    index: int
    key: Hashable
    param_indices: List[int] = field(default_factory=list)
    offsets: List[int] = field(default_factory=list)
    numels: List[int] = field(default_factory=list)
    numel: int = 0
    nbytes: int = 0

    def add(self, param_index: int, numel: int, itemsize: int) -> None:
        self.param_indices.append(param_index)
        self.offsets.append(self.numel)
        self.numels.append(numel)
        self.numel += numel
        self.nbytes += numel * itemsize


class _OpenBucket:
    __slots__ = ("spec", "limit_pos")

    def __init__(self, key: Hashable, limit_pos: int) -> None:
        self.spec = BucketSpec(index=-1, key=key)
        self.limit_pos = limit_pos


def plan_buckets(
    numels: Sequence[int],
    itemsizes: Sequence[int],
    keys: Sequence[Hashable],
    first_bucket_bytes: int = DEFAULT_FIRST_BUCKET_BYTES,
    bucket_cap_bytes: int = DEFAULT_BUCKET_CAP_BYTES,
) -> List[BucketSpec]:
    """numels/itemsizes/keys describe the trainable parameters in registration order; ``keys`` separates
    tensors that cannot share a flat buffer (dtype, device).  Returns buckets in launch order (bucket 0 holds
    the LAST parameters, whose gradients are produced first)."""
    limits = (first_bucket_bytes, bucket_cap_bytes)
    open_by_key: Dict[Hashable, _OpenBucket] = {}
    stage_by_key: Dict[Hashable, int] = {}
    closed: List[BucketSpec] = []
    for i in range(len(numels) - 1, -1, -1):
        key = keys[i]
        ob = open_by_key.get(key)
        if ob is None:
            ob = open_by_key[key] = _OpenBucket(key, stage_by_key.setdefault(key, 0))
        ob.spec.add(i, int(numels[i]), int(itemsizes[i]))
        if ob.spec.nbytes >= limits[ob.limit_pos]:
            closed.append(ob.spec)
            del open_by_key[key]
            stage_by_key[key] = min(ob.limit_pos + 1, len(limits) - 1)
    for ob in open_by_key.values():  # partially filled tails, in first-seen order
        closed.append(ob.spec)
    for k, spec in enumerate(closed):
        spec.index = k
    return closed
