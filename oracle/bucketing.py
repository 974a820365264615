"""Restatement of DDP's parameter -> bucket assignment (layout parity, bit-exact on indices/offsets).

Follows torch/csrc/distributed/c10d/reducer.cpp `compute_bucket_assignment_by_size` as driven by
torch/nn/parallel/distributed.py:1219-1257 (limits ``[dist._DEFAULT_FIRST_BUCKET_BYTES (1 MiB),
bucket_cap_mb * 2**20 (25 MiB)]``) and by ``Reducer::rebuild_buckets`` after the first backward, which
re-runs the same function over the parameters in gradient-ready order (approximated here, as in
distributed.py:1253-1257, by REVERSE registration order).  tests/test_bucketing.py checks it against
torch's own ``dist._compute_bucket_assignment_by_size`` and the committed layouts in tests/golden/.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

FIRST_BUCKET_BYTES = 1024 * 1024  # dist._DEFAULT_FIRST_BUCKET_BYTES
DEFAULT_CAP_BYTES = 25 * 1024 * 1024  # bucket_cap_mb=25


def compute_bucket_assignment_by_size(
    sizes: Sequence[Tuple[int, str]],
    limits: Sequence[int],
    tensor_indices: Sequence[int] = (),
) -> Tuple[List[List[int]], List[int]]:
    """sizes[i] = (nbytes, key) where key identifies (dtype, device).  Returns (bucket index lists,
    per-bucket size limits), exactly like the C++ function for dense gradients."""
    acc: Dict[str, Tuple[List[int], int]] = {}
    which: Dict[str, int] = {}
    order: List[str] = []
    result: List[Tuple[List[int], int]] = []
    for i, (nbytes, key) in enumerate(sizes):
        idx = tensor_indices[i] if tensor_indices else i
        if key not in acc:
            acc[key] = ([], 0)
            order.append(key)
        if key not in which:
            which[key] = 0
        indices, size = acc[key]
        indices.append(idx)
        size += nbytes
        limit = limits[which[key]]
        if size >= limit:
            result.append((indices, limit))
            acc[key] = ([], 0)
            if which[key] + 1 < len(limits):
                which[key] += 1
        else:
            acc[key] = (indices, size)
    for key in order:
        indices, _ = acc[key]
        if indices:
            result.append((indices, limits[which[key]]))
    if not tensor_indices:
        result.sort(key=lambda b: min(b[0]))
    return [b[0] for b in result], [b[1] for b in result]


def ddp_bucket_layout(
    param_nbytes: Sequence[int],
    key: str = "f32",
    first_bytes: int = FIRST_BUCKET_BYTES,
    cap_bytes: int = DEFAULT_CAP_BYTES,
) -> List[List[int]]:
    """Steady-state DDP layout: parameters visited in reverse registration order, first bucket capped at
    `first_bytes`, the rest at `cap_bytes`.  Bucket 0 is the first to become ready in backward."""
    n = len(param_nbytes)
    rev = list(range(n - 1, -1, -1))
    buckets, _ = compute_bucket_assignment_by_size(
        [(param_nbytes[i], key) for i in rev], [first_bytes, cap_bytes], tensor_indices=rev
    )
    return buckets
