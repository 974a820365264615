"""CPU oracle for the DDP bucket-allreduce hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  Nothing under ``torchx_b200/`` imports it; the product path fails loudly
when the CUDA extension is missing instead of falling back to this code.

Pinned against the reference: ``tests/golden/ddp_w{2,4}_*.npz`` hold inputs and outputs of stock
``DistributedDataParallel`` (no hook / ``allreduce_hook`` / ``bf16_compress_hook``) launched through the
reference's own ``torchx run -s local_cwd dist.ddp`` over gloo in the build container
(``tests/golden/make_golden.py``); ``tests/test_oracle.py`` checks this oracle against them (bit-exact at
W=2, toleranced at W=4 where the backend's reduction order is not rank order).
"""
from .ref import (  # noqa: F401
    B2O_BF16,
    B2O_F32,
    B2O_F32_WIRE_BF16,
    allreduce,
    bf16_bits_to_f32,
    build,
    compress,
    f32_to_bf16_bits,
    torch_hook_restatement,
)
from .bucketing import compute_bucket_assignment_by_size, ddp_bucket_layout  # noqa: F401
