"""BASELINE.json config #1: toy MLP, 2 ranks, CPU/gloo - pure plumbing check of the launch path.

    python -m torchx_b200.cli.main run -s local_cwd  dist.ddp -j 1x2 --script examples/toy_ddp.py
    python -m torchx_b200.cli.main run -s local_cuda dist.ddp -j 1x2 --script examples/toy_ddp.py

Every rank wraps the same Linear(32,64)-ReLU-Linear(64,8) in stock DistributedDataParallel, runs one backward on its
own batch (seed 100+rank) and checks that its synced gradients equal sum_r(g_r / W) in rank order bit for bit and that
all ranks hold the same SHA-256 (exact at W=2 on any backend).
"""
import hashlib
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.nn.parallel import DistributedDataParallel as DDP

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchx_b200.distributed import init_pg, local_rank, rank, world_size  # noqa: E402


def model() -> nn.Module:
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(32, 64), nn.ReLU(), nn.Linear(64, 8))


def flat(m: nn.Module) -> torch.Tensor:
    return torch.cat([p.grad.reshape(-1) for p in m.parameters()])


def main() -> None:
    device = init_pg("gloo")
    r, w = rank(), world_size()
    torch.manual_seed(100 + r)
    x = torch.randn(16, 32)
    local = model()
    local(x).sum().backward()
    mine = flat(local).clone()
    gathered = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(gathered, mine)
    ddp = DDP(model())
    ddp(x).sum().backward()
    got = flat(ddp.module)
    want = None
    for g in gathered:
        c = g * (1.0 / w)
        want = c if want is None else want + c
    digest = hashlib.sha256(got.numpy().tobytes()).hexdigest()
    digests = [None] * w
    dist.all_gather_object(digests, digest)
    ok = torch.equal(got, want) if w <= 2 else torch.allclose(got, want, rtol=1e-6, atol=1e-7)
    print(f"rank {r}/{w} local_rank {local_rank()} device {device} grad sha256 {digest[:16]} exact={ok} same_on_all_ranks={len(set(digests)) == 1}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok or len(set(digests)) != 1:
        sys.exit(3)


if __name__ == "__main__":
    main()
