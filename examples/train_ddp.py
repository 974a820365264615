"""Small end-to-end training job on the B200 data plane, launched through the scheduler:

    python -m torchx_b200.cli.main run -s local_cuda dist.ddp -j 1x8 --script examples/train_ddp.py -- --steps 20

No torch.distributed anywhere: rank/device come from the env contract, the communicator from the scheduler's shm
control block, gradients are averaged by the fused kernels.  ``--fail-rank R --fail-at-step K`` makes rank R exit(17)
at step K of the FIRST attempt only (BASELINE.json config #5: rank drop -> gang re-launch under a new epoch when the job
was submitted with ``--max_retries``).  At the end every rank prints a hash of its parameters; they must agree.
"""
import argparse
import hashlib
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchx_b200.ddp import DistributedDataParallel  # noqa: E402
from torchx_b200.distributed import communicator, init_pg, on_rank0_first  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--fail-rank", type=int, default=-1)
    ap.add_argument("--fail-at-step", type=int, default=-1)
    ap.add_argument("--max-ctas", type=int, default=0)
    a = ap.parse_args()

    device = init_pg("b200")
    comm = communicator()
    if a.max_ctas:
        comm.set_max_ctas(a.max_ctas)
    attempt = int(os.environ.get("TORCHELASTIC_RESTART_COUNT", "0"))
    torch.manual_seed(1234 + comm.rank)  # deliberately different init per rank: DDP must broadcast rank 0's
    model = nn.Sequential(nn.Linear(256, a.hidden), nn.ReLU(), nn.Linear(a.hidden, a.hidden), nn.ReLU(), nn.Linear(a.hidden, 10)).to(device)
    ddp = DistributedDataParallel(model, comm, bucket_cap_mb=1.0, first_bucket_mb=0.25)
    opt = torch.optim.SGD(ddp.parameters(), lr=0.05, momentum=0.9)
    gen = torch.Generator(device=device).manual_seed(99 + comm.rank)
    with on_rank0_first():
        pass  # e.g. dataset download
    for step in range(a.steps):
        if attempt == 0 and comm.rank == a.fail_rank and step == a.fail_at_step:
            print(f"rank {comm.rank}: injected failure at step {step}", flush=True)
            os._exit(17)
        x = torch.randn(64, 256, device=device, generator=gen)
        y = torch.randint(0, 10, (64,), device=device, generator=gen)
        loss = nn.functional.cross_entropy(ddp(x), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    torch.cuda.synchronize(device)
    comm.check()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu().numpy().tobytes()
    print(f"rank {comm.rank}/{comm.world} attempt {attempt} device {device} buckets {len(ddp.buckets)} launches {comm.launches} "
          f"loss {loss.item():.4f} params sha256 {hashlib.sha256(flat).hexdigest()[:16]}", flush=True)
    comm.close()


if __name__ == "__main__":
    main()
