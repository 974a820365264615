"""One rank of a bare multi-process allreduce (no torchrun, no torch.distributed): env RANK / WORLD_SIZE / LOCAL_RANK /
B2_SHM_NAME as the local_cuda scheduler would set them.  Used by tools/ncu_multirank.sh to put ONE rank under ncu while its
peers run free: `--warm` un-profiled collectives, then one more, then every rank idles for `--hold` seconds so that a
profiler replaying rank 0's kernel never overlaps a peer's next collective."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchx_b200.ddp import Communicator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=float, default=25.04)
ap.add_argument("--algo", default="auto")
ap.add_argument("--warm", type=int, default=3)
ap.add_argument("--hold", type=float, default=20.0)
a = ap.parse_args()
rank = int(os.environ["RANK"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
comm = Communicator.from_env()
comm.set_timeout(120.0)
n = int(a.mib * (1 << 20) / 4) // 8 * 8
x = torch.randn(n, device="cuda", generator=torch.Generator("cuda").manual_seed(1234 + rank))
for _ in range(a.warm):
    comm.allreduce_(x, algo=a.algo)
torch.cuda.synchronize()
comm.barrier()
torch.cuda.synchronize()
comm.allreduce_(x, algo=a.algo)  # the profiled launch (ncu: --launch-skip counts this rank's earlier kernels)
torch.cuda.synchronize()
print(f"rank {rank}: done algo={comm.last_algo} launches={comm.launches}", flush=True)
time.sleep(a.hold)
comm.check()
comm.close()
