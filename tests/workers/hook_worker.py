"""One rank of the comm-hook parity check (spawned by tests/test_hook_multirank_gpu.py): the fused kernel delivered as a
DDP communication hook on STOCK torch.nn.parallel.DistributedDataParallel at W > 1 (the A/B parity mode),
next to the mini-DDP and, when the ranks have a GPU each, next to the reference's own bf16_compress_hook over NCCL.
torch.distributed is used for DDP's own bookkeeping only (gloo when ranks share a GPU - NCCL refuses that - else nccl)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch import nn  # noqa: E402


def mlp(seed):
    torch.manual_seed(seed)
    # 37- and 13-wide layers: biases / weights whose sizes are not multiples of 8, so later parameters start at
    # bucket offsets that are not vec-aligned (exercises the straddling-vec path of the zero-copy bucket fill)
    return nn.Sequential(nn.Linear(64, 256), nn.ReLU(), nn.Linear(256, 37), nn.ReLU(), nn.Linear(37, 13)).cuda()


def flat_grads(m):
    return torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu().numpy()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--device", type=int, required=True)
    ap.add_argument("--shm", required=True)
    ap.add_argument("--port", type=int, required=True)
    ap.add_argument("--backend", required=True)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()

    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
    from torch.nn.parallel import DistributedDataParallel as TorchDDP

    from torchx_b200.ddp import B200HookState, Communicator, DistributedDataParallel, b200_bf16_compress_hook

    torch.cuda.set_device(a.device)
    dist.init_process_group(a.backend, init_method=f"tcp://127.0.0.1:{a.port}", rank=a.rank, world_size=a.world)
    comm = Communicator.create(a.rank, a.world, a.device, a.shm, stage_mb=8, timeout_s=60)
    comm.set_timeout(30.0)
    comm.set_max_ctas(8)
    res = {}
    x = torch.randn(32, 64, device="cuda", generator=torch.Generator("cuda").manual_seed(100 + a.rank))

    twin = mlp(0)
    twin(x).square().mean().backward()
    res["local"] = flat_grads(twin)

    m1 = mlp(0)
    d1 = TorchDDP(m1, device_ids=[a.device], bucket_cap_mb=0.05)  # several buckets
    d1.register_comm_hook(B200HookState(comm), b200_bf16_compress_hook)
    for _ in range(2):  # second iteration runs on the rebuilt bucket layout
        d1.zero_grad(set_to_none=True)
        d1(x).square().mean().backward()
    torch.cuda.synchronize()
    comm.check()
    res["hook"] = flat_grads(m1)

    m2 = mlp(0)
    d2 = DistributedDataParallel(m2, comm, bucket_cap_mb=0.05, first_bucket_mb=0.01)
    for _ in range(2):
        d2.zero_grad(set_to_none=True)
        d2(x).square().mean().backward()
    torch.cuda.synchronize()
    comm.check()
    res["mini"] = flat_grads(m2)
    res["mini_gathered"] = np.array([d2.gathered_buckets, d2.copied_in_buckets])

    # channels_last conv weights (gradient views with channels_last strides), ragged sizes, tiny buckets: zero-copy bucket fill
    # (kernel reads the gradients in place) against the copy-in path, bit for bit
    def convnet():
        torch.manual_seed(0)
        net = nn.Sequential(nn.Conv2d(3, 5, 3, bias=True), nn.ReLU(), nn.Conv2d(5, 7, 3), nn.ReLU(), nn.Flatten(), nn.Linear(7 * 4 * 4, 11))
        return net.cuda().to(memory_format=torch.channels_last)

    xc = torch.randn(4, 3, 8, 8, device="cuda", generator=torch.Generator("cuda").manual_seed(200 + a.rank)).contiguous(memory_format=torch.channels_last)
    tw = convnet()
    tw(xc).square().mean().backward()
    res["conv_local"] = flat_grads(tw)
    for tag, zc in (("conv_zero_copy", True), ("conv_copy_in", False)):
        m4 = convnet()
        d4 = DistributedDataParallel(m4, comm, bucket_cap_mb=0.002, first_bucket_mb=0.0005, zero_copy=zc)
        for step in range(3):
            d4.zero_grad(set_to_none=(step != 1))
            d4(xc).square().mean().backward()
        torch.cuda.synchronize()
        comm.check()
        res[tag] = flat_grads(m4)
        res[tag + "_counts"] = np.array([d4.gathered_buckets, d4.copied_in_buckets])

    if a.backend == "nccl":  # the reference's own hook over NCCL: at W = 2 one fp32 add, one rounding => same bits as ours
        m3 = mlp(0)
        d3 = TorchDDP(m3, device_ids=[a.device], bucket_cap_mb=0.05)
        d3.register_comm_hook(None, default_hooks.bf16_compress_hook)
        for _ in range(2):
            d3.zero_grad(set_to_none=True)
            d3(x).square().mean().backward()
        torch.cuda.synchronize()
        res["nccl_hook"] = flat_grads(m3)
        eq = []
        for n in (1000, 65536, (1 << 20) + 3, 6_563_840):  # ... and raw buffers up to a 25 MiB bucket
            g = torch.Generator("cuda").manual_seed(1234 + a.rank)
            buf = torch.randn(n, device="cuda", generator=g)
            ours = buf.clone()
            comm.allreduce_(ours)
            c = buf.to(torch.bfloat16).div_(a.world)
            dist.all_reduce(c)
            ref = buf.clone().copy_(c)
            torch.cuda.synchronize()
            eq.append(bool(torch.equal(ours, ref)))
        res["nccl_bit_equal"] = np.array(eq)
    comm.check()
    np.savez(a.out, **res)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    print(f"rank {a.rank} ok", flush=True)


if __name__ == "__main__":
    main()
