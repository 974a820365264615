"""Generates the committed golden fixtures by RUNNING THE REFERENCE in the build container.

  ddp_w{2,4}.npz   stock torch DistributedDataParallel (no hook / allreduce_hook / bf16_compress_hook) launched
                   through the reference's own launcher:  PYTHONPATH=/root/reference python -m torchx.cli.main
                   run -s local_cwd /root/reference/torchx/components/dist.py:ddp -j 1xW --script <worker>
                   (gloo/CPU - there is no GPU here).  Holds every rank's local
                   gradients and the three averaged results, so the oracle can be pinned against them.
  bucket_layouts.json   torch's own dist._compute_bucket_assignment_by_size over ResNet-50 / GPT-2-small /
                   BERT-base parameters (reverse order, limits [1 MiB, 25 MiB]) - layout parity fixture.

Run from the repo root:  python tests/golden/make_golden.py      (needs /root/reference; not needed at test time)
"""
import json
import os
import subprocess
import sys
import tempfile
import textwrap

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

WORKER = textwrap.dedent(
    '''
    import os, sys, argparse
    import numpy as np
    import torch, torch.nn as nn, torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks

    ap = argparse.ArgumentParser(); ap.add_argument("--out"); a = ap.parse_args()
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    def model():
        torch.manual_seed(0)
        return nn.Sequential(nn.Linear(64, 128), nn.ReLU(), nn.Linear(128, 16))

    torch.manual_seed(100 + rank)
    x = torch.randn(32, 64)

    def flat_grads(m):
        return torch.cat([p.grad.reshape(-1) for p in m.parameters()])

    m = model(); m(x).sum().backward(); local = flat_grads(m).clone()
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)

    res = {}
    for name, hook in (("none", None), ("allreduce", default_hooks.allreduce_hook), ("bf16_compress", default_hooks.bf16_compress_hook)):
        d = DDP(model())
        if hook is not None:
            d.register_comm_hook(None, hook)
        d(x).sum().backward()
        res[name] = flat_grads(d.module).clone()
        dist.barrier()
    outs = {}
    for k, v in res.items():
        g = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(g, v)
        outs[k] = torch.stack(g).numpy()
    if rank == 0:
        np.savez(a.out, local=torch.stack(gathered).numpy(), **{"ddp_" + k: v for k, v in outs.items()})
    dist.barrier()
    dist.destroy_process_group()
    '''
)


def run_reference_ddp(world: int) -> None:
    out = os.path.join(HERE, f"ddp_w{world}.npz")
    with tempfile.TemporaryDirectory() as td:
        script = os.path.join(td, "golden_worker.py")
        with open(script, "w") as f:
            f.write(WORKER)
        env = dict(os.environ, PYTHONPATH=REF)
        cmd = [sys.executable, "-m", "torchx.cli.main", "run", "-s", "local_cwd", f"{REF}/torchx/components/dist.py:ddp",
               "-j", f"1x{world}", "--script", script, "--", "--out", out]
        subprocess.run(cmd, check=True, cwd=td, env=env)
    assert os.path.exists(out), out
    print("wrote", out)


def bucket_layouts() -> None:
    import torch
    import torch.distributed as dist

    def layout(params):
        rev = list(reversed(range(len(params))))
        idx, _ = dist._compute_bucket_assignment_by_size([params[i] for i in rev], [1024 * 1024, 25 * 1024 * 1024], [False] * len(rev), rev)
        return idx

    out = {}
    import torchvision

    models = {"resnet50": lambda: torchvision.models.resnet50()}
    try:
        from transformers import BertConfig, BertForMaskedLM, GPT2Config, GPT2LMHeadModel

        models["gpt2_small"] = lambda: GPT2LMHeadModel(GPT2Config())
        models["bert_base"] = lambda: BertForMaskedLM(BertConfig())
    except Exception as e:  # pragma: no cover
        print("transformers unavailable:", e)
    for name, ctor in models.items():
        with torch.device("meta"):
            m = ctor()
        params = [p for p in m.parameters() if p.requires_grad]
        real = [torch.empty(p.shape, dtype=torch.float32, device="meta") for p in params]
        out[name] = {
            "param_numel": [p.numel() for p in params],
            "buckets_fp32": layout(real),
            "buckets_bf16": layout([torch.empty(p.shape, dtype=torch.bfloat16, device="meta") for p in params]),
        }
        print(name, "params", len(params), "numel", sum(out[name]["param_numel"]), "buckets", len(out[name]["buckets_fp32"]))
    with open(os.path.join(HERE, "bucket_layouts.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    for w in (2, 4):
        run_reference_ddp(w)
    bucket_layouts()
