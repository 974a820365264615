"""Training job for the elastic re-launch measurement (BASELINE.json config #5: BERT-base dist.ddp --nproc 4 with a rank
drop): every rank prints one JSON line per completed step with a wall-clock stamp, and rank --fail-rank kills itself at
step --fail-at-step of the first attempt.

    torchx run -s local_cuda dist.ddp -j 1x4 --max_retries 1 --script examples/train_elastic.py -- --model bert --impl b200
    torchx run -s local_cwd  dist.ddp -j 1x4 --script examples/train_elastic.py -- --model bert --impl nccl     # reference path

--impl b200: torchx_b200 DistributedDataParallel on the peer-buffer fabric (no torch.distributed).
--impl nccl: what the reference's workers run - stock DistributedDataParallel + bf16_compress_hook over NCCL.
"""
import argparse
import json
import os
import sys
import time

T_IMPORT0 = time.time()
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def emit(**kw):
    print("ELASTIC " + json.dumps(kw), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="bert", choices=["bert", "mlp"])
    ap.add_argument("--impl", default="b200", choices=["b200", "nccl", "gloo"])
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--fail-rank", type=int, default=1)
    ap.add_argument("--fail-at-step", type=int, default=30)
    ap.add_argument("--no-fail", action="store_true", help="a manual re-submission: do not inject the failure again")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    attempt = int(os.environ.get("TORCHELASTIC_RESTART_COUNT", "0"))
    emit(event="start", rank=rank, attempt=attempt, t=time.time(), t_proc_start=T_IMPORT0)
    if a.impl == "b200":
        from torchx_b200.ddp import DistributedDataParallel
        from torchx_b200.distributed import communicator, init_pg

        device = init_pg("b200")
        comm = communicator()
    elif a.impl == "gloo":  # CPU plumbing check of this script and of tools/elastic_recover.py (no GPU)
        import torch.distributed as dist

        device = torch.device("cpu")
        dist.init_process_group("gloo")
        comm = None
    else:
        import torch.distributed as dist
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks

        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        device = torch.device("cuda", local_rank)
        torch.cuda.set_device(device)
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=device)
        comm = None
    emit(event="comm_ready", rank=rank, attempt=attempt, t=time.time())
    torch.manual_seed(0)
    if a.model == "bert":
        import transformers

        transformers.logging.set_verbosity_error()
        model = transformers.BertForMaskedLM(transformers.BertConfig()).to(device)
        vocab, seq = 30522, 512
    else:
        model = nn.Sequential(nn.Linear(256, 1024), nn.ReLU(), nn.Linear(1024, 1024), nn.ReLU(), nn.Linear(1024, 10)).to(device)
    if a.impl == "b200":
        ddp = DistributedDataParallel(model, comm)
    elif a.impl == "gloo":
        ddp = torch.nn.parallel.DistributedDataParallel(model)
    else:
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index])
        ddp.register_comm_hook(None, default_hooks.bf16_compress_hook)
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4)
    gen = torch.Generator(device=device).manual_seed(99 + rank)
    sync = (lambda: torch.cuda.synchronize(device)) if device.type == "cuda" else (lambda: None)
    emit(event="model_ready", rank=rank, attempt=attempt, t=time.time())
    for step in range(a.steps):
        if attempt == 0 and not a.no_fail and rank == a.fail_rank and step == a.fail_at_step:
            emit(event="kill", rank=rank, attempt=attempt, step=step, t=time.time())
            os._exit(17)
        if a.model == "bert":
            x = torch.randint(0, vocab, (a.batch, seq), device=device, generator=gen)
            y = torch.randint(0, vocab, (a.batch, seq), device=device, generator=gen)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = ddp(input_ids=x, labels=y).loss
        else:
            x = torch.randn(64, 256, device=device, generator=gen)
            y = torch.randint(0, 10, (64,), device=device, generator=gen)
            loss = nn.functional.cross_entropy(ddp(x), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        sync()
        emit(event="step", rank=rank, attempt=attempt, step=step, t=time.time(), loss=round(float(loss.item()), 4))
    if comm is not None:
        comm.check()
        comm.close()
    else:
        import torch.distributed as dist

        dist.destroy_process_group()
    emit(event="done", rank=rank, attempt=attempt, t=time.time())


if __name__ == "__main__":
    main()
