#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): images/sec of ResNet-50 DDP training on N B200s of one box.

    python bench.py --gpus 1 --steps K --warmup W                      # N=1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                         # N>1, one rank per GPU
    ... --impl reference                                                # the reference path, same workload

A step = one optimizer step on a synthetic ImageNet-shaped batch (B=256/GPU, bf16 autocast, channels_last, SGD
momentum) of random-init torchvision ResNet-50: forward, backward with the DDP gradient-
bucket averaging (5 buckets, 97.5 MiB fp32) overlapped, optimizer step.

  default arm   torchx_b200.ddp.DistributedDataParallel: every bucket averaged by ONE fused sm_100a kernel
                (libb200ddp.so) over NVSwitch peer buffers; no torch.distributed / NCCL anywhere.
  reference arm what `torchx run -s local_cwd dist.ddp` workers run: stock torch DistributedDataParallel over NCCL
                with bf16_compress_hook (cast, div, ncclAllReduce, copy per bucket).

`value` is measured with the batch resident in HBM; `e2e` through the public API with the batch coming from pinned
host memory every step (H2D inside the timed region, overlapped on a copy stream) and the loss read back (D2H).
Both are device-timed with CUDA events, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--wire", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", default=True, help="(N>1) also report allreduce bus GB/s at the model's bucket sizes and 256 MiB")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(pw) if pw else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def build_model(name: str, device):
    import torch
    import torchvision

    torch.manual_seed(0)  # identical init on every rank (then rank-0 broadcast, as DDP does)
    if name != "resnet50":
        raise SystemExit(f"unknown model {name}")
    m = torchvision.models.resnet50()
    return m.to(device).to(memory_format=torch.channels_last)


def synthetic_batches(batch: int, rank: int, count: int, pinned: bool):
    import torch

    g = torch.Generator().manual_seed(1000 + rank)
    out = []
    for _ in range(count):
        x = torch.randn(batch, 3, 224, 224, generator=g)
        y = torch.randint(0, 1000, (batch,), generator=g)
        if pinned:
            x, y = x.pin_memory(), y.pin_memory()
        out.append((x, y))
    return out


class Trainer:
    """The user-level training loop, identical for both arms apart from how the model is wrapped."""

    def __init__(self, args, rank, world, local_rank):
        import torch

        self.torch = torch
        self.args, self.rank, self.world = args, rank, world
        self.device = torch.device("cuda", local_rank)
        torch.cuda.set_device(self.device)
        torch.backends.cudnn.benchmark = True
        self.comm = None
        model = build_model(args.model, self.device)
        if args.impl == "b200":
            from torchx_b200.ddp import Communicator, DistributedDataParallel

            self.comm = Communicator.from_env()
            self.ddp = DistributedDataParallel(model, self.comm, wire=args.wire)
        else:
            import torch.distributed as dist
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            from torch.nn.parallel import DistributedDataParallel as TorchDDP

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if "MASTER_PORT" not in os.environ:  # plain `python bench.py --impl reference` (N=1): any free port
                import socket

                with socket.socket() as sock:
                    sock.bind(("127.0.0.1", 0))
                    os.environ["MASTER_PORT"] = str(sock.getsockname()[1])
            # keep stdout to the ONE JSON line: this image's NCCL otherwise prints "NCCL version ..." there
            if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=self.device)
            self.ddp = TorchDDP(model, device_ids=[local_rank])
            if args.wire == "bf16":
                self.ddp.register_comm_hook(None, default_hooks.bf16_compress_hook)
        self.opt = torch.optim.SGD(self.ddp.parameters(), lr=0.1, momentum=0.9)
        self.loss_fn = torch.nn.CrossEntropyLoss()
        self.copy_stream = torch.cuda.Stream(device=self.device)

    # -- collective helpers that work for both arms ---------------------------------------------------------
    def barrier(self):
        torch = self.torch
        if self.comm is not None:
            self.comm.barrier()
        elif self.world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, value: float) -> float:
        torch = self.torch
        if self.world == 1:
            return value
        t = torch.zeros(self.world, device=self.device)
        t[self.rank] = value
        if self.comm is not None:
            self.comm.allreduce_(t, scale=1.0, wire="f32")
        else:
            import torch.distributed as dist

            dist.all_reduce(t)
        torch.cuda.synchronize()
        return float(t.max().item())

    # -- one optimizer step -----------------------------------------------------------------------------------
    def step(self, x, y):
        torch = self.torch
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = self.loss_fn(self.ddp(x), y)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss

    def run_resident(self, steps, warmup):
        """`value`: batch already in HBM."""
        torch = self.torch
        (xh, yh), = synthetic_batches(self.args.batch, self.rank, 1, pinned=False)
        x = xh.to(self.device).contiguous(memory_format=torch.channels_last)
        y = yh.to(self.device)
        for _ in range(warmup):
            self.step(x, y)
        self.barrier()
        launches0 = self.comm.launches if self.comm is not None else 0
        prof = getattr(self.ddp, "start_profile", None)
        if prof:
            prof()
        sampler = ClockSampler(self.device.index)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        profiled = bool(os.environ.get("BENCH_CUDA_PROFILER"))  # ncu --profile-from-start off: capture the timed region only
        if profiled:
            torch.cuda.profiler.start()
        e0.record()
        for _ in range(steps):
            loss = self.step(x, y)
        e1.record()
        self.barrier()
        if profiled:
            torch.cuda.profiler.stop()
        clocks = sampler.stop()
        elapsed = e0.elapsed_time(e1) * 1e-3
        launches = (self.comm.launches - launches0) if self.comm is not None else 0
        kernel = self.ddp.stop_profile() if prof else None
        return elapsed, float(loss.item()), clocks, launches, kernel

    def run_e2e(self, steps, warmup):
        """`e2e`: every step's batch is copied from pinned host memory inside the timed region (double-buffered on
        a copy stream) and the loss is read back to the host."""
        torch = self.torch
        host = synthetic_batches(self.args.batch, self.rank, 2, pinned=True)
        dev = [(torch.empty_like(xh, device=self.device), torch.empty_like(yh, device=self.device)) for xh, yh in host]
        ready = [torch.cuda.Event() for _ in host]
        consumed = [torch.cuda.Event() for _ in host]
        loss_host = torch.zeros(steps + warmup, pin_memory=True)

        def prefetch(i):
            k = i % 2
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(consumed[k])
                dev[k][0].copy_(host[k][0], non_blocking=True)
                dev[k][1].copy_(host[k][1], non_blocking=True)
                ready[k].record(self.copy_stream)

        def one(i):
            k = i % 2
            torch.cuda.current_stream().wait_event(ready[k])
            x = dev[k][0].contiguous(memory_format=torch.channels_last)
            y = dev[k][1].clone()
            consumed[k].record()  # the landing buffers may be overwritten by the next prefetch from here on
            loss = self.step(x, y)
            loss_host[i].copy_(loss.detach(), non_blocking=True)

        for k in range(2):
            consumed[k].record()
        total = steps + warmup
        prefetch(0)
        for i in range(warmup):
            prefetch(i + 1)
            one(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(warmup, total):
            if i + 1 < total:
                prefetch(i + 1)
            one(i)
        e1.record()
        self.barrier()
        elapsed = e0.elapsed_time(e1) * 1e-3
        h2d = host[0][0].numel() * host[0][0].element_size() + host[0][1].numel() * host[0][1].element_size()
        return elapsed, h2d, loss_host.element_size(), float(loss_host[-1].item())

    def allreduce_points(self):
        """Second half of BASELINE.json's metric: allreduce bus GB/s (fraction of 900 GB/s/dir) on this arm's data path, for
        fp32 buckets of the model's sizes and 256 MiB.  ours: one fused launch (bf16 wire, 1/W).  reference: the
        bf16_compress_hook sequence (cast, div, ncclAllReduce, copy).  busbw = wire bytes / t * 2(W-1)/W, wire = 2 B/element."""
        torch = self.torch
        sizes_mib = [7.82, 25.04, 30.04, 256.0]
        out = []
        stream = torch.cuda.Stream(device=self.device)
        k = 2.0 * (self.world - 1) / self.world
        for mib in sizes_mib:
            n = int(mib * (1 << 20) / 4) // 8 * 8
            bufs = [torch.randn(n, device=self.device) for _ in range(2 if mib >= 64 else 6)]
            iters, warm = (20, 5) if mib >= 64 else (100, 20)
            if self.comm is not None:
                def op(b):
                    self.comm.allreduce_(b, stream=stream)
            else:
                import torch.distributed as dist

                def op(b):
                    c = b.to(torch.bfloat16).div_(self.world)
                    dist.all_reduce(c)
                    b.copy_(c)
            with torch.cuda.stream(stream):
                for i in range(warm):
                    op(bufs[i % len(bufs)])
            stream.synchronize()
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                for i in range(iters):
                    op(bufs[i % len(bufs)])
                e1.record(stream)
            stream.synchronize()
            t = self.max_over_ranks(e0.elapsed_time(e1) * 1e-3 / iters)
            out.append({"bucket_mib_fp32": mib, "us": round(t * 1e6, 2), "busbw_gbs": round(2 * n / t * k / 1e9, 1),
                        "frac_of_900": round(2 * n / t * k / 1e9 / 900.0, 4)})
            del bufs
        return out

    def close(self):
        if self.comm is not None:
            self.comm.close()
        else:
            import torch.distributed as dist

            if dist.is_initialized():
                dist.destroy_process_group()


def cpu_baseline(world: int, batch: int, bucket_numels):
    """The oracle (CPU port of the bucket averaging, oracle/allreduce_oracle.c) timed on the host: one pass over
    ResNet-50's real buckets for `world` ranks, single thread.  A reported baseline, not a target."""
    import numpy as np

    import oracle

    rng = np.random.default_rng(0)
    n_total = int(sum(bucket_numels))
    w = max(world, 1)
    ins = [[rng.standard_normal(n, dtype=np.float32) for n in bucket_numels] for _ in range(w)]
    t0 = time.perf_counter()
    reps = 0
    while True:
        for b in range(len(bucket_numels)):
            oracle.allreduce(oracle.B2O_F32_WIRE_BF16, [ins[r][b] for r in range(w)], 1.0 / w)
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 20:
            break
    dt = (time.perf_counter() - t0) / reps
    return {
        "value": round(batch * w / dt, 1),
        "unit": "images/sec if a step were only the CPU gradient averaging (B*W images per averaging pass)",
        "cores": 1,
        "kind": "port",
        "sample": f"{reps} passes over ResNet-50's {len(bucket_numels)} gradient buckets ({n_total} fp32 elements) x {w} ranks, "
                  f"oracle/allreduce_oracle.c single-threaded; {round(8 * n_total / dt / 1e9, 2)} GB/s algorithmic",
        "host_cpus": os.cpu_count(),
    }


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus if "RANK" in os.environ else 1)))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")

    tr = Trainer(args, rank, world, local_rank)
    elapsed, last_loss, clocks, launches, kernel = tr.run_resident(args.steps, args.warmup)
    elapsed = tr.max_over_ranks(elapsed)
    images = args.batch * world * args.steps
    line = {
        "metric": "images/sec (max over ranks) ResNet-50 DDP",
        "value": round(images / elapsed, 1),
        "unit": "images/sec",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": round(elapsed / args.steps * 1e3, 3),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "impl": args.impl,
        "config": {
            "workload": "ResNet-50 bf16 dist.ddp training step (BASELINE.json configs[1]), B=%d/GPU, 3x224x224, channels_last, SGD momentum" % args.batch,
            "global_batch": args.batch * world,
            "parallelism": f"dp{world}",
            "precision": "bf16 autocast compute, fp32 master weights and gradient buckets, %s on the wire" % args.wire,
            "l2": "inputs larger than L2: 154 MB batch + 97.5 MiB of gradient buckets per step exceed the 126 MB L2",
            "gradient_buckets_mib": getattr(tr.ddp, "bucket_sizes_mib", lambda: None)(),
        },
        "clocks": clocks,
        "gpu_launches": launches,
        "last_loss": round(last_loss, 4),
    }
    if not args.no_e2e:
        e_elapsed, h2d, d2h, _ = tr.run_e2e(args.steps, args.warmup)
        e_elapsed = tr.max_over_ranks(e_elapsed)
        line["e2e"] = {"value": round(images / e_elapsed, 1), "unit": "images/sec", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": round(e_elapsed / args.steps * 1e3, 3)}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    if kernel is not None:
        line["roofline"] = {
            "bound": "hbm", "achieved": round(kernel["alg_bytes"] / kernel["seconds"] / 1e9, 1), "peak": peak, "unit": "GB/s",
            "frac": round(kernel["alg_bytes"] / kernel["seconds"] / 1e9 / peak, 4), "traffic": None,
            "kernel": kernel["name"], "launches_timed": kernel["launches"], "avg_us": round(kernel["seconds"] / max(kernel["launches"], 1) * 1e6, 2),
            "alg_bytes_per_launch": kernel["alg_bytes"] // max(kernel["launches"], 1), "peak_source": peak_src,
            "note": "8 B per gradient element (read fp32 once, write fp32 once), timed with CUDA events on the comm stream while backward runs beside it",
        }
        if world > 1:
            k = 2.0 * (world - 1) / world
            line["roofline"]["nvlink_busbw_gbs"] = round(kernel["alg_bytes"] / 4 / kernel["seconds"] * k / 1e9, 1)
            line["roofline"]["nvlink_frac_of_900"] = round(kernel["alg_bytes"] / 4 / kernel["seconds"] * k / 1e9 / 900.0, 4)
    if kernel is not None and world == 1:
        tpath = os.path.join(ROOT, "profiles", "r01_local_pass_traffic.json")
        if os.path.exists(tpath):
            t = json.load(open(tpath))
            # per launch, like `achieved`: scaled from the profiled 30 MiB launch to this run's mean bucket
            per_launch = line["roofline"]["alg_bytes_per_launch"]
            line["roofline"]["traffic"] = int((t["dram_bytes_read"] + t["dram_bytes_write"]) * per_launch / t["alg_bytes"])
            line["roofline"]["traffic_source"] = t["source"] + "; " + t["note"]
    if args.sweep and world > 1:
        line["allreduce"] = tr.allreduce_points()
    if args.impl == "reference":
        line["cpu_baseline"] = {"value": line["value"], "unit": "images/sec", "cores": os.cpu_count(), "kind": "reference",
                                "sample": "stock torch DistributedDataParallel + NCCL (bf16_compress_hook) - the path `torchx run -s local_cwd dist.ddp` "
                                          "workers execute; launcher and agents on the host cores, gradients on the GPUs"}
        line.setdefault("e2e", {"value": line["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    elif rank == 0 and world == 1 and not args.no_cpu_baseline:
        nums = [int(m * (1 << 20) / 4) for m in (line["config"]["gradient_buckets_mib"] or [97.5])]
        line["cpu_baseline"] = cpu_baseline(world, args.batch, nums)
    tr.close()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
