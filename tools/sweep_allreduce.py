"""Allreduce bandwidth sweep (BASELINE.json config #4): the fused B200 kernel vs the reference path's NCCL,
launched one process per GPU (torchrun or `torchx run -s local_cuda dist.ddp`).

For each message size S (bytes of the fp32 bucket) it times, device-side with CUDA events and max over ranks:
  ours_fused      b2_allreduce  fp32 bucket, bf16 wire, 1/W scale fused   (ONE launch)
  ref_hook_seq    buf.to(bf16).div_(W) -> dist.all_reduce(NCCL) -> buf.copy_()   (what bf16_compress_hook runs)
  nccl_bf16       dist.all_reduce on a bf16 tensor of the same element count (the wire-only part of the above)
  ours_f32 / nccl_f32   fp32 wire variants (DDP default semantics)
busbw = (wire bytes / t) * 2(W-1)/W, nccl-tests convention; wire bytes = 2N (bf16) or 4N (fp32).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from torchx_b200.ddp import Communicator  # noqa: E402


def time_op(fn, bufs, iters, warmup, stream):
    with torch.cuda.stream(stream):
        for i in range(warmup):
            fn(bufs[i % len(bufs)])
        stream.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(iters):
            fn(bufs[i % len(bufs)])
        e1.record(stream)
        stream.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3 / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-kib", type=int, default=4)
    ap.add_argument("--max-mib", type=int, default=1024)
    ap.add_argument("--ctas", default="0", help="comma list of max_ctas values to try for ours (0 = default)")
    ap.add_argument("--algos", default="auto")
    ap.add_argument("--extra-mib", default="7.82,30.04,25.04,25.32,9.27,27.04,168.27", help="DDP bucket sizes (SURVEY 8a)")
    ap.add_argument("--skip-nccl", action="store_true")
    ap.add_argument("--oneshot-max-mib", type=float, default=32.0, help="do not time the forced one-shot algorithm above this size")
    ap.add_argument("--out", default="")
    ap.add_argument("--trace", action="store_true", help="record the per-CTA phase breakdown of one launch per size/algo")
    a = ap.parse_args()

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    comm = Communicator.from_env(stage_mb=int(os.environ.get("B2_STAGE_MB", "0") or 0))
    stream = torch.cuda.Stream()
    sizes = []
    s = a.min_kib << 10
    while s <= a.max_mib << 20:
        sizes.append(s)
        s *= 2
    sizes += [int(float(m) * (1 << 20)) // 32 * 32 for m in a.extra_mib.split(",") if m]
    rows = []
    for S in sizes:
        n = S // 4
        nbuf = 2 if S >= (64 << 20) else min(16, max(2, (256 << 20) // S))
        g = torch.Generator(device="cuda").manual_seed(1234 + rank)
        bufs = [torch.randn(n, device="cuda", generator=g) for _ in range(nbuf)]
        iters = 20 if S >= (256 << 20) else (50 if S >= (16 << 20) else 200)
        warm = 5 if S >= (256 << 20) else 20
        row = {"bytes_fp32": S, "n": n, "world": world}
        k = 2.0 * (world - 1) / world

        # parity first (one fresh buffer): ours vs NCCL bf16 path
        x = bufs[0].clone()
        y = bufs[0].clone()
        with torch.cuda.stream(stream):
            comm.allreduce_(x, stream=stream)
            c = y.to(torch.bfloat16).div_(world)
            dist.all_reduce(c)
            y.copy_(c)
        stream.synchronize()
        comm.check()
        # NCCL's reduction order is its own for W >= 4, so compare against the magnitude of the data, not per element
        row["max_abs_diff_over_max_abs_vs_nccl_bf16"] = float(((x - y).abs().max() / y.abs().max().clamp_min(1e-30)).item()) if n else 0.0
        row["bit_equal_vs_nccl_bf16"] = bool(torch.equal(x, y))

        for ctas in [int(v) for v in a.ctas.split(",")]:
            comm.set_max_ctas(ctas)
            for algo in a.algos.split(","):
                if algo == "oneshot" and S > a.oneshot_max_mib * (1 << 20):
                    continue
                t = time_op(lambda b: comm.allreduce_(b, algo=algo, stream=stream), bufs, iters, warm, stream)
                key = f"ours_fused[{algo},ctas={ctas}]"
                row[key] = {"us": round(t * 1e6, 2), "busbw_gbs": round(2 * n / t * k / 1e9, 1), "hbm_alg_gbs": round(8 * n / t / 1e9, 1)}
                if a.trace and algo != "auto":
                    comm.trace(True)
                    dist.barrier()
                    comm.allreduce_(bufs[0], algo=algo, stream=stream)
                    stream.synchronize()
                    stamps = [st for st in comm.trace(False, read_ctas=296) if st[0]]
                    if stamps:
                        t0 = min(st[0] for st in stamps)
                        last = 3 if algo == "oneshot" else 5
                        names = {"oneshot": ["push", "bar1", "reduce"], "twoshot": ["scatter", "bar1", "reduce", "bar2", "gather"],
                                 "twoshot_pull": ["compress", "bar1", "pull_reduce", "bar2", "gather"]}[algo]
                        med = lambda xs: sorted(xs)[len(xs) // 2]  # noqa: E731
                        row[key]["trace_us"] = {
                            "ctas": len(stamps), "start_spread": round((max(st[0] for st in stamps) - t0) / 1e3, 2),
                            **{nm: round(med([st[i + 1] - st[i] for st in stamps]) / 1e3, 2) for i, nm in enumerate(names)},
                            "total": round((max(st[last] for st in stamps) - t0) / 1e3, 2),
                        }
        comm.set_max_ctas(0)
        t = time_op(lambda b: comm.allreduce_(b, wire="f32", stream=stream), bufs, iters, warm, stream)
        row["ours_f32"] = {"us": round(t * 1e6, 2), "busbw_gbs": round(4 * n / t * k / 1e9, 1)}
        if not a.skip_nccl:
            def hook_seq(b):
                c = b.to(torch.bfloat16).div_(world)
                dist.all_reduce(c)
                b.copy_(c)

            t = time_op(hook_seq, bufs, iters, warm, stream)
            row["ref_hook_seq"] = {"us": round(t * 1e6, 2), "busbw_gbs": round(2 * n / t * k / 1e9, 1)}
            hb = [b.to(torch.bfloat16) for b in bufs]
            t = time_op(lambda b: dist.all_reduce(b), hb, iters, warm, stream)
            row["nccl_bf16"] = {"us": round(t * 1e6, 2), "busbw_gbs": round(2 * n / t * k / 1e9, 1)}
            del hb
            t = time_op(lambda b: dist.all_reduce(b), bufs, iters, warm, stream)
            row["nccl_f32"] = {"us": round(t * 1e6, 2), "busbw_gbs": round(4 * n / t * k / 1e9, 1)}
        comm.check()
        if rank == 0:
            print(json.dumps(row), flush=True)
        rows.append(row)
        del bufs
        torch.cuda.empty_cache()
    if rank == 0 and a.out:
        with open(a.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
