"""Allreduce bandwidth sweep (BASELINE.json config #4): the fused B200 kernels vs the reference path's NCCL,
launched one process per GPU (torchrun or `torchx run -s local_cuda dist.ddp`).

For each message size S (bytes of the fp32 bucket) it times, device-side with CUDA events and max over ranks:
  ours[<variant>]  b2_allreduce  fp32 bucket, bf16 wire, 1/W scale fused   (ONE launch); a variant is
                   algo[:ctas=N][:chunk=KiB], e.g. "auto", "nvls:ctas=128:chunk=1024", "twoshot_pipe:chunk=4096"
  ref_hook_seq     buf.to(bf16).div_(W) -> dist.all_reduce(NCCL) -> buf.copy_()   (what bf16_compress_hook runs)
  nccl_bf16        dist.all_reduce on a bf16 tensor of the same element count (the wire-only part of the above)
  ours_f32 / nccl_f32   fp32 wire variants (DDP default semantics)
busbw = (wire bytes / t) * 2(W-1)/W, nccl-tests convention; wire bytes = 2N (bf16) or 4N (fp32).

--nvlink-counters: rank 0 reads `nvidia-smi nvlink -gt d` before and after one timed loop per size and reports the
NVLink bytes its GPU sent / received per collective (hardware counters, not a model).
"""
import argparse
import json
import os
import re
import subprocess
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


def nvlink_kib(gpu: int):
    """(tx KiB, rx KiB) summed over the GPU's links from `nvidia-smi nvlink -gt d`, or None."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return None
    tx = sum(int(m) for m in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(m) for m in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    return (tx, rx) if (tx or rx) else None


def parse_variant(v):
    parts = v.split(":")
    d = {"algo": parts[0], "ctas": 0, "chunk": 0}
    for p in parts[1:]:
        k, val = p.split("=")
        d[k] = int(val)
    return d


TRACE_NAMES = {
    "twoshot_ll": (["flow_control", "scatter_issue", "reduce_push", "(unused)", "widen"], 5),
    "oneshot": (["push", "bar1", "reduce"], 3),
    "twoshot": (["scatter", "bar1", "reduce", "bar2", "gather"], 5),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-kib", type=int, default=4)
    ap.add_argument("--max-mib", type=int, default=1024)
    ap.add_argument("--sizes-mib", default="", help="explicit comma list of fp32 bucket sizes in MiB (replaces the power-of-two ladder)")
    ap.add_argument("--variants", default="auto", help="';'-separated list of algo[:ctas=N][:chunk=KiB]")
    ap.add_argument("--extra-mib", default="7.82,30.04,25.04,25.32,9.27,27.04,168.27", help="DDP bucket sizes (SURVEY 8a)")
    ap.add_argument("--skip-nccl", action="store_true")
    ap.add_argument("--skip-f32", action="store_true")
    ap.add_argument("--oneshot-max-mib", type=float, default=32.0, help="do not time the forced one-shot algorithm above this size")
    ap.add_argument("--out", default="")
    ap.add_argument("--trace", action="store_true", help="record the per-CTA phase breakdown of one launch per size/variant")
    ap.add_argument("--nvlink-counters", action="store_true")
    ap.add_argument("--check-variants", action="store_true", help="compare EVERY variant's result with the NCCL hook sequence (count / first index of differing elements)")
    a = ap.parse_args()

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    comm = Communicator.from_env(stage_mb=int(os.environ.get("B2_STAGE_MB", "0") or 0))
    stream = torch.cuda.Stream()
    if a.sizes_mib:
        sizes = [int(float(m) * (1 << 20)) // 32 * 32 for m in a.sizes_mib.split(",") if m]
    else:
        sizes = []
        s = a.min_kib << 10
        while s <= a.max_mib << 20:
            sizes.append(s)
            s *= 2
        sizes += [int(float(m) * (1 << 20)) // 32 * 32 for m in a.extra_mib.split(",") if m]
    variants = [parse_variant(v) for v in a.variants.split(";") if v]
    if rank == 0:
        print(json.dumps({"world": world, "caps": comm.caps, "has_multicast": comm.has_multicast, "variants": a.variants}), flush=True)
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

        # parity first (one fresh buffer): ours (AUTO) vs NCCL bf16 path.  `pristine` keeps the un-reduced input: the timing
        # loops below allreduce `bufs` in place over and over (after the first pass every rank holds the same values).
        pristine = bufs[0].clone()
        x = pristine.clone()
        y = pristine.clone()
        torch.cuda.synchronize()
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

        for v in variants:
            algo = v["algo"]
            if algo == "oneshot" and S > a.oneshot_max_mib * (1 << 20):
                continue
            if algo == "nvls" and not comm.has_multicast:
                continue
            comm.set_max_ctas(v["ctas"])
            comm.set_param("pipe_chunk_bytes", (v["chunk"] or 2048) << 10)
            key = f"ours[{algo},ctas={v['ctas']},chunk={v['chunk']}]"
            chk = None
            if a.check_variants:
                xv = pristine.clone()
                torch.cuda.synchronize()
                with torch.cuda.stream(stream):
                    comm.allreduce_(xv, algo=algo, stream=stream)
                stream.synchronize()
                comm.check()
                d = (xv.view(torch.int32) != y.view(torch.int32)).nonzero()
                chk = {"n_diff": int(d.numel()), "first_diff": int(d[0].item()) if d.numel() else -1, "last_diff": int(d[-1].item()) if d.numel() else -1,
                       "max_abs_diff_over_max_abs": float(((xv - y).abs().max() / y.abs().max().clamp_min(1e-30)).item()), "ran": comm.last_algo}
                del xv, d
            nv0 = nvlink_kib(local_rank) if (a.nvlink_counters and rank == 0) else None
            t = time_op(lambda b: comm.allreduce_(b, algo=algo, stream=stream), bufs, iters, warm, stream)
            row[key] = {"us": round(t * 1e6, 2), "busbw_gbs": round(2 * n / t * k / 1e9, 1), "frac_of_900": round(2 * n / t * k / 1e9 / 900, 4),
                        "hbm_alg_gbs": round(8 * n / t / 1e9, 1)}
            if chk is not None:
                row[key]["vs_nccl_hook"] = chk
            if nv0 is not None:
                nv1 = nvlink_kib(local_rank)
                if nv1 is not None:
                    ops = iters + warm
                    row[key]["nvlink_tx_bytes_per_op"] = round((nv1[0] - nv0[0]) * 1024 / ops)
                    row[key]["nvlink_rx_bytes_per_op"] = round((nv1[1] - nv0[1]) * 1024 / ops)
                    row[key]["wire_bytes_S"] = 2 * n
            if a.trace and algo != "auto":
                comm.trace(True)
                dist.barrier()
                comm.allreduce_(bufs[0], algo=algo, stream=stream)
                stream.synchronize()
                stamps = [st for st in comm.trace(False, read_ctas=296) if st[0]]
                if stamps:
                    t0 = min(st[0] for st in stamps)
                    med = lambda xs: sorted(xs)[len(xs) // 2]  # noqa: E731
                    if algo in TRACE_NAMES:
                        names, last = TRACE_NAMES[algo]
                        row[key]["trace_us"] = {
                            "ctas": len(stamps), "start_spread": round((max(st[0] for st in stamps) - t0) / 1e3, 2),
                            **{nm: round(med([st[i + 1] - st[i] for st in stamps]) / 1e3, 2) for i, nm in enumerate(names)},
                            "total": round((max(st[last] for st in stamps) - t0) / 1e3, 2),
                        }
                    else:  # pipelined kernels: one stamp per role group, all relative to the earliest CTA start
                        rel = lambda i: round(med([st[i] - t0 for st in stamps if st[i]]) / 1e3, 2)  # noqa: E731
                        row[key]["trace_us"] = {"ctas": len(stamps), "roleA_done": rel(1), "roleB_first_wait_passed": rel(2), "roleB_done": rel(3),
                                                "roleC_first_wait_passed": rel(4), "roleC_done": rel(5),
                                                "total": round((max(st[5] for st in stamps) - t0) / 1e3, 2)}
        comm.set_max_ctas(0)
        comm.set_param("pipe_chunk_bytes", 2048 << 10)
        if not a.skip_f32:
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
            if not a.skip_f32:
                t = time_op(lambda b: dist.all_reduce(b), bufs, iters, warm, stream)
                row["nccl_f32"] = {"us": round(t * 1e6, 2), "busbw_gbs": round(4 * n / t * k / 1e9, 1)}
        comm.check()
        if rank == 0:
            print(json.dumps(row), flush=True)
        rows.append(row)
        del bufs, pristine, x, y
        torch.cuda.empty_cache()
    if rank == 0 and a.out:
        with open(a.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
