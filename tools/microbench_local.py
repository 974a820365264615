"""Single-GPU microbenchmark of the fused cast/scale pass (b2_local_pass == the W=1 bucket kernel) against
the HBM roofline, with torch's own 2-launch equivalent beside it.  Prints one JSON line per size.

Algorithmic bytes per element: 8 (read fp32 once, write fp32 once) for f32 buckets; 4 for bf16 buckets.
Inputs rotate through enough buffers to exceed the 126 MB L2 between timed launches.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from torchx_b200.ddp import local_pass_  # noqa: E402


def timed(fn, bufs, iters):
    for b in bufs[: min(3, len(bufs))]:
        fn(b)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(iters):
        fn(bufs[i % len(bufs)])
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) * 1e-3 / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes-mib", default="1,8,30,64,168,512,1024")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--peak-gbs", type=float, default=0.0)
    a = ap.parse_args()
    peak = a.peak_gbs
    if not peak and os.path.exists("MEASURED_PEAKS.json"):
        peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 0.0)
    peak = peak or 6650.0
    dev = torch.device("cuda:0")
    for mib in [float(s) for s in a.sizes_mib.split(",")]:
        n = int(mib * (1 << 20) / 4)
        nbuf = max(2, int((512 << 20) / (n * 4)) + 1)
        nbuf = min(nbuf, 64)
        bufs = [torch.randn(n, device=dev) for _ in range(nbuf)]
        iters = max(10, min(a.iters, int(2e10 / (n * 8))))
        t_ours = timed(lambda b: local_pass_(b, scale=0.125), bufs, iters)
        t_copy = timed(lambda b: b.copy_(bufs[0] if b is not bufs[0] else bufs[1]), bufs, iters)

        def torch_seq(b):
            c = b.to(torch.bfloat16)
            c.div_(8)
            b.copy_(c)

        t_torch = timed(torch_seq, bufs, iters)
        print(json.dumps({
            "bench": "local_pass", "mib": mib, "n": n, "us": round(t_ours * 1e6, 2),
            "gbs": round(8 * n / t_ours / 1e9, 1), "frac_of_measured_peak": round(8 * n / t_ours / 1e9 / peak, 3),
            "torch_copy_gbs": round(8 * n / t_copy / 1e9, 1),
            "torch_cast_div_copy_us": round(t_torch * 1e6, 2), "speedup_vs_torch_seq": round(t_torch / t_ours, 2),
        }), flush=True)


if __name__ == "__main__":
    main()
