"""Probe of the NVSwitch in-switch reduction (multimem.ld_reduce) that the NVLS allreduce path relies on: which
summation order / accumulator does the switch use, and is it reproducible?  Runs an in-process world over the first W
GPUs (b2_comm_create_local: VMM arenas + one multicast object), feeds crafted bf16 inputs whose sum depends on the
order of fp32 additions, runs the NVLS allreduce `--reps` times and saves inputs and every rank's outputs to an .npz
for offline analysis (tools/nvls_order_fit.py).  Output of this tool is data, not a pass/fail.

    python tools/nvls_probe.py --world 8 --out gpurun_out/nvls_probe_w8.npz
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from torchx_b200.ddp import Communicator  # noqa: E402


def bf16_bits(sign, exp, mant):
    """sign in {0,1}, unbiased exponent, 7-bit mantissa -> bf16 bit pattern (normal numbers)."""
    return ((sign.astype(np.uint32) << 15) | (((exp + 127).astype(np.uint32) & 0xFF) << 7) | (mant.astype(np.uint32) & 0x7F)).astype(np.uint16)


def f32_to_bf16_exact(x):
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    assert np.all((b & 0xFFFF) == 0), "value is not bf16-representable"
    return (b >> 16).astype(np.uint16)


def craft(world, per_class, seed=0):
    rng = np.random.default_rng(seed)
    cols = []
    names = []

    def add(name, arr):  # arr: (world, k) uint16
        names.append((name, sum(c.shape[1] for c in cols), arr.shape[1]))
        cols.append(arr)

    k = per_class
    add("wide", bf16_bits(rng.integers(0, 2, (world, k)), rng.integers(-14, 15, (world, k)), rng.integers(0, 128, (world, k))))
    add("narrow", bf16_bits(rng.integers(0, 2, (world, k)), rng.integers(-2, 3, (world, k)), rng.integers(0, 128, (world, k))))
    # tie-breakers: {1, 2^-8, 2^-30} on three random ranks, zeros elsewhere.  Exact sum rounds UP to 1 + 2^-7; any fp32
    # sequence that adds 2^-30 after 1 has been accumulated loses it and rounds to even = 1.0.
    tie = np.zeros((world, k), dtype=np.uint16)
    vals = f32_to_bf16_exact([1.0, 2.0 ** -8, 2.0 ** -30])
    for i in range(k):
        pos = rng.permutation(world)[:3] if world >= 3 else np.arange(world)
        for p, v in zip(pos, vals):
            tie[p, i] = v
    add("tie", tie)
    # cancellation: big, -big, and small values whose survival depends on when the big pair cancels
    canc = bf16_bits(rng.integers(0, 2, (world, k)), rng.integers(-4, 1, (world, k)), rng.integers(0, 128, (world, k)))
    big = f32_to_bf16_exact([2.0 ** 20, -(2.0 ** 20)])
    for i in range(k):
        pos = rng.permutation(world)[:2]
        canc[pos[0], i], canc[pos[1], i] = big[0], big[1]
    add("cancel", canc)
    # subnormal bf16 inputs (exponent field 0) mixed with tiny normals
    sub = ((rng.integers(0, 2, (world, k)).astype(np.uint32) << 15) | rng.integers(1, 128, (world, k)).astype(np.uint32)).astype(np.uint16)
    mask = rng.random((world, k)) < 0.3
    sub[mask] = bf16_bits(rng.integers(0, 2, (world, k)), np.full((world, k), -126), rng.integers(0, 128, (world, k)))[mask]
    add("subnormal", sub)
    # specials: +-inf, nan, +-0, max finite
    table = np.array([0x7F80, 0xFF80, 0x7FC0, 0x0000, 0x8000, 0x7F7F, 0xFF7F, 0x3F80], dtype=np.uint16)
    add("special", table[rng.integers(0, len(table), (world, k))])
    return np.concatenate(cols, axis=1), names


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=0)
    ap.add_argument("--per-class", type=int, default=8192)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/nvls_probe.npz")
    ap.add_argument("--stage-mb", type=int, default=8)
    a = ap.parse_args()
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return main_multiprocess(a)
    W = a.world or torch.cuda.device_count()
    comms = Communicator.create_local(list(range(W)), stage_mb=a.stage_mb)
    print(f"world={W} caps={comms[0].caps} multicast={comms[0].has_multicast}", flush=True)
    if not comms[0].has_multicast:
        print("NO MULTICAST on this box: nothing to probe")
        np.savez(a.out, multicast=np.array([0]))
        return
    streams = [torch.cuda.Stream(device=d) for d in range(W)]
    for c in comms:
        c.set_timeout(20.0)
    bits, names = craft(W, a.per_class)
    n = bits.shape[1]
    outs = np.zeros((a.reps, W, n), dtype=np.uint16)
    for rep in range(a.reps):
        tens = [torch.from_numpy(bits[r].view(np.int16).copy()).to(f"cuda:{r}").view(torch.bfloat16) for r in range(W)]
        for r, (c, s) in enumerate(zip(comms, streams)):
            c.allreduce_(tens[r], scale=1.0, wire="bf16", algo="nvls", stream=s)
        for s in streams:
            s.synchronize()
        for c in comms:
            c.check()
        for r in range(W):
            outs[rep, r] = tens[r].view(torch.int16).cpu().numpy().view(np.uint16)
    # the same bf16 values held in fp32 buckets (mode F32_WIRE_BF16, scale 1): the DDP path's own configuration
    outs_fw = np.zeros((a.reps, W, n), dtype=np.uint16)
    as_f32 = (bits.astype(np.uint32) << 16).view(np.float32)
    for rep in range(a.reps):
        tens = [torch.from_numpy(as_f32[r].copy()).to(f"cuda:{r}") for r in range(W)]
        for r, (c, s) in enumerate(zip(comms, streams)):
            c.allreduce_(tens[r], scale=1.0, wire="bf16", algo="nvls", stream=s)
        for s in streams:
            s.synchronize()
        for r in range(W):
            outs_fw[rep, r] = (tens[r].cpu().numpy().view(np.uint32) >> 16).astype(np.uint16)
    # ... and through the rank-order P2P kernel, as the on-device reference for the same inputs
    outs_p2p = np.zeros((W, n), dtype=np.uint16)
    tens = [torch.from_numpy(bits[r].view(np.int16).copy()).to(f"cuda:{r}").view(torch.bfloat16) for r in range(W)]
    for r, (c, s) in enumerate(zip(comms, streams)):
        c.allreduce_(tens[r], scale=1.0, wire="bf16", algo="twoshot", stream=s)
    for s in streams:
        s.synchronize()
    for r in range(W):
        outs_p2p[r] = tens[r].view(torch.int16).cpu().numpy().view(np.uint16)
    # fp32-wire NVLS (multimem.ld_reduce.add.f32): wide-spread fp32 inputs
    rng = np.random.default_rng(1)
    f_in = (rng.standard_normal((W, a.per_class)) * np.exp2(rng.integers(-20, 21, (W, a.per_class)))).astype(np.float32)
    f_out = np.zeros((a.reps, W, a.per_class), dtype=np.float32)
    for rep in range(a.reps):
        tens = [torch.from_numpy(f_in[r].copy()).to(f"cuda:{r}") for r in range(W)]
        for r, (c, s) in enumerate(zip(comms, streams)):
            c.allreduce_(tens[r], scale=1.0, wire="f32", algo="nvls", stream=s)
        for s in streams:
            s.synchronize()
        for r in range(W):
            f_out[rep, r] = tens[r].cpu().numpy()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    np.savez_compressed(a.out, multicast=np.array([1]), world=np.array([W]), bf16_in=bits, bf16_out=outs, bf16_out_f32wire=outs_fw, bf16_out_p2p=outs_p2p,
                        f32_in=f_in, f32_out=f_out,
                        class_names=np.array([nm for nm, _, _ in names]), class_start=np.array([s for _, s, _ in names]),
                        class_len=np.array([ln for _, _, ln in names]))
    same_reps = bool(np.all(outs == outs[0]))
    same_ranks = bool(np.all(outs[:, 1:] == outs[:, :1]))
    print(f"nvls(bf16 bucket) == p2p rank-order: {int((outs[0, 0] != outs_p2p[0]).sum())} of {n} differ; nvls(f32 bucket, bf16 wire) vs p2p: "
          f"{int((outs_fw[0, 0] != outs_p2p[0]).sum())} differ", flush=True)
    print(f"saved {a.out}: n={n}; identical across reps: {same_reps}; identical across ranks: {same_ranks}; "
          f"f32 identical across reps: {bool(np.array_equal(f_out[0], f_out[-1], equal_nan=True))}", flush=True)
    for c in comms:
        c.close()


def main_multiprocess(a):
    """Under torchrun: one process per GPU (the production topology); rank r feeds row r of the same crafted inputs."""
    rank, W = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    comm = Communicator.from_env(stage_mb=a.stage_mb)
    comm.set_timeout(30.0)
    bits, names = craft(W, a.per_class)
    n = bits.shape[1]
    res = {}
    for tag, algo in (("nvls", "nvls"), ("p2p", "twoshot")):
        if algo == "nvls" and not comm.has_multicast:
            continue
        t = torch.from_numpy(bits[rank].view(np.int16).copy()).cuda().view(torch.bfloat16)
        comm.allreduce_(t, scale=1.0, wire="bf16", algo=algo)
        torch.cuda.synchronize()
        res[f"bf16_out_{tag}"] = t.view(torch.int16).cpu().numpy().view(np.uint16)
        f = torch.from_numpy((bits[rank].astype(np.uint32) << 16).view(np.float32).copy()).cuda()
        comm.allreduce_(f, scale=1.0, wire="bf16", algo=algo)
        torch.cuda.synchronize()
        res[f"f32wire_out_{tag}"] = (f.cpu().numpy().view(np.uint32) >> 16).astype(np.uint16)
    comm.check()
    if rank == 0:
        np.savez_compressed(a.out, world=np.array([W]), bf16_in=bits, class_names=np.array([nm for nm, _, _ in names]),
                            class_start=np.array([s for _, s, _ in names]), class_len=np.array([ln for _, _, ln in names]), **res)
        if "bf16_out_nvls" in res:
            print(f"multi-process W={W}: nvls(bf16) vs p2p differ in {int((res['bf16_out_nvls'] != res['bf16_out_p2p']).sum())} of {n}; "
                  f"nvls(f32 bucket) vs p2p differ in {int((res['f32wire_out_nvls'] != res['f32wire_out_p2p']).sum())}", flush=True)
    comm.barrier()
    torch.cuda.synchronize()
    comm.close()


if __name__ == "__main__":
    main()
