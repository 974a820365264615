"""The comm hook on stock DistributedDataParallel at W > 1 (torch/nn/parallel/distributed.py:1987-2067), the mini-DDP with
its zero-copy bucket fill, and the reference's bf16_compress_hook over NCCL - one process per rank, compared bit for bit
with the CPU oracle on the ranks' local gradients."""
import os
import socket
import subprocess
import sys
import uuid

import numpy as np
import pytest

import oracle
from tests._util import assert_bits_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, devices, backend, tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    shm = f"/b2_hook_{uuid.uuid4().hex[:12]}"
    procs = []
    for r in range(world):
        cmd = [sys.executable, os.path.join(ROOT, "tests", "workers", "hook_worker.py"), "--rank", str(r), "--world", str(world),
               "--device", str(devices[r]), "--shm", shm, "--port", str(port), "--backend", backend, "--out", str(tmp_path / f"r{r}.npz")]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    try:
        for p in procs:
            o, _ = p.communicate(timeout=300)
            outs.append(o)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{outs[r]}"
    got = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    want = oracle.allreduce(oracle.B2O_F32_WIRE_BF16, [g["local"] for g in got], 1.0 / world)
    for r in range(world):
        assert_bits_equal(got[r]["hook"], want, f"comm hook on stock DDP, rank {r}")
        assert_bits_equal(got[r]["mini"], want, f"mini-DDP, rank {r}")
        gathered, copied = got[r]["mini_gathered"]
        assert gathered > 0 and copied == 0, (gathered, copied)  # every bucket was read in place, none copied in
    want_conv = oracle.allreduce(oracle.B2O_F32_WIRE_BF16, [g["conv_local"] for g in got], 1.0 / world)
    for r in range(world):
        assert_bits_equal(got[r]["conv_zero_copy"], want_conv, f"convnet, zero-copy bucket fill, rank {r}")
        assert_bits_equal(got[r]["conv_copy_in"], want_conv, f"convnet, copy-in, rank {r}")
        assert got[r]["conv_zero_copy_counts"][0] > 0 and got[r]["conv_copy_in_counts"][0] == 0
    return got


def test_hook_on_stock_ddp_two_ranks_one_gpu(tmp_path):
    """Runs on the 1-GPU box too: both ranks on cuda:0, gloo for DDP's bookkeeping."""
    _run(2, [0, 0], "gloo", tmp_path)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_hook_on_stock_ddp_one_gpu_per_rank_vs_nccl_hook(world, tmp_path, cuda_count):
    if cuda_count < world:
        pytest.skip(f"needs {world} GPUs")
    got = _run(world, list(range(world)), "nccl", tmp_path)
    if world == 2:
        # one fp32 add and one rounding: the reference's hook over NCCL and the fused kernel must agree bit for bit
        for r in range(world):
            assert_bits_equal(got[r]["nccl_hook"], got[r]["hook"], f"NCCL bf16_compress_hook vs ours, rank {r}")
            assert got[r]["nccl_bit_equal"].all(), got[r]["nccl_bit_equal"]
    else:
        # NCCL's own reduction order and bf16 partial sums: within (W-1) bf16 roundings of the fp32-accumulated value
        for r in range(world):
            a, b = got[r]["nccl_hook"].astype(np.float64), got[r]["hook"].astype(np.float64)
            scale = np.max(np.abs(b)) + 1e-30
            assert np.max(np.abs(a - b)) / scale < (world - 1) * 2.0 ** -8, np.max(np.abs(a - b)) / scale
