"""Parity of the CUDA hot path (through the C ABI) against the CPU oracle.  All ranks of a topology live
in this process on one or more devices (b2_comm_create_local), each launching on its own stream - the same
kernels, flags and staging layout as the multi-process path, minus CUDA IPC (covered in test_ipc_gpu.py).
"""
import numpy as np
import pytest
import torch

import oracle
from tests._util import assert_bits_equal, make_inputs

pytestmark = pytest.mark.gpu

MODES = {"f32_wire_bf16": oracle.B2O_F32_WIRE_BF16, "f32": oracle.B2O_F32, "bf16": oracle.B2O_BF16}
WIRE = {"f32_wire_bf16": "bf16", "f32": "f32", "bf16": "bf16"}
SIZES = [1, 7, 8, 9, 1023, 1024, 1025, 4099, 32771, (1 << 18) + 5]


def _devices(world, cuda_count, spread):
    if spread:
        if cuda_count < world:
            pytest.skip(f"needs {world} GPUs")
        return list(range(world))
    return [0] * world


class World:
    def __init__(self, devices, stage_mb=8, timeout_s=10.0):
        from torchx_b200.ddp import Communicator

        self.comms = Communicator.create_local(devices, stage_mb=stage_mb)
        self.streams = [torch.cuda.Stream(device=d) for d in devices]
        same_device = len(set(devices)) == 1
        for c in self.comms:
            c.set_timeout(timeout_s)
            if same_device:  # all kernels must be co-resident on one GPU: W ranks x 2 lanes x grid <= #SMs (1 CTA per SM)
                c.set_max_ctas(max(1, 64 // len(devices)))

    def run(self, fn):
        """fn(rank, comm, stream) launches that rank's work; then wait for all and check health."""
        for r, (c, s) in enumerate(zip(self.comms, self.streams)):
            fn(r, c, s)
        for s in self.streams:
            s.synchronize()
        for c in self.comms:
            c.check()

    def close(self):
        for c in self.comms:
            c.close()


def _to_dev(x, mode, device):
    if mode == "bf16":
        bits = oracle.f32_to_bf16_bits(x)
        return torch.from_numpy(bits.view(np.int16).copy()).to(f"cuda:{device}").view(torch.bfloat16), bits
    return torch.from_numpy(x.copy()).to(f"cuda:{device}"), x


def _to_host(t, mode):
    if mode == "bf16":
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


def _check_allreduce(world_obj, n, mode, algo, kind, seed, offset=0):
    W = len(world_obj.comms)
    xs = make_inputs(W, n + offset, seed, kind)
    tens, host = [], []
    for r, c in enumerate(world_obj.comms):
        t, h = _to_dev(xs[r], mode, c.device)
        tens.append(t[offset:])
        host.append(h[offset:])
    scale = 1.0 / W
    world_obj.run(lambda r, c, s: c.allreduce_(tens[r], scale=scale, wire=WIRE[mode], algo=algo, stream=s))
    want = oracle.allreduce(MODES[mode], host, scale)
    for r in range(W):
        assert_bits_equal(_to_host(tens[r], mode), want, f"W={W} n={n} mode={mode} algo={algo} kind={kind} rank={r}")


@pytest.mark.parametrize("mode", list(MODES))
def test_local_pass_matches_oracle(mode):
    from torchx_b200.ddp import local_pass_

    for n in SIZES + [(1 << 22) + 3]:
        for kind in ("randn", "special"):
            for offset in (0, 1):
                x = make_inputs(1, n + offset, 7, kind)[0]
                t, h = _to_dev(x, mode, 0)
                for scale in (1.0, 0.125, 1.0 / 3.0):
                    tt = t.clone()[offset:]
                    local_pass_(tt, scale=scale, wire=WIRE[mode])
                    torch.cuda.synchronize()
                    want = oracle.allreduce(MODES[mode], [h[offset:]], scale)
                    assert_bits_equal(_to_host(tt, mode), want, f"local n={n} mode={mode} scale={scale} off={offset}")


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("algo", ["oneshot", "twoshot", "twoshot_pull"])
def test_allreduce_matches_oracle_one_device(world, mode, algo):
    w = World([0] * world)
    try:
        for i, n in enumerate(SIZES):
            _check_allreduce(w, n, mode, algo, "randn" if i % 2 == 0 else "special", seed=i)
        _check_allreduce(w, 4099, mode, algo, "randn", seed=99, offset=1)  # misaligned base pointer
        _check_allreduce(w, 1 << 12, mode, algo, "onehot", seed=0)
        _check_allreduce(w, 1 << 12, mode, algo, "ints", seed=0)
    finally:
        w.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_allreduce_auto_and_chunking(world):
    """stage_mb=1 forces messages through several chunked launches; AUTO switches algorithm by size."""
    w = World([0] * world, stage_mb=1)
    try:
        for n in (100, 5000, 70001, (1 << 20) + 17):
            _check_allreduce(w, n, "f32_wire_bf16", "auto", "randn", seed=n)
        _check_allreduce(w, (1 << 19) + 3, "f32", "twoshot", "randn", seed=5)
        _check_allreduce(w, (1 << 20) + 9, "f32_wire_bf16", "twoshot_pull", "special", seed=6)
    finally:
        w.close()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("algo", ["twoshot", "twoshot_pull", "auto"])
def test_large_buckets_split_over_two_lanes(world, algo):
    """>= 4 MiB on the wire: the collective runs as two concurrent half-collectives (lane 1 on an internal stream)."""
    w = World([0] * world, stage_mb=32)
    try:
        before = w.comms[0].launches
        _check_allreduce(w, (1 << 21) + 8 * world + 5, "f32_wire_bf16", algo, "special", seed=21)
        # AUTO at W=2 still prefers one-shot at this size (one launch); the two-shot algorithms split
        assert w.comms[0].launches - before == (1 if (algo == "auto" and world == 2) else 2)
        _check_allreduce(w, (1 << 20) + 3, "f32", "twoshot", "randn", seed=22)  # 4 MiB of fp32 wire: split as well
        _check_allreduce(w, 3000, "f32_wire_bf16", "twoshot", "randn", seed=23)  # small again: single lane, same comm
    finally:
        w.close()


def test_back_to_back_ops_reuse_staging_safely():
    """40 collectives of mixed size/algorithm without host syncs in between: exercises the double-buffered
    staging + monotonically increasing flag sequence (no resets)."""
    W = 4
    w = World([0] * W)
    try:
        plan = [(1000 + 37 * i, ("oneshot", "twoshot", "twoshot_pull")[i % 3]) for i in range(42)]
        tens = [[None] * len(plan) for _ in range(W)]
        wants = []
        for k, (n, _) in enumerate(plan):
            xs = make_inputs(W, n, 1000 + k, "randn")
            for r in range(W):
                tens[r][k] = torch.from_numpy(xs[r]).to("cuda:0")
            wants.append(oracle.allreduce(oracle.B2O_F32_WIRE_BF16, xs, 1.0 / W))

        def launch(r, c, s):
            for k, (_, algo) in enumerate(plan):
                c.allreduce_(tens[r][k], algo=algo, stream=s)

        w.run(launch)
        for k in range(len(plan)):
            for r in range(W):
                assert_bits_equal(tens[r][k].cpu().numpy(), wants[k], f"op {k} rank {r}")
        assert w.comms[0].launches == len(plan)
    finally:
        w.close()


@pytest.mark.parametrize("world", [2, 8])
def test_broadcast_and_barrier(world):
    w = World([0] * world, stage_mb=1)
    try:
        for nbytes, root, off in ((1, 0, 0), (15, 1, 0), (4096, world - 1, 0), (100003, 0, 1), ((3 << 20) + 5, 1, 0)):
            src = np.random.default_rng(nbytes).integers(0, 256, size=nbytes + off, dtype=np.uint8)
            tens = []
            for r in range(world):
                data = src if r == root else np.full(nbytes + off, r, dtype=np.uint8)
                tens.append(torch.from_numpy(data.copy()).to("cuda:0")[off:])
            w.run(lambda r, c, s: c.broadcast_(tens[r], root=root, stream=s))
            for r in range(world):
                assert np.array_equal(tens[r].cpu().numpy(), src[off:]), (nbytes, root, r)
        w.run(lambda r, c, s: c.barrier(stream=s))
    finally:
        w.close()


def test_identical_inputs_property_full_bucket_sizes():
    """Size-independent property at the real DDP bucket sizes (ResNet-50's 30.04 MiB fp32 bucket): when every
    rank holds the same x and W is a power of two, every partial sum k * bf16(x)/W is exact, so the result
    must be float(bf16(x)) bit for bit - checked against torch's own cast, no oracle pass over 7.9M elements."""
    W = 4
    n = 7_875_584  # 30.04 MiB of fp32, SURVEY §8a
    w = World([0] * W, stage_mb=16)
    try:
        x = torch.randn(n, device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(3))
        tens = [x.clone() for _ in range(W)]
        w.run(lambda r, c, s: c.allreduce_(tens[r], stream=s))
        want = x.to(torch.bfloat16).float()
        for r in range(W):
            assert torch.equal(tens[r], want), r
    finally:
        w.close()


def test_dead_peer_times_out_instead_of_hanging():
    w = World([0, 0], timeout_s=0.3)
    try:
        t = torch.ones(4096, device="cuda:0")
        w.comms[0].allreduce_(t, stream=w.streams[0])  # rank 1 never shows up
        w.streams[0].synchronize()
        from torchx_b200.ddp._native import B2Error, B2_ETIMEOUT

        with pytest.raises(B2Error) as ei:
            w.comms[0].check()
        assert ei.value.code == B2_ETIMEOUT
        with pytest.raises(B2Error):  # poisoned: refuses further work
            w.comms[0].allreduce_(t, stream=w.streams[0])
    finally:
        w.close()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("algo", ["oneshot", "twoshot", "twoshot_pull"])
def test_allreduce_across_devices(world, algo, cuda_count):
    """Real NVLink/NVSwitch peers (skipped on a 1-GPU box)."""
    devs = _devices(world, cuda_count, spread=True)
    w = World(devs, stage_mb=64)
    try:
        for mode in MODES:
            for n in (9, 4099, (1 << 20) + 5):
                _check_allreduce(w, n, mode, algo, "special", seed=n)
    finally:
        w.close()
