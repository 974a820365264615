"""Pins the CPU oracle against outputs of the reference itself (tests/golden/ddp_w*.npz: stock DDP launched by
the reference's `torchx run -s local_cwd dist.ddp` over gloo, see make_golden.py) and against an independent
numpy restatement."""
import os

import numpy as np
import pytest

import oracle
from oracle.ref import allreduce_numpy
from tests._util import assert_bits_equal, make_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(w):
    return np.load(os.path.join(GOLD, f"ddp_w{w}.npz"))


def test_w2_bit_exact_vs_reference_ddp():
    g = _gold(2)
    local = [g["local"][r] for r in range(2)]
    for r in range(2):  # every rank of the reference run holds the same result
        assert_bits_equal(oracle.allreduce(oracle.B2O_F32, local, 0.5), g["ddp_none"][r], "no hook")
        assert_bits_equal(oracle.allreduce(oracle.B2O_F32, local, 0.5), g["ddp_allreduce"][r], "allreduce_hook")
        assert_bits_equal(oracle.allreduce(oracle.B2O_F32_WIRE_BF16, local, 0.5), g["ddp_bf16_compress"][r], "bf16_compress_hook")


def test_w4_within_tolerance_of_reference_ddp():
    """At W=4 gloo/NCCL reduce in their own order, so the pin is a tolerance: fp32 within 1e-6 relative of
    max|g| per SURVEY §7; bf16 within (W-1) bf16 ulps of the largest magnitude."""
    g = _gold(4)
    local = [g["local"][r] for r in range(4)]
    scale = float(np.abs(np.stack(local)).max())
    o32 = oracle.allreduce(oracle.B2O_F32, local, 0.25)
    assert np.abs(o32 - g["ddp_none"][0]).max() <= 1e-6 * scale
    assert np.abs(o32 - g["ddp_allreduce"][0]).max() <= 1e-6 * scale
    ob = oracle.allreduce(oracle.B2O_F32_WIRE_BF16, local, 0.25)
    ref = g["ddp_bf16_compress"][0]
    # every intermediate bf16 rounding of the reference's own reduction order errs by at most half a bf16 ulp
    # (2^-9 relative) of a partial sum, and |partial sum| <= sum_r |c_r|: (W-1) roundings in total
    mag = sum(np.abs(oracle.compress(oracle.B2O_F32_WIRE_BF16, x, 0.25)) for x in local)
    assert np.all(np.abs(ob - ref) <= 3 * 2.0 ** -8 * mag + 1e-30)
    # and the oracle (fp32 accumulate, one rounding) is at least as close to the exact mean as the reference
    exact = np.mean(np.stack(local).astype(np.float64), axis=0)
    assert np.abs(ob - exact).mean() <= np.abs(ref - exact).mean() * 1.0001


def test_torch_op_sequence_restatement_agrees_at_w2():
    torch = pytest.importorskip("torch")
    g = _gold(2)
    ins = [torch.from_numpy(g["local"][r].copy()) for r in range(2)]
    for hook, key in (("none", "ddp_none"), ("allreduce", "ddp_allreduce"), ("bf16_compress", "ddp_bf16_compress")):
        assert_bits_equal(oracle.torch_hook_restatement(ins, hook).numpy(), g[key][0], hook)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("kind", ["randn", "special", "onehot", "ints"])
def test_c_oracle_equals_numpy_twin(world, kind):
    for n in (0, 1, 7, 1025, 40001):
        xs = make_inputs(world, n, 3, kind)
        for mode in (oracle.B2O_F32_WIRE_BF16, oracle.B2O_F32):
            assert_bits_equal(oracle.allreduce(mode, xs, 1.0 / world), allreduce_numpy(mode, xs, 1.0 / world), f"{mode}")
        xb = [oracle.f32_to_bf16_bits(x) for x in xs]
        assert_bits_equal(oracle.allreduce(oracle.B2O_BF16, xb, 1.0 / world), allreduce_numpy(oracle.B2O_BF16, xb, 1.0 / world), "bf16")


def test_exactness_cases_any_order():
    """Inputs whose sum is exact in bf16 regardless of order (the reference's own one-hot trick,
    torchx/examples/apps/compute_world_size/module/util.py:30-37)."""
    for w in (2, 4, 8):
        xs = make_inputs(w, 4096, 0, "onehot")
        out = oracle.allreduce(oracle.B2O_F32_WIRE_BF16, xs, 1.0)
        assert np.array_equal(out, np.ones(4096, np.float32))
        out = oracle.allreduce(oracle.B2O_F32_WIRE_BF16, xs, 1.0 / w)
        assert np.array_equal(out, np.full(4096, 1.0 / w, np.float32))


def test_bf16_rounding_known_answers():
    x = np.array([1.0, 1.00390625, 1.005859375, 1.01171875, 3.4e38, -0.0, 1e-40], np.float32)
    bits = oracle.f32_to_bf16_bits(x)
    # ties-to-even: 1.00390625 (=1+2^-8) is exactly halfway between 1.0 and 1.0078125 -> even mantissa 1.0
    assert bits[1] == 0x3F80 and bits[2] == 0x3F81 and bits[3] == 0x3F82
    assert bits[4] == 0x7F80  # rounds up to +inf
    assert bits[5] == 0x8000
    torch = pytest.importorskip("torch")
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(bits, want)
