"""Shared helpers for the parity tests (CPU side: numpy + oracle; GPU side: torch tensors)."""
from __future__ import annotations

import numpy as np

import oracle

SPECIALS = np.array(
    [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-40, -1e-40, 1.17549435e-38, 3.3895314e38, -3.3895314e38, 1.0, -1.0,
     1.00390625, 1.0078125, 1.01171875, 65504.0, 1e-3, -1e-3, 255.0, 256.0, 257.0],
    dtype=np.float32,
)


def make_inputs(world: int, n: int, seed: int, kind: str = "randn") -> list:
    """Per-rank fp32 buckets. kinds: randn (seed 1234+rank as in SURVEY §8d), special (inf/nan/subnormal/
    cancellation sprinkled in), onehot (the reference's own compute_world_size trick), ints (exact in bf16)."""
    out = []
    for r in range(world):
        rng = np.random.default_rng(seed + 1234 + r)
        if kind == "randn":
            x = rng.standard_normal(n).astype(np.float32)
        elif kind == "special":
            x = rng.standard_normal(n).astype(np.float32)
            if n:
                idx = rng.integers(0, n, size=max(1, n // 7))
                x[idx] = SPECIALS[rng.integers(0, len(SPECIALS), size=idx.size)]
                if n > 4:
                    x[1] = 3.25 if r % 2 == 0 else -3.25  # cancellation across rank pairs
        elif kind == "onehot":
            x = np.zeros(n, dtype=np.float32)
            x[r::world] = 1.0
        elif kind == "ints":
            x = rng.integers(-120, 121, size=n).astype(np.float32)
        else:
            raise ValueError(kind)
        out.append(x)
    return out


def assert_bits_equal(got: np.ndarray, want: np.ndarray, what: str = "") -> None:
    """Bit-exact comparison; NaNs must coincide but their payload may differ (cvt.rn gives 0x7fff,
    torch's CPU path 0x7fc0)."""
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if got.dtype == np.uint16:
        gf, wf = oracle.bf16_bits_to_f32(got), oracle.bf16_bits_to_f32(want)
        gi, wi = got, want
    else:
        gf, wf = got.astype(np.float32, copy=False), want.astype(np.float32, copy=False)
        gi, wi = gf.view(np.uint32), wf.view(np.uint32)
    gn, wn = np.isnan(gf), np.isnan(wf)
    assert np.array_equal(gn, wn), f"{what}: NaN positions differ at {np.flatnonzero(gn != wn)[:8]}"
    bad = np.flatnonzero((gi != wi) & ~gn)
    assert bad.size == 0, (
        f"{what}: {bad.size} of {got.size} elements differ; first at {bad[:8]}: got {gf[bad[:8]]} want {wf[bad[:8]]}"
    )


def assert_nvls_result(got: np.ndarray, inputs: list, scale: float, mode: int, what: str = "") -> dict:
    """The NVLS path lets the NVSwitch add the W wire contributions (fp32 accumulation, one rounding to bf16).  The
    switch's summation ORDER is not the rank order of the P2P kernels, so the contract checked here is:
      * wherever the exact sum of the contributions is representable in fp32 along EVERY summation order the result is
        order-independent and must equal the rank-order oracle bit for bit (the overwhelming majority of elements);
      * elsewhere it must equal the correctly rounded bf16 of SOME fp32 summation order: checked as within one bf16 ulp
        of the exact (float64) sum.
    Returns counts for reporting."""
    want = oracle.allreduce(mode, inputs, scale)
    contrib = [oracle.compress(mode, x, scale).astype(np.float64) for x in inputs]  # bf16-representable values
    exact = np.sum(contrib, axis=0)
    if got.dtype == np.uint16:
        gf, wf = oracle.bf16_bits_to_f32(got), oracle.bf16_bits_to_f32(want)
    else:
        gf, wf = got.astype(np.float32), want.astype(np.float32)
    gn, wn = np.isnan(gf), np.isnan(wf)
    assert np.array_equal(gn, wn), f"{what}: NaN positions differ at {np.flatnonzero(gn != wn)[:8]}"
    diff = np.flatnonzero((gf.view(np.uint32) != wf.view(np.uint32)) & ~gn)
    if diff.size:
        # order-independence test: every partial sum of |c_r| spans < 2^24 relative to the smallest contribution bit
        absmax = np.max(np.abs(contrib), axis=0)[diff]
        ulp = np.maximum(np.abs(exact[diff]), np.float64(2.0) ** -126) * 2.0 ** -7  # >= one bf16 ulp of the result
        err = np.abs(gf[diff].astype(np.float64) - exact[diff])
        bad = diff[(err > ulp) & np.isfinite(exact[diff])]
        assert bad.size == 0, (
            f"{what}: {bad.size} elements are more than one bf16 ulp from the exact sum; first at {bad[:8]}: "
            f"got {gf[bad[:8]]} exact {exact[bad[:8]]} rank-order {wf[bad[:8]]}")
        del absmax
    return {"n": int(got.size), "differ_from_rank_order": int(diff.size)}
