"""Offline analysis of tools/nvls_probe.py output: how does the NVSwitch's multimem.ld_reduce(.acc::f32, bf16x2) round?

    python tools/nvls_order_fit.py profiles/r02_nvls_probe_mp_w2.npz profiles/r02_nvls_probe_mp_w8.npz

For each probe file and input class it prints: the share of elements equal to RNE(exact sum), equal to the rank-order P2P
kernel (fp32 accumulate in rank order, one RNE rounding), the largest error in bf16 ulps of the exact sum, the mean signed
error, and P(result rounded away from zero) as a function of the discarded remainder - the signature of the rounding rule."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (test infrastructure; this tool is an analysis script, not the product)


def analyse(path):
    d = np.load(path)
    W = int(d["world"][0])
    bits = d["bf16_in"]
    got_all = d["bf16_out_nvls"] if "bf16_out_nvls" in d.files else d["bf16_out"][0, 0]
    p2p = d["bf16_out_p2p"] if d["bf16_out_p2p"].ndim == 1 else d["bf16_out_p2p"][0]
    print(f"## {os.path.basename(path)}  (W = {W}, {bits.shape[1]} elements per rank)\n")
    print("| class | n | == RNE(exact) | == P2P rank-order | max err (bf16 ulp) | mean err (ulp) | P(away) for remainder in (0,1/8) .. (7/8,1) | exact ties: n, P(away) |")
    print("|---|---:|---:|---:|---:|---:|---|---|")
    for name, s0, ln in zip(d["class_names"], d["class_start"], d["class_len"]):
        name = str(name)
        vals = np.stack([oracle.bf16_bits_to_f32(bits[r, s0:s0 + ln]).astype(np.float64) for r in range(W)])
        with np.errstate(all="ignore"):
            exact = vals.sum(0)
        got_b = got_all[s0:s0 + ln]
        got = oracle.bf16_bits_to_f32(got_b).astype(np.float64)
        fin = np.isfinite(exact) & np.isfinite(got) & (exact != 0)
        if fin.sum() == 0:
            continue
        rne = oracle.bf16_bits_to_f32(oracle.f32_to_bf16_bits(exact.astype(np.float32))).astype(np.float64)
        ulp = 2.0 ** (np.floor(np.log2(np.abs(exact[fin]))) - 7)
        ulp = np.maximum(ulp, 2.0 ** -133)
        err = (got[fin] - exact[fin]) / ulp
        t = np.sign(exact[fin]) * np.floor(np.abs(exact[fin]) / ulp) * ulp
        frac = np.abs(exact[fin] - t) / ulp
        away = np.abs(got[fin]) > np.abs(t)
        cells = []
        for lo in np.arange(8) / 8.0:
            m = (frac > lo) & (frac < lo + 0.125)
            cells.append(f"{away[m].mean():.2f}" if m.sum() >= 20 else "-")
        ties = frac == 0.5
        print(f"| {name} | {ln} | {np.mean(got[fin] == rne[fin]):.3f} | {np.mean(got_b == p2p[s0:s0 + ln]):.3f} | {np.abs(err).max():.3f} | {err.mean():+.4f} | "
              f"{' '.join(cells)} | {int(ties.sum())}, {away[ties].mean():.2f} |" if ties.sum() else
              f"| {name} | {ln} | {np.mean(got[fin] == rne[fin]):.3f} | {np.mean(got_b == p2p[s0:s0 + ln]):.3f} | {np.abs(err).max():.3f} | {err.mean():+.4f} | {' '.join(cells)} | 0 |")
    print()


def structure(path):
    """W = 2 only: what does the rounding decision depend on?  Same (a, b) at two positions; same sum at two positions; the
    'tie' class (one input pair repeated at every position)."""
    from collections import defaultdict

    d = np.load(path)
    if int(d["world"][0]) != 2:
        return
    names = [str(x) for x in d["class_names"]]
    a, b = d["bf16_in"][0].astype(np.uint32), d["bf16_in"][1].astype(np.uint32)
    out = (d["bf16_out_nvls"] if "bf16_out_nvls" in d.files else d["bf16_out"][0, 0]).astype(np.uint32)
    with np.errstate(all="ignore"):
        s64 = (a << 16).view(np.float32).astype(np.float64) + (b << 16).view(np.float32).astype(np.float64)
        s32 = s64.astype(np.float32)
    sb = s32.view(np.uint32)
    rem, trunc = sb & 0xFFFF, sb >> 16
    cls = np.zeros(a.size, dtype=int)
    for i, (s0, ln) in enumerate(zip(d["class_start"], d["class_len"])):
        cls[s0:s0 + ln] = i
    inexact = np.isfinite(s64) & (s32.astype(np.float64) == s64) & (rem != 0) & ((out == trunc) | (out == trunc + 1))
    pool = inexact & np.isin(cls, [names.index(c) for c in ("wide", "narrow", "subnormal")])
    up = out == trunc + 1

    def inconsistent(key):
        groups = defaultdict(set)
        for i in np.flatnonzero(pool):
            groups[key(i)].add(bool(up[i]))
        multi = [v for v in groups.values()]
        sizes = defaultdict(int)
        for i in np.flatnonzero(pool):
            sizes[key(i)] += 1
        multi = [k for k, c in sizes.items() if c > 1]
        return sum(1 for k in multi if len(groups[k]) > 1), len(multi)

    print(f"### what the decision depends on ({os.path.basename(path)}, inexact sums of the wide / narrow / subnormal classes: {int(pool.sum())})\n")
    for label, key in (("same fp32 sum at several positions", lambda i: int(sb[i])),
                       ("same fp32 sum and same lane (position mod 8)", lambda i: (int(sb[i]), int(i) % 8)),
                       ("same input pair (a, b) at several positions", lambda i: (int(a[i]), int(b[i])))):
        bad, tot = inconsistent(key)
        print(f"* {label}: {tot} groups, {bad} of them rounded BOTH ways")
    t = cls == names.index("tie")
    print(f"* 'tie' class (the pair 1.0, 2^-8 at all {int(t.sum())} positions, remainder exactly 1/2): {int((out[t] == trunc[t] + 1).sum())} rounded up, "
          f"{int((out[t] == trunc[t]).sum())} down; ties elsewhere: {float(up[pool & (rem == 0x8000)].mean()):.2f} rounded up\n")


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] in ("-h", "--help"):
        print(__doc__)
        sys.exit(0)
    for p in sys.argv[1:]:
        analyse(p)
    for p in sys.argv[1:]:
        structure(p)
