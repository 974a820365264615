"""B2_ALGO_AUTO's policy table - a pure host function, no GPU: which algorithm a message of a
given size gets on a given world, with and without NVSwitch multicast.  The thresholds come from the measured sweeps in
profiles/r02_sweep_w8.md and r02_pipeline_and_ll_w2.md; this test pins them so that a change is a conscious one."""
import pytest

from torchx_b200.ddp import _native as N

MIB = 1 << 20
NAMES = {N.B2_ALGO_ONESHOT: "oneshot", N.B2_ALGO_TWOSHOT: "twoshot", N.B2_ALGO_TWOSHOT_LL: "twoshot_ll", N.B2_ALGO_NVLS: "nvls",
         N.B2_ALGO_TWOSHOT_PIPE: "twoshot_pipe", N.B2_ALGO_AUTO: "local"}


def pick(world, fp32_bucket_mib, multicast=True, mode=N.B2_F32_WIRE_BF16):
    return NAMES[N.lib().b2_auto_algo(world, mode, int(fp32_bucket_mib * MIB / 4), int(multicast))]


@pytest.fixture(autouse=True)
def _no_env_overrides(monkeypatch):
    for k in ("B2_ONESHOT_MAX_BYTES", "B2_PIPE_MIN_BYTES", "B2_NVLS_MIN_BYTES", "B2_NVLS_MIN_WORLD", "B2_LL_MIN_BYTES", "B2_LL_MAX_BYTES"):
        monkeypatch.delenv(k, raising=False)


def test_resnet50_gpt2_and_bert_buckets_at_w8():
    # ResNet-50: 7.82 / 9.27 MiB -> LL two-shot, 25-30 MiB -> single-pass two-shot (rank-order kernels: bit-exact vs the oracle)
    assert [pick(8, m) for m in (7.82, 9.27, 25.04, 25.32, 30.04)] == ["twoshot_ll", "twoshot_ll", "twoshot", "twoshot", "twoshot"]
    # GPT-2-small's 168 MiB and BERT-base's 91 MiB embedding buckets: the former is past the NVLS threshold (128 MiB of fp32)
    assert pick(8, 168.27) == "nvls" and pick(8, 90.93) == "twoshot" and pick(8, 27.04) == "twoshot"
    assert pick(8, 168.27, multicast=False) == "twoshot"            # no multicast on the box: rank-order kernel
    assert pick(8, 168.27, mode=N.B2_F32) == "twoshot"              # fp32 wire never goes through the switch under AUTO
    assert pick(8, 0.5) == "oneshot" and pick(8, 1.0) == "oneshot" and pick(8, 1.01) == "twoshot_ll"
    assert pick(8, 16.0) == "twoshot" and pick(8, 15.9) == "twoshot_ll"
    assert pick(8, 127.9) == "twoshot" and pick(8, 128.0) == "nvls" and pick(8, 1024.0) == "nvls"


def test_small_worlds_and_the_local_pass():
    assert pick(1, 25.0) == "local"
    # W=2: one-shot up to 16 MiB of wire data (32 MiB fp32), single-pass two-shot above; neither LL nor NVLS pays there
    assert [pick(2, m) for m in (0.25, 7.82, 30.04, 32.0, 64.0, 1024.0)] == ["oneshot"] * 4 + ["twoshot"] * 2
    # W=4: one-shot to 2 MiB of wire data, LL to 8 MiB, two-shot above, no NVLS (1.25 S through a ~430 GB/s path loses to 1.5 S P2P)
    assert [pick(4, m) for m in (1.0, 4.0, 7.82, 15.9, 25.04, 256.0)] == ["oneshot", "oneshot", "twoshot_ll", "twoshot_ll", "twoshot", "twoshot"]
    assert pick(7, 256.0) == "twoshot" and pick(7, 4.0) == "twoshot_ll" and pick(5, 1.0) == "oneshot" and pick(5, 1.5) == "twoshot_ll"


def test_environment_overrides_and_validation(monkeypatch):
    monkeypatch.setenv("B2_NVLS_MIN_BYTES", str(1 << 40))
    assert pick(8, 1024.0) == "twoshot"
    monkeypatch.setenv("B2_LL_MIN_BYTES", str(1 << 40))
    assert pick(8, 7.82) == "twoshot"
    monkeypatch.setenv("B2_PIPE_MIN_BYTES", str(4 * MIB))
    assert pick(8, 25.04) == "twoshot_pipe" and pick(8, 4.0) == "twoshot"
    L = N.lib()
    assert L.b2_auto_algo(9, 0, 100, 0) == N.B2_EINVAL and L.b2_auto_algo(2, 7, 100, 0) == N.B2_EINVAL
