"""Product bucket planner vs torch's own function (live) and the committed layouts (indices bit-exact)."""
import json
import os

import pytest

from torchx_b200.ddp.bucketing import plan_buckets

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bucket_layouts.json")


@pytest.mark.parametrize("model", ["resnet50", "gpt2_small", "bert_base"])
@pytest.mark.parametrize("dtype,size", [("fp32", 4), ("bf16", 2)])
def test_plan_matches_stock_ddp_layout(model, dtype, size):
    g = json.load(open(GOLD))[model]
    numels = g["param_numel"]
    specs = plan_buckets(numels, [size] * len(numels), [dtype] * len(numels))
    assert [s.param_indices for s in specs] == g[f"buckets_{dtype}"]
    for s in specs:  # offsets are the running sum in fill order
        off = 0
        for i, o, n in zip(s.param_indices, s.offsets, s.numels):
            assert o == off and n == numels[i]
            off += n
        assert s.numel == off and s.nbytes == off * size


def test_plan_matches_live_torch_random_shapes():
    torch = pytest.importorskip("torch")
    import torch.distributed as dist

    gen = torch.Generator().manual_seed(1)
    for _ in range(25):
        k = int(torch.randint(1, 80, (1,), generator=gen))
        numels = [int(x) for x in torch.randint(1, 4_000_000, (k,), generator=gen)]
        tens = [torch.empty(n, device="meta") for n in numels]
        rev = list(reversed(range(k)))
        want, _ = dist._compute_bucket_assignment_by_size([tens[i] for i in rev], [1 << 20, 25 << 20], [False] * k, rev)
        got = plan_buckets(numels, [4] * k, ["f32"] * k)
        assert [s.param_indices for s in got] == want
