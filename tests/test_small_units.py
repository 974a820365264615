"""ids, Tee, StructuredOpts docstrings, distributed helpers without a process group (CPU)."""
import io
import os
import re
import time

import pytest

from torchx_b200.schedulers.ids import make_unique, random_id
from torchx_b200.schedulers.streams import Tee


def test_unique_ids_shape():
    ids = {make_unique("job") for _ in range(200)}
    assert len(ids) == 200
    for i in ids:
        assert re.fullmatch(r"job-[bcdfghjklmnpqrstvwxz][bcdfghjklmnpqrstvwxz012345679]*", i) and 6 <= len(i) <= 20
    assert random_id(0) == "" and len(random_id(5)) <= 5 and make_unique("x", 4).startswith("x-")


def test_tee_merges_growing_sources_and_prefixes_lines(tmp_path):
    a, b, out = tmp_path / "a", tmp_path / "b", tmp_path / "out"
    a.write_bytes(b"")
    fa = open(a, "ab", buffering=0)
    tee = Tee(io.open(out, "wb", buffering=0), str(a), str(b), prefixes=[b"[0]:", b"[1]:"])
    fa.write(b"one\ntwo")
    time.sleep(0.2)
    fb = open(b, "ab", buffering=0)  # a source may appear later
    fb.write(b"uno\n")
    fa.write(b"-continued\n")
    fa.close(); fb.close()
    time.sleep(0.2)
    tee.close(); tee.close()
    lines = sorted(out.read_bytes().splitlines())
    assert lines == [b"[0]:one", b"[0]:two-continued", b"[1]:uno"]
    raw = tmp_path / "raw"
    t2 = Tee(io.open(raw, "wb", buffering=0), str(a))
    time.sleep(0.15)
    t2.close()
    assert raw.read_bytes() == b"one\ntwo-continued\n"
    with pytest.raises(ValueError):
        Tee(io.BytesIO())


def test_cuda_opts_help_texts_include_inherited_fields():
    from torchx_b200.schedulers.local_cuda_scheduler import CudaOpts

    ro = dict(CudaOpts.as_runopts())
    assert ro["log_dir"].help.startswith("Directory to write stdout/stderr")
    assert ro["pin_cpus"].help.startswith("Bind each worker") and ro["pin_cpus"].default is True
    assert ro["devices"].opt_type.__args__ == (str,)


def test_worker_helpers_without_a_launcher(monkeypatch):
    import torchx_b200.distributed as td

    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "B2_DEVICE", "TORCHELASTIC_RUN_ID"):
        monkeypatch.delenv(k, raising=False)
    assert (td.rank(), td.world_size(), td.local_rank()) == (0, 1, 0) and not td.is_torchelastic_launched()
    monkeypatch.setenv("LOCAL_RANK", "3")
    monkeypatch.setenv("B2_DEVICE", "5")
    assert td.local_rank() == 3 and td.local_cuda_device().index == 5  # the scheduler's pin wins over LOCAL_RANK
    with td.on_rank0_first():  # no group: plain pass-through
        pass
    with pytest.raises(RuntimeError):
        td.communicator()


def test_numa_cpulist_parser_and_device_pool():
    from torchx_b200.schedulers.local_cuda_scheduler import _parse_cpulist

    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11] and _parse_cpulist("") == []
