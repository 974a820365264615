"""Our launcher vs the REFERENCE launcher's own outputs (tests/golden/launcher.json, written by make_golden.py which
imports meta-pytorch/torchx from /root/reference): dist.ddp AppDefs, -j grammar, cfg strings, macros, the local_cwd
dry-run request and the CUDA_VISIBLE_DEVICES partitioning table.  Byte-for-byte equality is the bar."""
import json
import os
from dataclasses import asdict
from typing import Dict, List
from unittest import mock

import pytest

from torchx_b200.components.dist import ddp, parse_nnodes
from torchx_b200.components.structured_arg import StructuredNameArgument
from torchx_b200.schedulers.local_scheduler import create_scheduler
from torchx_b200.specs import AppDef, Resource, Role, macros, runopts
from torchx_b200.util.types import to_dict

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "launcher.json")))


def _norm_app(app) -> dict:
    d = asdict(app)
    for r in d["roles"]:
        for k in ("overrides", "workspace", "mounts", "image"):
            r.pop(k, None)
        r["resource"] = {k: r["resource"][k] for k in ("cpu", "gpu", "memMB")}
        r["retry_policy"] = str(r["retry_policy"].value)
    return d


@pytest.mark.parametrize("case", sorted(G["ddp"]))
def test_dist_ddp_appdef_is_byte_compatible(case):
    c = G["ddp"][case]
    with mock.patch.dict(os.environ, {"LOGLEVEL": "WARNING"}):
        app = ddp(*c["call"]["args"], **c["call"]["kw"])
    assert _norm_app(app) == c["app"]


def test_parse_nnodes_table():
    for j, want in G["parse_nnodes"].items():
        assert list(parse_nnodes(j)) == want, j
    with pytest.raises(ValueError, match="Invalid format for -j"):
        parse_nnodes("2x")


def test_structured_name_argument():
    for key, want in G["name_arg"].items():
        n, m, s = key.split("|")
        got = StructuredNameArgument.parse_from(name=n, m=None if m == "None" else m, script=None if s == "None" else s)
        assert asdict(got) == want, key
    with pytest.raises(ValueError):
        StructuredNameArgument.parse_from(name="a/b")
    with pytest.raises(ValueError):
        StructuredNameArgument.parse_from(name="a/b", m="x", script="y.py")


def test_to_dict_literals():
    for lit, want in G["to_dict"].items():
        assert to_dict(lit) == want, lit


def test_runopts_cfg_from_str_and_resolve():
    opts = runopts()
    opts.add("FOO", type_=List[str], default=["a"], help="list")
    opts.add("BAR", type_=str, required=True, help="str")
    opts.add("N", type_=int, default=3, help="int")
    opts.add("B", type_=bool, default=False, help="bool")
    opts.add("D", type_=Dict[str, str], default=None, help="dict")
    for lit, want in G["cfg_from_str"].items():
        assert opts.cfg_from_str(lit) == want, lit
    assert opts.resolve({"BAR": "z"}) == G["resolve"]
    from torchx_b200.specs import InvalidRunConfigException

    with pytest.raises(InvalidRunConfigException):
        opts.resolve({})  # BAR is required
    with pytest.raises(InvalidRunConfigException):
        opts.resolve({"BAR": "x", "N": "not-an-int"})
    assert opts.resolve({"BAR": "x", "n": 1})["N"] == 3  # unknown keys pass through, known keep defaults


def test_macro_substitution():
    role = Role(name="r", image="img", entrypoint="e", args=["${img_root}/x", "--id", "${app_id}", "${replica_id}", "$$lit", "${unknown}"],
                env={"H": "${rank0_env}"}, metadata={"k": {"a": ["${app_id}", {"b": "${replica_id}"}]}}, resource=Resource(1, 0, 1))
    rr = macros.Values(img_root="/img", app_id="app-1", replica_id="3", rank0_env="TORCHX_RANK0_HOST").apply(role)
    assert {"args": rr.args, "env": rr.env, "metadata": rr.metadata} == G["macros"]
    assert role.args[0] == "${img_root}/x"  # the original is untouched


def _cvd(device_count, roles, auto=True):
    sched = create_scheduler("golden")
    try:
        with mock.patch.object(sched, "_cuda_device_count", return_value=device_count):
            app = AppDef("a", roles=[Role(name=n, image="", entrypoint="e", num_replicas=k, resource=Resource(1, g, 1)) for n, k, g in roles])
            info = sched.submit_dryrun(app, {"auto_set_cuda_visible_devices": auto})
            return {n: [p.env.get("CUDA_VISIBLE_DEVICES") for p in info.request.role_params[n]] for n, _, _ in roles}
    finally:
        sched.close()


def test_cuda_visible_devices_partitioning_matches_reference():
    want = G["cuda_visible_devices"]
    assert _cvd(8, [("t", 2, 4)]) == want["8gpu_1role_2x4"]
    assert _cvd(8, [("a", 1, 2), ("b", 3, 2)]) == want["8gpu_2roles"]
    assert _cvd(8, [("t", 3, 4)]) == want["8gpu_too_many"]
    assert _cvd(0, [("t", 1, 2)]) == want["0gpu"]
    assert _cvd(16, [("cpu", 2, 0), ("g", 2, 8)]) == want["16gpu_cpu_and_gpu_roles"]
    assert _cvd(8, [("t", 2, 4)], auto=False) == want["8gpu_auto_off"]


def test_local_cwd_dryrun_request_matches_reference_and_creates_nothing(tmp_path):
    sched = create_scheduler("golden")
    try:
        app = ddp("--foo", "bar", script="toy_ddp.py", j="1x2")
        info = sched.submit_dryrun(app, {"log_dir": "/tmp/golden_logs"})
        req = info.request
        rp = req.role_params["toy_ddp"][0]
        sub = lambda s: s.replace(req.app_id, "<APP_ID>")  # noqa: E731
        got = {
            "args": [sub(a) for a in rp.args],
            "env_added": {k: sub(v) for k, v in rp.env.items() if k in ("TORCHX_RANK0_HOST", "TORCHELASTIC_ERROR_FILE", "PET_LOG_DIR")},
            "stdout": sub(rp.stdout), "stderr": sub(rp.stderr), "combined": sub(rp.combined), "log_dir": sub(req.log_dir),
            "created_dirs": os.path.exists(req.log_dir),
        }
        assert got == G["local_cwd_request"]
        assert "role_params" in repr(info)  # pretty-printable request (AppDryRunInfo.__repr__)
        assert req.app_id.startswith("toy_ddp-") and len(req.app_id) > len("toy_ddp-") + 8
    finally:
        sched.close()
