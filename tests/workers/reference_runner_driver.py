"""Runs INSIDE a subprocess whose sys.path has /root/reference first: the REFERENCE's own Runner / CLI machinery drives
this repo's `local_cuda` scheduler (tests/test_reference_dropin.py).  Prints one JSON line."""
import json
import os
import sys

mode, script, log_dir = sys.argv[1], sys.argv[2], sys.argv[3]

import torchx  # noqa: E402  (the reference package)
from torchx.runner.api import Runner  # noqa: E402
from torchx.specs import AppState  # noqa: E402

assert os.path.realpath(torchx.__file__).startswith("/root/reference"), torchx.__file__
out = {"torchx": torchx.__file__}
dist_ddp = "/root/reference/torchx/components/dist.py:ddp"  # builtin-by-name discovery needs hydra; address it by file

if mode == "factory":
    # torchx/runner/api.py:621-632: a Runner is handed {name: factory}; ours has the reference factory signature
    from torchx_b200.schedulers.local_cuda_scheduler import create_scheduler

    runner = Runner("torchx", {"local_cuda": create_scheduler})
elif mode == "plugin":
    # torchx/schedulers/__init__.py:40-60: the registry the CLI uses; the torchx_plugins namespace package on sys.path
    # (written by the test from INTEGRATION.md) registers local_cuda (and re-registers local_cwd)
    from torchx.schedulers import get_scheduler_factories

    factories = get_scheduler_factories()
    out["schedulers"] = sorted(factories)
    assert "local_cuda" in factories and "local_cwd" in factories, sorted(factories)
    runner = Runner("torchx", factories)
else:
    raise SystemExit(mode)

with runner:
    cfg = {"log_dir": log_dir}
    dry = runner.dryrun_component(dist_ddp, ["-j", "1x2", "--script", script], "local_cuda", cfg)
    out["dryrun_repr_has_workers"] = "RANK" in repr(dry) or "rank" in repr(dry).lower()
    handle = runner.run_component(dist_ddp, ["-j", "1x2", "--script", script], "local_cuda", cfg)
    out["handle"] = handle
    status = runner.wait(handle, wait_interval=0.2)
    out["state"] = str(status.state)
    out["ok"] = status.state == AppState.SUCCEEDED
    out["describe_roles"] = [r.name for r in runner.describe(handle).roles] if runner.describe(handle) else None
    out["log_tail"] = [ln.rstrip("\n") for ln in runner.log_lines(handle, "toy_ddp", 0)][-4:]
    out["list"] = [a.app_id for a in runner.list("local_cuda")][:3] if hasattr(runner, "list") else None
print(json.dumps(out))
