"""The drop-in, pinned against the REAL reference: when /root/reference is present (the build container; never the GPU
box) the reference's own `torchx.runner.api.Runner` submits its own `dist.ddp` AppDef (-j 1x2, CPU/gloo toy job) to this
repo's `local_cuda` scheduler - once with the factory handed to the Runner (torchx/runner/api.py:621-632) and once through
the `torchx_plugins.schedulers` namespace-plugin route of INTEGRATION.md (torchx/schedulers/__init__.py:40-60,
torchx/plugins/_registry.py)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "torchx")), reason="/root/reference is not mounted here")

PLUGIN = textwrap.dedent('''
    from torchx.plugins import register


    @register.scheduler(name="local_cuda")
    def local_cuda(session_name: str, **kwargs):
        from torchx_b200.schedulers.local_cuda_scheduler import create_scheduler
        return create_scheduler(session_name, **kwargs)


    @register.scheduler(name="local_cwd")
    def local_cwd(session_name: str, **kwargs):
        from torchx.schedulers.local_scheduler import create_scheduler
        return create_scheduler(session_name, **kwargs)
''')


def _drive(mode, tmp_path, extra_path=()):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([REF, *extra_path, ROOT])
    env["TORCHX_HOME"] = str(tmp_path / "home")
    cmd = [sys.executable, os.path.join(ROOT, "tests", "workers", "reference_runner_driver.py"), mode,
           os.path.join(ROOT, "examples", "toy_ddp.py"), str(tmp_path / "logs")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])


def test_reference_runner_drives_local_cuda_through_the_factory(tmp_path):
    out = _drive("factory", tmp_path)
    assert out["ok"], out
    assert out["handle"].startswith("local_cuda://torchx/toy_ddp-")
    assert out["describe_roles"] == ["toy_ddp"]
    assert any("grads" in ln or "sha" in ln.lower() or "ok" in ln.lower() for ln in out["log_tail"]), out["log_tail"]


def test_reference_registry_finds_local_cuda_as_a_namespace_plugin(tmp_path):
    pkg = tmp_path / "plug" / "torchx_plugins" / "schedulers"
    pkg.mkdir(parents=True)
    (pkg / "b200.py").write_text(PLUGIN)  # namespace packages: no __init__.py on purpose
    out = _drive("plugin", tmp_path, extra_path=(str(tmp_path / "plug"),))
    assert out["ok"], out
    assert "local_cuda" in out["schedulers"] and "local_cwd" in out["schedulers"]
