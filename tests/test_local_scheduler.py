"""Behavioural tests of the local schedulers with tiny bash stubs and a mocked device count - the reference's own
strategy (torchx/schedulers/test/local_scheduler_test.py:224-244 stubs, :52-55 device mock).  No GPU needed."""
import json
import os
import signal
import stat
import subprocess
import sys
import time

import pytest

from torchx_b200.schedulers.api import Stream
from torchx_b200.schedulers.local_scheduler import LocalDirectoryImageProvider, LocalScheduler, create_scheduler
from torchx_b200.specs import AppDef, AppState, Resource, Role, is_terminal, macros

SCRIPTS = {
    "touch.sh": "#!/bin/bash\ntouch $1\n",
    "env.sh": "#!/bin/bash\necho \"$1=${!1}\" > $2\n",
    "fail.sh": "#!/bin/bash\nexit 1\n",
    "sleep.sh": "#!/bin/bash\nsleep $1\n",
    "echo_stdout.sh": "#!/bin/bash\necho $1\n",
    "echo_stderr.sh": "#!/bin/bash\necho $1 1>&2\n",
    "echo_range.sh": "#!/bin/bash\nfor i in $(seq 0 $1); do echo $i 1>&2; done\n",
    "echo_env_foo.sh": "#!/bin/bash\necho $FOO 1>&2\n",
}


@pytest.fixture()
def image(tmp_path):
    d = tmp_path / "image"
    d.mkdir()
    for name, body in SCRIPTS.items():
        p = d / name
        p.write_text(body)
        p.chmod(p.stat().st_mode | stat.S_IEXEC)
    return str(d)


@pytest.fixture()
def sched():
    s = LocalScheduler("test_session", image_provider_class=LocalDirectoryImageProvider, cache_size=4)
    yield s
    s.close()


def _role(image, entrypoint, *args, name="trainer", n=1, env=None, **kw):
    return Role(name=name, image=image, entrypoint=entrypoint, args=list(args), env=dict(env or {}), num_replicas=n,
                resource=Resource(1, 0, 1), **kw)


def _wait(s, app_id, timeout=30):
    end = time.time() + timeout
    while time.time() < end:
        d = s.describe(app_id)
        if d is None or is_terminal(d.state):
            return d
        time.sleep(0.05)
    raise TimeoutError(app_id)


def test_submit_runs_replicas_and_substitutes_macros(sched, image, tmp_path):
    out = tmp_path / "out"
    out.mkdir()
    role = _role(image, "touch.sh", os.path.join(str(out), f"{macros.app_id}_{macros.replica_id}"), n=3)
    app_id = sched.submit(AppDef("t", roles=[role]), {"log_dir": str(tmp_path / "logs")})
    assert _wait(sched, app_id).state == AppState.SUCCEEDED
    assert sorted(os.listdir(out)) == [f"{app_id}_{k}" for k in range(3)]
    assert role.args[0].endswith("${app_id}_${replica_id}")  # caller's AppDef is not mutated by substitution


def test_child_env_inherits_parent_and_role_env_path_joined(sched, image, tmp_path, monkeypatch):
    monkeypatch.setenv("FROM_PARENT", "p1")
    f1, f2, f3 = (str(tmp_path / n) for n in ("e1", "e2", "e3"))
    roles = [
        _role(image, "env.sh", "FROM_PARENT", f1, name="a"),
        _role(image, "env.sh", "FROM_ROLE", f2, name="b", env={"FROM_ROLE": "r1"}),
        _role(image, "env.sh", "PATH", f3, name="c", env={"PATH": "/custom/bin"}),
    ]
    app_id = sched.submit(AppDef("t", roles=roles), {"log_dir": str(tmp_path / "logs")})
    assert _wait(sched, app_id).state == AppState.SUCCEEDED
    assert open(f1).read().strip() == "FROM_PARENT=p1"
    assert open(f2).read().strip() == "FROM_ROLE=r1"
    path = open(f3).read().strip().split("=", 1)[1].split(":")
    assert path[0] == "/custom/bin" and image in path and os.environ["PATH"].split(":")[0] in path


def test_failure_state_error_file_and_success_manifest(sched, image, tmp_path):
    roles = [_role(image, "fail.sh", name="bad"), _role(image, "sleep.sh", "60", name="slow")]
    app_id = sched.submit(AppDef("t", roles=roles), {"log_dir": str(tmp_path / "logs")})
    # a live sibling keeps the app RUNNING (reference describe semantics); cancel then tears it down
    time.sleep(0.5)
    assert sched.describe(app_id).state == AppState.RUNNING
    sched.cancel(app_id)
    d = sched.describe(app_id)
    assert d.state == AppState.CANCELLED
    app_dir = os.path.join(str(tmp_path / "logs"), "test_session", app_id)
    manifest = json.load(open(os.path.join(app_dir, "SUCCESS")))
    assert manifest["app_id"] == app_id and set(manifest["roles"]) == {"bad", "slow"}
    assert manifest["roles"]["bad"][0]["exitcode"] == 1

    app_id = sched.submit(AppDef("t", roles=[_role(image, "fail.sh", name="bad")]), {"log_dir": str(tmp_path / "logs")})
    d = _wait(sched, app_id)
    assert d.state == AppState.FAILED and d.num_restarts == 0 and d.ui_url.startswith("file://")


def test_structured_error_comes_from_oldest_error_file(sched, image, tmp_path):
    app = AppDef("t", roles=[_role(image, "sleep.sh", "0.3", n=2)])
    info = sched.submit_dryrun(app, {"log_dir": str(tmp_path / "logs")})
    app_id = sched.schedule(info)
    dirs = info.request.role_log_dirs["trainer"]
    json.dump({"message": {"message": "first", "errorCode": 7, "extraInfo": {"timestamp": 1}}}, open(os.path.join(dirs[1], "error.json"), "w"))
    time.sleep(0.05)
    json.dump({"message": {"message": "second", "errorCode": 9, "extraInfo": {"timestamp": 2}}}, open(os.path.join(dirs[0], "error.json"), "w"))
    d = _wait(sched, app_id)
    assert json.loads(d.structured_error_msg)["message"]["message"] == "first"


def test_logs_stdout_stderr_combined_and_tailing(sched, image, tmp_path):
    roles = [_role(image, "echo_stdout.sh", "hello_out", name="o"), _role(image, "echo_range.sh", "20", name="e")]
    app_id = sched.submit(AppDef("t", roles=roles), {"log_dir": str(tmp_path / "logs")})
    assert "".join(sched.log_iter(app_id, "o", 0, streams=Stream.STDOUT)) == "hello_out\n"
    lines = list(sched.log_iter(app_id, "e", 0, should_tail=True))  # combined is the default
    assert lines == [f"{i}\n" for i in range(21)]
    assert list(sched.log_iter(app_id, "e", 0, regex=r"^1\d$", streams=Stream.STDERR)) == [f"{i}\n" for i in range(10, 20)]
    assert list(sched.log_iter(app_id, "e", 0, streams=Stream.STDOUT)) == []
    with pytest.raises(RuntimeError):
        list(sched.log_iter(app_id, "e", 5))
    base = os.path.join(str(tmp_path / "logs"), "test_session", app_id, "e", "0")
    assert sorted(os.listdir(base)) == ["combined.log", "stderr.log", "stdout.log"]


def test_dryrun_is_pure_and_log_files_never_clobbered(sched, image, tmp_path):
    app = AppDef("t", roles=[_role(image, "echo_stdout.sh", "x")])
    info = sched.submit_dryrun(app, {"log_dir": str(tmp_path / "logs")})
    assert not os.path.exists(info.request.log_dir)
    app_id = sched.schedule(info)
    _wait(sched, app_id)
    with pytest.raises(FileExistsError):  # an existing log file is never overwritten
        sched._get_file_io(info.request.role_params["trainer"][0].stdout)


def test_cache_evicts_only_terminal_apps(image, tmp_path):
    s = LocalScheduler("c", image_provider_class=LocalDirectoryImageProvider, cache_size=2)
    try:
        cfg = {"log_dir": str(tmp_path / "logs")}
        a = s.submit(AppDef("a", roles=[_role(image, "echo_stdout.sh", "1")]), cfg)
        b = s.submit(AppDef("b", roles=[_role(image, "sleep.sh", "60")]), cfg)
        _wait(s, a)
        c = s.submit(AppDef("c", roles=[_role(image, "sleep.sh", "60")]), cfg)  # evicts finished `a`
        assert s.describe(a) is None and s.describe(b) is not None and s.describe(c) is not None
        with pytest.raises(IndexError):
            s.submit(AppDef("d", roles=[_role(image, "echo_stdout.sh", "1")]), cfg)  # both cached apps still run
        with pytest.raises(ValueError):
            LocalScheduler("x", image_provider_class=LocalDirectoryImageProvider, cache_size=0)
    finally:
        s.close()


def test_close_kills_children_is_idempotent_and_unknown_apps(sched, image, tmp_path):
    app_id = sched.submit(AppDef("t", roles=[_role(image, "sleep.sh", "60", n=2)]), {"log_dir": str(tmp_path / "logs")})
    pids = [r.proc.pid for r in sched._apps[app_id].replicas()]
    sched.close()
    sched.close()
    for pid in pids:
        with pytest.raises(ProcessLookupError):
            os.kill(pid, 0)
    assert sched.describe("nope") is None and not sched.exists("nope")
    sched.cancel("nope")  # no-op
    with pytest.raises(Exception):
        sched.list()


def test_temp_log_dir_created_when_unset_and_removed_on_close(image):
    s = LocalScheduler("t", image_provider_class=LocalDirectoryImageProvider)
    app_id = s.submit(AppDef("t", roles=[_role(image, "echo_stdout.sh", "x")]), {})
    _wait(s, app_id)
    base = s._base_log_dir
    assert base and os.path.isdir(base)
    s.close()
    assert not os.path.exists(base)


def test_sigterm_to_launcher_leaves_no_orphans(image, tmp_path):
    """A launcher killed with SIGTERM must take its replicas with it (reference local_scheduler_test.py:1115-1141)."""
    pidfile = tmp_path / "pids"
    code = f"""
import os, sys, time
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
from torchx_b200.schedulers.local_scheduler import LocalScheduler, LocalDirectoryImageProvider
from torchx_b200.specs import AppDef, Role, Resource
s = LocalScheduler("orphan", image_provider_class=LocalDirectoryImageProvider)
try:
    app_id = s.submit(AppDef("t", roles=[Role(name="r", image={image!r}, entrypoint="sleep.sh", args=["120"], num_replicas=2, resource=Resource(1,0,1))]),
                      {{"log_dir": {str(tmp_path / 'logs')!r}}})
    open({str(pidfile)!r}, "w").write(" ".join(str(r.proc.pid) for r in s._apps[app_id].replicas()))
    time.sleep(120)
finally:
    s.close()
"""
    p = subprocess.Popen([sys.executable, "-c", code])
    for _ in range(200):
        if pidfile.exists() and pidfile.read_text().strip():
            break
        time.sleep(0.05)
    pids = [int(x) for x in pidfile.read_text().split()]
    assert len(pids) == 2
    p.send_signal(signal.SIGTERM)
    p.wait(timeout=30)
    for pid in pids:
        for _ in range(100):
            try:
                os.kill(pid, 0)
            except ProcessLookupError:
                break
            time.sleep(0.05)
        else:
            pytest.fail(f"replica {pid} survived its launcher")


def test_factory_accepts_runner_kwargs():
    s = create_scheduler("sess", cache_size=7, some_torchx_env_param="ignored")
    try:
        assert s.session_name == "sess" and s.backend == "local"
        assert {k for k, _ in s.run_opts()} == {"log_dir", "prepend_cwd", "auto_set_cuda_visible_devices"}
    finally:
        s.close()
