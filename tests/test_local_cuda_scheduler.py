"""``local_cuda`` scheduler on CPU: device count / NUMA topology mocked, real worker processes (python stubs)."""
import json
import os
import sys
import time
from unittest import mock

import pytest

from torchx_b200.components.dist import ddp
from torchx_b200.components.utils import echo
from torchx_b200.runner import get_runner
from torchx_b200.schedulers import get_scheduler_factories
from torchx_b200.schedulers.api import Stream
from torchx_b200.schedulers.local_cuda_scheduler import LocalCudaScheduler, create_scheduler, parse_torchrun
from torchx_b200.specs import AppState, is_terminal

WORKER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "workers", "env_worker.py")


@pytest.fixture()
def sched():
    s = create_scheduler("sess")
    yield s
    s.close()


def _wait(s, app_id, timeout=60):
    end = time.time() + timeout
    while time.time() < end:
        d = s.describe(app_id)
        if d is None or is_terminal(d.state):
            return d
        time.sleep(0.05)
    raise TimeoutError(app_id)


def _envs(out_dir):
    return {f[:-5]: json.load(open(os.path.join(out_dir, f))) for f in sorted(os.listdir(out_dir)) if f.endswith(".json")}


def test_parse_torchrun_forms():
    s = parse_torchrun("torchrun --rdzv_backend c10d --rdzv_endpoint localhost:0 --rdzv_id 'app-1' --nnodes 1 --nproc_per_node 8 --tee 3 --role '' x.py --lr 0.1 -m notmodule")
    assert (s.nproc_per_node, s.min_nnodes, s.max_nnodes, s.rdzv_id, s.tee, s.role) == ("8", 1, 1, "app-1", 3, "")
    assert s.script == "x.py" and s.script_args == ["--lr", "0.1", "-m", "notmodule"] and s.worker_cmd()[:2] == [sys.executable, "-u"]
    s = parse_torchrun("torchrun --nnodes 1:2 --nproc-per-node=4 --max-restarts 3 -m pkg.train --epochs 2")
    assert (s.min_nnodes, s.max_nnodes, s.nproc_per_node, s.max_restarts, s.module, s.script_args) == (1, 2, "4", 3, "pkg.train", ["--epochs", "2"])
    s = parse_torchrun("torchrun --rdzv_endpoint ${TORCHX_RANK0_HOST:=localhost}:29500 --nnodes 2 --nproc_per_node 2 --node_rank 1 t.py")
    assert s.node_rank == 1 and s.script == "t.py"
    assert parse_torchrun("echo hi") is None and parse_torchrun("python train.py") is None
    assert parse_torchrun(f"{sys.executable} -m torch.distributed.run --nproc_per_node 2 t.py").script == "t.py"
    with pytest.raises(ValueError):
        parse_torchrun("torchrun --nproc_per_node 2")
    with pytest.raises(ValueError):
        parse_torchrun("torchrun --some_future_flag t.py")


def test_registered_as_default_scheduler_next_to_local_cwd():
    names = list(get_scheduler_factories())
    assert names[0] == "local_cuda" and "local_cwd" in names
    s = get_scheduler_factories()["local_cuda"]("x", foo="bar")  # Runner passes TORCHX_* env as kwargs
    assert isinstance(s, LocalCudaScheduler) and s.backend == "local_cuda"
    s.close()


def test_dryrun_pins_one_worker_per_gpu_and_is_pure(sched, tmp_path):
    numa = {d: list(range(0, 8)) if d < 4 else list(range(8, 16)) for d in range(8)}
    with mock.patch.object(sched, "_cuda_device_count", return_value=8), mock.patch.object(sched, "_numa_cpus", side_effect=lambda d: numa[d]):
        app = ddp("--x", "1", script=WORKER, j="1x8", gpu=8)
        info = sched.submit_dryrun(app, {"log_dir": str(tmp_path / "logs")})
    req = info.request
    assert not os.path.exists(req.log_dir)
    (g,) = req.groups["env_worker"]
    assert g.is_torchrun and g.nproc == 8 and g.devices == list(range(8)) and g.world_size == 8 and g.rank_offset == 0
    assert g.cmd == [sys.executable, "-u", WORKER, "--x", "1"]
    assert g.cpu_sets[0] == [0, 1] and g.cpu_sets[3] == [6, 7] and g.cpu_sets[4] == [8, 9] and g.cpu_sets[7] == [14, 15]
    assert "CUDA_VISIBLE_DEVICES" not in g.env  # one node: all peers stay visible, LOCAL_RANK == device ordinal
    assert req.shm_name == f"/b2_{req.app_id}" and "groups" in repr(info)
    # explicit device list + multi-node on one box: each node gets its own visible set
    with mock.patch.object(sched, "_cuda_device_count", return_value=8), mock.patch.object(sched, "_numa_cpus", return_value=[]):
        info = sched.submit_dryrun(ddp(script=WORKER, j="2x2", gpu=2), {"devices": ["4", "5", "6", "7"]})
    g0, g1 = info.request.groups["env_worker"]
    assert (g0.env["CUDA_VISIBLE_DEVICES"], g1.env["CUDA_VISIBLE_DEVICES"]) == ("4,5", "6,7")
    assert (g0.devices, g1.devices, g0.rank_offset, g1.rank_offset, g1.world_size, g1.group_world_size) == ([0, 1], [0, 1], 0, 2, 4, 2)
    with mock.patch.object(sched, "_cuda_device_count", return_value=2):
        with pytest.raises(ValueError, match="needs 4 GPUs"):
            sched.submit_dryrun(ddp(script=WORKER, j="1x4", gpu=4), {})


def test_workers_get_the_torchrun_env_contract_and_logs_are_merged(sched, tmp_path):
    out = tmp_path / "out"
    out.mkdir()
    with mock.patch.object(sched, "_cuda_device_count", return_value=0):
        app = ddp(str(out), script=WORKER, j="1x2", env={"MY_ENV": "v"}, name="exp/run7")
        info = sched.submit_dryrun(app, {"log_dir": str(tmp_path / "logs")})
    app_id = sched.schedule(info)
    d = _wait(sched, app_id)
    assert d.state == AppState.SUCCEEDED and d.num_restarts == 0
    envs = _envs(str(out))
    assert sorted(envs) == ["attempt0_rank0", "attempt0_rank1"]
    for r in (0, 1):
        e = envs[f"attempt0_rank{r}"]
        assert (e["RANK"], e["LOCAL_RANK"], e["WORLD_SIZE"], e["LOCAL_WORLD_SIZE"]) == (str(r), str(r), "2", "2")
        assert (e["GROUP_RANK"], e["GROUP_WORLD_SIZE"], e["ROLE_RANK"], e["ROLE_WORLD_SIZE"], e["ROLE_NAME"]) == ("0", "1", str(r), "2", "")
        assert e["MASTER_ADDR"] == "127.0.0.1" and int(e["MASTER_PORT"]) > 0
        assert (e["TORCHELASTIC_RESTART_COUNT"], e["TORCHELASTIC_MAX_RESTARTS"], e["TORCHELASTIC_RUN_ID"]) == ("0", "0", app_id)
        assert e["TORCHELASTIC_USE_AGENT_STORE"] == "False" and e["OMP_NUM_THREADS"] == "1"
        assert e["B2_SHM_NAME"] == f"/b2_{app_id}_env_worker" and e["B2_EPOCH"] == "0" and e["B2_DEVICE"] is None
        assert e["TORCHELASTIC_ERROR_FILE"].endswith(f"env_worker/0/attempt_0/rank_{r}/error.json")
        assert e["MY_ENV"] == "v" and e["TORCHX_RANK0_HOST"] == "localhost" and e["TORCHX_TRACKING_RUN_NAME"] == "run7"
        assert e["argv"] == [str(out)]
    assert envs["attempt0_rank0"]["MASTER_PORT"] == envs["attempt0_rank1"]["MASTER_PORT"]
    # replica-level logs carry the [rank]: prefix like torchrun --tee; per-worker files exist underneath
    outl = sorted(sched.log_iter(app_id, "env_worker", 0, streams=Stream.STDOUT))
    assert outl == ["[0]:hello from rank 0 attempt 0\n", "[1]:hello from rank 1 attempt 0\n"]
    errl = sorted(sched.log_iter(app_id, "env_worker", 0, streams=Stream.STDERR))
    assert errl == ["[0]:warn from rank 0\n", "[1]:warn from rank 1\n"]
    assert len(list(sched.log_iter(app_id, "env_worker", 0))) == 4
    rdir = info.request.role_log_dirs["env_worker"][0]
    assert open(os.path.join(rdir, "attempt_0", "rank_1", "stdout.log")).read() == "hello from rank 1 attempt 0\n"
    manifest = json.load(open(os.path.join(info.request.log_dir, "SUCCESS")))
    assert manifest["final_state"] == "SUCCEEDED" and len(manifest["roles"]["env_worker"]) == 2


def test_gang_is_relaunched_with_next_epoch_when_retries_remain(sched, tmp_path):
    out = tmp_path / "out"
    out.mkdir()
    with mock.patch.object(sched, "_cuda_device_count", return_value=0):
        app = ddp(str(out), "fail_rank1_first_attempt", script=WORKER, j="1x2", max_retries=2)
        app_id = sched.submit(app, {"log_dir": str(tmp_path / "logs")})
    t0 = time.time()
    d = _wait(sched, app_id)
    assert d.state == AppState.SUCCEEDED and d.num_restarts == 1
    assert time.time() - t0 < 30  # the hung survivor was torn down, not waited for
    envs = _envs(str(out))
    assert sorted(envs) == ["attempt0_rank0", "attempt0_rank1", "attempt1_rank0", "attempt1_rank1"]
    assert envs["attempt1_rank0"]["B2_EPOCH"] == "1" and envs["attempt1_rank0"]["TORCHELASTIC_RESTART_COUNT"] == "1"
    assert envs["attempt1_rank0"]["TORCHELASTIC_MAX_RESTARTS"] == "2"
    lines = list(sched.log_iter(app_id, "env_worker", 0, streams=Stream.STDOUT))
    assert "[1]:hello from rank 1 attempt 0\n" in lines and "[1]:hello from rank 1 attempt 1\n" in lines


def test_failure_without_retries_fails_fast_with_root_cause(sched, tmp_path):
    out = tmp_path / "out"
    out.mkdir()
    with mock.patch.object(sched, "_cuda_device_count", return_value=0):
        app_id = sched.submit(ddp(str(out), "fail_rank1_first_attempt", script=WORKER, j="1x2"), {"log_dir": str(tmp_path / "logs")})
    d = _wait(sched, app_id)
    assert d.state == AppState.FAILED and d.num_restarts == 0
    assert json.loads(d.structured_error_msg)["message"]["message"] == "boom on rank 1"
    assert "exited with code 13" in d.msg


def test_non_torchrun_roles_run_like_local_cwd_and_cancel_works(sched, tmp_path):
    app_id = sched.submit(echo(msg="plain", num_replicas=2), {"log_dir": str(tmp_path / "logs")})
    assert _wait(sched, app_id).state == AppState.SUCCEEDED
    assert list(sched.log_iter(app_id, "echo", 1, streams=Stream.STDOUT)) == ["plain\n"]
    out = tmp_path / "out"
    out.mkdir()
    with mock.patch.object(sched, "_cuda_device_count", return_value=0):
        app_id = sched.submit(ddp(str(out), "sleep", script=WORKER, j="1x2"), {"log_dir": str(tmp_path / "logs")})
    time.sleep(1.0)
    assert sched.describe(app_id).state == AppState.RUNNING
    pids = [r.proc.pid for r in sched._apps[app_id].replicas()]
    sched.cancel(app_id)
    assert sched.describe(app_id).state == AppState.CANCELLED
    for pid in pids:
        with pytest.raises(ProcessLookupError):
            os.kill(pid, 0)


def test_world_size_2_gloo_ddp_through_runner_and_cli_component(tmp_path):
    """BASELINE config #1 on the new scheduler: 2 ranks, CPU/gloo, stock DDP, bit-exact gradient check inside the
    worker (examples/toy_ddp.py exits non-zero on mismatch)."""
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "toy_ddp.py")
    with get_runner() as runner:
        handle = runner.run_component("dist.ddp", ["-j", "1x2", "--script", script], "local_cuda", cfg={"log_dir": str(tmp_path / "logs"), "devices": []})
        status = runner.wait(handle, wait_interval=0.2)
        assert status is not None and status.state == AppState.SUCCEEDED, status
        lines = list(runner.log_lines(handle, "toy_ddp", 0))
        assert sum("exact=True same_on_all_ranks=True" in ln for ln in lines) == 2
        assert handle.startswith("local_cuda://torchx/toy_ddp-")


def test_status_list_and_logs_resolve_from_another_process(tmp_path, monkeypatch):
    """The app registry: a second scheduler instance (another `torchx status` invocation) sees apps it did not launch."""
    monkeypatch.setenv("TORCHX_HOME", str(tmp_path / "home"))
    out = tmp_path / "out"
    out.mkdir()
    a = create_scheduler("sess")
    try:
        with mock.patch.object(a, "_cuda_device_count", return_value=0):
            app_id = a.submit(ddp(str(out), script=WORKER, j="1x2"), {"log_dir": str(tmp_path / "logs")})
        b = create_scheduler("sess")  # stands in for a different process
        try:
            assert b.describe(app_id).state in (AppState.RUNNING, AppState.SUCCEEDED)
            assert _wait(a, app_id).state == AppState.SUCCEEDED
            d = b.describe(app_id)
            assert d.state == AppState.SUCCEEDED and d.ui_url.endswith(app_id)
            assert sorted(b.log_iter(app_id, "env_worker", 0, streams=Stream.STDOUT)) == ["[0]:hello from rank 0 attempt 0\n", "[1]:hello from rank 1 attempt 0\n"]
            assert [(r.app_id, r.state) for r in b.list()] == [(app_id, AppState.SUCCEEDED)]
            assert b.describe("never-launched") is None
            b.cancel(app_id)  # already terminal: a no-op, also from here
            assert b.describe(app_id).state == AppState.SUCCEEDED
        finally:
            b.close()
        assert create_scheduler("other_session").list() == []
    finally:
        a.close()


def test_cancel_from_another_process_goes_through_the_registry(tmp_path, monkeypatch):
    """`torchx cancel local_cuda://...` typed in a second shell: the request file is picked up by the launcher's supervisor
    thread, which takes the gang down and records CANCELLED (the reference's local scheduler cannot do this:
    local_scheduler.py:1099-1102 knows only the apps of its own process)."""
    monkeypatch.setenv("TORCHX_HOME", str(tmp_path / "home"))
    out = tmp_path / "out"
    out.mkdir()
    a, b = create_scheduler("sess"), create_scheduler("sess")  # launcher / the other process
    try:
        with mock.patch.object(a, "_cuda_device_count", return_value=0):
            app_id = a.submit(ddp(str(out), "sleep", script=WORKER, j="1x2"), {"log_dir": str(tmp_path / "logs")})
        time.sleep(1.0)
        pids = [r.proc.pid for r in a._apps[app_id].replicas()]
        assert b.describe(app_id).state == AppState.RUNNING
        b.cancel(app_id)  # returns once the launcher has acknowledged, i.e. the workers are gone
        for pid in pids:
            with pytest.raises(ProcessLookupError):
                os.kill(pid, 0)
        assert a.describe(app_id).state == AppState.CANCELLED and b.describe(app_id).state == AppState.CANCELLED
        assert not os.path.exists(os.path.join(str(tmp_path / "home"), "apps", "sess", f"{app_id}.cancel"))
        with pytest.raises(RuntimeError, match="not in this session's registry"):
            b._cancel_existing("never-launched")
    finally:
        b.close()
        a.close()


def test_two_nodes_of_two_workers_on_one_box_gloo(tmp_path):
    """The reference's own distributed integration case is `dist.ddp -j 2x2` on one host
    (torchx/components/integration_tests/component_provider.py:39-51): two replicas ("nodes") x two workers, world 4."""
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "toy_ddp.py")
    with get_runner() as runner:
        handle = runner.run_component("dist.ddp", ["-j", "2x2", "--script", script], "local_cuda", cfg={"log_dir": str(tmp_path / "logs"), "devices": []})
        status = runner.wait(handle, wait_interval=0.2)
        assert status is not None and status.state == AppState.SUCCEEDED, status
        per_node = [list(runner.log_lines(handle, "toy_ddp", k)) for k in (0, 1)]
        ranks = sorted(int(m) for lines in per_node for ln in lines for m in __import__("re").findall(r"rank (\d)/4", ln))
        assert ranks == [0, 1, 2, 3]
        assert all("same_on_all_ranks=True" in ln for lines in per_node for ln in lines if "sha256" in ln)
        assert any(ln.startswith("[2]:") or ln.startswith("[3]:") for ln in per_node[1])  # node 1 hosts global ranks 2, 3


def test_gang_supervision_timeout_and_sentinel_files(sched, tmp_path):
    """process_monitor semantics for the whole gang (reference torchx/apps/utils/process_monitor.py:65-118): a job_timeout
    terminates every worker and fails the app; exit_on_file does the same on demand; start_on_file holds the launch."""
    out = tmp_path / "out"
    out.mkdir()
    with mock.patch.object(sched, "_cuda_device_count", return_value=0):
        t0 = time.time()
        app_id = sched.submit(ddp(str(out), "sleep", script=WORKER, j="1x2"), {"log_dir": str(tmp_path / "l1"), "job_timeout": 1.0})
        time.sleep(0.3)
        pids = [r.proc.pid for r in sched._apps[app_id].replicas()]
        d = _wait(sched, app_id)
        assert d.state == AppState.FAILED and "job timeout" in d.msg and time.time() - t0 < 30
        assert sched._apps[app_id].exit_code == 34
        for pid in pids:
            with pytest.raises(ProcessLookupError):
                os.kill(pid, 0)
        stop = tmp_path / "stop"
        app_id = sched.submit(ddp(str(out), "sleep", script=WORKER, j="1x2"), {"log_dir": str(tmp_path / "l2"), "exit_on_file": str(stop)})
        time.sleep(0.5)
        assert sched.describe(app_id).state == AppState.RUNNING
        stop.write_text("now")
        d = _wait(sched, app_id)
        assert d.state == AppState.FAILED and str(stop) in d.msg
        go = tmp_path / "go"
        app_id = sched.submit(ddp(str(out), "ok", script=WORKER, j="1x2"), {"log_dir": str(tmp_path / "l3"), "start_on_file": str(go)})
        time.sleep(0.5)
        assert sched.describe(app_id).state == AppState.PENDING and not sched._apps[app_id].replicas()
        go.write_text("go")
        assert _wait(sched, app_id).state == AppState.SUCCEEDED
        app_id = sched.submit(ddp(str(out), "ok", script=WORKER, j="1x2"),
                              {"log_dir": str(tmp_path / "l4"), "start_on_file": str(tmp_path / "never"), "job_timeout": 0.5})
        d = _wait(sched, app_id)
        assert d.state == AppState.FAILED and "before launching" in d.msg


def test_process_monitor_module(tmp_path):
    from torchx_b200.apps.utils import process_monitor as pm

    assert pm.supervise(sys.executable, ["-c", "import sys; sys.exit(7)"], poll_rate=0.1) == 7
    t0 = time.time()
    rc = pm.supervise(sys.executable, ["-c", "import time; time.sleep(60)"], timeout=0.5, poll_rate=0.1, kill_timeout=5)
    assert rc != 0 and time.time() - t0 < 20
    assert pm.supervise(sys.executable, ["-c", "pass"], timeout=0.3, start_on_file=str(tmp_path / "nope"), poll_rate=0.1) == pm.TIMEOUT_EXIT_CODE
    stop = tmp_path / "stop"
    stop.write_text("x")
    assert pm.supervise(sys.executable, ["-c", "import time; time.sleep(60)"], exit_on_file=str(stop), poll_rate=0.1, kill_timeout=5) != 0
