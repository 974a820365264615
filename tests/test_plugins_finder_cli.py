"""Plugin registry, component finder, CLI argument plumbing and the small spec additions (CPU).  Modelled on the
reference's plugins/test/{register,registry}_test.py, specs/test/finder_test.py, cli/test/{argparse_util,cmd_run}_test.py,
util/test/entrypoints_test.py and specs/test/api_test.py; tools/run_reference_tests.py runs those files themselves."""
import argparse
import asyncio
import io
import json
import os
import sys
import textwrap
import warnings
from importlib.metadata import EntryPoint, EntryPoints
from unittest import mock

import pytest
import yaml

from torchx_b200 import plugins, specs
from torchx_b200.cli import argparse_util
from torchx_b200.cli.cmd_run import CmdRun, TorchXRunArgs, torchx_run_args_from_json
from torchx_b200.plugins import PluginRegistry, PluginSource, PluginType, register, resource_tags
from torchx_b200.specs import finder
from torchx_b200.specs.api import AppState, AppStatus, Resource, Role, macros
from torchx_b200.util import entrypoints
from torchx_b200.util.log_tee_helpers import _find_role_replicas, _prefix_line, tee_logs


# ---- plugin registry --------------------------------------------------------------------------------------------
def _write(root, rel, body):
    path = root / rel
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(textwrap.dedent(body))


@pytest.fixture
def plugin_tree(tmp_path, monkeypatch):
    """Two 'distributions' contributing to the torchx_b200_plugins namespace (PEP 420: no __init__.py anywhere)."""
    a, b = tmp_path / "dist_a", tmp_path / "dist_b"
    _write(a, "torchx_b200_plugins/schedulers/local.py", """
        from torchx_b200.plugins import register
        @register.scheduler()
        def my_local(session_name, **kwargs):
            return ("my_local", session_name)
        @register.scheduler(name="renamed")
        def create_other(session_name, **kwargs):
            return ("renamed", session_name)
        def not_a_plugin():
            pass
    """)
    _write(a, "torchx_b200_plugins/schedulers/_private.py", "raise RuntimeError('private modules are never imported')\n")
    _write(a, "torchx_b200_plugins/schedulers/broken.py", "raise RuntimeError('cannot reach the cluster API')\n")
    _write(a, "torchx_b200_plugins/schedulers/misplaced.py", """
        from torchx_b200.plugins import register
        from torchx_b200.specs import Resource
        @register.named_resource()
        def stray():
            return Resource(cpu=1, gpu=0, memMB=1)
    """)
    _write(b, "torchx_b200_plugins/schedulers/cloud/k8s.py", """
        from torchx_b200.plugins import register
        @register.scheduler()
        def cloud_k8s(session_name, **kwargs):
            return ("cloud_k8s", session_name)
    """)
    _write(b, "torchx_b200_plugins/named_resources/box.py", """
        from torchx_b200.plugins import register, powers_of_two_gpus
        from torchx_b200.specs import Resource
        @register.named_resource(aliases=["hgx"], fractionals=powers_of_two_gpus)
        def b200_box(fractional: float = 1.0):
            return Resource(cpu=int(192 * fractional), gpu=int(8 * fractional), memMB=int(2048 * 1024 * fractional))
    """)
    monkeypatch.syspath_prepend(str(a))
    monkeypatch.syspath_prepend(str(b))
    plugins.reset_for_tests()
    yield tmp_path
    plugins.reset_for_tests()


def test_registry_discovers_namespace_plugins_and_reports_problems(plugin_tree):
    reg = PluginRegistry(plugin_sources=PluginSource.NAMESPACE_PKG)
    scheds = reg.get(PluginType.SCHEDULER)
    assert sorted(scheds) == ["cloud_k8s", "my_local", "renamed"]  # nested implicit namespace sub-package found too
    assert scheds["renamed"]("s") == ("renamed", "s")
    assert reg.get(PluginType.SCHEDULER) is scheds  # cached
    errs = {(e.module.rsplit(".", 1)[-1], e.name) for e in reg.errors}
    assert ("broken", None) in errs and ("misplaced", "stray") in errs and not any(m == "_private" for m, _ in errs)
    report = yaml.safe_load(str(reg))  # the printable report is valid YAML
    assert sorted(s["name"] for s in report["scheduler"]) == ["cloud_k8s", "my_local", "renamed"]
    assert report["errors"][0]["module"].endswith("schedulers.broken") and "cluster API" in report["errors"][0]["error"]
    stray = [r for r in report["named_resource"] if r.get("error")][0]
    assert stray["name"] == "stray" and "under the scheduler namespace" in stray["error"]
    box = [r for r in report["named_resource"] if r["name"] == "b200_box"][0]
    assert box["aliases"] == ["hgx"] and sorted(box["fractionals"]) == ["b200_box_1", "b200_box_2", "b200_box_4", "b200_box_8"]


def test_named_resource_aliases_fractionals_and_tags(plugin_tree):
    table = plugins.registry().get(PluginType.NAMED_RESOURCE)
    whole, half, alias = table["b200_box"](), table["b200_box_4"](), table["hgx"]()
    assert (whole.gpu, half.gpu, half.cpu) == (8, 4, 96) and alias == whole
    assert whole.get_resource_name() == "b200_box" and not whole.is_fractional()
    assert half.get_resource_name() == "b200_box_4" and half.is_fractional()
    assert not table["b200_box_8"]().is_fractional()  # the 1.0 "slice" is the whole host
    assert specs.named_resources["b200_box_2"].gpu == 2 and "hgx" in specs.named_resources  # visible through specs
    assert specs.resource(h="b200_box_1").tags[resource_tags.RESOURCE_NAME] == "b200_box_1"
    with pytest.raises(KeyError, match="Did you mean `b200_box_8`"):
        specs.named_resources["b200_box_88"]
    assert Resource(cpu=1, gpu=0, memMB=1).get_resource_name() is None


def test_registered_schedulers_replace_the_defaults_and_sources_are_selectable(plugin_tree, monkeypatch):
    from torchx_b200.schedulers import get_scheduler_factories

    assert sorted(get_scheduler_factories()) == ["cloud_k8s", "my_local", "renamed"]
    ep = {"ep_only": lambda session_name, **kw: ("ep", session_name), "my_local": lambda session_name, **kw: ("ep wins", session_name)}
    for value, want_ns, want_ep in (("0", False, False), ("1", True, False), ("2", False, True), ("3", True, True)):
        plugins.registry().clear()
        monkeypatch.setenv("TORCHX_PLUGINS_SOURCE", value)
        with mock.patch.object(entrypoints, "load_group", return_value=ep):
            found = plugins.registry().get(PluginType.SCHEDULER)
        assert ("renamed" in found) == want_ns and ("ep_only" in found) == want_ep
        if want_ns and want_ep:
            assert found["my_local"]("s") == ("ep wins", "s")  # entry points override namespace plugins
    for bad in ("namespace", "-1", "4"):
        plugins.registry.cache_clear()
        monkeypatch.setenv("TORCHX_PLUGINS_SOURCE", bad)
        with pytest.raises(ValueError, match="TORCHX_PLUGINS_SOURCE"):
            plugins.registry()
    monkeypatch.delenv("TORCHX_PLUGINS_SOURCE")
    plugins.registry.cache_clear()


def test_fractional_generators_and_duplicate_registration():
    box = Resource(cpu=8, gpu=8, memMB=64 * 1024)
    assert plugins.powers_of_two_gpus(box) == {1.0: "8", 0.5: "4", 0.25: "2", 0.125: "1"}
    assert plugins.halve_mem_down_to(minGiB=8)(box) == {1.0: "64", 0.5: "32", 0.25: "16", 0.125: "8"}
    assert plugins.halve_mem_down_to(minGiB=3)(Resource(cpu=1, gpu=0, memMB=48 * 1024)) == {1.0: "48", 0.5: "24", 0.25: "12", 0.125: "6", 0.0625: "3"}
    for bad in (Resource(cpu=1, gpu=0, memMB=1024), Resource(cpu=1, gpu=6, memMB=1024)):
        with pytest.raises(ValueError):
            plugins.powers_of_two_gpus(bad)
    with pytest.raises(ValueError, match="odd part"):
        plugins.halve_mem_down_to(minGiB=2)(Resource(cpu=1, gpu=0, memMB=48 * 1024))
    with pytest.raises(ValueError, match="whole number of GiB"):
        plugins.halve_mem_down_to(minGiB=1)(Resource(cpu=1, gpu=0, memMB=1000))

    def res():
        return Resource(cpu=1, gpu=0, memMB=1)

    register.named_resource(name="dup_check")(res)
    try:
        with pytest.raises(ValueError, match="duplicate named resource `dup_check`"):
            register.named_resource(name="dup_check")(res)
        assert sys.modules[__name__].NAMED_RESOURCES["dup_check"]().get_resource_name() == "dup_check"  # legacy per-module table
    finally:
        del sys.modules[__name__].NAMED_RESOURCES, sys.modules[__name__].dup_check


# ---- entry points -----------------------------------------------------------------------------------------------
def _eps(text):
    import configparser

    cp = configparser.ConfigParser(delimiters="=")
    cp.read_string(textwrap.dedent(text))
    return EntryPoints(EntryPoint(n, v, g) for g in cp.sections() for n, v in cp.items(g))


def helper_for_entry_points(x="nothing"):
    return f"called with {x}"


def test_entry_point_groups_load_lazily():
    me = __name__
    eps = _eps(f"""
        [grp.test]
        fn = {me}:helper_for_entry_points
        mod = {me}
        missing = {me}.no_such_module
    """)
    with mock.patch("torchx_b200.util.entrypoints.metadata.entry_points", return_value=eps):
        group = entrypoints.load_group("grp.test")
        assert sorted(group) == ["fn", "missing", "mod"]  # nothing imported yet, so the broken one is listed too
        assert group["fn"]("an arg") == "called with an arg" and group["mod"]("ignored") is sys.modules[me]
        with pytest.raises(ModuleNotFoundError):
            group["missing"]()
        assert entrypoints.load_group("no.such.group") is None and entrypoints.load_group("no.such.group", default={"a": 1}) == {"a": 1}
        assert entrypoints.load("grp.test", "fn")() == "called with nothing"
        assert entrypoints.load("grp.test", "absent", default="dflt") == "dflt"
        with pytest.raises(KeyError):
            entrypoints.load("grp.test", "absent")


# ---- component finder -------------------------------------------------------------------------------------------
@pytest.fixture
def fresh_finder():
    finder._components = None
    yield
    finder._components = None


def test_builtin_table_and_source_template(fresh_finder, tmp_path):
    table = finder.get_components()
    assert {"dist.ddp", "utils.echo", "utils.sh"} <= set(table) and all(not c.validation_errors for c in table.values())
    assert table.keys() == {c.name for c in finder.ModuleComponentsFinder("torchx_b200.components", group="").find(None)}
    ddp = finder.get_component("dist.ddp")
    assert (ddp.fn_name, ddp.name) == ("ddp", "dist.ddp") and ddp.description
    # `torchx builtins --print utils.echo > copy.py` gives a file that works as a component on its own
    copy = tmp_path / "echo_copy.py"
    copy.write_text(finder.get_builtin_source("utils.echo"))
    app = finder.get_component(f"{copy}:echo").fn(msg="from the copy")
    assert app.roles[0].args[-1] == "from the copy" or "from the copy" in " ".join(app.roles[0].args)


def test_registered_component_modules_replace_builtins(fresh_finder, tmp_path, monkeypatch):
    _write(tmp_path, "mycomps/__init__.py", "")
    _write(tmp_path, "mycomps/train.py", """
        from typing import Dict, List, Optional
        from torchx_b200.specs import AppDef, Role
        def good(name: str, nodes: int = 1, tags: Optional[List[str]] = None, env: Optional[Dict[str, str]] = None) -> AppDef:
            \"\"\"A good one.

            Args:
                name: app name

            Returns:
                the app
            \"\"\"
            return AppDef(name, roles=[Role(name="r", image="i", entrypoint="e")])
        def untyped(name, nodes: int = 1) -> AppDef:
            return AppDef("x")
        def bad_type(when: complex) -> AppDef:
            return AppDef("x")
        def kwargs_only(**kw: str) -> AppDef:
            return AppDef("x")
        def helper(x: int) -> int:
            return x
    """)
    monkeypatch.syspath_prepend(str(tmp_path))
    eps = _eps("""
        [torchx_b200.components]
        mine = mycomps
        _flat = mycomps.train
    """)
    with mock.patch("torchx_b200.util.entrypoints.metadata.entry_points", return_value=eps):
        table = finder._load_components(None)
        assert sorted(table) == ["bad_type", "good", "kwargs_only", "mine.train.bad_type", "mine.train.good", "mine.train.kwargs_only",
                                 "mine.train.untyped", "untyped"]  # builtins are NOT merged in; `helper` is not a component
        assert table["good"].description == "A good one." and table["good"].validation_errors == []
        assert "Missing type annotation for argument 'name'" in table["untyped"].validation_errors[0]
        assert "Unsupported argument type" in table["bad_type"].validation_errors[0] and "**kw" in table["kwargs_only"].validation_errors[0]
        assert sorted(finder.get_components()) == ["good", "mine.train.good"]
        with pytest.raises(finder.ComponentValidationException, match="Missing type annotation"):
            finder.get_component("untyped")
        with pytest.raises(finder.ComponentNotFoundException):
            finder.get_component("dist.ddp")
    import mycomps
    import mycomps.train

    assert finder.module_relname(mycomps.train, relative_to=mycomps) == "train" and finder.module_relname(mycomps, relative_to=mycomps) == ""
    with pytest.raises(ValueError):
        finder.module_relname(mycomps, relative_to=mycomps.train)
    with pytest.raises(finder.ComponentNotFoundException):
        finder.get_component(f"{tmp_path}/mycomps/train.py:nope")
    with pytest.raises(finder.ComponentValidationException):
        finder.get_component(f"{tmp_path}/mycomps/train.py:untyped")
    assert finder.get_component("mycomps.train:good").fn_name == "good"  # importable-module form


# ---- CLI argument plumbing --------------------------------------------------------------------------------------
@pytest.fixture
def clean_actions(tmp_path, monkeypatch):
    monkeypatch.setenv("TORCHXCONFIG", str(tmp_path / ".torchxconfig"))
    (tmp_path / ".torchxconfig").write_text("")
    for cls in (argparse_util.ArgOnceAction, argparse_util.torchxconfig):
        cls.called_args = set()
    argparse_util.torchxconfig._subcmd_configs.clear()
    yield tmp_path
    argparse_util.torchxconfig._subcmd_configs.clear()


def test_torchxconfig_action_defaults_and_once_only(clean_actions):
    (clean_actions / ".torchxconfig").write_text("[cli:run]\nworkspace = from-config\nneeded = also-from-config\n")
    p = argparse.ArgumentParser()
    p.add_argument("--workspace", default="argparse-default", action=argparse_util.torchxconfig_run)
    p.add_argument("--other", default="argparse-default", action=argparse_util.torchxconfig_run)
    p.add_argument("--needed", required=True, action=argparse_util.torchxconfig_run)  # satisfied by the config file
    p.add_argument("--once", action=argparse_util.ArgOnceAction)
    ns = p.parse_args([])
    assert (ns.workspace, ns.other, ns.needed) == ("from-config", "argparse-default", "also-from-config")
    assert p.parse_args(["--workspace", "cli"]).workspace == "cli"
    with pytest.raises(SystemExit):
        p.parse_args(["--once", "a", "--once", "b"])


def _run_parser():
    cmd, p = CmdRun(), argparse.ArgumentParser()
    cmd.add_arguments(p)
    return cmd, p


def test_run_stdin_json_request(clean_actions, capsys):
    cmd, p = _run_parser()
    req = {"scheduler": "local_cuda", "scheduler_args": {"pin_cpus": False}, "component_name": "utils.echo", "component_args": {"msg": "hi"}}
    with mock.patch("sys.stdin", io.StringIO(json.dumps(req))):
        cmd.run(p.parse_args(["--stdin", "--dryrun"]))
    out = capsys.readouterr().out
    assert "=== SCHEDULER REQUEST ===" in out and "'hi'" in out and "pin_cpus" not in out or "hi" in out
    for extra in (["--wait"], ["-cfg", "a=b"], ["utils.echo"], ["--workspace", "/x"], ["-s", "local_cwd"]):
        cmd, p = _run_parser()
        for cls in (argparse_util.ArgOnceAction, argparse_util.torchxconfig):
            cls.called_args = set()
        with pytest.raises(SystemExit) as e:
            cmd.verify_no_extra_args(p.parse_args(["--stdin", *extra]))
        assert e.value.code == 2
    assert "when using --stdin" in capsys.readouterr().err
    cmd, p = _run_parser()
    with mock.patch("sys.stdin", io.StringIO("not json")), pytest.raises(SystemExit):
        cmd.torchx_json_from_stdin()
    with mock.patch("sys.stdin", io.StringIO("[1, 2]")), pytest.raises(SystemExit):
        cmd.torchx_json_from_stdin()


def test_run_args_from_json_validation():
    ok = torchx_run_args_from_json({"scheduler": "local_cuda", "scheduler_args": {}, "component_name": "dist.ddp", "tee_logs": True})
    assert isinstance(ok, TorchXRunArgs) and ok.tee_logs and ok.workspace == os.getcwd() and ok.component_args == {}
    with pytest.raises(ValueError, match="required fields are missing"):
        torchx_run_args_from_json({"scheduler": "local_cuda"})
    with pytest.raises(ValueError, match="not part of the run command"):
        torchx_run_args_from_json({"scheduler": "s", "scheduler_args": {}, "component_name": "c", "bogus": 1})


def test_sub_commands_can_be_overridden_by_entry_points():
    from torchx_b200.cli import main as cli_main
    from torchx_b200.cli.cmd_base import SubCommand

    class MyRun(SubCommand):
        def add_arguments(self, subparser):
            pass

        def run(self, args):
            pass

    with mock.patch.object(cli_main, "load_group", return_value={"run": MyRun, "extra": MyRun}):
        cmds = cli_main.get_sub_cmds()
    assert isinstance(cmds["run"], MyRun) and isinstance(cmds["extra"], MyRun) and "status" in cmds and "delete" in cmds


# ---- log tee ----------------------------------------------------------------------------------------------------
def test_prefix_line_survives_carriage_returns():
    assert _prefix_line("P ", "abc\n") == "P abc\n" and _prefix_line("P ", "a\nb\n") == "P a\nP b\n"
    assert _prefix_line("P ", "10%\r20%\r30%") == "P 10%\rP 20%\rP 30%" and _prefix_line("P ", "\rdone") == "\rP done"


def test_tee_logs_fans_out_one_reader_per_replica():
    app = specs.AppDef("a", roles=[Role("trainer", "img", num_replicas=2), Role("reader", "img", num_replicas=1)])
    assert _find_role_replicas(app, None) == [("trainer", 0), ("trainer", 1), ("reader", 0)] and _find_role_replicas(app, "reader") == [("reader", 0)]
    runner = mock.MagicMock()
    runner.describe.return_value = app
    runner.log_lines.side_effect = lambda handle, role, k, regex, should_tail, streams: [f"hello from {role} {k}\n"]
    dst = io.StringIO()
    t = tee_logs(dst, "local_cuda://s/a", None, runner)
    t.start()
    t.join(10)
    assert sorted(dst.getvalue().splitlines()) == ["reader/0 hello from reader 0", "trainer/0 hello from trainer 0", "trainer/1 hello from trainer 1"]
    assert all(c.kwargs["should_tail"] for c in runner.log_lines.call_args_list)


# ---- spec additions ---------------------------------------------------------------------------------------------
def test_torchx_home(tmp_path, monkeypatch):
    monkeypatch.setenv("TORCHX_HOME", str(tmp_path / "h"))
    assert specs.TORCHX_HOME() == tmp_path / "h" and specs.TORCHX_HOME("a", "b") == tmp_path / "h" / "a" / "b" and (tmp_path / "h" / "a" / "b").is_dir()
    monkeypatch.delenv("TORCHX_HOME")
    with mock.patch("pathlib.Path.home", return_value=tmp_path / "sally"):
        assert specs.TORCHX_HOME() == tmp_path / "sally" / ".torchx"


def test_role_lazy_overrides():
    calls = []

    def image():
        calls.append(1)
        return "resolved"

    async def entry():
        await asyncio.sleep(0)
        return "main.py"

    role = Role("r", "placeholder", overrides={"image": image, "entrypoint": entry()})
    assert calls == [] and role.image == "resolved" and role.image == "resolved" and calls == [1]  # evaluated once, on first read
    assert role.entrypoint == "main.py" and role.name == "r"
    applied = macros.Values(img_root="/i", app_id="app", replica_id="0", rank0_env="R").apply(Role("r", "x", args=["${app_id}"], overrides={"image": lambda: "late"}))
    assert applied.args == ["app"] and applied.image == "late"


def test_status_error_message_wrapping():
    st = AppStatus(state=AppState.FAILED)
    rpc = ("RuntimeError('On WorkerInfo(id=1, name=trainer:0:0):\nRuntimeError(ShardingError('Table of size 715.26GB cannot be added to any rank'))\n"
           "Traceback (most recent call last):\n..\n')\nTraceback (most recent call last):\n  File \"x.py\", line 190, in _run_function\n")
    out = st._format_error_message(rpc, header="", width=80)
    assert out.endswith("..\n')") and "x.py" not in out  # only the RPC envelope is kept
    assert out.splitlines()[1] == "RuntimeError(ShardingError('Table" and out.splitlines()[2].startswith(" of size 715.26GB")
    wrapped = st._format_error_message("word " * 40, header="    error_msg: ", width=40).splitlines()
    assert wrapped[0].startswith("    error_msg: word") and all(ln.startswith(" " * 15) for ln in wrapped[1:]) and len(wrapped) > 3
    assert st._format_error_message("boom happened\nException raised from foo at bar.cpp:1", header="") == "boom happened"


def test_local_rank_warns_when_group_is_up_without_launcher_env(monkeypatch):
    from torchx_b200 import distributed as d

    monkeypatch.delenv("LOCAL_RANK", raising=False)
    with mock.patch("torchx_b200.distributed.dist.is_initialized", return_value=True), warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert d.local_rank() == 0 and len(w) == 1 and "LOCAL_RANK" in str(w[0].message)
    with mock.patch("torchx_b200.distributed.dist.is_initialized", return_value=False), warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert d.local_rank() == 0 and not w
    monkeypatch.setenv("LOCAL_RANK", "3")
    assert d.local_rank() == 3
    with mock.patch("torch.cuda.is_available", return_value=True), mock.patch("torch.cuda.device_count", return_value=0), \
            mock.patch("torchx_b200.distributed.dist.init_process_group") as init:
        assert d.init_pg("auto").type == "cpu" and init.call_args.kwargs["backend"] == "gloo"  # CUDA build of torch on a GPU-less host


# ---- runner telemetry -------------------------------------------------------------------------------------------
def test_runner_calls_are_recorded_as_events(tmp_path):
    """Every Runner API call yields one TorchxEvent (reference torchx/runner/events): scheduler, app id, image, run config,
    timings, and the exception when the call raised.  Nothing is emitted unless a handler is installed."""
    import logging

    from torchx_b200.runner import events, get_runner

    seen = []

    class Sink(logging.Handler):
        def emit(self, record):
            seen.append(events.TorchxEvent.deserialize(record.getMessage()))

    old_handler, old_logger = events.handlers["null"], events._events_logger
    events.handlers["null"], events._events_logger = Sink(), None
    try:
        with get_runner() as runner:
            handle = runner.run_component("utils.echo", ["--msg", "x"], "local_cwd", cfg={"log_dir": str(tmp_path)})
            runner.wait(handle, wait_interval=0.1)
            with pytest.raises(specs.MalformedAppHandleException):
                runner.status("not-a-handle")
    finally:
        events.handlers["null"], events._events_logger = old_handler, old_logger
    by_api = {e.api: e for e in seen}
    app_id = handle.rsplit("/", 1)[1]
    assert {"dryrun", "schedule", "run_component", "wait", "status"} <= set(by_api)
    assert by_api["schedule"].app_id == app_id and by_api["schedule"].scheduler == "local_cwd" and by_api["schedule"].app_image
    assert json.loads(by_api["dryrun"].runcfg) == {"log_dir": str(tmp_path)}
    failed = [e for e in seen if e.exception_type]
    assert [e.exception_type for e in failed] == ["MalformedAppHandleException"] and "not-a-handle" in failed[0].exception_message
    assert json.loads(failed[0].exception_source_location)["name"] == "parse_app_handle" and "Traceback" in failed[0].raw_exception
    assert all(e.wall_time_usec >= 0 and e.cpu_time_usec >= 0 and e.session for e in seen)
    assert events.TorchxEvent.deserialize(str(seen[0])) == seen[0]


def test_tracker_configuration_is_forwarded_to_the_workers(tmp_path, monkeypatch):
    """.torchxconfig [torchx:tracker] / [tracker:<name>] (or the TORCHX_TRACKERS* variables) end up in every role's environment
    (reference torchx/runner/api.py:68-87, 386-391); the backends themselves live in the worker."""
    from torchx_b200.runner import get_runner
    from torchx_b200.runner.api import get_configured_trackers

    (tmp_path / ".torchxconfig").write_text("[torchx:tracker]\nfsspec =\nmlflow =\n\n[tracker:fsspec]\nconfig = /tmp/tracker/root\n")
    monkeypatch.setenv("TORCHXCONFIG", str(tmp_path / ".torchxconfig"))
    monkeypatch.delenv("TORCHX_TRACKERS", raising=False)
    assert get_configured_trackers() == {"fsspec": "/tmp/tracker/root", "mlflow": None}
    with get_runner() as runner:
        info = runner.dryrun_component("dist.ddp", ["-j", "1x2", "--script", "t.py"], "local_cuda", parent_run_id="exp7")
    env = info._app.roles[0].env
    assert env["TORCHX_TRACKERS"] == "fsspec,mlflow" and env["TORCHX_TRACKER_FSSPEC_CONFIG"] == "/tmp/tracker/root"
    assert "TORCHX_TRACKER_MLFLOW_CONFIG" not in env and env["TORCHX_PARENT_RUN_ID"] == "exp7"
    monkeypatch.setenv("TORCHX_TRACKERS", "only")
    monkeypatch.setenv("TORCHX_TRACKER_ONLY_CONFIG", "s3://bucket/cfg")
    assert get_configured_trackers() == {"only": "s3://bucket/cfg"}
