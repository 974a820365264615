"""``torchx run``: argv -> (scheduler, cfg, component, component args) -> submit; for ``local*`` schedulers it then
blocks, streams the replicas' logs to stderr and exits non-zero unless the app SUCCEEDED
(reference torchx/cli/cmd_run.py: _parse_component_name_and_args:120, CmdRun:206, _run_inner:281, _wait_and_exit:453)."""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys
import threading
from collections import Counter
from dataclasses import MISSING as _NO_DEFAULT
from dataclasses import asdict, dataclass, field, fields
from itertools import groupby
from pathlib import Path
from pprint import pformat
from typing import Any, Dict, List, Optional, Tuple

from torchx_b200 import specs
from torchx_b200.cli.argparse_util import ArgOnceAction, torchxconfig_run
from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.cli.cmd_log import get_logs
from torchx_b200.cli.cmd_misc import CmdBuiltins  # noqa: F401 - the reference keeps `builtins` next to `run`
from torchx_b200.runner import Runner, config, get_runner
from torchx_b200.schedulers import get_default_scheduler_name, get_scheduler_factories
from torchx_b200.specs import CfgVal
from torchx_b200.specs.finder import ComponentNotFoundException, ComponentValidationException
from torchx_b200.util.log_tee_helpers import tee_logs

logger = logging.getLogger(__name__)

MISSING_COMPONENT_ERROR_MSG = "missing component name, either provide it from the CLI or in .torchxconfig"


@dataclass
class TorchXRunArgs:
    """One ``torchx run`` request, whichever way it arrived: argv, or a JSON object on stdin (``--stdin``) whose keys are
    these field names (reference cmd_run.py:56-69).  ``scheduler_args`` is the raw JSON form of ``-cfg``;
    ``scheduler_cfg`` the resolved one.  ``component_args`` (dict) is the JSON form, ``component_args_str`` the argv form."""

    component_name: str
    scheduler: str
    scheduler_args: Dict[str, Any]
    scheduler_cfg: Dict[str, CfgVal] = field(default_factory=dict)
    dryrun: bool = False
    wait: bool = False
    log: bool = False
    workspace: str = ""
    parent_run_id: Optional[str] = None
    tee_logs: bool = False
    component_args: Dict[str, Any] = field(default_factory=dict)
    component_args_str: List[str] = field(default_factory=list)


def torchx_run_args_from_json(json_data: Dict[str, Any]) -> TorchXRunArgs:
    known = [f.name for f in fields(TorchXRunArgs)]
    required = {f.name for f in fields(TorchXRunArgs) if f.default is _NO_DEFAULT and f.default_factory is _NO_DEFAULT}
    missing = required - json_data.keys()
    if missing:
        raise ValueError(f"The following required fields are missing: {', '.join(missing)}")
    unknown = set(json_data) - set(known)
    if unknown:
        raise ValueError(f"The following fields are not part of the run command: {', '.join(unknown)}.",
                         "Please check your JSON and try launching again.")
    run_args = TorchXRunArgs(**json_data)
    if run_args.workspace == "":
        run_args.workspace = f"{Path.cwd()}"
    return run_args


def torchx_run_args_from_argparse(args: argparse.Namespace, component_name: str, component_args: List[str],
                                  scheduler_cfg: Dict[str, CfgVal]) -> TorchXRunArgs:
    return TorchXRunArgs(component_name=component_name, scheduler=args.scheduler, scheduler_args={}, scheduler_cfg=scheduler_cfg,
                         dryrun=args.dryrun, wait=args.wait, log=args.log, workspace=args.workspace, parent_run_id=args.parent_run_id,
                         tee_logs=args.tee_logs, component_args_str=component_args)


def _parse_component_name_and_args(tokens: List[str], subparser: argparse.ArgumentParser, dirs: Optional[List[str]] = None) -> Tuple[str, List[str]]:
    """``[component, *args]`` or just ``[*args]`` (component then comes from ``[cli:run] component`` in .torchxconfig).
    A leading ``--`` (argparse's options/positionals delimiter) is dropped; an option repeated within one ``--``-
    delimited group is an error."""
    component = config.get_config(prefix="cli", name="run", key="component", dirs=dirs)
    args = list(tokens)
    if args[:1] == ["--"]:
        args = args[1:]
    component_args: List[str] = []
    if args:
        if args[0].startswith("-"):
            component_args = args
        else:
            component, component_args = args[0], args[1:]
    for is_delim, group in groupby(component_args, key=lambda tok: tok == "--"):
        if is_delim:
            continue
        opts = Counter(tok for tok in group if tok.startswith("-") and tok.strip() not in ("-", "--"))
        dup = [tok for tok, n in opts.items() if n > 1]
        if dup:
            subparser.error(f"Repeated Command Line Arguments: {dup}")
    if not component:
        subparser.error(MISSING_COMPONENT_ERROR_MSG)
    return component, component_args  # type: ignore[return-value]


class CmdRun(SubCommand):
    def __init__(self) -> None:
        self._subparser: Optional[argparse.ArgumentParser] = None
        self._stdin_data_json: Optional[Dict[str, Any]] = None

    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        self._subparser = subparser
        names = list(get_scheduler_factories())
        subparser.add_argument("-s", "--scheduler", type=str, default=get_default_scheduler_name(), choices=names, action=torchxconfig_run,
                               help="Name of the scheduler to use.")
        subparser.add_argument("-cfg", "--scheduler_args", type=str, action=ArgOnceAction,
                               help="Arguments to pass to the scheduler (Ex:`log_dir=/tmp/x,pin_cpus=False`). See `torchx runopts`")
        subparser.add_argument("--dryrun", action="store_true", default=False, help="Does not actually submit the app, just prints the scheduler request")
        subparser.add_argument("--wait", action="store_true", default=False, help="Wait for the app to finish before exiting.")
        subparser.add_argument("--log", action="store_true", default=False, help="Stream logs while waiting for app to finish.")
        subparser.add_argument("--workspace", "--buck-target", default=f"{Path.cwd()}", action=torchxconfig_run,
                               help="local workspace to build/patch; the local schedulers run from the cwd and do not use it")
        subparser.add_argument("--parent_run_id", type=str, action=ArgOnceAction, help="optional parent run ID that this run belongs to")
        subparser.add_argument("--tee_logs", action="store_true", default=False,
                               help="Add additional prefix to log lines to indicate which replica is printing the log")
        subparser.add_argument("--stdin", action="store_true", default=False,
                               help="Read JSON input from stdin to parse into torchx run args and run the component.")
        subparser.add_argument("component_name_and_args", nargs=argparse.REMAINDER)

    def run(self, args: argparse.Namespace) -> None:
        os.environ.setdefault("TORCHX_CONTEXT_NAME", "cli_run")
        with get_runner(component_defaults=config.load_sections(prefix="component")) as runner:
            self._run(runner, args)

    # -- the two ways a request arrives -----------------------------------------------------------------------------
    def _run(self, runner: Runner, args: argparse.Namespace) -> None:
        self.verify_no_extra_args(args)
        if args.stdin:
            data = self._get_torchx_stdin_args(args)
            if data is not None:
                self._run_from_stdin_args(runner, data)
        else:
            self._run_from_cli_args(runner, args)

    def _run_from_cli_args(self, runner: Runner, args: argparse.Namespace) -> None:
        cfg = runner.scheduler_run_opts(args.scheduler).cfg_from_str(args.scheduler_args or "")
        assert self._subparser is not None
        component, component_args = _parse_component_name_and_args(args.component_name_and_args, self._subparser)
        self._run_inner(runner, torchx_run_args_from_argparse(args, component, component_args, cfg))

    def _run_from_stdin_args(self, runner: Runner, stdin_data: Dict[str, Any]) -> None:
        run_args = torchx_run_args_from_json(stdin_data)
        run_args.scheduler_cfg = runner.scheduler_run_opts(run_args.scheduler).cfg_from_json_repr(json.dumps(run_args.scheduler_args))
        self._run_inner(runner, run_args)

    def _get_torchx_stdin_args(self, args: argparse.Namespace) -> Optional[Dict[str, Any]]:
        if not args.stdin:
            return None
        if self._stdin_data_json is None:
            self._stdin_data_json = self.torchx_json_from_stdin(args)
        return self._stdin_data_json

    def torchx_json_from_stdin(self, args: Optional[argparse.Namespace] = None) -> Dict[str, Any]:
        try:
            data = json.load(sys.stdin)
        except (json.JSONDecodeError, EOFError):
            logger.error("Unable to parse JSON input for `torchx run` command, please make sure it's a valid JSON input.")
            sys.exit(1)
        if not isinstance(data, dict):
            logger.error("Invalid JSON input for `torchx run` command. Expected a dictionary.")
            sys.exit(1)
        if args and args.dryrun:
            data["dryrun"] = True
        return data

    def verify_no_extra_args(self, args: argparse.Namespace) -> None:
        """With ``--stdin`` the JSON is the whole request: any other option (except ``--dryrun``) is an error."""
        if not args.stdin:
            return
        assert self._subparser is not None
        clash = []
        for action in self._subparser._actions:
            if action.dest in ("stdin", "help", "dryrun"):
                continue
            value = getattr(args, action.dest, None)
            if value == action.default or (action.dest == "component_name_and_args" and value == []):
                continue
            clash.append(f"--{action.dest.replace('_', '-')}")
        if clash:
            self._subparser.error(f"Cannot specify {', '.join(clash)} when using --stdin. All configuration should be provided in JSON input.")

    # -- submit -----------------------------------------------------------------------------------------------------
    def _run_inner(self, runner: Runner, args: TorchXRunArgs) -> None:
        config.apply(scheduler=args.scheduler, cfg=args.scheduler_cfg)
        component_args: Any = args.component_args_str if args.component_args_str != [] else args.component_args
        try:
            if args.dryrun:
                info = runner.dryrun_component(args.component_name, component_args, args.scheduler, workspace=args.workspace,
                                               cfg=args.scheduler_cfg, parent_run_id=args.parent_run_id)
                print(f"\n=== APPLICATION ===\n{pformat(asdict(info._app), indent=2, width=80)}")
                print(f"\n=== SCHEDULER REQUEST ===\n{info}")
                return
            handle = runner.run_component(args.component_name, component_args, args.scheduler, workspace=args.workspace,
                                          cfg=args.scheduler_cfg, parent_run_id=args.parent_run_id)
            print(handle, flush=True)  # scripts read the handle from the first stdout line
            if args.scheduler.startswith("local"):
                self._wait_and_exit(runner, handle, log=True, tee_logs=args.tee_logs)
            else:
                logger.info(f"Launched app: {handle}")
                status = runner.status(handle)
                if status:
                    logger.info(status.format())
                if args.wait or args.log:
                    self._wait_and_exit(runner, handle, log=args.log, tee_logs=args.tee_logs)
        except (ComponentValidationException, ComponentNotFoundException) as e:
            logger.error(f"\nFailed to run component `{args.component_name}` got errors: \n {e}")
            sys.exit(1)
        except specs.InvalidRunConfigException as e:
            print(f"Invalid scheduler configuration: {e}\nUse `-cfg key=value,...` or a `.torchxconfig` file; run `torchx runopts "
                  f"{args.scheduler}` to list the `{args.scheduler}` scheduler's options.", file=sys.stderr)
            sys.exit(1)

    def _wait_and_exit(self, runner: Runner, app_handle: str, log: bool, tee_logs: bool = False) -> None:
        logger.info("Waiting for the app to finish...")
        thread = self._start_log_thread(runner, app_handle, tee_logs_enabled=tee_logs) if log else None
        status = runner.wait(app_handle, wait_interval=1)
        if not status:
            raise RuntimeError(f"unknown status, wait returned {status}")
        logger.info(f"Job finished: {status.state}")
        if thread:
            thread.join()
        if status.state != specs.AppState.SUCCEEDED:
            logger.error(status)
            sys.exit(1)
        logger.debug(status)

    def _start_log_thread(self, runner: Runner, app_handle: str, tee_logs_enabled: bool = False) -> threading.Thread:
        if tee_logs_enabled:
            thread = tee_logs(dst=sys.stderr, app_handle=app_handle, regex=None, runner=runner, should_tail=True, streams=None,
                              colorize=not sys.stderr.closed and sys.stderr.isatty())
        else:
            thread = threading.Thread(target=get_logs, daemon=True,
                                      kwargs={"file": sys.stderr, "runner": runner, "identifier": app_handle, "regex": None, "should_tail": True})
        thread.start()
        return thread
