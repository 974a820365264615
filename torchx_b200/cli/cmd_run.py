"""``torchx run``: argv -> (scheduler, cfg, component, component args) -> submit; for ``local*`` schedulers it then
blocks, streams the replicas' logs to stderr and exits non-zero unless the app SUCCEEDED
(reference torchx/cli/cmd_run.py: _parse_component_name_and_args:120, CmdRun:206, _run_inner:281, _wait_and_exit:453)."""
from __future__ import annotations

import argparse
import logging
import os
import sys
import threading
from collections import Counter
from dataclasses import asdict
from itertools import groupby
from pprint import pformat
from typing import Dict, List, Optional, Tuple

from torchx_b200 import specs
from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.cli.cmd_log import get_logs
from torchx_b200.runner import Runner, config, get_runner
from torchx_b200.schedulers import get_default_scheduler_name, get_scheduler_factories
from torchx_b200.specs.finder import ComponentNotFoundException, ComponentValidationException

logger = logging.getLogger(__name__)

MISSING_COMPONENT_ERROR_MSG = "missing component name, either provide it from the CLI or in .torchxconfig"


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

    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        self._subparser = subparser
        names = list(get_scheduler_factories())
        default_sched = config.get_config(prefix="cli", name="run", key="scheduler") or get_default_scheduler_name()
        subparser.add_argument("-s", "--scheduler", type=str, default=default_sched, choices=names, help="Name of the scheduler to use.")
        subparser.add_argument("-cfg", "--scheduler_args", type=str, default="",
                               help="Arguments to pass to the scheduler (Ex:`log_dir=/tmp/x,pin_cpus=False`). See `torchx runopts`")
        subparser.add_argument("--dryrun", action="store_true", default=False, help="Does not actually submit the app, just prints the scheduler request")
        subparser.add_argument("--wait", action="store_true", default=False, help="Wait for the app to finish before exiting.")
        subparser.add_argument("--log", action="store_true", default=False, help="Stream logs while waiting for app to finish.")
        subparser.add_argument("--parent_run_id", type=str, default=None, help="optional parent run ID that this run belongs to")
        subparser.add_argument("component_name_and_args", nargs=argparse.REMAINDER)

    def run(self, args: argparse.Namespace) -> None:
        os.environ.setdefault("TORCHX_CONTEXT_NAME", "cli_run")
        with get_runner(component_defaults=config.load_sections(prefix="component")) as runner:
            self._run(runner, args)

    def _run(self, runner: Runner, args: argparse.Namespace) -> None:
        opts = runner.scheduler_run_opts(args.scheduler)
        cfg = opts.cfg_from_str(args.scheduler_args or "")
        config.apply(scheduler=args.scheduler, cfg=cfg)
        assert self._subparser is not None
        component, component_args = _parse_component_name_and_args(args.component_name_and_args, self._subparser)
        try:
            if args.dryrun:
                info = runner.dryrun_component(component, component_args, args.scheduler, cfg=cfg, parent_run_id=args.parent_run_id)
                print(f"\n=== APPLICATION ===\n{pformat(asdict(info._app), indent=2, width=80)}")
                print(f"\n=== SCHEDULER REQUEST ===\n{info}")
                return
            handle = runner.run_component(component, component_args, args.scheduler, cfg=cfg, parent_run_id=args.parent_run_id)
            print(handle, flush=True)
            if args.scheduler.startswith("local"):
                self._wait_and_exit(runner, handle, log=True)
            else:
                logger.info(f"Launched app: {handle}")
                if args.wait or args.log:
                    self._wait_and_exit(runner, handle, log=args.log)
        except (ComponentValidationException, ComponentNotFoundException) as e:
            logger.error(f"\nFailed to run component `{component}` got errors: \n {e}")
            sys.exit(1)
        except specs.InvalidRunConfigException as e:
            print(f"Invalid scheduler configuration: {e}\nUse `-cfg key=value,...` or a `.torchxconfig` file; run `torchx runopts "
                  f"{args.scheduler}` to list the `{args.scheduler}` scheduler's options.", file=sys.stderr)
            sys.exit(1)

    def _wait_and_exit(self, runner: Runner, app_handle: str, log: bool) -> None:
        logger.info("Waiting for the app to finish...")
        thread = None
        if log:
            thread = threading.Thread(target=get_logs, kwargs={"file": sys.stderr, "runner": runner, "identifier": app_handle,
                                                               "regex": None, "should_tail": True}, daemon=True)
            thread.start()
        status = runner.wait(app_handle, wait_interval=1)
        if not status:
            raise RuntimeError(f"unknown status, wait returned {status}")
        logger.info(f"Job finished: {status.state}")
        if thread:
            thread.join()
        if status.state != specs.AppState.SUCCEEDED:
            logger.error(status)
            sys.exit(1)
