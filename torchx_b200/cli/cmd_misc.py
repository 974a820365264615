"""The small sub-commands: status, describe, cancel, runopts, builtins, configure
(reference torchx/cli/cmd_status.py, cmd_describe.py, cmd_cancel.py, cmd_runopts.py, cmd_run.py:CmdBuiltins, cmd_configure.py).
Local schedulers keep app state in the launching process, so status/describe/cancel of a handle only resolve from the
process that submitted it (true for the reference as well: local_scheduler.py:615, 1099-1102)."""
from __future__ import annotations

import argparse
import inspect
import json
import logging
import sys
from dataclasses import asdict
from pprint import pformat

from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.runner import config, get_runner
from torchx_b200.schedulers import get_scheduler_factories
from torchx_b200.specs.finder import get_builtin_components, get_component

logger = logging.getLogger(__name__)


class CmdStatus(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("app_handle", type=str, help="torchx app handle (e.g. local_cuda://torchx/app_id)")
        subparser.add_argument("--roles", type=str, default="", help="comma separated roles to filter")
        subparser.add_argument("--json", action="store_true", help="output the status as JSON")

    def run(self, args: argparse.Namespace) -> None:
        with get_runner() as runner:
            status = runner.status(args.app_handle)
            roles = [r for r in args.roles.split(",") if r]
            if status is None:
                logger.error(f"AppDef: {args.app_handle}, does not exist or has been removed from the scheduler's data plane")
                sys.exit(1)
            print(json.dumps(status.to_json(roles)) if args.json else status.format(roles))


class CmdDescribe(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("app_handle", type=str, help="torchx app handle (e.g. local_cuda://torchx/app_id)")

    def run(self, args: argparse.Namespace) -> None:
        with get_runner() as runner:
            app = runner.describe(args.app_handle)
            if app is None:
                logger.error(f"AppDef: {args.app_handle}, does not exist or has been removed from the scheduler's data plane")
                sys.exit(1)
            print(pformat(asdict(app), indent=2, width=80))


class CmdCancel(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("app_handle", type=str, help="torchx app handle (e.g. local_cuda://torchx/app_id)")

    def run(self, args: argparse.Namespace) -> None:
        with get_runner() as runner:
            runner.cancel(args.app_handle)


class CmdList(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("-s", "--scheduler", type=str, default="local_cuda", help="scheduler whose apps to list")

    def run(self, args: argparse.Namespace) -> None:
        with get_runner() as runner:
            apps = runner.list(args.scheduler)
            print(f"{'APP HANDLE':60s} APP STATUS")
            for a in apps:
                print(f"{a.app_handle:60s} {a.state}")


class CmdRunopts(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("scheduler", type=str, nargs="?", help="scheduler to dump the runopts for, dumps for all schedulers if not specified")

    def run(self, args: argparse.Namespace) -> None:
        factories = get_scheduler_factories()
        names = [args.scheduler] if args.scheduler else list(factories)
        for name in names:
            if name not in factories:
                logger.error(f"unknown scheduler `{name}`; choose from {list(factories)}")
                sys.exit(1)
            sched = factories[name]("runopts")
            try:
                print(f"{name}:\n{sched.run_opts()!r}\n")
            finally:
                sched.close()


class CmdBuiltins(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("--print", type=str, help="prints the builtin's component def to stdout")

    def run(self, args: argparse.Namespace) -> None:
        if args.print:
            print(inspect.getsource(get_component(args.print).fn))
            return
        comps = get_builtin_components()
        print(f"Found {len(comps)} builtin components:")
        for i, c in enumerate(comps.values()):
            print(f" {i + 1:2d}. {c.name}")


class CmdConfigure(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("-s", "--schedulers", type=str, help="comma delimited list of schedulers to dump runopts for, if not specified, dumps all")
        subparser.add_argument("--print", action="store_true", help="if specified, prints the config file to stdout instead of saving it to a file")
        subparser.add_argument("-a", "--all", action="store_true", help="if specified, includes required and optional runopts (default only dumps required)")

    def run(self, args: argparse.Namespace) -> None:
        scheds = args.schedulers.split(",") if args.schedulers else None
        if args.print:
            config.dump(sys.stdout, scheds, required_only=not args.all)
        else:
            with open(config.CONFIG_FILE, "w") as f:
                config.dump(f, scheds, required_only=not args.all)
