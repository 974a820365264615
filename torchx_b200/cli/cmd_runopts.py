"""``torchx runopts [scheduler]``: the ``-cfg`` options each scheduler takes (reference torchx/cli/cmd_runopts.py)."""
from __future__ import annotations

import argparse
import logging

from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.runner import get_runner

logger = logging.getLogger(__name__)
GREEN, ENDC = "\033[92m", "\033[0m"


class CmdRunopts(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("scheduler", type=str, nargs="?", help="scheduler to dump the runopts for, dumps for all schedulers if not specified")

    def run(self, args: argparse.Namespace) -> None:
        with get_runner() as runner:
            for scheduler in runner.scheduler_backends():
                if args.scheduler and scheduler != args.scheduler:
                    continue
                try:
                    print(f"{GREEN}{scheduler}{ENDC}:\n{runner.scheduler_run_opts(scheduler)!r}\n")
                except ModuleNotFoundError as e:  # a scheduler whose optional dependency is not installed
                    print(f"{GREEN}{scheduler}{ENDC}: {e}\n")
