"""``torchx configure [-s a,b] [--print] [-a]``: write a ``.torchxconfig`` template with every scheduler's runopts
(reference torchx/cli/cmd_configure.py:23-58)."""
from __future__ import annotations

import argparse
import logging
import sys

from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.runner.config import dump
from torchx_b200.schedulers import get_scheduler_factories

logger = logging.getLogger(__name__)


class CmdConfigure(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("-s", "--schedulers", type=str, help="comma delimited list of schedulers to dump runopts for, if not specified, dumps for all schedulers")
        subparser.add_argument("--print", action="store_true", help="if specified, prints the config file to stdout instead of saving it to a file")
        subparser.add_argument("-a", "--all", action="store_true", help="if specified, includes required and optional runopts (default only dumps required)")

    def run(self, args: argparse.Namespace) -> None:
        schedulers = args.schedulers.split(",") if args.schedulers else list(get_scheduler_factories())
        if args.print:
            dump(f=sys.stdout, schedulers=schedulers, required_only=not args.all)
        else:
            with open(".torchxconfig", "w") as f:
                dump(f=f, schedulers=schedulers, required_only=not args.all)
