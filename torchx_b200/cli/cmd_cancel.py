"""``torchx cancel <handle>`` (reference torchx/cli/cmd_cancel.py).  For ``local_cuda`` this SIGTERMs the whole gang,
also from a process other than the submitting one (via the app registry)."""
from __future__ import annotations

import argparse

from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.runner import get_runner


class CmdCancel(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("app_handle", type=str, help="torchx app handle (e.g. local_cuda://torchx/app_id)")

    def run(self, args: argparse.Namespace) -> None:
        get_runner().cancel(args.app_handle)
