"""``torchx delete <handle>``: remove the app from the scheduler's records (reference torchx/cli/cmd_delete.py)."""
from __future__ import annotations

import argparse

from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.runner import get_runner


class CmdDelete(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("app_handle", type=str, help="torchx app handle (e.g. local_cuda://torchx/app_id)")

    def run(self, args: argparse.Namespace) -> None:
        get_runner().delete(args.app_handle)
