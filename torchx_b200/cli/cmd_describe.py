"""``torchx describe <handle>``: the AppDef as the scheduler reconstructs it (reference torchx/cli/cmd_describe.py)."""
from __future__ import annotations

import argparse
import logging
import sys
from dataclasses import asdict
from pprint import pformat

from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.runner import get_runner
from torchx_b200.specs.api import parse_app_handle

logger = logging.getLogger(__name__)


class CmdDescribe(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("app_handle", type=str, help="torchx app handle (e.g. local_cuda://torchx/app_id)")

    def run(self, args: argparse.Namespace) -> None:
        scheduler, _, app_id = parse_app_handle(args.app_handle)
        app = get_runner().describe(args.app_handle)
        if not app:
            logger.error(f"AppDef: {app_id}, does not exist or has been removed from {scheduler}'s data plane")
            sys.exit(1)
        print(pformat(asdict(app), indent=2, width=80))
