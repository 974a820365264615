"""``torchx status <handle> [--roles a,b] [--json]`` (reference torchx/cli/cmd_status.py:43-79).  ``local_cuda`` apps
resolve from any process through the scheduler's app registry; the reference's local schedulers only from the
submitting one (local_scheduler.py:1099-1102)."""
from __future__ import annotations

import argparse
import json
import logging
import sys
from typing import List, Optional

from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.runner import get_runner
from torchx_b200.specs.api import parse_app_handle

logger = logging.getLogger(__name__)


def parse_list_arg(arg: str) -> Optional[List[str]]:
    return arg.split(",") if arg else None


class CmdStatus(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("app_handle", type=str, help="torchx app handle (e.g. local_cuda://torchx/app_id)")
        subparser.add_argument("--roles", type=str, default="", help="comma separated roles to filter")
        subparser.add_argument("--json", action="store_true", help="output the status in JSON format")

    def run(self, args: argparse.Namespace) -> None:
        scheduler, _, app_id = parse_app_handle(args.app_handle)
        status = get_runner().status(args.app_handle)
        if not status:
            logger.error(f"AppDef: {app_id}, does not exist or has been removed from {scheduler}'s data plane")
            sys.exit(1)
        roles = parse_list_arg(args.roles)
        print(json.dumps(status.to_json(roles)) if args.json else status.format(roles))
