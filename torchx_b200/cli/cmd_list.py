"""``torchx list [-s scheduler] [-cfg k=v,...]``: handle, name and state of the scheduler's apps
(reference torchx/cli/cmd_list.py:25-68).  Defaults come from ``[cli:list]`` in .torchxconfig."""
from __future__ import annotations

import argparse
import logging
from typing import Dict

from tabulate import tabulate

from torchx_b200.cli.argparse_util import ArgOnceAction, torchxconfig_list
from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.runner import config, get_runner
from torchx_b200.schedulers import get_default_scheduler_name, get_scheduler_factories
from torchx_b200.specs import CfgVal

logger = logging.getLogger(__name__)

HANDLE_HEADER = "APP HANDLE"
STATUS_HEADER = "APP STATUS"
NAME_HEADER = "APP NAME"


class CmdList(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        names = list(get_scheduler_factories())
        subparser.add_argument("-s", "--scheduler", type=str, default=get_default_scheduler_name(), choices=names, action=torchxconfig_list,
                               help=f"Name of the scheduler to use. One of: [{','.join(names)}].")
        subparser.add_argument("-cfg", "--scheduler_args", type=str, action=ArgOnceAction,
                               help="Arguments to pass to the scheduler (Ex: `log_dir=/tmp/x`). See `torchx runopts`")

    def run(self, args: argparse.Namespace) -> None:
        with get_runner() as runner:
            cfg: Dict[str, CfgVal] = {}
            if args.scheduler_args:  # the command line wins; .torchxconfig only fills what it left open
                cfg = runner.scheduler_run_opts(args.scheduler).cfg_from_str(args.scheduler_args)
            config.apply(scheduler=args.scheduler, cfg=cfg)
            apps = runner.list(args.scheduler, cfg if cfg else None)
            print(tabulate([[a.app_handle, a.name, str(a.state)] for a in apps], headers=[HANDLE_HEADER, NAME_HEADER, STATUS_HEADER]))
