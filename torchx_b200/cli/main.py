"""Entry point: ``python -m torchx_b200.cli.main run -s local_cuda dist.ddp -j 1x8 --script train.py -- --lr 0.1``."""
from __future__ import annotations

import argparse
import logging
import os
import sys
from typing import Dict, List, Optional

import torchx_b200
from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.cli.cmd_misc import CmdBuiltins, CmdCancel, CmdConfigure, CmdDescribe, CmdList, CmdRunopts, CmdStatus
from torchx_b200.cli.cmd_log import CmdLog
from torchx_b200.cli.cmd_run import CmdRun


def get_sub_cmds() -> Dict[str, SubCommand]:
    return {
        "builtins": CmdBuiltins(),
        "cancel": CmdCancel(),
        "configure": CmdConfigure(),
        "describe": CmdDescribe(),
        "list": CmdList(),
        "log": CmdLog(),
        "run": CmdRun(),
        "runopts": CmdRunopts(),
        "status": CmdStatus(),
    }


def create_parser(subcmds: Dict[str, SubCommand]) -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="torchx", description="torchx_b200: B200-native single-box DDP launcher",
                                     formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--log_level", type=str, default="INFO", help="Python logging log level")
    parser.add_argument("--version", action="version", version=f"torchx_b200-{torchx_b200.__version__}")

    def default_help(args: argparse.Namespace) -> None:
        parser.print_help()

    parser.set_defaults(func=default_help)
    sub = parser.add_subparsers(title="sub-commands", description="Use the following commands to run and manage apps", help="sub-command help")
    for name, cmd in subcmds.items():
        p = sub.add_parser(name)
        cmd.add_arguments(p)
        p.set_defaults(func=cmd.run)
    return parser


def run_main(subcmds: Dict[str, SubCommand], argv: Optional[List[str]] = None) -> None:
    parser = create_parser(subcmds)
    args = parser.parse_args(sys.argv[1:] if argv is None else argv)
    logging.basicConfig(level=getattr(logging, str(args.log_level).upper(), logging.INFO),
                        format="torchx %(asctime)s %(levelname)-8s %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    if "func" not in args:
        parser.print_help()
        sys.exit(1)
    args.func(args)


def main(argv: Optional[List[str]] = None) -> None:
    run_main(get_sub_cmds(), argv)


if __name__ == "__main__":
    main()
