"""Entry point: ``python -m torchx_b200.cli.main run -s local_cuda dist.ddp -j 1x8 --script train.py -- --lr 0.1``."""
from __future__ import annotations

import argparse
import logging
import os
import sys
from typing import Dict, List, Optional

import torchx_b200
from torchx_b200.cli.argparse_util import ArgOnceAction, torchxconfig
from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.cli.cmd_cancel import CmdCancel
from torchx_b200.cli.cmd_configure import CmdConfigure
from torchx_b200.cli.cmd_delete import CmdDelete
from torchx_b200.cli.cmd_describe import CmdDescribe
from torchx_b200.cli.cmd_list import CmdList
from torchx_b200.cli.cmd_log import CmdLog
from torchx_b200.cli.cmd_run import CmdBuiltins, CmdRun
from torchx_b200.cli.cmd_runopts import CmdRunopts
from torchx_b200.cli.cmd_status import CmdStatus
from torchx_b200.util.entrypoints import load_group


def get_default_sub_cmds() -> Dict[str, SubCommand]:
    return {
        "builtins": CmdBuiltins(),
        "cancel": CmdCancel(),
        "configure": CmdConfigure(),
        "delete": CmdDelete(),
        "describe": CmdDescribe(),
        "list": CmdList(),
        "log": CmdLog(),
        "run": CmdRun(),
        "runopts": CmdRunopts(),
        "status": CmdStatus(),
    }


def get_sub_cmds() -> Dict[str, SubCommand]:
    """The built-in sub-commands, overlaid by SubCommand classes published under the ``torchx_b200.cli.cmds`` entry-point
    group (reference cli/main.py:55-77 does the same with ``torchx.cli.cmds``)."""
    cmds = get_default_sub_cmds()
    for name, make in (load_group("torchx_b200.cli.cmds", default={}) or {}).items():
        cmds[name] = make()
    return cmds


def create_parser(subcmds: Dict[str, SubCommand]) -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(prog="torchx", description="torchx_b200: B200-native single-box DDP launcher",
                                     formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("--log_level", type=str, default=os.getenv("LOGLEVEL", "INFO"), help="Python logging log level")
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
    # the once-only bookkeeping is per command line, not per process (main() may be called repeatedly from one program)
    ArgOnceAction.called_args = set()
    torchxconfig.called_args = set()
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
