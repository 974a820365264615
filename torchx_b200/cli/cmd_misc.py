"""``torchx builtins`` (reference torchx/cli/cmd_run.py:183-203) and, for code written against round 1 of this package,
the other small commands re-exported from their own modules."""
from __future__ import annotations

import argparse
import inspect

from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.cli.cmd_cancel import CmdCancel  # noqa: F401
from torchx_b200.cli.cmd_configure import CmdConfigure  # noqa: F401
from torchx_b200.cli.cmd_describe import CmdDescribe  # noqa: F401
from torchx_b200.cli.cmd_list import CmdList  # noqa: F401
from torchx_b200.cli.cmd_runopts import CmdRunopts  # noqa: F401
from torchx_b200.cli.cmd_status import CmdStatus  # noqa: F401
from torchx_b200.specs.finder import get_builtin_components, get_component


class CmdBuiltins(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("--print", type=str, help="prints the builtin's component def to stdout")

    def _builtins(self):
        return get_builtin_components()

    def run(self, args: argparse.Namespace) -> None:
        if args.print:
            print(inspect.getsource(get_component(args.print).fn))
            return
        comps = self._builtins()
        print(f"Found {len(comps)} builtin components:")
        for i, c in enumerate(comps.values()):
            print(f" {i + 1:2d}. {c.name}")
