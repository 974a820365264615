"""Contract of a ``torchx`` sub-command (reference torchx/cli/cmd_base.py): ``main`` gives each one its own sub-parser to
fill, and dispatches the parsed namespace to ``run``.  Third-party commands subclass this and register under the
``torchx_b200.cli.cmds`` entry-point group."""
from __future__ import annotations

import argparse
from abc import ABC, abstractmethod


class SubCommand(ABC):
    @abstractmethod
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        """Declare this command's options on ``subparser``."""

    @abstractmethod
    def run(self, args: argparse.Namespace) -> None:
        """Execute with the parsed options; ``sys.exit(code)`` for a non-zero result."""
