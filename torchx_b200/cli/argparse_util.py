"""argparse actions shared by the sub-commands: defaults taken from ``.torchxconfig`` ``[cli:<subcmd>]`` sections and
options that may be given only once (reference torchx/cli/argparse_util.py:20-153)."""
from __future__ import annotations

import logging
import sys
from argparse import Action, ArgumentParser, Namespace
from typing import Any, Dict, Optional, Sequence, Set

from torchx_b200.runner import config

logger = logging.getLogger(__name__)


def _once(seen: Set[str], option_string: Optional[str]) -> None:
    if option_string is None:
        return
    if option_string in seen:
        logger.error(f"{option_string} is specified more than once")
        sys.exit(1)
    seen.add(option_string)


class ArgOnceAction(Action):
    """``store`` that refuses a second occurrence of the same option string."""

    called_args: Set[str] = set()

    def __call__(self, parser: ArgumentParser, namespace: Namespace, values: Any, option_string: Optional[str] = None) -> None:
        _once(self.called_args, option_string)
        setattr(namespace, self.dest, values)


class torchxconfig(Action):
    """``store`` whose default comes from ``[cli:<subcmd>] <dest> = ...`` when the config files define it; an option
    that has such a default is no longer ``required`` on the command line.  Also once-only."""

    called_args: Set[str] = set()
    _subcmd_configs: Dict[str, Dict[str, str]] = {}  # one config read per sub-command, shared by its options

    def __init__(self, subcmd: str, dest: str, option_strings: Sequence[str], required: bool = False, default: Any = None, **kwargs: Any) -> None:
        if subcmd not in self._subcmd_configs:
            self._subcmd_configs[subcmd] = config.get_configs(prefix="cli", name=subcmd)
        default = self._subcmd_configs[subcmd].get(dest, default)
        super().__init__(dest=dest, option_strings=option_strings, default=default, required=required and not default, **kwargs)

    def __call__(self, parser: ArgumentParser, namespace: Namespace, values: Any, option_string: Optional[str] = None) -> None:
        _once(self.called_args, option_string)
        setattr(namespace, self.dest, values)


class torchxconfig_run(torchxconfig):
    def __init__(self, dest: str, option_strings: Sequence[str], required: bool = False, default: Any = None, **kwargs: Any) -> None:
        super().__init__("run", dest=dest, option_strings=option_strings, required=required, default=default, **kwargs)


class torchxconfig_list(torchxconfig):
    def __init__(self, dest: str, option_strings: Sequence[str], required: bool = False, default: Any = None, **kwargs: Any) -> None:
        super().__init__("list", dest=dest, option_strings=option_strings, required=required, default=default, **kwargs)
