"""``join``: the inverse of ``shlex.split`` (reference torchx/util/shlex.py) - how component args become one ``bash -c`` line."""
import shlex as _shlex
from typing import Iterable


def join(args: Iterable[str]) -> str:
    return _shlex.join(list(args))
