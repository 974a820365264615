"""ANSI colours for CLI output; empty strings when stdout is not a terminal (reference torchx/util/colors.py)."""
import sys

_TTY = not sys.stdout.closed and sys.stdout.isatty()


def _code(seq: str) -> str:
    return seq if _TTY else ""


GREEN = _code("\033[32m")
BLUE = _code("\033[34m")
ORANGE = _code("\033[38:2:238:76:44m")
GRAY = _code("\033[2m")
ENDC = _code("\033[0m")
