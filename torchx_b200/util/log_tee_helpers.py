"""Stream every replica's log of an app into one text stream, each line tagged ``<role>/<replica> ``.

For CLIs built on the Runner that want ``torchx log``'s output inline (what ``torchx run --tee_logs`` uses); public
entry point and behaviour follow reference torchx/util/log_tee_helpers.py:174-210 (``tee_logs`` returns an unstarted
daemon thread that fans out into one reader per (role, replica) and re-raises the first reader error on join).  With
``local_cuda`` a replica's stream is already the ``[rank]:``-prefixed merge of its workers, so a tee-d line reads
``trainer/0 [3]:loss=...``.
"""
from __future__ import annotations

import logging
import threading
from queue import Queue
from typing import TYPE_CHECKING, List, Optional, TextIO, Tuple

if TYPE_CHECKING:
    from torchx_b200.runner.api import Runner
    from torchx_b200.schedulers.api import Stream
    from torchx_b200.specs.api import AppDef

logger = logging.getLogger(__name__)
_GREEN, _RESET = "\033[32m", "\033[0m"


def _find_role_replicas(app: "AppDef", role_name: Optional[str]) -> List[Tuple[str, int]]:
    """All ``(role, replica id)`` pairs of ``app`` (only ``role_name``'s when given)."""
    return [(r.name, k) for r in app.roles if role_name in (None, r.name) for k in range(r.num_replicas)]


def _prefix_line(prefix: str, line: str) -> str:
    """Tag ``line``; progress-bar style lines that rewind with ``\\r`` or carry inner newlines get the tag after every
    rewind / newline so it never disappears from the terminal."""
    head, last = line[:-1], line[-1:]
    head = head.replace("\r", "\r" + prefix).replace("\n", "\n" + prefix)
    if last == "\r":
        last = "\r" + prefix
    out = head + last
    return out if out.startswith("\r") else prefix + out


def _print_log_lines_for_role_replica(dst: TextIO, app_handle: str, regex: Optional[str], runner: "Runner", which_role: str,
                                      which_replica: int, exceptions: "Queue[Exception]", should_tail: bool,
                                      streams: "Optional[Stream]", colorize: bool = False) -> None:
    tag = f"{_GREEN}{which_role}/{which_replica}{_RESET} " if colorize else f"{which_role}/{which_replica} "
    try:
        for line in runner.log_lines(app_handle, which_role, which_replica, regex, should_tail=should_tail, streams=streams):
            print(_prefix_line(tag, line.strip()), file=dst, end="\n", flush=True)
    except Exception as e:  # noqa: BLE001 - handed to the joining thread
        exceptions.put(e)
        raise


def _start_threads_to_monitor_role_replicas(dst: TextIO, app_handle: str, regex: Optional[str], runner: "Runner",
                                            which_role: Optional[str] = None, should_tail: bool = False,
                                            streams: "Optional[Stream]" = None, colorize: bool = False) -> None:
    app = runner.describe(app_handle)
    if app is None:
        raise ValueError(f"unknown app: {app_handle}")
    targets = _find_role_replicas(app, which_role)
    if not targets:
        raise ValueError(f"{which_role} is not a valid role name. Available: {[r.name for r in app.roles]}")
    errors: "Queue[Exception]" = Queue()
    readers = [threading.Thread(target=_print_log_lines_for_role_replica, daemon=True,
                                kwargs=dict(dst=dst, app_handle=app_handle, regex=regex, runner=runner, which_role=role,
                                            which_replica=k, exceptions=errors, should_tail=should_tail, streams=streams,
                                            colorize=colorize))
               for role, k in targets]
    for t in readers:
        t.start()
    for t in readers:
        t.join()
    raised = []
    while not errors.empty():
        raised.append(errors.get())
    for extra in raised[1:]:
        logger.error(extra)
    if raised:
        raise raised[0]


def tee_logs(dst: TextIO, app_handle: str, regex: Optional[str], runner: "Runner", should_tail: bool = False,
             streams: "Optional[Stream]" = None, colorize: bool = False) -> threading.Thread:
    """An unstarted daemon thread that copies all replicas' logs of ``app_handle`` to ``dst`` until they end; the caller
    starts and joins it.  Like the reference it always tails and ignores ``regex`` / ``streams`` (combined stream)."""
    return threading.Thread(target=_start_threads_to_monitor_role_replicas, daemon=True,
                            kwargs=dict(dst=dst, app_handle=app_handle, regex=None, runner=runner, should_tail=True, colorize=colorize))
