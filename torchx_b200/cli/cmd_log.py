"""``torchx log <scheduler>://<session>/<app_id>[/<role>[/<k1,k2>]]`` (reference torchx/cli/cmd_log.py:75-193): one
reader thread per (role, replica), lines prefixed ``role/replica``.  For ``local_cuda`` a replica's log is the
``[rank]:``-prefixed merge of its workers' output."""
from __future__ import annotations

import argparse
import logging
import re
import sys
import threading
import time
from queue import Queue
from typing import Optional, TextIO

from torchx_b200.cli.cmd_base import SubCommand
from torchx_b200.runner import Runner, get_runner
from torchx_b200.schedulers.api import Stream
from torchx_b200.specs.api import is_started, make_app_handle
from torchx_b200.util.log_tee_helpers import _find_role_replicas as find_role_replicas
from torchx_b200.util.log_tee_helpers import _prefix_line

logger = logging.getLogger(__name__)
ID_FORMAT = "SCHEDULER://[SESSION_NAME]/APP_ID/[ROLE_NAME/[REPLICA_IDS,...]]"
_ID = re.compile(r"^\w+://[^/]*/[^/]+(/[^/]+(/(\d+,?)+)?)?$")
GREEN, ENDC = "\033[92m", "\033[0m"


def validate(identifier: str) -> None:
    if not _ID.match(identifier):
        logger.error(f"{identifier} is not of the form {ID_FORMAT}")
        sys.exit(1)


def print_log_lines(file: TextIO, runner: Runner, app_handle: str, role_name: str, replica_id: int, regex: Optional[str],
                    should_tail: bool, exceptions: "Queue[Exception]", streams: Optional[Stream]) -> None:
    prefix = f"{GREEN}{role_name}/{replica_id}{ENDC} "  # always coloured, as the reference prints it (cmd_log.py:66)
    try:
        for line in runner.log_lines(app_handle, role_name, replica_id, regex, should_tail=should_tail, streams=streams):
            try:
                print(_prefix_line(prefix, line), file=file, end="", flush=True)
            except BrokenPipeError:
                return
    except Exception as e:  # noqa: BLE001 - surfaced by get_logs in the caller's thread
        exceptions.put(e)
        raise


def get_logs(file: TextIO, identifier: str, regex: Optional[str], should_tail: bool = False, runner: Optional[Runner] = None,
             streams: Optional[Stream] = None) -> None:
    validate(identifier)
    backend, _, rest = identifier.partition("://")
    parts = rest.split("/")
    session, app_id = parts[0] or "default", parts[1]
    role_name = parts[2] if len(parts) > 2 else None
    runner = runner or get_runner()
    handle = make_app_handle(backend, session, app_id)
    if len(parts) == 4:
        targets = [(role_name, int(k)) for k in parts[3].split(",") if k]
    else:
        announced = False
        while True:
            st = runner.status(handle)
            if st and is_started(st.state):
                break
            if not announced:
                logger.info("Waiting for app state response before fetching logs...")
                announced = True
            time.sleep(1)
        app = runner.describe(handle)
        assert app is not None
        targets = find_role_replicas(app, role_name)
        if not targets:
            logger.error(f"No role [{role_name}] found for app: {app.name}. Roles: {[r.name for r in app.roles]}")
            sys.exit(1)
    errors: "Queue[Exception]" = Queue()
    threads = [threading.Thread(target=print_log_lines, args=(file, runner, handle, r, k, regex, should_tail, errors, streams), daemon=True)
               for r, k in targets]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if not errors.empty():
        raise errors.get()


class CmdLog(SubCommand):
    def add_arguments(self, subparser: argparse.ArgumentParser) -> None:
        subparser.add_argument("--regex", type=str, help="regex filter")
        subparser.add_argument("-t", "--tail", action="store_true", help="Tail logs")
        subparser.add_argument("--streams", type=Stream, choices=list(Stream), default=None, help="IO streams to use. Default is combined.")
        subparser.add_argument("identifier", type=str, metavar=ID_FORMAT, help="identifiers for the roles and replicas to log")

    def run(self, args: argparse.Namespace) -> None:
        get_logs(sys.stdout, args.identifier, args.regex, args.tail, streams=args.streams)
