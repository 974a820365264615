"""Run a program under supervision: optional wall-clock limit, optional start / stop sentinel files.

Stand-alone counterpart of the gang-level ``-cfg job_timeout=...,exit_on_file=...,start_on_file=...`` options of the
``local_cuda`` scheduler, for wrapping a single role's entrypoint the way the reference's
``python -m torchx.apps.utils.process_monitor`` is used (torchx/apps/utils/process_monitor.py:20-118): same flags, same exit
code 34 when the limit is hit before the program was started, SIGTERM first and SIGKILL after ``--kill_timeout``.
Sentinels are local paths, or ``scheme://`` URLs resolved through fsspec as in the reference (fsspec must then be
installed).  The progress lines printed to stdout are the reference's, so log greps written for it keep working.

    python -m torchx_b200.apps.utils.process_monitor --timeout 3600 --exit_on_file /tmp/stop -- python train.py
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import time
from typing import List, Optional

TIMEOUT_EXIT_CODE = 34


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="process_monitor", description="supervise a process: time limit and sentinel files")
    p.add_argument("--timeout", type=float, help="seconds after which the process is terminated")
    p.add_argument("--start_on_file", type=str, help="do not start the process before this path exists")
    p.add_argument("--exit_on_file", type=str, help="terminate the process as soon as this path exists")
    p.add_argument("--poll_rate", type=float, default=5, help="seconds between checks")
    p.add_argument("--kill_timeout", type=float, default=60, help="grace period between SIGTERM and SIGKILL")
    p.add_argument("entrypoint", type=str)
    p.add_argument("args", type=str, nargs=argparse.REMAINDER)
    return p


def _exists(path: str) -> bool:
    if "://" not in path:
        return os.path.exists(path)
    import fsspec  # only needed for URL sentinels; a missing fsspec is an error, not "file absent"

    fs, fs_path = fsspec.core.url_to_fs(path)
    return bool(fs.exists(fs_path))


def supervise(entrypoint: str, args: List[str], timeout: Optional[float] = None, start_on_file: Optional[str] = None,
              exit_on_file: Optional[str] = None, poll_rate: float = 5.0, kill_timeout: float = 60.0) -> int:
    """Returns the exit code the monitor itself should exit with."""
    t0 = time.monotonic()

    def expired() -> bool:
        return bool(timeout) and time.monotonic() - t0 > timeout  # type: ignore[operator]

    while start_on_file:
        if _exists(start_on_file):
            print(f"{start_on_file} exists, starting process...", flush=True)
            break
        if expired():
            print("reached timeout before launching, terminating...", flush=True)
            return TIMEOUT_EXIT_CODE
        time.sleep(poll_rate)
    if args and args[0] == "--":
        args = args[1:]
    proc = subprocess.Popen([entrypoint, *args])
    print(f"started process {proc.pid}", flush=True)
    while True:
        try:
            rc = proc.wait(poll_rate)
            print(f"process exited with exit code {rc}", flush=True)
            return rc
        except subprocess.TimeoutExpired:
            if expired():
                print("reached timeout, terminating...", flush=True)
                break
            if exit_on_file and _exists(exit_on_file):
                print(f"{exit_on_file} exists, terminating...", flush=True)
                break
    proc.terminate()
    print("issued terminate, waiting for exit...", flush=True)
    try:
        proc.wait(kill_timeout)
    except subprocess.TimeoutExpired:
        print("reached safe termination timeout, killing...", flush=True)
        proc.kill()
    rc = proc.wait()
    print(f"process exited with exit code {rc}", flush=True)
    return rc


def main(argv: Optional[List[str]] = None) -> None:
    a = build_parser().parse_args(sys.argv[1:] if argv is None else argv)
    sys.exit(supervise(a.entrypoint, a.args, a.timeout, a.start_on_file, a.exit_on_file, a.poll_rate, a.kill_timeout))


if __name__ == "__main__":
    main()
