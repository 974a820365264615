"""No-op shim of the reference's per-call telemetry (torchx/runner/events/__init__.py:79-175, api.py:24).

The reference wraps every Runner API call in ``log_event`` which builds a ``TorchxEvent`` (wall/CPU time, scheduler, app
id, exception) and hands it to a python logger that has a NullHandler by default.  Telemetry is out of scope for the
single-box path; this module keeps the import surface (``record``, ``log_event``, ``TorchxEvent``)
so code and tests that patch ``runner.events.record`` keep working, and does nothing else."""
from __future__ import annotations

import time
from contextlib import contextmanager
from dataclasses import dataclass
from typing import Iterator, Optional


@dataclass
class TorchxEvent:
    session: str = ""
    scheduler: str = ""
    api: str = ""
    app_id: Optional[str] = None
    app_image: Optional[str] = None
    runcfg: Optional[str] = None
    workspace: Optional[str] = None
    exception_type: Optional[str] = None
    exception_message: Optional[str] = None
    wall_time_usec: Optional[int] = None


def record(event: TorchxEvent, destination: str = "null") -> None:
    """Sink for events: intentionally does nothing."""


class _Ctx:
    def __init__(self, event: TorchxEvent) -> None:
        self._torchx_event = event


@contextmanager
def log_event(api: str, scheduler: Optional[str] = None, app_id: Optional[str] = None, **kwargs: object) -> Iterator[_Ctx]:
    from torchx_b200.util.session import get_session_id_or_create_new

    ev = TorchxEvent(session=get_session_id_or_create_new(), api=api, scheduler=scheduler or "", app_id=app_id)
    t0 = time.perf_counter_ns()
    try:
        yield _Ctx(ev)
    except Exception as e:
        ev.exception_type, ev.exception_message = type(e).__name__, str(e)
        raise
    finally:
        ev.wall_time_usec = (time.perf_counter_ns() - t0) // 1000
        record(ev)
