"""Per-call telemetry of the Runner (reference torchx/runner/events/__init__.py:39-175, api.py:24-80, handlers.py).

Every Runner API call runs inside ``log_event(api, scheduler, app_id, ...)``; on exit a :class:`TorchxEvent` with the wall
and CPU time of the call, and the exception if it raised, is serialised to JSON and handed to a dedicated, non-propagating
python logger.  The default destination is a ``NullHandler`` - nothing is written anywhere - until a deployment registers a
handler in :data:`handlers` (``handlers["console"]`` prints to stderr) or patches :func:`record`.
"""
from __future__ import annotations

import json
import logging
import time
import traceback
from dataclasses import asdict, dataclass
from enum import Enum
from types import TracebackType
from typing import Dict, Optional, Type, Union

from torchx_b200.util.session import get_session_id_or_create_new

log = logging.getLogger(__name__)


class SourceType(str, Enum):
    UNKNOWN = "<unknown>"
    INTERNAL = "INTERNAL"
    EXTERNAL = "EXTERNAL"


@dataclass
class TorchxEvent:
    """One Runner API call: who (``session``), what (``api``, ``scheduler``, ``app_id``, ``app_image``, ``runcfg`` as JSON,
    ``workspace``), how long (``cpu_time_usec``, ``wall_time_usec``, ``start_epoch_time_usec``) and how it failed
    (``exception_*``, ``raw_exception`` = the formatted traceback)."""

    session: str
    scheduler: str
    api: str
    app_id: Optional[str] = None
    app_image: Optional[str] = None
    app_metadata: Optional[Dict[str, str]] = None
    runcfg: Optional[str] = None
    raw_exception: Optional[str] = None
    source: SourceType = SourceType.UNKNOWN
    cpu_time_usec: Optional[int] = None
    wall_time_usec: Optional[int] = None
    start_epoch_time_usec: Optional[int] = None
    workspace: Optional[str] = None
    exception_type: Optional[str] = None
    exception_message: Optional[str] = None
    exception_source_location: Optional[str] = None

    def serialize(self) -> str:
        return json.dumps(asdict(self))

    __str__ = serialize

    @staticmethod
    def deserialize(data: Union[str, "TorchxEvent"]) -> "TorchxEvent":
        if isinstance(data, TorchxEvent):
            return data
        fields = json.loads(data)
        if "source" in fields:
            try:
                fields["source"] = SourceType(fields["source"])
            except ValueError:  # written by a newer version: fall back to the default
                del fields["source"]
        return TorchxEvent(**fields)


# destination name -> logging.Handler; deployments add their own sink here before the first event is recorded
handlers: Dict[str, logging.Handler] = {"console": logging.StreamHandler(), "null": logging.NullHandler()}


def get_logging_handler(destination: str = "null") -> logging.Handler:
    return handlers[destination]


_events_logger: Optional[logging.Logger] = None


def _get_or_create_logger(destination: str = "null") -> logging.Logger:
    """The events logger: created once, for the first destination asked for; does not propagate to the root logger, so an
    event is handled exactly once."""
    global _events_logger
    if _events_logger is None:
        handler = get_logging_handler(destination)
        handler.setLevel(logging.DEBUG)
        _events_logger = logging.getLogger(f"torchx-events-{destination}")
        _events_logger.propagate = False
        _events_logger.setLevel(logging.DEBUG)
        _events_logger.addHandler(handler)
    return _events_logger


def record(event: TorchxEvent, destination: str = "null") -> None:
    try:
        payload = event.serialize()
    except Exception:  # noqa: BLE001 - telemetry never breaks the call it describes
        log.exception("failed to serialize event, will not record event")
        return
    _get_or_create_logger(destination).info(payload)


class log_event:
    """``with log_event("schedule", scheduler, app_id) as ctx: ...`` - ``ctx._torchx_event`` can be amended inside the block
    (the Runner fills in ``app_id`` / ``app_image`` once it knows them); the event is recorded on exit, raised or not."""

    def __init__(self, api: str, scheduler: Optional[str] = None, app_id: Optional[str] = None, app_image: Optional[str] = None,
                 app_metadata: Optional[Dict[str, str]] = None, runcfg: Optional[str] = None, workspace: Optional[str] = None) -> None:
        self._torchx_event = self._generate_torchx_event(api, scheduler or "", app_id, app_image=app_image, app_metadata=app_metadata,
                                                         runcfg=runcfg, workspace=workspace)
        self._cpu0 = self._wall0 = 0

    def _generate_torchx_event(self, api: str, scheduler: str, app_id: Optional[str] = None, app_image: Optional[str] = None,
                               app_metadata: Optional[Dict[str, str]] = None, runcfg: Optional[str] = None,
                               source: SourceType = SourceType.UNKNOWN, workspace: Optional[str] = None) -> TorchxEvent:
        return TorchxEvent(session=get_session_id_or_create_new(), scheduler=scheduler, api=api, app_id=app_id, app_image=app_image,
                           app_metadata=app_metadata, runcfg=runcfg, source=source, workspace=workspace)

    def __enter__(self) -> "log_event":
        self._cpu0, self._wall0 = time.process_time_ns(), time.perf_counter_ns()
        self._torchx_event.start_epoch_time_usec = int(time.time() * 1_000_000)
        return self

    def __exit__(self, exc_type: Optional[Type[BaseException]], exc: Optional[BaseException], tb: Optional[TracebackType]) -> Optional[bool]:
        ev = self._torchx_event
        ev.cpu_time_usec = (time.process_time_ns() - self._cpu0) // 1000
        ev.wall_time_usec = (time.perf_counter_ns() - self._wall0) // 1000
        if tb is not None:
            ev.raw_exception = traceback.format_exc()
            where = traceback.extract_tb(tb)[-1]
            ev.exception_source_location = json.dumps({"filename": where.filename, "lineno": where.lineno, "name": where.name})
        if exc_type is not None:
            ev.exception_type = exc_type.__name__
        if exc is not None:
            ev.exception_message = str(exc)
        record(ev)
        return None
