"""Destination name -> ``logging.Handler`` for Runner events (reference torchx/runner/events/handlers.py)."""
from torchx_b200.runner.events import get_logging_handler, handlers as _log_handlers  # noqa: F401
