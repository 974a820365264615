"""``importlib.metadata`` entry-point groups as ``{name: deferred loader}`` maps
(reference torchx/util/entrypoints.py:12-123).  A loader imports its target only when called: calling it with arguments
calls the loaded attribute with them (``ep.load()(*args)``); an entry point that names a module just imports it."""
from __future__ import annotations

from importlib import metadata
from typing import Any, Callable, Dict, Optional


def load(group: str, name: str, default: Any = None) -> Any:
    """The object behind ``[group] name = pkg.mod:attr``; ``default`` (when given) if there is no such entry."""
    found = metadata.entry_points().select(group=group, name=name)
    if not found:
        if default is not None:
            return default
        raise KeyError(f"entrypoint {group}.{name} not found")
    return next(iter(found)).load()


def _deferred(ep: metadata.EntryPoint) -> Callable[..., Any]:
    def run(*args: Any, **kwargs: Any) -> Any:
        target = ep.load()
        return target if ep.attr is None else target(*args, **kwargs)

    run.__qualname__ = f"entrypoint<{ep.value}>"
    return run


def load_group(group: str, default: Optional[Dict[str, Any]] = None) -> Optional[Dict[str, Any]]:
    """``{entry name: deferred loader}`` for ``group``; ``default`` when the group is empty or absent."""
    eps = metadata.entry_points().select(group=group)
    if not eps:
        return default
    return {ep.name: _deferred(ep) for ep in eps}
