"""Plugins: add scheduler names and named resources without editing this package.

Put ``@register``-decorated factories in modules under the ``torchx_b200_plugins.{schedulers,named_resources}`` namespace
packages anywhere on ``sys.path`` (or publish them as ``torchx_b200.schedulers`` / ``torchx_b200.named_resources`` entry
points); ``plugins.registry()`` discovers them lazily, and ``print(plugins.registry())`` is a YAML report including the
modules that failed to import.  Registered schedulers REPLACE the built-in scheduler map, as in TorchX.

Same public names as reference torchx/plugins/__init__.py:36-70 (``register``, ``resource_tags``, the fractional helpers,
``PluginRegistry``, ``PluginType``, ``RegistrationError``, ``registry``); tracker plugins are out of scope.
"""
from __future__ import annotations

import sys
from typing import Any, Callable, Dict, List

from torchx_b200.plugins._registration import (  # noqa: F401
    EIGHTH,
    HALF,
    QUARTER,
    SIXTEENTH,
    WHOLE,
    halve_mem_down_to,
    powers_of_two_gpus,
    register,
    resource_tags,
)
from torchx_b200.plugins._registry import (  # noqa: F401
    NAMESPACE,
    PluginRegistry,
    PluginSource,
    PluginType,
    RegistrationError,
    registry,
)

__all__ = ["register", "resource_tags", "powers_of_two_gpus", "halve_mem_down_to", "WHOLE", "HALF", "QUARTER", "EIGHTH", "SIXTEENTH",
           "RegistrationError", "PluginType", "PluginSource", "PluginRegistry", "registry"]


# -- convenience wrappers used inside this package ------------------------------------------------------------------
def registered_schedulers() -> Dict[str, Callable[..., Any]]:
    return dict(registry().get(PluginType.SCHEDULER))


def registered_named_resources() -> Dict[str, Callable[[], Any]]:
    return dict(registry().get(PluginType.NAMED_RESOURCE))


def errors() -> List[Dict[str, str]]:
    """Plugin modules that failed to import / plugins that were rejected, as plain dicts."""
    reg = registry()
    reg.info()
    return [{"module": e.module, "error": e.error, **({"name": e.name} if e.name else {})} for e in reg.errors]


def reset_for_tests() -> None:
    """Drop the cached registry and the imported plugin namespace (so a changed ``sys.path`` is re-scanned)."""
    registry().clear()
    for name in [m for m in sys.modules if m == NAMESPACE or m.startswith(NAMESPACE + ".")]:
        del sys.modules[name]
