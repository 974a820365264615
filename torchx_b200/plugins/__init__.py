"""Minimal plugin registry: ``@register.scheduler()`` / ``@register.named_resource()`` plus discovery of the
``torchx_b200_plugins.{schedulers,named_resources}`` namespace packages on ``sys.path``.

Mirrors the part of reference torchx/plugins (_registration.py:187-274 decorators, _registry.py:83-515 discovery with
error capture) that the launch path touches: a scheduler module anywhere on the path can add or override scheduler
names without editing this package; modules that fail to import are recorded, not fatal.  Tracker plugins and the
fractional named-resource machinery are out of scope.
"""
from __future__ import annotations

import importlib
import pkgutil
import sys
import traceback
from typing import Any, Callable, Dict, List, Optional

from torchx_b200.plugins._registration import NAMED_RESOURCES as _NAMED_RESOURCES
from torchx_b200.plugins._registration import SCHEDULERS as _SCHEDULERS
from torchx_b200.plugins._registration import register, resource_tags  # noqa: F401

_ERRORS: List[Dict[str, str]] = []
_DISCOVERED = False
NAMESPACE = "torchx_b200_plugins"


def _discover() -> None:
    global _DISCOVERED
    if _DISCOVERED:
        return
    _DISCOVERED = True
    for group in ("schedulers", "named_resources"):
        pkg_name = f"{NAMESPACE}.{group}"
        try:
            pkg = importlib.import_module(pkg_name)
        except ModuleNotFoundError:
            continue
        except Exception:  # a broken namespace package must not take the launcher down
            _ERRORS.append({"module": pkg_name, "error": traceback.format_exc()})
            continue
        for info in pkgutil.iter_modules(getattr(pkg, "__path__", [])):
            if info.name.startswith("_"):
                continue
            mod = f"{pkg_name}.{info.name}"
            try:
                importlib.import_module(mod)
            except Exception:
                _ERRORS.append({"module": mod, "error": traceback.format_exc()})


def registered_schedulers() -> Dict[str, Callable[..., Any]]:
    _discover()
    return dict(_SCHEDULERS)


def registered_named_resources() -> Dict[str, Callable[[], Any]]:
    _discover()
    return dict(_NAMED_RESOURCES)


def errors() -> List[Dict[str, str]]:
    """Plugin modules that failed to import (for ``torchx runopts`` / diagnostics)."""
    _discover()
    return list(_ERRORS)


def reset_for_tests() -> None:
    global _DISCOVERED
    _SCHEDULERS.clear()
    _NAMED_RESOURCES.clear()
    _ERRORS.clear()
    _DISCOVERED = False
    for name in [m for m in sys.modules if m == NAMESPACE or m.startswith(NAMESPACE + ".")]:
        del sys.modules[name]
