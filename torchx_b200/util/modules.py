"""Import-by-name helpers for optional dependencies and ``module:attr`` strings (reference torchx/util/modules.py:13-65)."""
from __future__ import annotations

import importlib
from types import ModuleType
from typing import Any, Optional, TypeVar, Union

T = TypeVar("T")


def load_module(path: str) -> Optional[Union[ModuleType, Any]]:
    """``"pkg.mod"`` -> the module, ``"pkg.mod:attr"`` -> that attribute, ``None`` if either cannot be had.  Parent
    packages are imported outermost first, so a failure in ``pkg/__init__`` is reported as "not loadable" as well."""
    module_path, _, attr = path.partition(":")
    attr = attr.split(":", 1)[0]
    try:
        module = None
        prefix = ""
        for part in module_path.split("."):
            prefix = f"{prefix}.{part}" if prefix else part
            module = importlib.import_module(prefix)
        return getattr(module, attr) if attr else module
    except Exception:  # noqa: BLE001 - "cannot be loaded" is the answer, whatever the reason
        return None


def import_attr(name: str, attr: str, default: T) -> T:
    """``getattr(import_module(name), attr)``, or ``default`` when module ``name`` is not installed.  A module that IS
    there but lacks ``attr`` raises AttributeError: that is a bug, not a missing optional dependency."""
    try:
        return getattr(importlib.import_module(name), attr)
    except ModuleNotFoundError:
        return default
