"""Decorators a plugin module uses, and the ``Resource.tags`` keys they set.

Import from ``torchx_b200.plugins``; this module exists under the reference's name so that
``from torchx.plugins._registration import resource_tags`` written against TorchX resolves
(reference torchx/plugins/_registration.py:46-68 tag keys, :187-274 decorators).  The fractional-slice generators of the
reference (``powers_of_two_gpus`` ...) describe cloud instance types and are not part of the single-box path.
"""
from __future__ import annotations

import functools
from typing import Any, Callable, Dict, Optional

SCHEDULERS: Dict[str, Callable[..., Any]] = {}
NAMED_RESOURCES: Dict[str, Callable[[], Any]] = {}


class resource_tags:
    """Keys ``register.named_resource`` writes into ``Resource.tags`` (read back by ``Resource.get_resource_name`` /
    ``Resource.is_fractional``).  The key strings are TorchX's, so resources cross between the two packages."""

    RESOURCE_NAME: str = "torchx/named_resources.name"
    IS_FRACTIONAL: str = "torchx/named_resources.is_fractional"


class register:
    """Decorators used INSIDE plugin modules."""

    @staticmethod
    def scheduler(name: Optional[str] = None) -> Callable[[Callable[..., Any]], Callable[..., Any]]:
        """Register ``fn(session_name, **kwargs) -> Scheduler`` under ``name`` (default: the function name)."""

        def deco(fn: Callable[..., Any]) -> Callable[..., Any]:
            SCHEDULERS[name or fn.__name__] = fn
            return fn

        return deco

    @staticmethod
    def named_resource(name: Optional[str] = None) -> Callable[[Callable[[], Any]], Callable[[], Any]]:
        """Register ``fn() -> Resource`` under ``name``; every Resource it returns carries its registered name in
        ``tags[resource_tags.RESOURCE_NAME]``."""

        def deco(fn: Callable[[], Any]) -> Callable[[], Any]:
            key = name or fn.__name__

            @functools.wraps(fn)
            def tagged() -> Any:
                res = fn()
                res.tags.setdefault(resource_tags.RESOURCE_NAME, key)
                return res

            NAMED_RESOURCES[key] = tagged
            return tagged

        return deco
