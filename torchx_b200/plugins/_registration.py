"""``@register`` decorators for plugin modules, the ``Resource.tags`` keys they set, and helpers that derive the
fractional variants of a whole-host named resource (reference torchx/plugins/_registration.py:36-427).

Import from ``torchx_b200.plugins``.  A decorator only TAGS the function (``_plugin_type``, ``_plugin_name``); discovery
(:mod:`torchx_b200.plugins._registry`) imports the modules under the ``torchx_b200_plugins.*`` namespace packages and
collects what is tagged, so importing a plugin module has no global side effect.

    # torchx_b200_plugins/named_resources/my_box.py
    from torchx_b200.plugins import register, powers_of_two_gpus
    from torchx_b200.specs import Resource

    @register.named_resource(aliases=["hgx"], fractionals=powers_of_two_gpus)
    def b200_box(fractional: float = 1.0) -> Resource:      # -> b200_box, hgx, b200_box_8, b200_box_4, b200_box_2, b200_box_1
        return Resource(cpu=int(192 * fractional), gpu=int(8 * fractional), memMB=int(2048 * 1024 * fractional))
"""
from __future__ import annotations

import functools
import sys
from typing import TYPE_CHECKING, Any, Callable, Dict, List, Optional, Union

from torchx_b200.plugins._registry import NAMED_RESOURCES_ATTR, PluginType

if TYPE_CHECKING:
    from torchx_b200.specs.api import Resource

WHOLE: float = 1.0
HALF: float = 0.5
QUARTER: float = 0.25
EIGHTH: float = 0.125
SIXTEENTH: float = 0.0625

Fractionals = Union[Callable[[Any], Dict[float, str]], Dict[float, str]]


class resource_tags:
    """Keys ``register.named_resource`` writes into ``Resource.tags`` (read back by ``Resource.get_resource_name`` /
    ``Resource.is_fractional``).  The key strings are TorchX's, so resources cross between the two packages."""

    RESOURCE_NAME: str = "torchx/named_resources.name"
    IS_FRACTIONAL: str = "torchx/named_resources.is_fractional"


def powers_of_two_gpus(resource: Any) -> Dict[float, str]:
    """``{1.0: "8", 0.5: "4", 0.25: "2", 0.125: "1"}`` for an 8-GPU host: one slice per power of two of its GPU count."""
    gpus = resource.gpu
    if gpus <= 0:
        raise ValueError(f"resource must have gpu > 0 to generate power-of-two slices (got {gpus})")
    if gpus & (gpus - 1):
        raise ValueError(f"resource.gpu must be a power of two to generate slices (got {gpus})")
    out: Dict[float, str] = {}
    share = gpus
    while share >= 1:
        out[share / gpus] = str(share)
        share //= 2
    return out


def halve_mem_down_to(*, minGiB: int) -> Callable[[Any], Dict[float, str]]:
    """Slices by halving the host's memory: a 64 GiB host with ``minGiB=8`` gives ``{1.0: "64", 0.5: "32", 0.25: "16",
    0.125: "8"}``.  The memory must be whole GiB and ``minGiB`` no smaller than its odd factor (else a halving step would
    not be a whole GiB)."""

    def slices(resource: Any) -> Dict[float, str]:
        mem_mb = resource.memMB
        if mem_mb <= 0:
            raise ValueError(f"resource must have memMB > 0 to generate memory slices (got {mem_mb})")
        if mem_mb % 1024:
            raise ValueError(f"resource.memMB must be a whole number of GiB (got {mem_mb} MB)")
        gib = mem_mb // 1024
        odd = gib
        while odd % 2 == 0:
            odd //= 2
        if minGiB < odd:
            raise ValueError(f"`minGiB` must be >= the odd part of `memGiB` ({odd}) because halving {gib} GiB below {odd} "
                             f"produces non-integer GiB values (got minGiB={minGiB})")
        out: Dict[float, str] = {}
        fraction, share = 1.0, gib
        while share >= minGiB and share >= 1:
            out[fraction] = str(share)
            if share % 2:
                break
            fraction, share = fraction / 2, share // 2
        return out

    return slices


class register:
    """``@register.scheduler()`` / ``@register.named_resource()``; ``register(PluginType.X, name=...)`` is the explicit form."""

    def __init__(self, type: PluginType, name: Optional[str] = None) -> None:  # noqa: A002 - the reference's keyword
        self._type = type
        self._name = name

    def __call__(self, fn: Callable[..., Any]) -> Callable[..., Any]:
        fn._plugin_type = self._type  # type: ignore[attr-defined]
        fn._plugin_name = self._name or fn.__name__  # type: ignore[attr-defined]
        return fn

    @classmethod
    def scheduler(cls, name: Optional[str] = None) -> "register":
        """Tag ``fn(session_name, **kwargs) -> Scheduler`` as the factory of scheduler ``name`` (default: ``fn.__name__``)."""
        return cls(PluginType.SCHEDULER, name=name)

    @classmethod
    def tracker(cls, name: Optional[str] = None) -> "register":
        """Tag a tracker factory (discovered and reported like the others; nothing in this package consumes trackers)."""
        return cls(PluginType.TRACKER, name=name)

    @classmethod
    def named_resource(cls, name: Optional[str] = None, aliases: Optional[List[str]] = None,
                       fractionals: Optional[Fractionals] = None) -> "_register_named_resource":
        """Tag ``fn() -> Resource`` (``fn(fractional: float = 1.0)`` when ``fractionals`` is given) as named resource
        ``name``, plus ``aliases`` and one generated ``<name>_<suffix>`` factory per ``{fraction: suffix}`` entry."""
        return _register_named_resource(name=name, aliases=aliases, fractionals=fractionals)


class _register_named_resource(register):
    """Every factory it produces returns Resources tagged with the name they were asked for by and whether they are a
    slice; the factories are also set as module attributes and listed in the module's ``NAMED_RESOURCES`` dict, the older
    discovery convention."""

    def __init__(self, name: Optional[str] = None, aliases: Optional[List[str]] = None, fractionals: Optional[Fractionals] = None) -> None:
        super().__init__(PluginType.NAMED_RESOURCE, name=name)
        self._aliases = list(aliases or [])
        self._fractionals: Optional[Callable[[Any], Dict[float, str]]]
        if fractionals is None or callable(fractionals):
            self._fractionals = fractionals  # type: ignore[assignment]
        else:
            table = dict(fractionals)
            self._fractionals = lambda _resource: table

    # -- hooks a subclass may override (e.g. to stamp platform metadata on the Resources) -----------------------------
    def _make_factory(self, fn: Callable[..., "Resource"], name: str) -> Callable[..., "Resource"]:
        """The factory registered under the base ``name`` (and, re-labelled, under each alias)."""
        return _tagging(fn, name, is_fractional=False)

    def _make_fractional(self, fn: Callable[..., "Resource"], fraction: float, frac_name: str) -> Callable[[], "Resource"]:
        """The zero-argument factory registered as ``frac_name`` = ``fn(fraction)``."""
        return _tagging(functools.partial(fn, fraction), frac_name, is_fractional=fraction != WHOLE, like=fn)

    def __call__(self, fn: Callable[..., "Resource"]) -> Callable[..., "Resource"]:
        mod = sys.modules[fn.__module__]
        name = self._name or fn.__name__
        table = getattr(mod, NAMED_RESOURCES_ATTR, None)
        if table is None:
            table = {}
            setattr(mod, NAMED_RESOURCES_ATTR, table)

        def publish(reg_name: str, factory: Callable[..., Any], base: Optional[str] = None, alias: bool = False) -> None:
            if reg_name in table:
                raise ValueError(f"duplicate named resource `{reg_name}` in module `{mod.__name__}`")
            if reg_name != fn.__name__ and hasattr(mod, reg_name):
                raise AttributeError(f"`{reg_name}()` already exists in `{mod.__name__}`")
            factory._plugin_type = PluginType.NAMED_RESOURCE  # type: ignore[attr-defined]
            factory._plugin_name = reg_name  # type: ignore[attr-defined]
            if base is not None:
                factory._plugin_base_name = base  # type: ignore[attr-defined]
                factory._plugin_is_alias = alias  # type: ignore[attr-defined]
            table[reg_name] = factory
            if reg_name != fn.__name__:
                setattr(mod, reg_name, factory)

        whole = self._make_factory(fn, name)
        publish(name, whole)
        for alias in self._aliases:  # another name for the SAME resource: same tags, no fractional variants of its own

            def aliased(*args: Any, _to: Callable[..., Any] = whole, **kwargs: Any) -> Any:
                return _to(*args, **kwargs)

            aliased.__module__, aliased.__qualname__, aliased.__name__ = whole.__module__, alias, alias
            publish(alias, aliased, base=name, alias=True)
        if self._fractionals:
            for fraction, suffix in self._fractionals(fn()).items():
                frac_name = f"{name}_{suffix}"
                publish(frac_name, self._make_fractional(fn, fraction, frac_name), base=name)
        return whole


def _tagging(produce: Callable[..., "Resource"], reg_name: str, is_fractional: bool, like: Optional[Callable[..., Any]] = None) -> Callable[..., "Resource"]:
    """``produce`` wrapped so that what it returns says which registered name made it and whether it is a slice."""
    like = like or produce

    def factory(*args: Any, **kwargs: Any) -> "Resource":
        res = produce(*args, **kwargs)
        res.tags.setdefault(resource_tags.RESOURCE_NAME, reg_name)
        res.tags.setdefault(resource_tags.IS_FRACTIONAL, is_fractional)
        return res

    factory.__module__ = like.__module__
    factory.__qualname__ = factory.__name__ = reg_name
    factory.__doc__ = like.__doc__
    return factory
