"""Plugin discovery, caching and diagnostics (reference torchx/plugins/_registry.py:38-552).

Two discovery channels, selectable with the ``TORCHX_PLUGINS_SOURCE`` bitmask (1 = namespace packages, 2 = entry points,
default both): modules under the ``torchx_b200_plugins.{schedulers,named_resources}`` namespace packages anywhere on
``sys.path`` whose functions were tagged by ``@register...``, and ``importlib.metadata`` entry-point groups named like the
:class:`PluginType` values (these win over namespace plugins of the same name).  A module that fails to import, a plugin in
the wrong namespace or a duplicate name never takes the launcher down: it is recorded in :attr:`PluginRegistry.errors`
and shows up in ``print(plugins.registry())``.
"""
from __future__ import annotations

import dataclasses
import enum
import functools
import importlib
import logging
import os
import pathlib
import pkgutil
from types import ModuleType
from typing import Any, Callable, Dict, Iterator, List, Optional, Set, Tuple, Union

from torchx_b200.util import entrypoints

logger = logging.getLogger(__name__)

NAMESPACE = "torchx_b200_plugins"
NAMED_RESOURCES_ATTR = "NAMED_RESOURCES"
ENV_PLUGINS_SOURCE = "TORCHX_PLUGINS_SOURCE"


class PluginType(str, enum.Enum):
    """Kinds of plugin; the value is the entry-point group, its last component the namespace sub-package."""

    SCHEDULER = "torchx_b200.schedulers"
    NAMED_RESOURCE = "torchx_b200.named_resources"
    TRACKER = "torchx_b200.tracker"  # discoverable for API parity; this package has no tracker subsystem consuming it


class PluginSource(enum.IntFlag):
    NONE = 0
    NAMESPACE_PKG = enum.auto()
    ENTRYPOINT = enum.auto()


@dataclasses.dataclass(frozen=True)
class RegistrationError:
    """``module`` failed to import (``name`` unset), or plugin ``name`` in it was rejected."""

    module: str
    error: str
    name: Optional[str] = None
    plugin_type: Optional[str] = None


Factory = Callable[..., Any]


class PluginRegistry:
    """Lazily discovered, cached ``{PluginType: {name: factory}}``."""

    def __init__(self, *, plugin_sources: PluginSource = PluginSource.NAMESPACE_PKG | PluginSource.ENTRYPOINT) -> None:
        self._plugin_sources = plugin_sources
        self._cache: Dict[PluginType, Dict[str, Factory]] = {}
        self._errors: List[RegistrationError] = []

    # -- namespace-package channel ---------------------------------------------------------------------------------
    @staticmethod
    def _namespace_for_type(pt: PluginType) -> str:
        return f"{NAMESPACE}.{pt.value.rsplit('.', 1)[-1]}"

    def _import(self, fqn: str) -> Optional[ModuleType]:
        try:
            return importlib.import_module(fqn)
        except Exception as e:  # noqa: BLE001 - a broken plugin must not break the launcher
            logger.warning("failed to import `%s`: %s", fqn, e)
            self._errors.append(RegistrationError(module=fqn, error=f"{type(e).__name__}: {e}"))
            return None

    def _walk(self, pkg: ModuleType, namespace: str) -> Iterator[Tuple[str, ModuleType]]:
        """Every importable public module below ``pkg``, depth first: regular sub-modules and sub-packages, then
        directories without ``__init__.py`` that hold ``.py`` files (implicit namespace sub-packages)."""
        try:
            listed = list(pkgutil.iter_modules(pkg.__path__))
        except Exception as e:  # noqa: BLE001
            logger.warning("failed to scan `%s`: %s", namespace, e)
            self._errors.append(RegistrationError(module=namespace, error=f"failed to scan `{namespace}`: {e}"))
            return
        seen: Set[str] = {info.name for info in listed}
        children: List[Tuple[str, bool]] = [(info.name, info.ispkg) for info in listed]
        for search_path in pkg.__path__:
            root = pathlib.Path(search_path)
            if root.is_dir():
                for child in sorted(root.iterdir()):
                    if child.is_dir() and child.name not in seen and any(child.glob("*.py")):
                        seen.add(child.name)
                        children.append((child.name, True))
        for name, is_pkg in children:
            if name.startswith("_"):
                continue
            fqn = f"{namespace}.{name}"
            mod = self._import(fqn)
            if mod is None:
                continue
            yield fqn, mod
            if is_pkg and hasattr(mod, "__path__"):
                yield from self._walk(mod, fqn)

    def _harvest(self, mod: ModuleType, fqn: str, expected: Optional[PluginType], into: Dict[str, Factory]) -> None:
        """Tagged callables DEFINED in ``mod`` (not merely imported into it)."""
        for attr in dir(mod):
            obj = getattr(mod, attr)
            tagged = getattr(obj, "_plugin_type", None)
            if tagged is None or not callable(obj) or getattr(obj, "__module__", None) != fqn:
                continue
            name = getattr(obj, "_plugin_name", attr)
            if expected is not None and tagged != expected:
                why = (f"is a {tagged.name.lower()} but is under the {expected.name.lower()} namespace — use "
                       f"@register.{expected.name.lower()}() or move to `{self._namespace_for_type(tagged)}`")
                logger.warning("`%s` in `%s` %s", name, fqn, why)
                self._errors.append(RegistrationError(module=fqn, error=why, name=name, plugin_type=tagged.name.lower()))
            elif name in into and into[name] is not obj:
                logger.warning("duplicate plugin `%s` in `%s`", name, fqn)
                self._errors.append(RegistrationError(module=fqn, error="duplicate — already discovered, keeping first occurrence",
                                                      name=name, plugin_type=(expected or tagged).name.lower()))
            else:
                into[name] = obj

    def _find_namespace_plugins(self, namespace: str, expected_type: Optional[PluginType] = None) -> Dict[str, Factory]:
        try:
            pkg = importlib.import_module(namespace)
        except ImportError:
            return {}
        if not hasattr(pkg, "__path__"):
            return {}
        found: Dict[str, Factory] = {}
        for fqn, mod in self._walk(pkg, namespace):
            self._harvest(mod, fqn, expected_type, found)
        return found

    def _find(self, plugin_type: PluginType) -> Dict[str, Factory]:
        """Namespace plugins, overlaid by entry points of the same name (reference _registry.py:493-515)."""
        found: Dict[str, Factory] = {}
        if PluginSource.NAMESPACE_PKG in self._plugin_sources:
            found = self._find_namespace_plugins(self._namespace_for_type(plugin_type), plugin_type)
        if PluginSource.ENTRYPOINT in self._plugin_sources:
            found.update(entrypoints.load_group(plugin_type.value) or {})
        return found

    # -- public ----------------------------------------------------------------------------------------------------
    def get(self, plugin_type: PluginType) -> Dict[str, Factory]:
        """``{name: factory}`` for one plugin type (``{}`` when there are none); discovered once."""
        if plugin_type not in self._cache:
            self._cache[plugin_type] = self._find(plugin_type)
        return self._cache[plugin_type]

    def info(self, plugin_type: Optional[PluginType] = None) -> Union[Dict[PluginType, Dict[str, Factory]], Dict[str, Factory]]:
        """Copies: of one type's plugins, or (no argument) of everything, discovering all types first."""
        if plugin_type is not None:
            return dict(self.get(plugin_type))
        return {pt: dict(self.get(pt)) for pt in PluginType}

    def clear(self) -> None:
        """Forget everything, including the :func:`registry` singleton."""
        self._cache.clear()
        self._errors.clear()
        registry.cache_clear()

    @property
    def errors(self) -> List[RegistrationError]:
        return list(self._errors)

    def to_dict(self) -> Dict[str, Any]:
        """Plain data for ``json`` / ``yaml``: one list per plugin type (aliases and fractional variants folded into
        their base entry, rejected plugins listed with their ``error``) plus import-level ``errors``."""
        everything = self.info()
        rejected: Dict[str, List[RegistrationError]] = {}
        import_errors = []
        for err in self._errors:
            if err.plugin_type is None:
                import_errors.append({"module": err.module, "error": err.error})
            else:
                rejected.setdefault(err.plugin_type, []).append(err)
        data: Dict[str, Any] = {}
        for pt in PluginType:
            group = everything[pt]  # type: ignore[index]
            aliases: Dict[str, List[str]] = {}
            fractionals: Dict[str, List[str]] = {}
            for name, fn in group.items():
                base = getattr(fn, "_plugin_base_name", None)
                if base is not None:
                    (aliases if getattr(fn, "_plugin_is_alias", False) else fractionals).setdefault(base, []).append(name)
            items: List[Dict[str, Any]] = []
            for name, fn in group.items():
                if getattr(fn, "_plugin_base_name", None) is not None:
                    continue
                entry: Dict[str, Any] = {"name": name, "module": getattr(fn, "__module__", "unknown")}
                if name in aliases:
                    entry["aliases"] = aliases[name]
                if name in fractionals:
                    entry["fractionals"] = fractionals[name]
                items.append(entry)
            items += [{"name": e.name, "module": e.module, "error": e.error} for e in rejected.get(pt.name.lower(), [])]
            data[pt.name.lower()] = items
        data["errors"] = import_errors
        return data

    def __str__(self) -> str:
        """The :meth:`to_dict` data as YAML (loadable with ``yaml.safe_load``); free-text values double-quoted."""

        def quoted(text: str) -> str:
            return '"' + text.replace("\\", "\\\\").replace('"', '\\"') + '"'

        data = self.to_dict()
        out: List[str] = []
        for pt in PluginType:
            key = pt.name.lower()
            if not data[key]:
                out.append(f"{key}: []")
                continue
            out.append(f"{key}:")
            for item in data[key]:
                out += [f"  - name: {item['name']}", f"    module: {item['module']}"]
                for extra in ("aliases", "fractionals"):
                    if extra in item:
                        out.append(f"    {extra}: [{', '.join(item[extra])}]")
                if "error" in item:
                    out.append(f"    error: {quoted(item['error'])}")
        if not data["errors"]:
            out.append("errors: []")
        else:
            out.append("errors:")
            for err in data["errors"]:
                out += [f"  - module: {err['module']}", f"    error: {quoted(err['error'])}"]
        return "\n".join(out)


@functools.lru_cache(maxsize=1)
def registry() -> PluginRegistry:
    """The process-wide registry.  ``TORCHX_PLUGINS_SOURCE`` (integer :class:`PluginSource` bitmask) selects channels."""
    everything = PluginSource.NAMESPACE_PKG | PluginSource.ENTRYPOINT
    raw = os.environ.get(ENV_PLUGINS_SOURCE)
    if raw is None:
        return PluginRegistry(plugin_sources=everything)
    try:
        bits = int(raw)
    except ValueError:
        bits = -1
    if not 0 <= bits <= int(everything):
        raise ValueError(f"{ENV_PLUGINS_SOURCE}={raw!r}: expected an integer bitmask 0..{int(everything)} "
                         f"(1 = namespace packages, 2 = entry points)")
    return PluginRegistry(plugin_sources=PluginSource(bits))
