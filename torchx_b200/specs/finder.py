"""Locate a component function by name (reference torchx/specs/finder.py get_component:437, CustomComponentsFinder:267).

Accepted forms:
  ``dist.ddp``                      builtin: module ``torchx_b200.components.dist``, function ``ddp``
  ``path/to/file.py:fn``            any python file (relative to cwd or absolute)
  ``my.pkg.module:fn``              any importable module
Builtin discovery lists the public, AppDef-annotated functions of the modules under ``torchx_b200.components`` - test
modules are never imported (the reference imports everything, which is why its builtin names need ``hydra`` here).
"""
from __future__ import annotations

import importlib
import importlib.util
import inspect
import os
import pkgutil
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional

from torchx_b200.specs.api import AppDef


class ComponentNotFoundException(Exception):
    pass


class ComponentValidationException(Exception):
    pass


@dataclass
class _Component:
    name: str
    description: str
    fn_name: str
    fn: Callable[..., AppDef]
    validation_errors: List[str]


def _returns_appdef(fn: Callable[..., object]) -> bool:
    ann = inspect.signature(fn).return_annotation
    return ann is AppDef or (isinstance(ann, str) and ann.split(".")[-1] == "AppDef")


def _validate(fn: Callable[..., object]) -> List[str]:
    errs = []
    sig = inspect.signature(fn)
    if not _returns_appdef(fn):
        errs.append(f"function {fn.__name__} must be annotated to return AppDef")
    for name, p in sig.parameters.items():
        if p.annotation is inspect.Parameter.empty:
            errs.append(f"parameter `{name}` of {fn.__name__} has no type annotation")
        if p.kind is inspect.Parameter.VAR_KEYWORD:
            errs.append(f"parameter `**{name}` of {fn.__name__}: **kwargs are not supported")
    return errs


def _wrap(name: str, fn: Callable[..., AppDef]) -> _Component:
    doc = (inspect.getdoc(fn) or "").strip().splitlines()
    return _Component(name=name, description=doc[0] if doc else "", fn_name=fn.__name__, fn=fn, validation_errors=_validate(fn))


def _module_components(module, prefix: str) -> Dict[str, _Component]:
    out: Dict[str, _Component] = {}
    for attr, fn in inspect.getmembers(module, inspect.isfunction):
        if attr.startswith("_") or fn.__module__ != module.__name__ or not _returns_appdef(fn):
            continue
        out[f"{prefix}.{attr}" if prefix else attr] = _wrap(f"{prefix}.{attr}" if prefix else attr, fn)
    return out


def get_builtin_components() -> Dict[str, _Component]:
    import torchx_b200.components as pkg

    found: Dict[str, _Component] = {}
    for info in pkgutil.iter_modules(pkg.__path__):
        if info.ispkg or info.name.startswith("_") or info.name in ("structured_arg",) or "test" in info.name:
            continue
        mod = importlib.import_module(f"{pkg.__name__}.{info.name}")
        found.update(_module_components(mod, info.name))
    return dict(sorted(found.items()))


get_components = get_builtin_components


def _load_from_file(path: str, fn_name: str) -> _Component:
    full = path if os.path.isabs(path) else os.path.join(os.getcwd(), path)
    if not os.path.isfile(full):
        raise ComponentNotFoundException(f"component file `{path}` does not exist")
    spec = importlib.util.spec_from_file_location(f"_torchx_component_{abs(hash(full))}", full)
    assert spec and spec.loader
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fn = getattr(mod, fn_name, None)
    if fn is None or not callable(fn):
        raise ComponentNotFoundException(f"function `{fn_name}` not found in `{path}`")
    return _wrap(f"{path}:{fn_name}", fn)


def get_component(name: str) -> _Component:
    if ":" in name:
        target, _, fn_name = name.rpartition(":")
        if target.endswith(".py") or os.sep in target:
            comp = _load_from_file(target, fn_name)
        else:
            try:
                mod = importlib.import_module(target)
            except ModuleNotFoundError as e:
                raise ComponentNotFoundException(f"cannot import module `{target}` for component `{name}`: {e}") from e
            fn = getattr(mod, fn_name, None)
            if fn is None:
                raise ComponentNotFoundException(f"function `{fn_name}` not found in module `{target}`")
            comp = _wrap(name, fn)
    else:
        builtins = get_builtin_components()
        if name not in builtins:
            raise ComponentNotFoundException(f"Component `{name}` not found. Please make sure it is one of the builtins: `torchx builtins`."
                                             f" Or registered via a python file: `path/to/file.py:fn`. Available: {list(builtins)}")
        comp = builtins[name]
    if comp.validation_errors:
        raise ComponentValidationException(f"Component {name} has validation errors: " + "; ".join(comp.validation_errors))
    return comp
