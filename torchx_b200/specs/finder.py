"""Locate component functions (reference torchx/specs/finder.py: ModuleComponentsFinder:148, CustomComponentsFinder:267,
get_components:382, get_component:437, get_builtin_source:466).

A component name is either
  ``dist.ddp``                      a name from the component table (below), or
  ``path/to/file.py:fn``            function ``fn`` of any python file (relative to cwd or absolute).
The component table is what the ``torchx_b200.components`` entry-point group registers (``alias = some.module``: every
AppDef-returning function of the module and its sub-modules, named ``alias.<relative module>.<fn>``; an alias starting
with ``_`` adds no prefix) - or, when nothing is registered, the builtins under ``torchx_b200.components``.

Validation is done on the live function objects (resolved annotations) rather than on the file's AST as the reference's
linter does: every parameter typed, types limited to what the CLI can parse (primitives, Optional of them, list / dict /
tuple of primitives), no ``**kwargs``, return annotated ``AppDef``.  A component with errors stays listed (so that
``get_component`` can explain what is wrong) but is not returned by ``get_components``.
"""
from __future__ import annotations

import abc
import importlib
import inspect
import os
import pkgutil
import typing
from dataclasses import dataclass
from types import ModuleType
from typing import Any, Callable, Dict, Iterator, List, Optional, Union

from torchx_b200.specs.api import AppDef
from torchx_b200.util import entrypoints

COMPONENTS_GROUP = "torchx_b200.components"
BUILTINS_MODULE = "torchx_b200.components"

# fn -> list of problems; the stand-in for the reference's AST ``ComponentFunctionValidator`` objects
Validator = Callable[[Callable[..., object]], List[str]]


class ComponentValidationException(Exception):
    pass


class ComponentNotFoundException(Exception):
    pass


@dataclass
class _Component:
    """``name`` is how the CLI addresses it; ``description`` the docstring's summary; ``validation_errors`` why it cannot
    be run (empty when it can)."""

    name: str
    description: str
    fn_name: str
    fn: Callable[..., AppDef]
    validation_errors: List[str]


# ---------------------------------------------------------------------------------------------------------------
# validation
# ---------------------------------------------------------------------------------------------------------------
_PRIMITIVES = (int, float, str, bool)


def _returns_appdef(fn: Callable[..., object]) -> bool:
    ann = inspect.signature(fn).return_annotation
    if isinstance(ann, str):
        return ann.replace(" ", "").split(".")[-1] == "AppDef"
    return ann is AppDef or getattr(ann, "__name__", None) == "AppDef"


def _type_problem(tp: Any) -> Optional[str]:
    """None when the CLI can parse a value of type ``tp`` from a string."""
    if tp is inspect.Parameter.empty:
        return "Missing type annotation"
    if isinstance(tp, str):  # an unresolved forward reference: judge by spelling
        bare = tp.replace("typing.", "").replace(" ", "")
        ok = bare.split("[")[0].split("|")[0] in ("int", "float", "str", "bool", "Optional", "List", "list", "Dict", "dict", "Tuple", "tuple")
        return None if ok else f"Unsupported argument type {tp!r}"
    if typing.get_origin(tp) is typing.Annotated:
        tp = typing.get_args(tp)[0]
    origin, args = typing.get_origin(tp), typing.get_args(tp)
    if origin is Union or (origin is not None and getattr(origin, "__name__", "") == "UnionType"):
        rest = [a for a in args if a is not type(None)]
        if len(rest) != 1:
            return f"Unsupported argument type {tp!r}"
        tp = rest[0]
        origin, args = typing.get_origin(tp), typing.get_args(tp)
    if tp in _PRIMITIVES:
        return None
    if origin in (list, dict, tuple):
        bad = [a for a in args if a is not Ellipsis and a not in _PRIMITIVES]
        return f"Non-primitive element type {bad[0]!r}" if bad else None
    return f"Unsupported argument type {tp!r}"


def _validate(fn: Callable[..., object], validators: Optional[List[Validator]] = None) -> List[str]:
    problems: List[str] = []
    try:
        hints = typing.get_type_hints(fn, include_extras=True)
    except Exception:  # noqa: BLE001 - unresolvable forward references: use the raw annotations
        hints = {}
    for name, p in inspect.signature(fn).parameters.items():
        if p.kind is inspect.Parameter.VAR_KEYWORD:
            problems.append(f"`**{name}` in function {fn.__name__!r}: keyword catch-alls cannot be filled from the command line")
            continue
        why = _type_problem(hints.get(name, p.annotation))
        if why:
            problems.append(f"{why} for argument {name!r} in function {fn.__name__!r}")
    if not _returns_appdef(fn):
        problems.append(f"Function: {fn.__name__} missing return annotation or has unsupported annotation: the function must return AppDef")
    for extra in validators or []:
        problems += list(extra(fn))
    return problems


def _describe(name: str, fn: Callable[..., AppDef], validators: Optional[List[Validator]]) -> _Component:
    from torchx_b200.specs.builders import get_fn_docstring

    return _Component(name=name, description=get_fn_docstring(fn)[0], fn_name=fn.__name__, fn=fn, validation_errors=_validate(fn, validators))


# ---------------------------------------------------------------------------------------------------------------
# finders
# ---------------------------------------------------------------------------------------------------------------
class ComponentsFinder(abc.ABC):
    @abc.abstractmethod
    def find(self, validators: Optional[List[Validator]]) -> List[_Component]:
        """All components this finder knows, valid or not."""


def is_namespace_package(module: ModuleType) -> bool:
    """A package without ``__init__.py`` (PEP 420)."""
    return hasattr(module, "__path__") and getattr(module, "__file__", None) is None


def is_package(module: ModuleType) -> bool:
    return hasattr(module, "__path__")


def module_relname(module: ModuleType, relative_to: ModuleType) -> str:
    """``a.b.c`` relative to ``a.b`` is ``c`` (``""`` for the module itself); ValueError when it is not below it."""
    name, base = module.__name__, relative_to.__name__
    if name == base:
        return ""
    if not name.startswith(base + "."):
        raise ValueError(f"`{base}` is not a parent of `{name}`")
    return name[len(base) + 1:]


class ModuleComponentsFinder(ComponentsFinder):
    """The components of ``module`` and, recursively, of its sub-modules (namespace packages are not entered), named
    ``<group>.<module path below the base>.<function>`` with empty parts dropped."""

    def __init__(self, module: Union[str, ModuleType], group: str) -> None:
        self.base_module: ModuleType = self._try_import(module)
        self.group = group

    @staticmethod
    def _try_import(module: Union[str, ModuleType]) -> ModuleType:
        return importlib.import_module(module) if isinstance(module, str) else module

    def _iter_modules_recursive(self, module: Union[str, ModuleType]) -> Iterator[ModuleType]:
        module = self._try_import(module)
        if not is_namespace_package(module):
            yield module
        if is_package(module):
            for info in pkgutil.iter_modules(module.__path__, prefix=f"{module.__name__}."):
                if info.ispkg:
                    yield from self._iter_modules_recursive(info.name)
                else:
                    yield self._try_import(info.name)

    def _get_components_from_module(self, module: ModuleType, validators: Optional[List[Validator]]) -> List[_Component]:
        rel = module_relname(module, relative_to=self.base_module)
        found = []
        for fn_name, fn in inspect.getmembers(module, inspect.isfunction):
            # a component is a function the module itself defines and annotates as returning an AppDef; imported helpers
            # and other functions are not listed (the reference lists them all and lets its linter reject them)
            if fn.__module__ != module.__name__ or not _returns_appdef(fn):
                continue
            found.append(_describe(".".join(p for p in (self.group, rel, fn_name) if p), fn, validators))
        return found

    def find(self, validators: Optional[List[Validator]]) -> List[_Component]:
        return [c for m in self._iter_modules_recursive(self.base_module) for c in self._get_components_from_module(m, validators)]


class CustomComponentsFinder(ComponentsFinder):
    """``function_name`` of the python file ``filepath``; the file is executed as its own module with ``__file__`` set, so
    components can locate resources next to themselves."""

    def __init__(self, filepath: str, function_name: str) -> None:
        self._filepath = filepath
        self._function_name = function_name

    def find(self, validators: Optional[List[Validator]]) -> List[_Component]:
        full = os.path.abspath(self._filepath)
        if not os.path.isfile(full):
            raise ComponentNotFoundException(f"component file `{self._filepath}` does not exist")
        spec = importlib.util.spec_from_file_location(f"_torchx_component_{abs(hash(full))}", full)
        assert spec and spec.loader
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        fn = getattr(module, self._function_name, None)
        if fn is None or not callable(fn):
            raise ComponentNotFoundException(f"Function {self._function_name} does not exist in file {self._filepath}")
        return [_describe(f"{self._filepath}:{self._function_name}", fn, validators)]


# ---------------------------------------------------------------------------------------------------------------
# the component table
# ---------------------------------------------------------------------------------------------------------------
def _load_custom_components(validators: Optional[List[Validator]]) -> List[_Component]:
    found: List[_Component] = []
    for alias, load in (entrypoints.load_group(COMPONENTS_GROUP, default={}) or {}).items():
        found += ModuleComponentsFinder(load(), "" if alias.startswith("_") else alias).find(validators)
    return found


def _load_components(validators: Optional[List[Validator]]) -> Dict[str, _Component]:
    """Registered component modules if there are any, otherwise the builtins - never both (a deployment that registers
    its own table lists the builtins it wants in it)."""
    found = _load_custom_components(validators) or ModuleComponentsFinder(BUILTINS_MODULE, "").find(validators)
    return {c.name: c for c in found}


_components: Optional[Dict[str, _Component]] = None  # reset to None to force a re-scan


def _find_components(validators: Optional[List[Validator]]) -> Dict[str, _Component]:
    global _components
    if not _components:
        _components = _load_components(validators)
    return _components


def _is_custom_component(component_name: str) -> bool:
    return ":" in component_name


def _find_custom_components(name: str, validators: Optional[List[Validator]]) -> Dict[str, _Component]:
    target, sep, fn_name = name.rpartition(":")
    if not sep or not target or not fn_name:
        raise ValueError(f"Invalid custom component: {name}, valid template : `FILEPATH`:`FUNCTION_NAME`")
    if target.endswith(".py") or os.sep in target or os.path.isfile(target):
        return {c.name: c for c in CustomComponentsFinder(target, fn_name).find(validators)}
    try:  # ``pkg.module:fn`` - an importable module instead of a file
        module = importlib.import_module(target)
    except ModuleNotFoundError as e:
        raise ComponentNotFoundException(f"cannot import module `{target}` for component `{name}`: {e}") from e
    fn = getattr(module, fn_name, None)
    if fn is None or not callable(fn):
        raise ComponentNotFoundException(f"Function {fn_name} does not exist in module {target}")
    return {name: _describe(name, fn, validators)}


def get_components(validators: Optional[List[Validator]] = None) -> Dict[str, _Component]:
    """The runnable entries of the component table (those without validation errors)."""
    return {name: c for name, c in _find_components(validators).items() if not c.validation_errors}


get_builtin_components = get_components  # round-1 name


def get_component(name: str, validators: Optional[List[Validator]] = None) -> _Component:
    """The component called ``name``; ComponentNotFoundException / ComponentValidationException explain a failure."""
    table = _find_custom_components(name, validators) if _is_custom_component(name) else _find_components(validators)
    if name not in table:
        raise ComponentNotFoundException(
            f"Component `{name}` not found. Please make sure it is one of the builtins: `torchx builtins`. Or registered via the "
            f"`[{COMPONENTS_GROUP}]` entry point, or given as `path/to/file.py:function`")
    comp = table[name]
    if comp.validation_errors:
        raise ComponentValidationException(f"Component {name} has validation errors: \n " + "\n".join(comp.validation_errors))
    return comp


def get_builtin_source(name: str, validators: Optional[List[Validator]] = None) -> str:
    """The source of component ``name`` preceded by the literal import lines of its file that come before it: a
    self-contained starting point for a customised copy (``torchx builtins --print dist.ddp > my_ddp.py``).  Imports the
    copy does not need may be included when the file defines several components."""
    comp = get_component(name, validators)
    imports: List[str] = []
    with open(inspect.getfile(comp.fn)) as f:
        for line in f:
            if line.startswith(("import ", "from ")):
                imports.append(line.rstrip("\n"))
            elif line.startswith(f"def {comp.fn_name}("):
                break
    return "\n".join([*imports, "", "", inspect.getsource(comp.fn).rstrip("\n"), ""])
