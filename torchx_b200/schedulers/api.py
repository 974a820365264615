"""The scheduler plugin contract (launcher-side drop-in boundary).

Names and call protocol follow reference torchx/schedulers/api.py (StructuredOpts:79, Stream:323,
DescribeAppResponse:330, ListAppResponse:346, Scheduler:364, filter_regex:528, split_lines:535): a scheduler written
for TorchX subclasses ``Scheduler`` here unchanged, and ``Runner`` drives both ``local_cwd`` and ``local_cuda``
through exactly these methods.
"""
from __future__ import annotations

import abc
import dataclasses
import inspect
import re
import typing
from dataclasses import dataclass, field
from datetime import datetime
from enum import Enum
from typing import Any, Dict, Generic, Iterable, Iterator, List, Mapping, Optional, TypeVar, Union

from torchx_b200.specs.api import (
    NONE,
    NULL_RESOURCE,
    AppDef,
    AppDryRunInfo,
    AppState,
    CfgVal,
    Role,
    RoleStatus,
    cases,
    runopts,
)


# ---------------------------------------------------------------------------------------------------------------
# typed run options
# ---------------------------------------------------------------------------------------------------------------
def _plain(tp: Any) -> Any:
    """``Optional[X]`` / ``X | None`` -> ``X``."""
    args = typing.get_args(tp)
    if args and type(None) in args and (typing.get_origin(tp) is Union or "UnionType" in type(tp).__name__):
        rest = [a for a in args if a is not type(None)]
        if len(rest) == 1:
            return rest[0]
    return tp


def _is_opts_class(tp: Any) -> bool:
    return inspect.isclass(tp) and tp is not StructuredOpts and issubclass(tp, StructuredOpts)


_FIELD_DOC = re.compile(r'^[ \t]+(\w+)[ \t]*:[^\n]*\n[ \t]+(?:"""|\'\'\')(.+?)(?:"""|\'\'\')', re.M | re.S)


class StructuredOpts(Mapping[str, CfgVal]):
    """Declare a scheduler's ``-cfg`` options as dataclass fields.

    ``as_runopts()`` derives the ``runopts`` schema (type from the annotation, default from the field, required iff
    no default, help from the attribute docstring); ``from_cfg()`` turns a resolved cfg mapping back into a typed
    instance (camelCase keys accepted; nested ``StructuredOpts`` fields use dotted keys).  Instances also behave as a
    read-only mapping so code written against raw cfg dicts keeps working.
    """

    # -- schema ------------------------------------------------------------------------------------------------
    @classmethod
    def _typed_fields(cls):
        hints = typing.get_type_hints(cls)
        for f in dataclasses.fields(cls):  # type: ignore[arg-type]
            yield f, _plain(hints.get(f.name, str))

    @classmethod
    def get_docstrings(cls) -> Dict[str, str]:
        docs: Dict[str, str] = {}
        for klass in reversed(cls.__mro__):  # inherited option groups keep their help text
            if not _is_opts_class(klass):
                continue
            try:
                src = inspect.getsource(klass)
            except (OSError, TypeError):
                continue
            for name, text in _FIELD_DOC.findall(src):
                docs[name] = " ".join(text.split())
        for f, tp in cls._typed_fields():
            if _is_opts_class(tp):
                docs.update({f"{f.name}.{k}": v for k, v in tp.get_docstrings().items()})
        return docs

    @classmethod
    def as_runopts(cls) -> runopts:
        opts = runopts()
        docs = cls.get_docstrings()
        for f, tp in cls._typed_fields():
            if _is_opts_class(tp):
                for key, sub in tp.as_runopts():
                    opts.add(f"{f.name}.{key}", type_=sub.opt_type, default=sub.default, required=sub.is_required, help=sub.help)
                continue
            has_value = f.default is not dataclasses.MISSING
            has_factory = f.default_factory is not dataclasses.MISSING  # type: ignore[misc]
            opts.add(f.name, type_=tp, default=f.default if has_value else None, required=not (has_value or has_factory),
                     help=docs.get(f.name, f.name))
        return opts

    @classmethod
    def from_cfg(cls, cfg: Mapping[str, CfgVal]):
        kwargs: Dict[str, Any] = {}
        for f, tp in cls._typed_fields():
            if _is_opts_class(tp):
                prefix = f.name + "."
                sub = {k[len(prefix):]: v for k, v in cfg.items() if k.startswith(prefix)}
                no_default = f.default is dataclasses.MISSING and f.default_factory is dataclasses.MISSING  # type: ignore[misc]
                if sub or no_default:
                    kwargs[f.name] = tp.from_cfg(sub)
                continue
            for key in (f.name, cases.snake_to_camel(f.name)):
                if key in cfg:
                    kwargs[f.name] = cfg[key]
                    break
        return cls(**kwargs)

    # -- read-only mapping view ----------------------------------------------------------------------------------
    def __getitem__(self, key: str) -> CfgVal:
        head, _, rest = key.partition(".")
        attr = cases.camel_to_snake(head)
        if not hasattr(self, attr):
            raise KeyError(key)
        val = getattr(self, attr)
        if rest:
            if isinstance(val, StructuredOpts):
                return val[rest]
            raise KeyError(key)
        return val

    def get(self, key: str, default: CfgVal = None) -> CfgVal:  # type: ignore[override]
        try:
            return self[key]
        except KeyError:
            return default

    def __iter__(self) -> Iterator[str]:
        for f, tp in type(self)._typed_fields():
            val = getattr(self, f.name)
            if _is_opts_class(tp):
                if val is not None:
                    yield from (f"{f.name}.{k}" for k in val)
            else:
                yield f.name

    def __len__(self) -> int:
        return sum(1 for _ in self)

    def __contains__(self, key: object) -> bool:
        if not isinstance(key, str):
            return False
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __or__(self, other: "StructuredOpts") -> Dict[str, CfgVal]:  # type: ignore[override]
        merged = {k: self[k] for k in self}
        merged.update({k: other[k] for k in other})
        return merged


# ---------------------------------------------------------------------------------------------------------------
# responses
# ---------------------------------------------------------------------------------------------------------------
class Stream(str, Enum):
    STDOUT = "stdout"
    STDERR = "stderr"
    COMBINED = "combined"


@dataclass
class DescribeAppResponse:
    app_id: str = "<NOT_SET>"
    state: AppState = AppState.UNSUBMITTED
    num_restarts: int = -1
    msg: str = NONE
    structured_error_msg: str = NONE
    ui_url: Optional[str] = None
    metadata: Dict[str, str] = field(default_factory=dict)
    roles_statuses: List[RoleStatus] = field(default_factory=list)
    roles: List[Role] = field(default_factory=list)


@dataclass
class ListAppResponse:
    app_id: str
    state: AppState
    app_handle: str = "<NOT_SET>"
    name: str = ""

    def __hash__(self) -> int:
        return hash((self.app_id, self.app_handle, self.state))


# ---------------------------------------------------------------------------------------------------------------
# the ABC
# ---------------------------------------------------------------------------------------------------------------
T = TypeVar("T")


class Scheduler(abc.ABC, Generic[T]):
    """Implement ``_submit_dryrun`` (pure: builds the request, creates nothing), ``schedule`` (launches it, returns
    the app id), ``describe`` (``None`` for unknown apps; must be thread-safe - log threads run beside status polls),
    ``list``, ``_cancel_existing`` and optionally ``log_iter`` / ``_run_opts`` / ``close``."""

    def __init__(self, backend: str, session_name: str) -> None:
        self.backend = backend
        self.session_name = session_name

    def close(self) -> None:  # idempotent
        pass

    # -- submission -------------------------------------------------------------------------------------------------
    def submit(self, app: AppDef, cfg: T, workspace: Optional[Any] = None) -> str:
        """Submit directly (the Runner is the production route).  ``workspace`` is only meaningful for schedulers that mix in
        ``WorkspaceMixin`` (reference schedulers/api.py:383-403); it becomes ``roles[0].workspace`` and is built first."""
        if workspace:
            from torchx_b200.specs.api import Workspace
            from torchx_b200.workspace.api import WorkspaceMixin

            assert isinstance(self, WorkspaceMixin), f"{type(self).__name__} does not build workspaces"
            app.roles[0].workspace = Workspace.from_str(workspace) if isinstance(workspace, str) else workspace
            self.build_workspaces(app.roles, self.run_opts().resolve(cfg))  # type: ignore[arg-type]
        return self.schedule(self.submit_dryrun(app, cfg))

    def submit_dryrun(self, app: AppDef, cfg: T) -> AppDryRunInfo:
        resolved = self.run_opts().resolve(cfg)  # type: ignore[arg-type]
        info = self._submit_dryrun(app, resolved)  # type: ignore[arg-type]
        for role in app.roles:
            info = role.pre_proc(self.backend, info)
        info._app = app
        info._cfg = resolved
        return info

    @abc.abstractmethod
    def _submit_dryrun(self, app: AppDef, cfg: T) -> AppDryRunInfo:
        raise NotImplementedError

    @abc.abstractmethod
    def schedule(self, dryrun_info: AppDryRunInfo) -> str:
        raise NotImplementedError

    # -- options ----------------------------------------------------------------------------------------------------
    def run_opts(self) -> runopts:
        return self._run_opts()

    def _run_opts(self) -> runopts:
        return runopts()

    # -- lifecycle --------------------------------------------------------------------------------------------------
    @abc.abstractmethod
    def describe(self, app_id: str) -> Optional[DescribeAppResponse]:
        raise NotImplementedError

    @abc.abstractmethod
    def list(self, cfg: Optional[Mapping[str, CfgVal]] = None) -> List[ListAppResponse]:
        raise NotImplementedError

    def exists(self, app_id: str) -> bool:
        return self.describe(app_id) is not None

    @abc.abstractmethod
    def _cancel_existing(self, app_id: str) -> None:
        raise NotImplementedError

    def cancel(self, app_id: str) -> None:
        """Idempotent, non-blocking."""
        if self.exists(app_id):
            self._cancel_existing(app_id)

    def _delete_existing(self, app_id: str) -> None:
        self._cancel_existing(app_id)

    def delete(self, app_id: str) -> None:
        if self.exists(app_id):
            self._delete_existing(app_id)

    def log_iter(self, app_id: str, role_name: str, k: int = 0, regex: Optional[str] = None, since: Optional[datetime] = None,
                 until: Optional[datetime] = None, should_tail: bool = False, streams: Optional[Stream] = None) -> Iterable[str]:
        """Lines (with their trailing newline) of replica ``k`` of ``role_name``; blocks until the app is terminal
        when ``should_tail``."""
        raise NotImplementedError(f"{self.__class__.__qualname__} does not support application log iteration")

    # -- validation hooks --------------------------------------------------------------------------------------------
    def _pre_build_validate(self, app: AppDef, scheduler: str, cfg: T) -> None:
        pass

    def _validate(self, app: AppDef, scheduler: str, cfg: T) -> None:
        for role in app.roles:
            if role.resource == NULL_RESOURCE:
                raise ValueError(f"No resource for role: {role.image}. Did you forget to attach resource to the role")


# ---------------------------------------------------------------------------------------------------------------
# log helpers
# ---------------------------------------------------------------------------------------------------------------
def filter_regex(regex: str, data: Iterable[str]) -> Iterable[str]:
    pat = re.compile(regex)
    return (line for line in data if pat.search(line))


def split_lines(text: str) -> List[str]:
    """Split on newlines, keeping them."""
    return text.splitlines(keepends=True) if "\r" not in text else re.findall(r"[^\n]*\n|[^\n]+", text)


def split_lines_iterator(chunks: Iterable[str]) -> Iterable[str]:
    for chunk in chunks:
        yield from split_lines(chunk)
