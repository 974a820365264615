"""``torchx_b200.specs`` - the data model components and schedulers share (reference torchx/specs/__init__.py)."""
import difflib
from typing import Callable, Dict, Optional

from .api import (  # noqa: F401
    ALL,
    MISSING,
    NONE,
    NULL_RESOURCE,
    TORCHX_HOME,
    AppDef,
    AppDryRunInfo,
    AppHandle,
    AppState,
    AppStatus,
    AppStatusError,
    BindMount,
    CfgVal,
    DeviceMount,
    InvalidRunConfigException,
    MalformedAppHandleException,
    ParsedAppHandle,
    ReplicaState,
    ReplicaStatus,
    Resource,
    RetryPolicy,
    Role,
    RoleStatus,
    UnknownAppException,
    UnknownSchedulerException,
    VolumeMount,
    Workspace,
    cases,
    get_type_name,
    is_started,
    is_terminal,
    macros,
    make_app_handle,
    parse_app_handle,
    runopt,
    runopts,
)
from .builders import materialize_appdef, parse_mounts  # noqa: F401,E402
from .named_resources_generic import NAMED_RESOURCES as _GENERIC

GiB: int = 1024


def _all_named_resources() -> Dict[str, Callable[[], Resource]]:
    merged: Dict[str, Callable[[], Resource]] = dict(_GENERIC)
    try:
        from torchx_b200.plugins import registered_named_resources

        merged.update(registered_named_resources())
    except Exception:  # pragma: no cover - plugin discovery must never break the core
        pass
    return merged


class _NamedResources:
    """Lazy mapping so plugin-registered resources show up without import-order games."""

    def __getitem__(self, key: str) -> Resource:
        table = _all_named_resources()
        if key.upper() in ("MISSING", "NULL"):
            return NULL_RESOURCE
        if key not in table:
            close = difflib.get_close_matches(key, table.keys(), n=1)
            hint = f"Did you mean `{close[0]}`?" if close else f"Registered named resources: {sorted(table)}"
            raise KeyError(f"No named resource found for `{key}`. {hint}")
        return table[key]()

    def __contains__(self, key: str) -> bool:
        return key in _all_named_resources()

    def keys(self):
        return _all_named_resources().keys()

    def __iter__(self):
        return iter(_all_named_resources())

    def __len__(self) -> int:
        return len(_all_named_resources())

    def items(self):
        return ((k, f()) for k, f in _all_named_resources().items())


named_resources = _NamedResources()


def resource(cpu: Optional[int] = None, gpu: Optional[int] = None, memMB: Optional[int] = None, h: Optional[str] = None) -> Resource:
    """``h`` (a named resource) wins over the raw values; unset (or zero) raw values default to cpu=2, gpu=0,
    memMB=1024 (reference torchx/specs/__init__.py:148-181)."""
    if h:
        return named_resources[h]
    return Resource(cpu=cpu or 2, gpu=gpu or 0, memMB=memMB or 1024)


def get_named_resources(res: str) -> Resource:
    return named_resources[res]
