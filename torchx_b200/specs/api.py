"""Job-definition vocabulary of the launcher: the plugin ABI's data model.

Same public names and semantics as reference torchx/specs/api.py (Resource:98, macros:183, RetryPolicy:277, Role:415,
AppDef:505, AppState:525, AppStatus:628, AppDryRunInfo:784, runopt:835, runopts:885, parse_app_handle:1204) so that
components, schedulers and the CLI written against TorchX keep working; the implementation is independent.
Mounts and ``Workspace`` are part of the data model (AppDefs stay interchangeable with TorchX's) but only container /
image-building schedulers act on them; the local schedulers of this package, like the reference's, do not.
"""
from __future__ import annotations

import asyncio
import copy
import inspect
import json
import logging
import os
import pathlib
import re
import shutil
import string
from dataclasses import asdict, dataclass, field
from datetime import datetime
from enum import Enum
from typing import Any, Callable, Dict, Generic, Iterator, List, Mapping, NamedTuple, Optional, Tuple, TypeVar, Union

from torchx_b200.util.types import to_dict

# ---------------------------------------------------------------------------------------------------------------
# sentinels
# ---------------------------------------------------------------------------------------------------------------
ALL: str = "all"  # "any scheduler backend"
MISSING: str = "<MISSING>"  # a required string attribute that was not provided
NONE: str = "<NONE>"  # an optional string attribute that is unset

logger = logging.getLogger(__name__)


def TORCHX_HOME(*subdir_paths: str) -> pathlib.Path:
    """The launcher's dot-directory (``$TORCHX_HOME``, else ``~/.torchx``), optionally a sub-directory of it; created on
    demand (reference torchx/specs/api.py:68-89)."""
    home = pathlib.Path(os.environ.get("TORCHX_HOME") or pathlib.Path.home() / ".torchx").joinpath(*subdir_paths)
    home.mkdir(parents=True, exist_ok=True)
    return home


# ---------------------------------------------------------------------------------------------------------------
# resources
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class Resource:
    """Per-replica resource request.  ``gpu`` is what ``local_cuda`` turns into one worker process per device."""

    cpu: int
    gpu: int
    memMB: int
    capabilities: Dict[str, Any] = field(default_factory=dict)
    devices: Dict[str, int] = field(default_factory=dict)
    tags: Dict[str, object] = field(default_factory=dict)

    def is_fractional(self) -> bool:
        """True when a named-resource registration tagged this as a slice of a bigger host (reference api.py:127-135)."""
        from torchx_b200.plugins._registration import resource_tags

        return bool(self.tags.get(resource_tags.IS_FRACTIONAL, False))

    def get_resource_name(self) -> Optional[str]:
        """The name this resource was registered under, if it came from ``register.named_resource`` (api.py:137-146)."""
        from torchx_b200.plugins._registration import resource_tags

        name = self.tags.get(resource_tags.RESOURCE_NAME)
        return None if name is None else str(name)

    @staticmethod
    def copy(original: "Resource", **capabilities: Any) -> "Resource":
        merged = {**original.capabilities, **capabilities}
        return Resource(cpu=original.cpu, gpu=original.gpu, memMB=original.memMB, capabilities=merged, devices=original.devices)


NULL_RESOURCE: Resource = Resource(cpu=-1, gpu=-1, memMB=-1)


def _null_resource() -> Resource:
    return NULL_RESOURCE


# ---------------------------------------------------------------------------------------------------------------
# mounts / workspace: carried in the AppDef for container schedulers; the local schedulers do not act on them
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class BindMount:
    """A host path made visible in the worker's container (reference specs/api.py:313-319)."""

    src_path: str
    dst_path: str
    read_only: bool = False


@dataclass
class VolumeMount:
    """A named persistent volume mounted at ``dst_path``."""

    src: str
    dst_path: str
    read_only: bool = False


@dataclass
class DeviceMount:
    """A host device node; ``permissions`` is a subset of ``rwm`` (read, write, mknod)."""

    src_path: str
    dst_path: str
    permissions: str = "rwm"


@dataclass
class Workspace:
    """``{local path: sub-directory in the job's image}`` - what an image-building scheduler copies before launch
    (reference specs/api.py:340-411).  An empty sub-directory means "into the image root"."""

    projects: Dict[str, str]

    def __bool__(self) -> bool:
        return bool(self.projects)

    def __eq__(self, other: object) -> bool:
        return isinstance(other, Workspace) and self.projects == other.projects

    def __hash__(self) -> int:
        return hash(frozenset(self.projects.items()))

    def is_unmapped_single_project(self) -> bool:
        """One project, copied to the image root."""
        return len(self.projects) == 1 and not next(iter(self.projects.values()))

    def merge_into(self, outdir: Union[str, "os.PathLike[str]"]) -> None:
        """Materialise the mapping under ``outdir`` (files copied, directories merged)."""
        for src, sub in self.projects.items():
            dst = pathlib.Path(outdir) / sub
            if pathlib.Path(src).is_file():
                shutil.copy2(src, dst)
            else:
                shutil.copytree(src, dst, dirs_exist_ok=True)

    @staticmethod
    def from_str(workspace: Optional[str]) -> "Workspace":
        """``"/one/dir"`` or a YAML mapping ``{"/a": "sub", "/b": null}`` (null = image root); empty -> no projects."""
        if not workspace:
            return Workspace({})
        import yaml

        parsed = yaml.safe_load(workspace)
        if isinstance(parsed, str):
            return Workspace({parsed: ""})
        return Workspace({k: (v or "") for k, v in parsed.items()})

    def __str__(self) -> str:  # for log lines; from_str() does not read this format back
        if self.is_unmapped_single_project():
            return next(iter(self.projects))
        return ";".join(f"{k}:{v}" if v else k for k, v in self.projects.items())


# ---------------------------------------------------------------------------------------------------------------
# macros
# ---------------------------------------------------------------------------------------------------------------
class macros:
    """``${...}`` placeholders a scheduler substitutes in ``Role.args``, ``Role.env`` and ``Role.metadata``
    just before launch (nowhere else)."""

    img_root = "${img_root}"
    base_img_root = "${base_img_root}"
    app_id = "${app_id}"
    replica_id = "${replica_id}"
    rank0_env = "${rank0_env}"  # NAME of the env var that holds rank 0's host; the app resolves it

    @dataclass
    class Values:
        img_root: str
        app_id: str
        replica_id: str
        rank0_env: str
        base_img_root: str = "DEPRECATED"

        def to_dict(self) -> Dict[str, Any]:
            return asdict(self)

        def substitute(self, arg: str) -> str:
            return string.Template(arg).safe_substitute(**self.to_dict())

        def _walk(self, node: Any) -> Any:
            if isinstance(node, str):
                return self.substitute(node)
            if isinstance(node, dict):
                for k in list(node):
                    node[k] = self._walk(node[k])
            elif isinstance(node, list):
                for i in range(len(node)):
                    node[i] = self._walk(node[i])
            return node

        def apply(self, role: "Role") -> "Role":
            """A deep copy of ``role`` with every macro resolved.  Lazy ``overrides`` (callables / awaitables) cannot be
            deep-copied: they are carried over by reference, as the reference does (api.py:232-243)."""
            lazy = role.overrides
            if lazy:
                logger.warning("Role overrides are not supported for macros. Overrides will not be copied")
            role.overrides = {}
            try:
                out = copy.deepcopy(role)
            finally:
                role.overrides = lazy
            out.overrides = lazy
            out.args = [self.substitute(a) for a in out.args]
            out.env = {k: self.substitute(v) for k, v in out.env.items()}
            out.metadata = self._walk(out.metadata)
            return out


# ---------------------------------------------------------------------------------------------------------------
# roles / apps
# ---------------------------------------------------------------------------------------------------------------
class RetryPolicy(str, Enum):
    """What to restart when a replica fails: just it (REPLICA), its role (ROLE) or everything (APPLICATION)."""

    REPLICA = "REPLICA"
    APPLICATION = "APPLICATION"
    ROLE = "ROLE"


def _resolve_awaitable(pending: Any) -> Any:
    """Drive ``pending`` to completion on a private event loop (attribute reads are synchronous)."""
    loop = asyncio.new_event_loop()
    try:
        return loop.run_until_complete(pending)
    finally:
        loop.close()


@dataclass
class Role:
    """A homogeneous group of replicas (for ``dist.ddp``: the "nodes", each running ``nproc_per_node`` workers)."""

    name: str
    image: str
    min_replicas: Optional[int] = None
    entrypoint: str = MISSING
    args: List[str] = field(default_factory=list)
    env: Dict[str, str] = field(default_factory=dict)
    num_replicas: int = 1
    max_retries: int = 0
    retry_policy: RetryPolicy = RetryPolicy.APPLICATION
    resource: Resource = field(default_factory=_null_resource)
    port_map: Dict[str, int] = field(default_factory=dict)
    metadata: Dict[str, Any] = field(default_factory=dict)
    mounts: List[Union[BindMount, VolumeMount, DeviceMount]] = field(default_factory=list)
    workspace: Optional[Workspace] = None
    # Deprecated in TorchX but still honoured: {attribute name: zero-arg callable | awaitable} evaluated on first read
    # of that attribute, then cached (reference api.py:469-487).
    overrides: Dict[str, Any] = field(default_factory=dict)

    def __getattribute__(self, attrname: str) -> Any:
        get = super().__getattribute__
        if attrname != "overrides" and not attrname.startswith("__"):
            try:
                lazy = get("overrides")
            except AttributeError:  # during dataclass __init__, before the field is assigned
                lazy = None
            if lazy and attrname in lazy:
                pending = lazy[attrname]
                value = _resolve_awaitable(pending) if inspect.isawaitable(pending) else pending()
                object.__setattr__(self, attrname, value)
                lazy[attrname] = lambda: value
        return get(attrname)

    def pre_proc(self, scheduler: str, dryrun_info: "AppDryRunInfo") -> "AppDryRunInfo":
        """Per-role hook to amend the scheduler request; called by ``Scheduler.submit_dryrun`` in role order."""
        return dryrun_info


@dataclass
class AppDef:
    name: str
    roles: List[Role] = field(default_factory=list)
    metadata: Dict[str, str] = field(default_factory=dict)


# ---------------------------------------------------------------------------------------------------------------
# state / status
# ---------------------------------------------------------------------------------------------------------------
class AppState(int, Enum):
    UNSUBMITTED = 0
    SUBMITTED = 1
    PENDING = 2
    RUNNING = 3
    SUCCEEDED = 4
    FAILED = 5
    CANCELLED = 6
    UNKNOWN = 7

    def __str__(self) -> str:
        return self.name

    def __repr__(self) -> str:
        return f"{self.name} ({self.value})"


ReplicaState = AppState
_TERMINAL_STATES: List[AppState] = [AppState.SUCCEEDED, AppState.FAILED, AppState.CANCELLED]
_STARTED_STATES: List[AppState] = _TERMINAL_STATES + [AppState.RUNNING]


def is_terminal(state: AppState) -> bool:
    return state in _TERMINAL_STATES


def is_started(state: AppState) -> bool:
    return state in _STARTED_STATES


@dataclass
class ReplicaStatus:
    id: int
    state: ReplicaState
    role: str
    hostname: str
    structured_error_msg: str = NONE
    hostaddr: Optional[str] = None

    def __post_init__(self) -> None:
        if self.hostaddr is None:
            self.hostaddr = self.hostname


@dataclass
class RoleStatus:
    role: str
    replicas: List[ReplicaStatus]

    def to_json(self) -> Dict[str, Any]:
        return {"role": self.role, "replicas": [asdict(r) for r in self.replicas]}


# ``ExcType('On WorkerInfo(...):\n<lines>\n')`` - the envelope torch.distributed.rpc wraps remote errors in; and a message
# followed by a C++ ``Exception raised from ...`` trailer.  Only the informative part is shown by ``torchx status``.
_RPC_ENVELOPE = re.compile(r"\w*\('On WorkerInfo\(.+\):\n(.*\n)*'\)")
_CPP_TRAILER = re.compile(r"(?P<msg>.+)\nException.*")


def _wrap(text: str, header: str, width: int = 80) -> str:
    """Soft-wrap as the reference's status printer does (api.py:663-684): a line runs to the first space at or after
    ``width`` characters, the space starts the next line; ``header`` leads the first line, blanks of its width the rest."""
    assert len(header) < width
    m = _RPC_ENVELOPE.search(text)
    if m:
        text = m.group(0)
    m = _CPP_TRAILER.search(text)
    if m:
        text = m.group("msg")
    pieces = re.findall(r"(?s).{%d}[^ ]*|.+" % width, text) or [""]
    pad = " " * len(header)
    return "\n".join((header if i == 0 else pad) + piece for i, piece in enumerate(pieces))


@dataclass
class AppStatus:
    """Runtime status; ``roles`` describes the most recent attempt only."""

    state: AppState
    num_restarts: int = 0
    msg: str = ""
    structured_error_msg: str = NONE
    ui_url: Optional[str] = None
    roles: List[RoleStatus] = field(default_factory=list)

    def is_terminal(self) -> bool:
        return is_terminal(self.state)

    def raise_for_status(self) -> None:
        if self.state != AppState.SUCCEEDED:
            raise AppStatusError(self, f"job did not succeed: {self}")

    def to_json(self, filter_roles: Optional[List[str]] = None) -> Dict[str, Any]:
        return {
            "state": str(self.state),
            "num_restarts": self.num_restarts,
            "roles": [r.to_json() for r in self._select(filter_roles)],
            "msg": self.msg,
            "structured_error_msg": self.structured_error_msg,
            "url": self.ui_url,
        }

    def _format_error_message(self, msg: str, header: str, width: int = 80) -> str:
        return _wrap(msg, header, width)

    def _select(self, filter_roles: Optional[List[str]]) -> List[RoleStatus]:
        return [r for r in self.roles if not filter_roles or r.role in filter_roles]

    @staticmethod
    def _describe_replica(rs: ReplicaStatus) -> str:
        text = str(rs.state)
        if rs.structured_error_msg != NONE:
            try:
                err = json.loads(rs.structured_error_msg)["message"]
                code = err.get("errorCode") or "<N/A>"
                when = datetime.fromtimestamp(int(err["extraInfo"]["timestamp"]))
                text += f" (exitcode: {code})\n        timestamp: {when}\n        hostname: {rs.hostname}\n"
                text += _wrap(str(err.get("message", "")), "    error_msg: ")
            except (ValueError, KeyError, TypeError):
                text = rs.structured_error_msg
        elif rs.state in (ReplicaState.CANCELLED, ReplicaState.FAILED):
            text += " (no reply file)"
        marker = "*" if rs.id == 0 else " "
        return f"\n {marker}{rs.role}[{rs.id}]:{text}"

    def format(self, filter_roles: Optional[List[str]] = None) -> str:
        body = "".join(
            self._describe_replica(rep) for role in self._select(filter_roles) for rep in sorted(role.replicas, key=lambda r: r.id)
        )
        return (
            f"AppStatus:\n  State: {self.state}\n  Num Restarts: {self.num_restarts}\n  Roles: {body}\n"
            f"  Msg: {self.msg}\n  Structured Error Msg: {self.structured_error_msg}\n  UI URL: {self.ui_url}\n"
        )

    def __repr__(self) -> str:
        d = asdict(self)
        sem = d.pop("structured_error_msg")
        try:
            d["structured_error_msg"] = json.loads(sem) if sem != NONE else NONE
        except ValueError:
            d["structured_error_msg"] = sem
        d["state"] = repr(self.state)
        for role in d["roles"]:
            for rep in role["replicas"]:
                rep["state"] = repr(AppState(rep["state"]))
        try:
            import yaml

            return yaml.safe_dump({"AppStatus": d})
        except Exception:
            return json.dumps({"AppStatus": d}, indent=2, default=str)


class AppStatusError(Exception):
    def __init__(self, status: AppStatus, *args: object) -> None:
        super().__init__(*args)
        self.status = status


# ---------------------------------------------------------------------------------------------------------------
# dry-run info
# ---------------------------------------------------------------------------------------------------------------
CfgVal = Union[str, int, float, bool, List[str], Dict[str, str], None]
T = TypeVar("T")


class AppDryRunInfo(Generic[T]):
    """What ``Scheduler.submit_dryrun`` returns: the native request that WOULD be submitted, printable."""

    def __init__(self, request: T, fmt: Callable[[T], str]) -> None:
        self.request = request
        self._fmt = fmt
        self._app: Optional[AppDef] = None
        self._cfg: Mapping[str, CfgVal] = {}
        self._scheduler: Optional[str] = None

    def __repr__(self) -> str:
        return self._fmt(self.request)


# ---------------------------------------------------------------------------------------------------------------
# run options (scheduler cfg schema)
# ---------------------------------------------------------------------------------------------------------------
def get_type_name(tp: Any) -> str:
    """``int`` -> "int", ``list[str]`` -> "list", ``typing.List[str]`` -> "typing.List[str]" (reference api.py:822-829)."""
    if getattr(tp, "__module__", "typing") != "typing" and hasattr(tp, "__name__"):
        return tp.__name__
    return str(tp)


class cases:
    @staticmethod
    def snake_to_camel(name: str) -> str:
        head, *rest = name.split("_")
        return head + "".join(part.title() for part in rest)

    @staticmethod
    def camel_to_snake(name: str) -> str:
        return re.sub(r"([a-z0-9])([A-Z])", r"\1_\2", name).lower()


def _is_list_of_str(tp: Any) -> bool:
    return tp in (List[str], list[str])


def _is_dict_of_str(tp: Any) -> bool:
    return tp in (Dict[str, str], dict[str, str])


@dataclass
class runopt:
    default: CfgVal
    opt_type: Any
    is_required: bool
    help: str

    @property
    def is_type_list_of_str(self) -> bool:
        return _is_list_of_str(self.opt_type)

    @property
    def is_type_dict_of_str(self) -> bool:
        return _is_dict_of_str(self.opt_type)

    def cast_to_type(self, value: str) -> CfgVal:
        """CLI literal -> typed value: ``"True"``->bool, ``"a,b"``/``"a;b"``->list, ``"k:v,k2:v2"``->dict
        (``:`` because ``=`` already separates cfg keys from values)."""
        tp = self.opt_type
        if tp is None:
            raise ValueError("runopt's opt_type cannot be `None`")
        if tp is bool:
            return value.lower() == "true"
        if _is_list_of_str(tp):
            return [v for v in value.replace(";", ",").split(",") if v]
        if _is_dict_of_str(tp):
            pairs = [kv.split(":", 1) for kv in value.replace(";", ",").split(",") if kv]
            return {k: v for k, v in pairs}
        if tp not in (str, int, float):
            raise ValueError(f"unsupported run option type {tp}")
        return tp(value)


class InvalidRunConfigException(Exception):
    def __init__(self, invalid_reason: str, cfg_key: str, cfg: Mapping[str, CfgVal]) -> None:
        super().__init__(f"{invalid_reason}. Given: {str(cfg) if cfg else '<EMPTY>'}")
        self.cfg_key = cfg_key


class runopts:
    """Schema of a scheduler's ``-cfg`` options: accepted keys, types, defaults, help."""

    def __init__(self) -> None:
        self._opts: Dict[str, runopt] = {}

    def __iter__(self) -> Iterator[Tuple[str, runopt]]:
        return iter(self._opts.items())

    def __len__(self) -> int:
        return len(self._opts)

    @staticmethod
    def is_type(obj: CfgVal, tp: Any) -> bool:
        """``isinstance`` that tolerates generic aliases: for ``List[str]`` / ``Dict[str, str]`` style types (where
        ``isinstance`` raises) a list passes when all its elements are str and a dict when all keys and values are -
        the reference is equally lenient about WHICH generic it was (api.py:912-927), and callers rely on that for
        empty defaults."""
        try:
            return isinstance(obj, tp)
        except TypeError:
            if isinstance(obj, list):
                return all(isinstance(e, str) for e in obj)
            if isinstance(obj, dict):
                return all(isinstance(k, str) and isinstance(v, str) for k, v in obj.items())
            return False

    def add(self, cfg_key: str, type_: Any, help: str, default: CfgVal = None, required: bool = False) -> None:
        if required and default is not None:
            raise ValueError(f"Required option: {cfg_key} must not specify default value. Given: {default}")
        if default is not None and not runopts.is_type(default, type_):
            raise TypeError(f"Option: {cfg_key}, must be of type: {type_}. Given: {default} ({type(default).__name__})")
        self._opts[cfg_key] = runopt(default, type_, required, help)

    def get(self, name: str) -> Optional[runopt]:
        """Lookup by snake_case or camelCase name."""
        return self._opts.get(name) or self._opts.get(cases.camel_to_snake(name))

    def update(self, other: "runopts") -> None:
        self._opts.update(other._opts)

    def __or__(self, other: "runopts") -> "runopts":
        merged = runopts()
        merged.update(self)
        merged.update(other)
        return merged

    def resolve(self, cfg: Mapping[str, CfgVal]) -> Dict[str, CfgVal]:
        """Validate ``cfg`` (types, required keys), accept camelCase aliases, fill defaults.  Unknown keys pass
        through untouched."""
        out: Dict[str, CfgVal] = dict(cfg)
        for key, opt in self._opts.items():
            if key not in out:
                alias = cases.snake_to_camel(key)
                if alias != key and alias in out:
                    out[key] = out.pop(alias)
            val = out.get(key)
            if opt.is_required and val is None:
                raise InvalidRunConfigException(f"Required run option: {key}, must be provided and not `None`", key, cfg)
            if val is not None and not runopts.is_type(val, opt.opt_type):
                raise InvalidRunConfigException(
                    f"Run option: {key}, must be of type: {get_type_name(opt.opt_type)}, but was: {val} ({type(val).__name__})", key, cfg)
            if key not in out:
                out[key] = opt.default
        return out

    def cfg_from_str(self, cfg_str: str) -> Dict[str, CfgVal]:
        """``"k1=v1,k2=a;b"`` -> typed dict; unknown keys are dropped (with a warning).  Does not fill defaults:
        chain with :py:meth:`resolve`."""
        import logging

        out: Dict[str, CfgVal] = {}
        for key, val in to_dict(cfg_str).items():
            opt = self.get(key)
            if opt is None:
                logging.getLogger(__name__).warning("Unknown run option passed to scheduler: %s=%s", key, val)
                continue
            out[key] = opt.cast_to_type(val)
        return out

    def cfg_from_json_repr(self, json_repr: str) -> Dict[str, CfgVal]:
        out: Dict[str, CfgVal] = {}
        for key, val in json.loads(json_repr).items():
            opt = self.get(key)
            if opt is None:
                continue
            if val is not None and opt.is_type_list_of_str:
                val = [str(v) for v in val]
            elif val is not None and opt.is_type_dict_of_str:
                val = {str(k): str(v) for k, v in val.items()}
            out[key] = val
        return out

    def __repr__(self) -> str:
        req = [(k, o) for k, o in self._opts.items() if o.is_required]
        opt = [(k, o) for k, o in self._opts.items() if not o.is_required]
        usage = ",".join((f"{k}={k.upper()}" if o.is_required else f"[{k}={k.upper()}]") for k, o in req + opt)
        out = f"    usage:\n        {usage}"
        for title, group in (("required", req), ("optional", opt)):
            if not group:
                continue
            out += f"\n\n    {title} arguments:"
            for k, o in group:
                dflt = "" if o.is_required else f", {o.default}"
                out += f"\n        {k}={k.upper()} ({get_type_name(o.opt_type)}{dflt})\n            {o.help}"
        return out


# ---------------------------------------------------------------------------------------------------------------
# app handles
# ---------------------------------------------------------------------------------------------------------------
AppHandle = str


class ParsedAppHandle(NamedTuple):
    scheduler_backend: str
    session_name: str
    app_id: str


class MalformedAppHandleException(Exception):
    def __init__(self, app_handle: str) -> None:
        super().__init__(f"{app_handle} is not of the form: <scheduler_backend>://<session_name>/<app_id>")


class UnknownSchedulerException(Exception):
    def __init__(self, scheduler_backend: str) -> None:
        super().__init__(f"Scheduler backend: {scheduler_backend} does not exist. Use session.scheduler_backends() to see all supported schedulers")


class UnknownAppException(Exception):
    def __init__(self, app_handle: "AppHandle") -> None:
        super().__init__(f"Unknown app = {app_handle}. Did you forget to call session.run()? Otherwise, the app may have already finished and purged by the scheduler")


_HANDLE = re.compile(r"(?P<scheduler_backend>.+)://(?P<session_name>.*)/(?P<app_id>.+)")


def parse_app_handle(app_handle: AppHandle) -> ParsedAppHandle:
    """``"local_cuda://torchx/train-abc"`` -> ``("local_cuda", "torchx", "train-abc")``."""
    m = _HANDLE.match(app_handle)
    if not m:
        raise MalformedAppHandleException(app_handle)
    return ParsedAppHandle(m["scheduler_backend"], m["session_name"], m["app_id"])


def make_app_handle(scheduler_backend: str, session_name: str, app_id: str) -> AppHandle:
    return f"{scheduler_backend}://{session_name}/{app_id}"


__all__ = [n for n in dir() if not n.startswith("_")] + ["_TERMINAL_STATES", "_STARTED_STATES"]
