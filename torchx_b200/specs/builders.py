"""Component function + CLI tokens -> ``AppDef``.

A component is a plain function returning ``AppDef`` whose parameters are type-annotated and documented in a
google-style ``Args:`` block.  This module derives an ``argparse`` parser from the signature (``--name`` per parameter,
``-x`` as well for one-letter names, ``*args`` as trailing REMAINDER), decodes the strings into the annotated types and
calls the function - the same contract as reference torchx/specs/builders.py (_create_args_parser:41,
component_args_from_str:155, materialize_appdef:244), so ``dist.ddp -j 1x8 --script x.py -- --lr 0.1`` parses alike.
"""
from __future__ import annotations

import argparse
import inspect
import os
import re
import typing
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Callable, Dict, List, Mapping, Optional, Tuple, Union

from torchx_b200.specs.api import AppDef, BindMount, DeviceMount, VolumeMount, make_app_handle  # noqa: F401
from torchx_b200.util.types import decode, decode_optional, get_argparse_param_type, is_bool

_ARGS_HEADER = re.compile(r"^\s*(Args|Arguments|Parameters)\s*:\s*$")
_SECTION = re.compile(r"^\s*(Returns|Return|Raises|Yields|Example|Examples|Note|Notes|Usage)\b.*:\s*$")
_PARAM = re.compile(r"^(\s*)\*{0,2}(\w+)\s*(\([^)]*\))?\s*:\s*(.*)$")


def get_fn_docstring(fn: Callable[..., object]) -> Tuple[str, Dict[str, str]]:
    """(function description, {param: help}) from a google-style docstring; undocumented params get a placeholder."""
    doc = inspect.getdoc(fn) or ""
    lines = doc.splitlines()
    desc_lines: List[str] = []
    params: Dict[str, str] = {}
    in_args, in_desc, indent, current = False, True, None, None
    for line in lines:
        if _ARGS_HEADER.match(line):
            in_args, in_desc = True, False
            continue
        if not in_args:
            if _SECTION.match(line):  # Returns: / Raises: / ... - the description is what precedes the first section
                in_desc = False
            if in_desc:
                desc_lines.append(line)
            continue
        if _SECTION.match(line) and (indent is None or len(line) - len(line.lstrip()) < indent):
            in_args = False
            continue
        m = _PARAM.match(line)
        if m and (indent is None or len(m.group(1)) <= indent):
            indent = len(m.group(1))
            current = m.group(2)
            params[current] = m.group(4).strip()
        elif current is not None and line.strip():
            params[current] = (params[current] + " " + line.strip()).strip()
    desc = "\n".join(desc_lines).strip() or f"{fn.__name__} TIP: improve this help string by adding a docstring to your component"
    for name in inspect.signature(fn).parameters:
        params.setdefault(name, " ")
    return desc, params


def _annotations(fn: Callable[..., object]) -> Dict[str, Any]:
    """Resolved parameter annotations (component modules may use ``from __future__ import annotations``)."""
    try:
        hints = typing.get_type_hints(fn)
    except Exception:  # noqa: BLE001 - unresolvable forward refs: fall back to the raw annotations
        hints = {}
    return {name: hints.get(name, p.annotation) for name, p in inspect.signature(fn).parameters.items()}


class _RemainderWithDefault(argparse.Action):
    def __call__(self, parser, namespace, values, option_string=None):  # type: ignore[override]
        setattr(namespace, self.dest, (self.default or "").split() if len(values) == 0 else values)


def _create_args_parser(cmpnt_fn: Callable[..., AppDef], cmpnt_defaults: Optional[Dict[str, str]] = None,
                        config: Optional[Dict[str, Any]] = None) -> argparse.ArgumentParser:
    desc, help_of = get_fn_docstring(cmpnt_fn)
    parser = argparse.ArgumentParser(prog=f"torchx run <run args...> {cmpnt_fn.__name__} ", description=desc,
                                     formatter_class=argparse.RawDescriptionHelpFormatter, add_help=False)
    parser.add_argument("--help", action="help", default=argparse.SUPPRESS, help="show this help message and exit")
    ann = _annotations(cmpnt_fn)
    for name, param in inspect.signature(cmpnt_fn).parameters.items():
        kw: Dict[str, Any] = {"help": help_of.get(name, " "), "type": get_argparse_param_type(ann[name])}
        if param.default is not inspect.Parameter.empty:
            kw["default"] = str(param.default) if is_bool(type(param.default)) else param.default
        if cmpnt_defaults and name in cmpnt_defaults:
            kw["default"] = cmpnt_defaults[name]
        if param.kind is inspect.Parameter.VAR_POSITIONAL:
            kw.update(nargs=argparse.REMAINDER, action=_RemainderWithDefault)
            parser.add_argument(name, **kw)
            continue
        if param.kind is inspect.Parameter.VAR_KEYWORD:
            raise TypeError(f"component fn param `{name}` is a '**kwargs' which is not supported; consider changing the"
                            f" type to a dict or explicitly declare the params")
        flags = [f"--{name}"]
        if len(name) == 1:
            flags.insert(0, f"-{name}")
        if "default" not in kw and not (config and name in config):
            kw["required"] = True
        parser.add_argument(*flags, **kw)
    return parser


def parse_args(cmpnt_fn: Callable[..., AppDef], cmpnt_args: List[str], cmpnt_defaults: Optional[Dict[str, Any]] = None,
               config: Optional[Dict[str, Any]] = None) -> argparse.Namespace:
    ns = _create_args_parser(cmpnt_fn, cmpnt_defaults, config).parse_args(cmpnt_args)
    for key, val in (config or {}).items():
        if key in ns:
            setattr(ns, key, val)
    return ns


@dataclass
class ComponentArgs:
    positional_args: Dict[str, Any] = field(default_factory=dict)
    var_args: List[str] = field(default_factory=list)
    kwargs: Dict[str, Any] = field(default_factory=dict)


def component_args_from_str(cmpnt_fn: Callable[..., AppDef], cmpnt_args: List[str], cmpnt_args_defaults: Optional[Dict[str, Any]] = None,
                            config: Optional[Dict[str, Any]] = None) -> ComponentArgs:
    ns = parse_args(cmpnt_fn, cmpnt_args, cmpnt_args_defaults, config)
    out = ComponentArgs()
    ann = _annotations(cmpnt_fn)
    for name, param in inspect.signature(cmpnt_fn).parameters.items():
        raw = getattr(ns, name)
        if param.kind is inspect.Parameter.VAR_POSITIONAL:
            out.var_args = list(raw[1:] if raw and raw[0] == "--" else raw)
            continue
        want = decode_optional(ann[name])
        value = decode(raw, want) if want != raw.__class__ else raw
        if param.kind is inspect.Parameter.KEYWORD_ONLY:
            out.kwargs[name] = value
        else:
            out.positional_args[name] = value
    return out


def materialize_appdef(cmpnt_fn: Callable[..., AppDef], cmpnt_args: List[str], cmpnt_defaults: Optional[Dict[str, Any]] = None,
                       config: Optional[Dict[str, Any]] = None) -> AppDef:
    """Parse ``cmpnt_args`` against the component's signature and call it."""
    ca = component_args_from_str(cmpnt_fn, cmpnt_args, cmpnt_defaults, config)
    app = cmpnt_fn(*ca.positional_args.values(), *ca.var_args, **ca.kwargs)
    if not isinstance(app, AppDef):
        raise TypeError(f"Expected a component that returns `AppDef`, but got `{type(app)}`")
    return app


# ---------------------------------------------------------------------------------------------------------------
# mounts
# ---------------------------------------------------------------------------------------------------------------
class MountType(str, Enum):
    BIND = "bind"
    VOLUME = "volume"
    DEVICE = "device"


# accepted spellings -> canonical option name (docker's --mount vocabulary)
_MOUNT_OPT_MAP: Mapping[str, str] = {"type": "type", "src": "src", "source": "src", "dst": "dst", "destination": "dst", "target": "dst",
                                     "readonly": "readonly", "read_only": "readonly", "perm": "perm"}


def parse_mounts(opts: List[str]) -> List[Union[BindMount, VolumeMount, DeviceMount]]:
    """``["type=bind", "src=/host", "dst=/job", "readonly", "type=device", "src=/dev/infiniband"]`` -> typed mounts
    (reference specs/builders.py:311-376).  Every mount starts with its ``type=``; a device's ``dst`` defaults to its
    ``src`` and its ``perm`` to ``rwm``; a leading ``~`` in a bind source is expanded."""
    groups: List[Dict[str, str]] = []
    for opt in opts:
        key, _, val = opt.partition("=")
        if key not in _MOUNT_OPT_MAP:
            raise KeyError(f"unknown mount option {key}, must be one of {list(_MOUNT_OPT_MAP.keys())}")
        key = _MOUNT_OPT_MAP[key]
        if key == "type":
            groups.append({})
        elif not groups:
            raise KeyError("type must be specified first")
        groups[-1][key] = val
    mounts: List[Union[BindMount, VolumeMount, DeviceMount]] = []
    for g in groups:
        kind = g.get("type")
        if kind == MountType.BIND:
            src = os.path.expanduser(g["src"]) if g["src"].startswith("~") else g["src"]
            mounts.append(BindMount(src_path=src, dst_path=g["dst"], read_only="readonly" in g))
        elif kind == MountType.VOLUME:
            mounts.append(VolumeMount(src=g["src"], dst_path=g["dst"], read_only="readonly" in g))
        elif kind == MountType.DEVICE:
            perm = g.get("perm", "rwm")
            wrong = [c for c in perm if c not in "rwm"]
            if wrong:
                raise ValueError(f"{wrong[0]} is not a valid permission flags must one of r,w,m")
            mounts.append(DeviceMount(src_path=g["src"], dst_path=g.get("dst", g["src"]), permissions=perm))
        else:
            raise ValueError(f"invalid mount type {kind!r}, must be one of {[m.value for m in MountType]}")
    return mounts
