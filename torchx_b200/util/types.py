"""String <-> typed-value helpers for component arguments and scheduler cfg strings.

Behavioural twin of reference torchx/util/types.py (to_dict:24, decode:163, decode_optional:231): the same CLI
literals must parse to the same Python values so `torchx run ... dist.ddp -j 1x8 --env A=1,B=2` is a drop-in.
"""
from __future__ import annotations

import inspect
import typing
from typing import Any, Callable, Dict, List, Optional, Tuple, Union, get_args, get_origin

_DELIMS = ",;"


def _scan(text: str) -> Tuple[List[int], List[int]]:
    """Positions of '=' signs and of ',' / ';' delimiters that sit outside single/double quotes."""
    eqs: List[int] = []
    delims: List[int] = []
    quote = ""
    i = 0
    while i < len(text):
        ch = text[i]
        if quote:
            if ch == "\\" and i + 1 < len(text):
                i += 1
            elif ch == quote:
                quote = ""
        elif ch in "'\"":
            quote = ch
        elif ch == "=":
            eqs.append(i)
        elif ch in _DELIMS:
            delims.append(i)
        i += 1
    return eqs, delims


def _unquote(val: str) -> str:
    if len(val) >= 2 and val[0] == val[-1] and val[0] in "'\"":
        return val[1:-1]
    return val


def to_dict(arg: str) -> Dict[str, str]:
    """``"K1=v1,v2,K2=v3"`` -> ``{"K1": "v1,v2", "K2": "v3"}``.

    Pairs are separated by ``,`` or ``;`` and a value may itself contain those characters (list literals are
    returned verbatim), so the LAST delimiter before the next ``=`` is the one that ends a pair.  Values may be
    quoted (``'...'`` / ``"..."``) to protect ``= , ;``.  Empty input gives ``{}``; input without any
    ``key=value`` raises ``ValueError`` (same contract as reference torchx/util/types.py:24-129).
    """
    if not arg or not arg.strip():
        return {}
    eqs, delims = _scan(arg)
    # Runs of '=' act as one separator and '=' signs with nothing after them are ignored, so a value cannot END in '='
    # ("BAR=v3==" parses as BAR=v3) - a documented limitation of the reference's grammar (types_test.py:168-179).
    merged: List[int] = []
    for e in eqs:
        if not arg[e + 1:].strip(" ="):
            continue
        if merged and not arg[merged[-1] + 1:e].strip(" ="):
            continue
        merged.append(e)
    eqs = merged
    if not eqs:
        raise ValueError(f"`{arg}` does not have at least one `key=value` pair")
    key_starts = [0]
    for prev, cur in zip(eqs, eqs[1:]):
        between = [d for d in delims if prev < d < cur]
        if not between:
            raise ValueError(f"`{arg[prev + 1:cur]}` cannot be split into `val<delim>key` with delims={list(_DELIMS)}")
        key_starts.append(between[-1] + 1)
    out: Dict[str, str] = {}
    for i, eq in enumerate(eqs):
        key = arg[key_starts[i]:eq].strip()
        stop = key_starts[i + 1] - 1 if i + 1 < len(eqs) else len(arg)
        raw = arg[eq + 1:stop].strip().lstrip("=").strip() if i + 1 < len(eqs) else arg[eq + 1:stop].strip().strip("=").strip()
        if not key or not raw:
            raise ValueError(f"malformed key-value pair near `{arg[key_starts[i]:stop]}` in `{arg}`")
        out[key] = _unquote(raw)
    return out


def to_list(arg: str) -> List[str]:
    return [tok.strip() for tok in arg.split(",")] if arg.strip() else []


def _strip_optional(tp: Any) -> Tuple[Any, bool]:
    """``Optional[X]`` / ``Union[X, None]`` / ``X | None`` -> (X, True); anything else -> (tp, False)."""
    import types as _types

    if get_origin(tp) is typing.Annotated:
        tp = get_args(tp)[0]
    if get_origin(tp) is Union or isinstance(tp, _types.UnionType):
        args = get_args(tp)
        if len(args) == 2 and args[1] is type(None):
            return args[0], True
    return tp, False


def is_bool(tp: Any) -> bool:
    return _strip_optional(tp)[0] is bool


def is_primitive(tp: Any) -> bool:
    return _strip_optional(tp)[0] in (int, float, str, bool)


def _to_bool(text: str) -> bool:
    return text.strip().lower() == "true"  # anything else is False, as in the reference (types.py:169-170)


def decode(value: Any, annotation: Any) -> Any:
    """Decode a CLI string into ``annotation`` (primitives, ``list[T]``, ``dict[K, V]``, and their Optionals)."""
    if value is None:
        return None
    tp, _ = _strip_optional(annotation)
    if not isinstance(value, str):
        return value
    origin = get_origin(tp)
    if tp is bool:
        return _to_bool(value)
    if tp in (int, float, str):
        return tp(value)
    if origin in (list, List):
        (elem,) = get_args(tp) or (str,)
        if not is_primitive(elem):
            raise ValueError("List types support only primitives: int, str, float")
        return [decode(v, elem) for v in to_list(value)]
    if origin in (dict, Dict):
        kt, vt = get_args(tp) or (str, str)
        return {decode(k, kt): decode(v, vt) for k, v in to_dict(value).items()}
    if tp is inspect.Parameter.empty or tp is Any:
        return value
    raise ValueError(f"unsupported argument type {annotation!r} for value {value!r}")


def decode_optional(annotation: Any) -> Any:
    return _strip_optional(annotation)[0]


def none_throws(x: Optional[Any], msg: str = "Unexpected `None`") -> Any:
    if x is None:
        raise AssertionError(msg)
    return x


def get_argparse_param_type(tp: Any) -> Callable[[str], Any]:
    """argparse ``type=`` for a parameter: the primitive itself for int / float / str, ``str`` for everything that is
    decoded later (lists, dicts, bools, Optionals).  Accepts an annotation or an ``inspect.Parameter``."""
    if isinstance(tp, inspect.Parameter):
        tp = tp.annotation
    if get_origin(tp) is typing.Annotated:
        tp = get_args(tp)[0]
    return tp if tp in (int, float, str) else str


def decode_from_string(encoded_value: str, annotation: Any) -> Any:
    """``"a,b"`` -> list / ``"k=v,k2=v2"`` -> dict according to ``annotation``; empty input gives None; anything that is
    not a list or dict type raises ``ValueError`` (reference torchx/util/types.py:174-207)."""
    if not encoded_value:
        return None
    tp = annotation
    if get_origin(tp) is typing.Annotated:
        tp = get_args(tp)[0]
    if get_origin(tp) in (list, List, dict, Dict):
        return decode(encoded_value, tp)
    raise ValueError(f"cannot decode {encoded_value!r} as {annotation!r}: only list and dict types are string-encoded")
