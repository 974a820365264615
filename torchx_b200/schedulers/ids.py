"""Unique app ids: ``{name}-{suffix}`` where the suffix is 64 bits from /dev/urandom spelled with characters that are
valid in hostnames / k8s object names and never form words (no vowels) - same scheme as reference
torchx/schedulers/ids.py:18-76 so ids look and sort alike."""
import os

_FIRST = "bcdfghjklmnpqrstvwxz"  # an id never starts with a digit
_REST = _FIRST + "012345679"


def random_uint64() -> int:
    return int.from_bytes(os.urandom(8), "big")


def random_id(max_length=None) -> str:
    if max_length is not None and max_length <= 0:
        return ""
    value, chars = random_uint64(), []
    while value > 0 and (max_length is None or len(chars) < max_length):
        alphabet = _REST if chars else _FIRST
        value, digit = divmod(value, len(alphabet))
        chars.append(alphabet[digit])
    return "".join(chars)


def make_unique(name: str, string_length: int = 0) -> str:
    return f"{name}-{random_id(string_length or None)}"
