"""String normalisation for identifiers that end up in resource names (reference torchx/util/strings.py)."""
import re

_KEEP = re.compile(r"[a-z0-9\-]+")


def normalize_str(data: str) -> str:
    """Lower-case, drop one leading dash and everything outside ``[a-z0-9-]`` (DNS-label-safe app names)."""
    if data[:1] == "-":
        data = data[1:]
    return "".join(_KEEP.findall(data.lower()))
