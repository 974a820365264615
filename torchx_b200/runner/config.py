"""``.torchxconfig``: INI defaults for scheduler cfg, the ``run`` sub-command and component arguments.

Same file format, sections and lookup order as reference torchx/runner/config.py (apply:295, load_sections:336,
get_config:446): ``$TORCHXCONFIG`` if set, else ``$HOME/.torchxconfig`` overlaid by ``./.torchxconfig`` (cwd wins);
values given on the command line always win over the file.

    [local_cuda]
    log_dir = /tmp/torchx_logs
    pin_cpus = True

    [cli:run]
    scheduler = local_cuda
    component = dist.ddp

    [component:dist.ddp]
    j = 1x8
"""
from __future__ import annotations

import configparser
import os
from pathlib import Path
from typing import Dict, List, Optional

from torchx_b200.settings import ENV_TORCHXCONFIG
from torchx_b200.specs.api import CfgVal, runopts

CONFIG_FILE = ".torchxconfig"
_NONE = "None"


def find_configs(dirs: Optional[List[str]] = None) -> List[str]:
    """Config files to read, lowest priority first."""
    explicit = os.environ.get(ENV_TORCHXCONFIG)
    if explicit:
        if not Path(explicit).is_file():
            raise FileNotFoundError(f"`{ENV_TORCHXCONFIG}={explicit}` does not exist or is not a file")
        return [explicit]
    roots = dirs if dirs is not None else [str(Path.home()), str(Path.cwd())]
    return [str(Path(d) / CONFIG_FILE) for d in roots if (Path(d) / CONFIG_FILE).is_file()]


def _parser(dirs: Optional[List[str]] = None) -> configparser.ConfigParser:
    cp = configparser.ConfigParser()
    cp.optionxform = str  # type: ignore[assignment]  # keys are case sensitive
    for path in find_configs(dirs):
        cp.read(path)
    return cp


def get_config(prefix: Optional[str], name: str, key: str, dirs: Optional[List[str]] = None) -> Optional[str]:
    """Value of ``key`` in section ``[prefix:name]`` (or ``[name]`` when prefix is empty), else None."""
    section = f"{prefix}:{name}" if prefix else name
    cp = _parser(dirs)
    if cp.has_option(section, key):
        val = cp.get(section, key)
        return None if val == _NONE else val
    return None


def load_sections(prefix: str, dirs: Optional[List[str]] = None) -> Dict[str, Dict[str, str]]:
    """All ``[prefix:*]`` sections as {name: {key: value}} (used for ``[component:dist.ddp]`` defaults)."""
    cp = _parser(dirs)
    out: Dict[str, Dict[str, str]] = {}
    for section in cp.sections():
        head, sep, name = section.partition(":")
        if sep and head == prefix:
            out[name] = {k: v for k, v in cp.items(section) if v != _NONE}
    return out


def apply(scheduler: str, cfg: Dict[str, CfgVal], dirs: Optional[List[str]] = None, opts: Optional[runopts] = None) -> None:
    """Fill ``cfg`` IN PLACE with the ``[scheduler]`` section's values for keys the caller did not set; literals are
    cast with the scheduler's runopts when given."""
    cp = _parser(dirs)
    if not cp.has_section(scheduler):
        return
    for key, raw in cp.items(scheduler):
        if key in cfg or raw == _NONE:
            continue
        opt = opts.get(key) if opts is not None else None
        cfg[key] = opt.cast_to_type(raw) if opt is not None else raw


def dump(f, schedulers: Optional[List[str]] = None, required_only: bool = False) -> None:
    """Write a template config with every scheduler's options (``torchx configure``)."""
    from torchx_b200.schedulers import get_scheduler_factories

    cp = configparser.ConfigParser()
    cp.optionxform = str  # type: ignore[assignment]
    for name, factory in get_scheduler_factories().items():
        if schedulers and name not in schedulers:
            continue
        sched = factory("")
        try:
            section = {k: str(o.default) for k, o in sched.run_opts() if o.is_required or not required_only}
        finally:
            sched.close()
        if section:
            cp[name] = section
    cp.write(f)
