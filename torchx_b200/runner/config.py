"""``.torchxconfig``: INI defaults for scheduler cfg, CLI sub-commands and component arguments.

Same file format, section names and precedence as reference torchx/runner/config.py (dump:220, apply:295,
load_sections:336, get_configs:421, get_config:446, find_configs:473, load:509).  Precedence, high to low:

  1. values given on the command line / in code (``cfg`` entries are never overwritten);
  2. the file named by ``$TORCHXCONFIG`` - if set it is the ONLY file read (empty value: no file at all);
  3. otherwise ``$HOME/.torchxconfig`` (user level) overlaid on ``./.torchxconfig`` (project level): files are read in
     that order and the FIRST file to define a key wins;
  4. the defaults declared by the scheduler's runopts.

    [local_cuda]                     scheduler section: keys are the scheduler's runopts
    log_dir = /tmp/torchx_logs       str / int / float literals
    pin_cpus = True                  bool (configparser spellings: true/false/yes/no/on/off/1/0)
    devices = 0;1;2;3                list of str, ';'-separated
    stage_mb = None                  explicit None

    [cli:run]                        defaults of a CLI sub-command's options (+ ``component`` for run)
    scheduler = local_cuda
    component = dist.ddp

    [component:dist.ddp]             defaults of a component's parameters (strings, decoded like CLI arguments)
    j = 1x8
"""
from __future__ import annotations

import configparser
import logging
import os
from pathlib import Path
from typing import Dict, Iterable, List, Optional, TextIO

from torchx_b200 import settings
from torchx_b200.schedulers import get_scheduler_factories
from torchx_b200.schedulers.api import Scheduler
from torchx_b200.specs.api import CfgVal, get_type_name, runopt

CONFIG_FILE = ".torchxconfig"
CONFIG_PREFIX_DELIM = ":"
ENV_TORCHXCONFIG: str = settings.ENV_TORCHXCONFIG
_NONE = "None"

log = logging.getLogger(__name__)


DEFAULT_CONFIG_DIRS = [str(Path.home()), str(Path.cwd())]  # user level first: its keys win over the project's


def _new_parser() -> configparser.ConfigParser:
    cp = configparser.ConfigParser()
    cp.optionxform = lambda option: option  # type: ignore[assignment]  # option names are case sensitive
    return cp


def _read(path: str) -> configparser.ConfigParser:
    cp = _new_parser()
    with open(path, "r") as f:
        cp.read_file(f)
    return cp


def _scheduler(name: str) -> Scheduler:
    factories = get_scheduler_factories()
    if name not in factories:
        raise ValueError(f"`{name}` is not a registered scheduler. Valid scheduler names: {factories.keys()}")
    return factories[name](session_name="_")


def find_configs(dirs: Optional[Iterable[str]] = None) -> List[str]:
    """Readable config files in reading order.  ``$TORCHXCONFIG`` short-circuits the directory search."""
    explicit = os.getenv(ENV_TORCHXCONFIG)
    if explicit is not None:
        if not explicit:
            return []
        if not Path(explicit).is_file():
            raise FileNotFoundError(f"`{ENV_TORCHXCONFIG}={explicit}` does not exist or is not a file.")
        return [str(Path(explicit))]
    found = []
    for d in (list(dirs) if dirs else DEFAULT_CONFIG_DIRS):
        candidate = Path(d) / CONFIG_FILE
        if os.access(candidate, os.R_OK):
            found.append(str(candidate))
    return found


def load(scheduler: str, f: TextIO, cfg: Dict[str, CfgVal]) -> None:
    """Merge the ``[scheduler]`` section of the INI stream ``f`` into ``cfg`` without overriding keys that are already
    there; literals are typed by the scheduler's runopts, unknown options are reported and skipped."""
    cp = _new_parser()
    cp.read_file(f)
    if not cp.has_section(scheduler):
        return
    sched = _scheduler(scheduler)
    try:
        opts = sched.run_opts()
    finally:
        sched.close()
    for name, raw in cp.items(scheduler):
        if name in cfg:
            continue
        if raw == _NONE:
            cfg[name] = None
            continue
        opt = opts.get(name)
        if opt is None:
            log.warning(f"`{name} = {raw}` was declared in the [{scheduler}] section of the config file but is not a runopt of"
                        f" `{scheduler}` scheduler. Remove the entry from the config file to no longer see this warning")
        elif opt.opt_type is bool:
            cfg[name] = cp.getboolean(scheduler, name)
        elif opt.is_type_list_of_str:
            cfg[name] = raw.split(";")
        elif opt.is_type_dict_of_str:
            cfg[name] = dict(pair.split(":", 1) for pair in raw.replace(",", ";").split(";"))
        else:
            cfg[name] = opt.opt_type(raw)


def apply(scheduler: str, cfg: Dict[str, CfgVal], dirs: Optional[List[str]] = None) -> None:
    """Fill ``cfg`` IN PLACE from the config files (see the module docstring for precedence)."""
    for path in find_configs(dirs):
        with open(path, "r") as f:
            load(scheduler, f, cfg)
        log.info(f"loaded configs from {path}")


def load_sections(prefix: str, dirs: Optional[List[str]] = None) -> Dict[str, Dict[str, str]]:
    """All ``[prefix:name]`` sections as ``{name: {key: raw string}}``; across files the first definition of a key wins."""
    out: Dict[str, Dict[str, str]] = {}
    for path in find_configs(dirs):
        cp = _read(path)
        for section in cp.sections():
            head, sep, name = section.partition(CONFIG_PREFIX_DELIM)
            if not (sep and head == prefix and name):
                continue
            merged = out.setdefault(name, {})
            for key, value in cp.items(section):
                merged.setdefault(key, value)
    return out


def get_configs(prefix: str, name: str, dirs: Optional[List[str]] = None) -> Dict[str, str]:
    return load_sections(prefix, dirs).get(name, {})


def get_config(prefix: str, name: str, key: str, dirs: Optional[List[str]] = None) -> Optional[str]:
    return get_configs(prefix, name, dirs).get(key)


def _fixme_placeholder(opt: runopt, max_len: int = 60) -> str:
    text = f"#FIXME:({get_type_name(opt.opt_type)}) {opt.help}"
    return text if len(text) <= max_len else f"{text[:max_len]}..."


def dump(f: TextIO, schedulers: Optional[List[str]] = None, required_only: bool = False) -> None:
    """Write a template: one section per scheduler, optional runopts pre-filled with their defaults, required ones with a
    ``#FIXME`` placeholder (``required_only`` drops the optional ones).  Unknown scheduler names raise ``ValueError``."""
    cp = _new_parser()
    for name in (schedulers or list(get_scheduler_factories())):
        try:
            sched = _scheduler(name)
        except ModuleNotFoundError:  # a scheduler whose optional dependency is not installed
            continue
        try:
            cp.add_section(name)
            for key, opt in sched.run_opts():
                if opt.is_required:
                    val = _fixme_placeholder(opt)
                elif required_only:
                    continue
                elif opt.is_type_list_of_str:
                    val = ";".join(opt.default) if opt.default else _NONE  # type: ignore[arg-type]
                elif opt.is_type_dict_of_str:
                    val = ";".join(f"{k}:{v}" for k, v in opt.default.items()) if opt.default else _NONE  # type: ignore[union-attr]
                else:
                    val = f"{opt.default}"
                cp.set(name, key, val)
        finally:
            sched.close()
    cp.write(f, space_around_delimiters=True)
