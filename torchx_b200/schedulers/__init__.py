"""Scheduler name -> lazily imported factory (reference torchx/schedulers/__init__.py:16-68).

``local_cwd`` is the reference's behaviour (one ``bash -c torchrun ...`` per replica); ``local_cuda`` is the
B200-native replacement (one pinned worker per GPU, CUDA-IPC rendezvous).  Both names start with ``local`` because
the CLI auto-waits and tails logs for ``local*`` schedulers (reference cli/cmd_run.py:321).  As in the reference, if
any scheduler plugins are registered (``torchx_b200.plugins``) they REPLACE this default map.
"""
import importlib
from typing import Callable, Dict, Mapping

from torchx_b200.schedulers.api import Scheduler

DEFAULT_SCHEDULER_MODULES: Mapping[str, str] = {
    "local_cuda": "torchx_b200.schedulers.local_cuda_scheduler",
    "local_cwd": "torchx_b200.schedulers.local_scheduler",
}

SchedulerFactory = Callable[..., Scheduler]


def _deferred(module_path: str) -> SchedulerFactory:
    def factory(*args: object, **kwargs: object) -> Scheduler:
        return importlib.import_module(module_path).create_scheduler(*args, **kwargs)

    factory.__qualname__ = f"create_scheduler<{module_path}>"
    return factory


def get_scheduler_factories(*, skip_defaults: bool = False) -> Dict[str, SchedulerFactory]:
    """The first entry is the default scheduler."""
    from torchx_b200 import plugins

    registered = plugins.registered_schedulers()
    if registered:
        return dict(registered)
    if skip_defaults:
        return {}
    return {name: _deferred(path) for name, path in DEFAULT_SCHEDULER_MODULES.items()}


def get_default_scheduler_name() -> str:
    return next(iter(get_scheduler_factories()))
