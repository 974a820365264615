"""Runner: the programmatic front door (``get_runner().run_component("dist.ddp", [...], "local_cuda")``).

Behaviour follows reference torchx/runner/api.py (Runner:90, run_component:159, dryrun:333-437, schedule:299, status:468,
wait:493, cancel, describe, log_lines:557, list, _scheduler:621, get_runner:649): validates the AppDef, injects
``TORCHX_JOB_ID`` / ``TORCHX_INTERNAL_SESSION_ID`` into every role, resolves the scheduler cfg, dry-runs, schedules, polls.
Scheduler factories receive ``TORCHX_*`` environment variables (lower-cased, prefix stripped) as keyword arguments.
Tracker backends of the reference are out of scope; per-call telemetry events go to ``runner/events`` (a NullHandler by default).
"""
from __future__ import annotations

import json
import logging
import os
import time
from datetime import datetime
from typing import Any, Dict, Iterable, List, Mapping, Optional, Tuple, Union

from torchx_b200.runner import events
from torchx_b200.schedulers import SchedulerFactory, get_scheduler_factories
from torchx_b200.schedulers.api import ListAppResponse, Scheduler, Stream
from torchx_b200.runner import config
from torchx_b200.settings import (
    ENV_TORCHX_JOB_ID,
    ENV_TORCHX_PARENT_RUN_ID,
    ENV_TORCHX_TRACKERS,
    TORCHX_INTERNAL_SESSION_ID,
    tracker_config_env_var_name,
)
from torchx_b200.specs.api import (
    AppDef,
    AppDryRunInfo,
    AppHandle,
    AppStatus,
    CfgVal,
    UnknownAppException,
    macros,
    make_app_handle,
    parse_app_handle,
    runopts,
)
from torchx_b200.specs.builders import materialize_appdef
from torchx_b200.specs.finder import get_component
from torchx_b200.specs.api import Workspace
from torchx_b200.util.session import get_session_id_or_create_new
from torchx_b200.workspace.api import WorkspaceMixin

logger = logging.getLogger(__name__)


def get_configured_trackers() -> Dict[str, Optional[str]]:
    """``{tracker name: its config string}`` for the job being submitted: names from ``[torchx:tracker]`` in .torchxconfig
    (or ``$TORCHX_TRACKERS``, which replaces the list), each one's config from ``[tracker:<name>] config = ...`` (or
    ``$TORCHX_TRACKER_<NAME>_CONFIG``).  The Runner only FORWARDS this to the workers' environment; tracker backends
    themselves run in the worker process and are not part of this package (reference torchx/runner/api.py:68-87)."""
    names = list(config.get_configs(prefix="torchx", name="tracker").keys())
    if ENV_TORCHX_TRACKERS in os.environ:
        names = [n for n in os.environ[ENV_TORCHX_TRACKERS].split(",") if n]
        logger.info(f"Using {ENV_TORCHX_TRACKERS}={names} as tracker names")
    configured: Dict[str, Optional[str]] = {}
    for name in names:
        value = config.get_config(prefix="tracker", name=name, key="config")
        env_name = tracker_config_env_var_name(name)
        if env_name in os.environ:
            value = os.environ[env_name]
            logger.info(f"Using {env_name}={value} for `{name}` tracker")
        configured[name] = value
    return configured


def _logged(api: str):
    """Run a Runner method inside its telemetry event (runner/events): scheduler, app id, image, run config and workspace are
    taken from the call's own arguments / result, as the reference fills them in by hand in each method
    (torchx/runner/api.py:176-191, 388-440 ...)."""

    def deco(fn):
        import functools
        import inspect

        sig = inspect.signature(fn)

        @functools.wraps(fn)
        def wrapper(self, *args, **kwargs):
            try:
                bound = sig.bind_partial(self, *args, **kwargs).arguments
            except TypeError:  # let the real call raise the proper error
                bound = {}
            scheduler, app_id, image = bound.get("scheduler"), None, None
            handle = bound.get("app_handle")
            if isinstance(handle, str):
                try:
                    scheduler, _, app_id = parse_app_handle(handle)
                except Exception:  # noqa: BLE001 - a malformed handle is the wrapped method's business
                    pass
            info = bound.get("dryrun_info")
            if info is not None:
                scheduler = getattr(info, "_scheduler", None) or scheduler
            app = bound.get("app") if bound.get("app") is not None else getattr(info, "_app", None)
            if app is not None and getattr(app, "roles", None):
                image = app.roles[0].image
            cfg, workspace = bound.get("cfg"), bound.get("workspace")
            with events.log_event(api, scheduler if isinstance(scheduler, str) else None, app_id, app_image=image,
                                  runcfg=json.dumps(dict(cfg)) if cfg else None, workspace=str(workspace) if workspace else None) as ctx:
                out = fn(self, *args, **kwargs)
                if isinstance(out, str) and "://" in out:  # schedule / run / run_component return the new app's handle
                    try:
                        ctx._torchx_event.app_id = parse_app_handle(out)[2]
                    except Exception:  # noqa: BLE001
                        pass
                return out

        return wrapper

    return deco


class Runner:
    def __init__(self, name: str = "", scheduler_factories: Optional[Dict[str, SchedulerFactory]] = None,
                 component_defaults: Optional[Dict[str, Dict[str, str]]] = None, scheduler_params: Optional[Dict[str, object]] = None) -> None:
        self._name = name
        self._scheduler_factories = scheduler_factories if scheduler_factories is not None else get_scheduler_factories()
        self._scheduler_params: Dict[str, Any] = {**self._get_scheduler_params_from_env(), **(scheduler_params or {})}
        self._scheduler_instances: Dict[str, Scheduler] = {}
        self._apps: Dict[AppHandle, AppDef] = {}
        self._component_defaults = component_defaults or {}

    @staticmethod
    def _get_scheduler_params_from_env() -> Dict[str, str]:
        return {k.lower()[len("torchx_"):]: v for k, v in os.environ.items() if k.lower().startswith("torchx_")}

    # -- lifecycle --------------------------------------------------------------------------------------------------
    def __enter__(self) -> "Runner":
        return self

    def __exit__(self, *exc: object) -> bool:
        self.close()
        return False

    def close(self) -> None:
        for sched in self._scheduler_instances.values():
            sched.close()

    def name(self) -> str:
        return self._name

    def scheduler_backends(self) -> List[str]:
        return list(self._scheduler_factories)

    def _scheduler(self, scheduler: str) -> Scheduler:
        inst = self._scheduler_instances.get(scheduler)
        if inst is None:
            factory = self._scheduler_factories.get(scheduler)
            if factory is None:
                raise KeyError(f"Undefined scheduler backend: {scheduler}. Use one of: {self._scheduler_factories.keys()}")
            inst = self._scheduler_instances[scheduler] = factory(self._name, **self._scheduler_params)
        return inst

    def _scheduler_app_id(self, app_handle: AppHandle) -> Tuple[Scheduler, str, str]:
        backend, _, app_id = parse_app_handle(app_handle)
        return self._scheduler(backend), backend, app_id

    def scheduler_run_opts(self, scheduler: str) -> runopts:
        return self._scheduler(scheduler).run_opts()

    def cfg_from_str(self, scheduler: str, *cfg_literal: str) -> Mapping[str, CfgVal]:
        opts = self._scheduler(scheduler).run_opts()
        cfg: Dict[str, CfgVal] = {}
        for lit in cfg_literal:
            cfg.update(opts.cfg_from_str(lit))
        return opts.resolve(cfg)

    # -- submission -------------------------------------------------------------------------------------------------
    @_logged("dryrun_component")
    def dryrun_component(self, component: str, component_args: Union[List[str], Dict[str, Any]], scheduler: str,
                         cfg: Optional[Mapping[str, CfgVal]] = None, workspace: Optional[object] = None,
                         parent_run_id: Optional[str] = None) -> AppDryRunInfo:
        comp = get_component(component)
        cli_args = component_args if isinstance(component_args, list) else []
        json_args = component_args if isinstance(component_args, dict) else {}
        app = materialize_appdef(comp.fn, cli_args, self._component_defaults.get(component), json_args)
        return self.dryrun(app, scheduler, cfg=cfg, workspace=workspace, parent_run_id=parent_run_id)

    @_logged("run_component")
    def run_component(self, component: str, component_args: Union[List[str], Dict[str, Any]], scheduler: str,
                      cfg: Optional[Mapping[str, CfgVal]] = None, workspace: Optional[object] = None,
                      parent_run_id: Optional[str] = None) -> AppHandle:
        return self.schedule(self.dryrun_component(component, component_args, scheduler, cfg, workspace, parent_run_id))

    @_logged("dryrun")
    def dryrun(self, app: AppDef, scheduler: str, cfg: Optional[Mapping[str, CfgVal]] = None, workspace: Optional[object] = None,
               parent_run_id: Optional[str] = None) -> AppDryRunInfo:
        if not app.roles:
            raise ValueError(f"No roles for app: {app.name}. Did you forget to add roles to AppDef?")
        parent_run_id = os.environ.get(ENV_TORCHX_PARENT_RUN_ID, parent_run_id)
        trackers = get_configured_trackers()
        for role in app.roles:
            if not role.entrypoint:
                raise ValueError(f"No entrypoint for role: {role.name}. Did you forget to call role.runs(entrypoint, args, env)?")
            if role.num_replicas <= 0:
                raise ValueError(f"Non-positive replicas for role: {role.name}. Did you forget to set role.num_replicas?")
            role.env[ENV_TORCHX_JOB_ID] = make_app_handle(scheduler, self._name, macros.app_id)
            role.env[TORCHX_INTERNAL_SESSION_ID] = get_session_id_or_create_new()
            if parent_run_id:
                role.env[ENV_TORCHX_PARENT_RUN_ID] = parent_run_id
            if trackers:
                role.env[ENV_TORCHX_TRACKERS] = ",".join(trackers)
            for tracker_name, tracker_config in trackers.items():
                if tracker_config:
                    role.env[tracker_config_env_var_name(tracker_name)] = tracker_config
        sched = self._scheduler(scheduler)
        resolved = sched.run_opts().resolve(cfg or {})
        sched._pre_build_validate(app, scheduler, resolved)
        if isinstance(sched, WorkspaceMixin):
            # only image-building (plugin) schedulers consume a workspace (reference runner/api.py:405-424); the argument wins
            # over roles[0].workspace for backwards compatibility.  The local schedulers run from the cwd and ignore it.
            if workspace:
                app.roles[0].workspace = Workspace.from_str(workspace) if isinstance(workspace, str) else workspace
            sched.build_workspaces(app.roles, resolved)
        elif workspace:
            logger.debug("workspace `%s` ignored: `%s` runs from the current directory", workspace, scheduler)
        sched._validate(app, scheduler, resolved)
        info = sched.submit_dryrun(app, resolved)
        info._scheduler = scheduler
        return info

    @_logged("schedule")
    def schedule(self, dryrun_info: AppDryRunInfo) -> AppHandle:
        scheduler = dryrun_info._scheduler
        assert scheduler is not None, "dryrun_info was not produced by Runner.dryrun"
        app_id = self._scheduler(scheduler).schedule(dryrun_info)
        handle = make_app_handle(scheduler, self._name, app_id)
        if dryrun_info._app is not None:
            self._apps[handle] = dryrun_info._app
        return handle

    @_logged("run")
    def run(self, app: AppDef, scheduler: str, cfg: Optional[Mapping[str, CfgVal]] = None, workspace: Optional[object] = None,
            parent_run_id: Optional[str] = None, *, dryrun: bool = False) -> Union[AppHandle, AppDryRunInfo]:
        info = self.dryrun(app, scheduler, cfg=cfg, workspace=workspace, parent_run_id=parent_run_id)
        return info if dryrun else self.schedule(info)

    # -- monitoring -------------------------------------------------------------------------------------------------
    @_logged("status")
    def status(self, app_handle: AppHandle) -> Optional[AppStatus]:
        sched, _, app_id = self._scheduler_app_id(app_handle)
        desc = sched.describe(app_id)
        if desc is None:
            self._apps.pop(app_handle, None)
            return None
        return AppStatus(desc.state, desc.num_restarts, msg=desc.msg, structured_error_msg=desc.structured_error_msg,
                         roles=desc.roles_statuses, ui_url=desc.ui_url)

    @_logged("wait")
    def wait(self, app_handle: AppHandle, wait_interval: float = 10) -> Optional[AppStatus]:
        while True:
            st = self.status(app_handle)
            if st is None or st.is_terminal():
                return st
            time.sleep(wait_interval)

    @_logged("cancel")
    def cancel(self, app_handle: AppHandle) -> None:
        sched, _, app_id = self._scheduler_app_id(app_handle)
        st = self.status(app_handle)
        if st is not None and not st.is_terminal():
            sched.cancel(app_id)

    stop = cancel

    @_logged("delete")
    def delete(self, app_handle: AppHandle) -> None:
        sched, _, app_id = self._scheduler_app_id(app_handle)
        if self.status(app_handle) is not None:
            sched.delete(app_id)

    @_logged("describe")
    def describe(self, app_handle: AppHandle) -> Optional[AppDef]:
        sched, _, app_id = self._scheduler_app_id(app_handle)
        app = self._apps.get(app_handle)
        if app is None:
            desc = sched.describe(app_id)
            if desc is not None:
                app = AppDef(name=app_id, roles=desc.roles, metadata=desc.metadata)
        return app

    @_logged("log_lines")
    def log_lines(self, app_handle: AppHandle, role_name: str, k: int = 0, regex: Optional[str] = None, since: Optional[datetime] = None,
                  until: Optional[datetime] = None, should_tail: bool = False, streams: Optional[Stream] = None) -> Iterable[str]:
        """Lines keep their trailing newline; ``k`` is the replica (node) index, not the worker rank."""
        sched, _, app_id = self._scheduler_app_id(app_handle)
        if not self.status(app_handle):
            raise UnknownAppException(app_handle)
        return sched.log_iter(app_id, role_name, k, regex, since, until, should_tail, streams=streams)

    @_logged("list")
    def list(self, scheduler: str, cfg: Optional[Mapping[str, CfgVal]] = None) -> List[ListAppResponse]:
        apps = self._scheduler(scheduler).list(cfg)
        for a in apps:
            a.app_handle = make_app_handle(scheduler, self._name, a.app_id)
        return apps

    def __repr__(self) -> str:
        return f"Runner(name={self._name}, schedulers={self._scheduler_factories}, apps={self._apps})"


def get_runner(name: Optional[str] = None, component_defaults: Optional[Dict[str, Dict[str, str]]] = None, **scheduler_params: Any) -> Runner:
    """A Runner wired with every registered scheduler.  The default session name is ``torchx`` (it shows up in app
    handles ``local_cuda://torchx/<app_id>`` and in the log tree)."""
    return Runner(name or "torchx", get_scheduler_factories(), component_defaults, scheduler_params=scheduler_params)
