"""``local_cwd``: replicas are plain subprocesses on this host, the current directory is the "image".

This is the reference's launcher behaviour (torchx/schedulers/local_scheduler.py: LocalScheduler:552,
_to_popen_request:947, schedule:795, _popen:681, auto_set_CUDA_VISIBLE_DEVICES:855, describe:1024, log_iter:1061,
close:1110, LogIterator:1130, create_scheduler:1199) re-implemented so that config #1 of BASELINE.json
(``torchx run -s local_cwd dist.ddp --nproc 2`` on CPU/gloo) runs here unchanged, and it is the base class the
B200-native ``local_cuda`` scheduler extends.  Contract points kept exactly:

  * request object ``PopenRequest{app_id, log_dir, role_params{role: [ReplicaParam]}, role_log_dirs}``; building it
    (``_submit_dryrun``) creates NO directories or processes;
  * macros ``${img_root} ${app_id} ${replica_id} ${rank0_env}`` substituted per replica; env additions
    ``TORCHX_RANK0_HOST=localhost``, ``TORCHELASTIC_ERROR_FILE=<replica log dir>/error.json``,
    ``PET_LOG_DIR=<app log dir>/torchelastic/<role>``; child env = parent env overlaid with role env, PATH joined;
  * log tree ``<log_dir>/<session>/<app_id>/<role>/<replica>/{stdout,stderr,combined}.log`` and a ``SUCCESS``
    manifest when the app is closed;
  * each replica is its own session/process group; teardown is SIGTERM, 10 s grace, SIGKILL;
  * state lives in this process only (``list`` raises), terminal apps are evicted LRU beyond ``cache_size``.
"""
from __future__ import annotations

import abc
import io
import json
import logging
import os
import pprint
import shutil
import signal
import subprocess
import tempfile
import threading
import time
import warnings
from dataclasses import asdict, dataclass
from datetime import datetime
from types import FrameType
from typing import Any, BinaryIO, Callable, Dict, Iterable, List, Mapping, Optional, TextIO, Tuple

from torchx_b200.schedulers.api import (
    DescribeAppResponse,
    ListAppResponse,
    Scheduler,
    Stream,
    StructuredOpts,
    filter_regex,
    split_lines_iterator,
)
from torchx_b200.schedulers.ids import make_unique
from torchx_b200.schedulers.streams import Tee
from torchx_b200.specs.api import NONE, AppDef, AppDryRunInfo, AppState, CfgVal, Role, is_terminal, macros, runopts

log = logging.getLogger(__name__)

STDOUT_LOG = "stdout.log"
STDERR_LOG = "stderr.log"
COMBINED_LOG = "combined.log"
NA = "<N/A>"
ENV_CUDA_VISIBLE_DEVICES = "CUDA_VISIBLE_DEVICES"
KILL_GRACE_S = 10.0


# ---------------------------------------------------------------------------------------------------------------
# signals: make SIGTERM/SIGINT unwind the launcher so close() kills the children (no orphans)
# ---------------------------------------------------------------------------------------------------------------
class SignalException(Exception):
    def __init__(self, msg: str, sigval: signal.Signals) -> None:
        super().__init__(msg)
        self.sigval = sigval


def _terminate_process_handler(signum: int, frame: Optional[FrameType]) -> None:
    sigval = signal.Signals(signum)
    raise SignalException(f"Process {os.getpid()} got signal: {sigval}", sigval=sigval)


def _register_termination_signals() -> None:
    if threading.current_thread() is threading.main_thread():
        signal.signal(signal.SIGTERM, _terminate_process_handler)
        signal.signal(signal.SIGINT, _terminate_process_handler)


# ---------------------------------------------------------------------------------------------------------------
# request types
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class ReplicaParam:
    """Everything needed to ``Popen`` one replica."""

    args: List[str]
    env: Dict[str, str]
    stdout: Optional[str] = None
    stderr: Optional[str] = None
    combined: Optional[str] = None
    cwd: Optional[str] = None


@dataclass
class PopenRequest:
    app_id: str
    log_dir: str
    role_params: Dict[str, List[ReplicaParam]]
    role_log_dirs: Dict[str, List[str]]


@dataclass
class Opts(StructuredOpts):
    log_dir: Optional[str] = None
    """Directory to write stdout/stderr log files of replicas."""

    prepend_cwd: bool = False
    """If set, prepends CWD to replica's PATH env var making binaries in CWD take precedence."""

    auto_set_cuda_visible_devices: bool = False
    """Sets CUDA_VISIBLE_DEVICES for roles that request GPU resources."""


LocalOpts = Mapping[str, CfgVal]


# ---------------------------------------------------------------------------------------------------------------
# image providers: what "image" means on a local host
# ---------------------------------------------------------------------------------------------------------------
class ImageProvider(abc.ABC):
    @abc.abstractmethod
    def fetch(self, image: str) -> str:
        """Make the image available locally and return its root directory."""

    def fetch_role(self, role: Role) -> str:
        return self.fetch(role.image)

    def get_cwd(self, image: str) -> Optional[str]:
        return None

    def get_entrypoint(self, img_root: str, role: Role) -> str:
        return os.path.join(img_root, role.entrypoint)

    def get_replica_param(self, img_root: str, role: Role, stdout: Optional[str] = None, stderr: Optional[str] = None,
                          combined: Optional[str] = None) -> ReplicaParam:
        return ReplicaParam([self.get_entrypoint(img_root, role), *role.args], dict(role.env), stdout, stderr, combined,
                            self.get_cwd(role.image))


class LocalDirectoryImageProvider(ImageProvider):
    """The image name is an existing directory; it becomes the child's cwd and relative entrypoints resolve in it."""

    def __init__(self, cfg: LocalOpts) -> None:
        pass

    def fetch(self, image: str) -> str:
        if not os.path.isdir(image):
            raise ValueError(f"Invalid image name: {image}, does not exist or is not a directory")
        return image

    def get_cwd(self, image: str) -> Optional[str]:
        return image

    def get_entrypoint(self, img_root: str, role: Role) -> str:
        return role.entrypoint


class CWDImageProvider(ImageProvider):
    """Ignores the image name: the launcher's working directory is the image (fast local iteration)."""

    def __init__(self, cfg: LocalOpts) -> None:
        pass

    def fetch(self, image: str) -> str:
        return os.getcwd()

    def get_cwd(self, image: str) -> Optional[str]:
        return os.getcwd()

    def get_entrypoint(self, img_root: str, role: Role) -> str:
        return role.entrypoint


# ---------------------------------------------------------------------------------------------------------------
# running apps
# ---------------------------------------------------------------------------------------------------------------
def _join_PATH(*paths: Optional[str]) -> str:
    return os.pathsep.join(p.strip(os.pathsep) for p in paths if p)


@dataclass
class _LocalReplica:
    """One spawned process (a replica for ``local_cwd``; a worker rank for ``local_cuda``) and its log handles."""

    role_name: str
    replica_id: int
    proc: "subprocess.Popen[bytes]"
    stdout: Optional[BinaryIO]
    stderr: Optional[BinaryIO]
    combined: Optional[Tee]
    error_file: str

    def terminate(self) -> None:
        """SIGTERM the replica's process group and close its log handles; repeatable."""
        try:
            os.killpg(self.proc.pid, signal.SIGTERM)
        except (ProcessLookupError, PermissionError):
            pass
        for handle in (self.stdout, self.stderr, self.combined):
            if handle is not None:
                try:
                    handle.close()
                except Exception:  # pragma: no cover - best effort
                    pass

    def is_alive(self) -> bool:
        return self.proc.poll() is None

    def failed(self) -> bool:
        return (not self.is_alive()) and self.proc.returncode != 0


class _LocalAppDef:
    """The processes of one app plus its state; every mutation happens under ``lock`` (describe() is called from
    the CLI's log threads as well as the wait loop)."""

    def __init__(self, id: str, log_dir: str) -> None:
        self.id = id
        self.log_dir = log_dir
        self.role_replicas: Dict[str, List[_LocalReplica]] = {}
        self.state: AppState = AppState.PENDING
        self.last_updated: float = -1
        self.num_restarts = 0
        self.lock = threading.RLock()
        self._closed = False
        self.extra_closers: List[Callable[[], None]] = []

    def add_replica(self, role_name: str, replica: _LocalReplica) -> None:
        self.role_replicas.setdefault(role_name, []).append(replica)

    def replicas(self) -> List[_LocalReplica]:
        return [r for group in self.role_replicas.values() for r in group]

    def set_state(self, state: AppState) -> None:
        self.last_updated = time.time()
        self.state = state

    def kill(self) -> None:
        """SIGTERM everything, wait up to KILL_GRACE_S in total, SIGKILL stragglers, reap.  Repeatable."""
        reps = self.replicas()
        for r in reps:
            r.terminate()
        deadline = time.monotonic() + KILL_GRACE_S
        for r in reps:
            remaining = deadline - time.monotonic()
            if remaining <= 0:
                break
            try:
                r.proc.wait(remaining)
            except subprocess.TimeoutExpired:
                pass
        for r in reps:
            if r.proc.poll() is None:
                try:
                    os.killpg(r.proc.pid, signal.SIGKILL)
                except (ProcessLookupError, PermissionError):
                    r.proc.kill()
        for r in reps:
            r.proc.wait()
            r.terminate()

    def first_error_file(self) -> Optional[str]:
        """The OLDEST error file is the root cause; later ones are usually collateral (peers torn down)."""
        best, best_mtime = None, float("inf")
        for r in self.replicas():
            for path in _error_file_candidates(r.error_file):
                try:
                    mtime = os.path.getmtime(path)
                except OSError:
                    continue
                if mtime < best_mtime:
                    best, best_mtime = path, mtime
        return best

    def get_structured_error_msg(self) -> str:
        path = self.first_error_file()
        if not path:
            return NONE
        try:
            with open(path, "r") as f:
                return json.dumps(json.load(f))
        except (OSError, ValueError):
            return NONE

    def close(self) -> None:
        """kill() + write the ``SUCCESS`` manifest (meaning: logs are flushed and complete, not that the job
        succeeded)."""
        with self.lock:
            self.kill()
            for fn in self.extra_closers:
                try:
                    fn()
                except Exception:  # pragma: no cover
                    log.exception("closer failed")
            self.extra_closers = []
            if self._closed:
                return
            self._closed = True
            manifest = {
                "app_id": self.id,
                "log_dir": self.log_dir,
                "final_state": self.state.name,
                "last_updated": self.last_updated,
                "num_restarts": self.num_restarts,
                "roles": {
                    role: [
                        {
                            "replica_id": r.replica_id,
                            "pid": r.proc.pid,
                            "exitcode": r.proc.returncode,
                            "stdout": getattr(r.stdout, "name", "<CONSOLE>") if r.stdout else "<CONSOLE>",
                            "stderr": getattr(r.stderr, "name", "<CONSOLE>") if r.stderr else "<CONSOLE>",
                            "error_file": r.error_file,
                        }
                        for r in reps
                    ]
                    for role, reps in self.role_replicas.items()
                },
            }
            try:
                with open(os.path.join(self.log_dir, "SUCCESS"), "w") as fp:
                    json.dump(manifest, fp, indent=2)
            except OSError:  # log dir already removed (temp dir on scheduler close)
                pass

    def __repr__(self) -> str:
        pids = {role: [r.proc.pid for r in reps] for role, reps in self.role_replicas.items()}
        return f"{{app_id:{self.id}, state:{self.state}, pid_map:{pids}}}"


def _error_file_candidates(error_file: str) -> List[str]:
    return [error_file] if error_file and error_file != NA else []


# ---------------------------------------------------------------------------------------------------------------
# the scheduler
# ---------------------------------------------------------------------------------------------------------------
class LocalScheduler(Scheduler[LocalOpts]):
    """Processes on localhost.  Ignores resource limits, retry policy and retry counts (no retries: the reference
    says so at local_scheduler.py:559-563); ``local_cuda`` adds GPU pinning, retries and the CUDA-IPC rendezvous."""

    def __init__(self, session_name: str, image_provider_class: Callable[[LocalOpts], ImageProvider], cache_size: int = 100,
                 extra_paths: Optional[List[str]] = None, backend: str = "local") -> None:
        super().__init__(backend, session_name)
        if cache_size <= 0:
            raise ValueError("cache size must be greater than zero")
        self._apps: Dict[str, _LocalAppDef] = {}
        self._apps_lock = threading.RLock()
        self._image_provider_class = image_provider_class
        self._cache_size = cache_size
        self._extra_paths: List[str] = list(extra_paths or [])
        self._base_log_dir: Optional[str] = None
        self._created_tmp_log_dir = False
        _register_termination_signals()

    # -- options ----------------------------------------------------------------------------------------------------
    def _run_opts(self) -> runopts:
        return Opts.as_runopts()

    def _opts(self, cfg: LocalOpts) -> Opts:
        return cfg if isinstance(cfg, Opts) else Opts.from_cfg(cfg)

    def _validate(self, app: AppDef, scheduler: str, cfg: LocalOpts) -> None:
        pass  # resources are advisory on a local host

    # -- request construction (pure) ----------------------------------------------------------------------------------
    def _get_app_log_dir(self, app_id: str, cfg: Opts) -> str:
        self._base_log_dir = cfg.log_dir
        if not self._base_log_dir:
            self._base_log_dir = tempfile.mkdtemp(prefix="torchx_")
            self._created_tmp_log_dir = True
            log.info("Log directory not set in scheduler cfg. Creating a temporary log dir that will be deleted on exit."
                     " To preserve log directory set the `log_dir` cfg option")
        log.info(f"Log directory is: {self._base_log_dir}")
        return os.path.join(str(self._base_log_dir), self.session_name, app_id)

    def _role_path_env(self, role: Role, cwd: Optional[str], opts: Opts) -> str:
        path = _join_PATH(*self._extra_paths, role.env.get("PATH"))
        if cwd:
            path = _join_PATH(cwd, path) if opts.prepend_cwd else _join_PATH(path, cwd)
        return path

    def _replica_role(self, role: Role, img_root: str, app_id: str, replica_id: int, app_log_dir: str) -> Tuple[Role, str]:
        """Macro-substituted copy of ``role`` for one replica, with the local-scheduler env additions."""
        values = macros.Values(img_root=img_root, app_id=app_id, replica_id=str(replica_id), rank0_env="TORCHX_RANK0_HOST")
        rr = values.apply(role)
        replica_log_dir = os.path.join(app_log_dir, role.name, str(replica_id))
        rr.env["TORCHX_RANK0_HOST"] = "localhost"
        rr.env.setdefault("TORCHELASTIC_ERROR_FILE", os.path.join(replica_log_dir, "error.json"))
        rr.env.setdefault("PET_LOG_DIR", os.path.join(app_log_dir, "torchelastic", role.name))
        return rr, replica_log_dir

    def _to_popen_request(self, app: AppDef, cfg: LocalOpts) -> PopenRequest:
        opts = self._opts(cfg)
        app_id = make_unique(app.name)
        provider = self._image_provider_class(cfg)
        app_log_dir = self._get_app_log_dir(app_id, opts)
        role_params: Dict[str, List[ReplicaParam]] = {}
        role_log_dirs: Dict[str, List[str]] = {}
        for role in app.roles:
            params = role_params.setdefault(role.name, [])
            log_dirs = role_log_dirs.setdefault(role.name, [])
            img_root = provider.fetch_role(role)
            role.env["PATH"] = self._role_path_env(role, provider.get_cwd(role.image), opts)
            for replica_id in range(role.num_replicas):
                rr, replica_log_dir = self._replica_role(role, img_root, app_id, replica_id, app_log_dir)
                params.append(provider.get_replica_param(
                    img_root, rr, os.path.join(replica_log_dir, STDOUT_LOG), os.path.join(replica_log_dir, STDERR_LOG),
                    os.path.join(replica_log_dir, COMBINED_LOG)))
                log_dirs.append(replica_log_dir)
        self.auto_set_CUDA_VISIBLE_DEVICES(role_params, app, opts)
        return PopenRequest(app_id, app_log_dir, role_params, role_log_dirs)

    def _submit_dryrun(self, app: AppDef, cfg: LocalOpts) -> AppDryRunInfo[PopenRequest]:
        return AppDryRunInfo(self._to_popen_request(app, cfg), lambda req: pprint.pformat(asdict(req), indent=2, width=80))

    # -- GPU partitioning ---------------------------------------------------------------------------------------------
    def _cuda_device_count(self) -> int:
        """GPUs on this host per ``nvidia-smi -L`` (0 when there is no driver).  Tests mock this."""
        try:
            res = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, check=True)
        except (OSError, subprocess.CalledProcessError) as e:
            log.debug("nvidia-smi -L failed: %s", e)
            return 0
        return sum(1 for line in res.stdout.splitlines() if line.strip())

    def auto_set_CUDA_VISIBLE_DEVICES(self, role_params: Dict[str, List[ReplicaParam]], app: AppDef, cfg: Opts) -> None:
        """With ``auto_set_cuda_visible_devices=True`` hand each GPU-requesting replica a contiguous, disjoint index
        range (role order, then replica order), overwriting any value in the role env.  All or nothing: if the host
        has fewer devices than the app requests in total, nothing is set."""
        wanted = sum(role.num_replicas * role.resource.gpu for role in app.roles)
        if wanted <= 0:
            return
        if not cfg.auto_set_cuda_visible_devices:
            log.warning("Running role replicas that require GPUs without setting `CUDA_VISIBLE_DEVICES` may put several processes"
                        " on the same GPU (CUDA OutOfMemory, ...). Set the `auto_set_cuda_visible_devices = True` scheduler runopt"
                        " to divide this host's GPUs among the replicas.")
            return
        have = self._cuda_device_count()
        if wanted > have:
            log.warning(f"Cannot auto-set `CUDA_VISIBLE_DEVICES`: available GPUs: {have} is less than the number of requested"
                        f" GPUs: {wanted}. Reduce requested GPU resources or use a host with more GPUs")
            return
        cursor = 0
        for role in app.roles:
            if role.resource.gpu <= 0:
                continue
            for replica in role_params[role.name]:
                replica.env[ENV_CUDA_VISIBLE_DEVICES] = ",".join(str(i) for i in range(cursor, cursor + role.resource.gpu))
                cursor += role.resource.gpu

    # -- launching ----------------------------------------------------------------------------------------------------
    def _get_file_io(self, file: Optional[str]) -> Optional[io.FileIO]:
        if not file:
            return None
        if os.path.isfile(file):
            raise FileExistsError(f"log file: {file} already exists, specify a different log_dir, app_name, or remove the file and retry")
        os.makedirs(os.path.dirname(file), exist_ok=True)
        return io.open(file, mode="wb", buffering=0)

    def _get_replica_output_handles(self, p: ReplicaParam) -> Tuple[Optional[io.FileIO], Optional[io.FileIO], Optional[Tee]]:
        out, err = self._get_file_io(p.stdout), self._get_file_io(p.stderr)
        comb_file = self._get_file_io(p.combined)
        comb = Tee(comb_file, p.stdout, p.stderr) if comb_file and p.stdout and p.stderr else None
        return out, err, comb

    def _get_replica_env(self, p: ReplicaParam) -> Dict[str, str]:
        env = os.environ.copy()
        env.update(p.env)
        env["PATH"] = _join_PATH(p.env.get("PATH"), os.getenv("PATH"))
        env.setdefault("PYTHONUNBUFFERED", "x")
        return env

    def run_local_job(self, args: List[str], env: Dict[str, str], stdout: Optional[io.FileIO], stderr: Optional[io.FileIO],
                      cwd: Optional[str] = None, preexec_fn: Optional[Callable[[], None]] = None) -> "subprocess.Popen[bytes]":
        return subprocess.Popen(args=args, env=env, stdout=stdout, stderr=stderr, start_new_session=True, cwd=cwd,
                                preexec_fn=preexec_fn)

    def _popen(self, role_name: str, replica_id: int, replica_params: ReplicaParam,
               preexec_fn: Optional[Callable[[], None]] = None) -> _LocalReplica:
        p = replica_params
        out, err, comb = self._get_replica_output_handles(p)
        env = self._get_replica_env(p)
        log.debug("Running %s (replica %s):\n %s", role_name, replica_id, pprint.pformat(asdict(p), indent=2, width=80))
        proc = self.run_local_job(args=p.args, env=env, stdout=out, stderr=err, cwd=p.cwd, preexec_fn=preexec_fn)
        return _LocalReplica(role_name, replica_id, proc, out, err, comb, error_file=env.get("TORCHELASTIC_ERROR_FILE", NA))

    def _evict_lru(self) -> bool:
        victim, oldest = None, float("inf")
        for app_id, app in self._apps.items():
            if is_terminal(app.state) and app.last_updated <= oldest:
                victim, oldest = app_id, app.last_updated
        if victim is None:
            return False
        del self._apps[victim]
        return True

    def _reserve_slot(self, app_id: str) -> None:
        if len(self._apps) >= self._cache_size and not self._evict_lru():
            raise IndexError(f"App cache size ({self._cache_size}) exceeded. Increase the cache size")
        assert app_id not in self._apps, "app ids carry 64 random bits; a collision means a bug"

    def schedule(self, dryrun_info: AppDryRunInfo[PopenRequest]) -> str:
        req: PopenRequest = dryrun_info.request
        with self._apps_lock:
            self._reserve_slot(req.app_id)
            os.makedirs(req.log_dir)
            app = _LocalAppDef(req.app_id, req.log_dir)
            for role_name, params in req.role_params.items():
                for replica_id, p in enumerate(params):
                    os.makedirs(req.role_log_dirs[role_name][replica_id])
                    app.add_replica(role_name, self._popen(role_name, replica_id, p))
            self._apps[req.app_id] = app
        return req.app_id

    # -- status -------------------------------------------------------------------------------------------------------
    def _refresh_state(self, app: _LocalAppDef) -> AppState:
        """RUNNING while anything is alive, else FAILED if anything exited non-zero, else SUCCEEDED."""
        reps = app.replicas()
        if any(r.is_alive() for r in reps):
            return AppState.RUNNING
        return AppState.FAILED if any(r.failed() for r in reps) else AppState.SUCCEEDED

    def describe(self, app_id: str) -> Optional[DescribeAppResponse]:
        app = self._apps.get(app_id)
        if app is None:
            return None
        with app.lock:
            err = app.get_structured_error_msg()
            if not is_terminal(app.state):
                app.set_state(self._refresh_state(app))
            if is_terminal(app.state):
                app.close()
            return DescribeAppResponse(app_id=app_id, state=app.state, num_restarts=app.num_restarts, structured_error_msg=err,
                                       ui_url=f"file://{app.log_dir}")

    def list(self, cfg: Optional[Mapping[str, CfgVal]] = None) -> List[ListAppResponse]:
        raise Exception("App handles cannot be listed for local scheduler as they are not persisted by torchx")

    def _cancel_existing(self, app_id: str) -> None:
        app = self._apps[app_id]
        with app.lock:
            app.close()
            app.state = AppState.CANCELLED

    # -- logs ---------------------------------------------------------------------------------------------------------
    def log_iter(self, app_id: str, role_name: str, k: int = 0, regex: Optional[str] = None, since: Optional[datetime] = None,
                 until: Optional[datetime] = None, should_tail: bool = False, streams: Optional[Stream] = None) -> Iterable[str]:
        if since or until:
            warnings.warn("Since and/or until times specified for LocalScheduler.log_iter. These will be ignored and all log lines will be returned")
        app = self._apps[app_id]
        name = {None: COMBINED_LOG, Stream.COMBINED: COMBINED_LOG, Stream.STDOUT: STDOUT_LOG, Stream.STDERR: STDERR_LOG}[streams]
        log_file = os.path.join(app.log_dir, role_name, str(k), name)
        if not os.path.isfile(log_file):
            raise RuntimeError(f"app: {app_id} was not configured to log into a file. Did you run it with log_dir set in Dict[str, CfgVal]?")
        lines: Iterable[str] = split_lines_iterator(LogIterator(app_id, log_file, self))
        return filter_regex(regex, lines) if regex else lines

    # -- shutdown -----------------------------------------------------------------------------------------------------
    def close(self) -> None:
        for app in list(self._apps.values()):
            app.kill()
            for fn in app.extra_closers:
                try:
                    fn()
                except Exception:  # pragma: no cover
                    pass
            app.extra_closers = []
        if self._base_log_dir and self._created_tmp_log_dir:
            shutil.rmtree(self._base_log_dir, ignore_errors=True)

    def __del__(self) -> None:
        try:
            self.close()
        except Exception as e:  # pragma: no cover
            log.warning(f"Exception {e} occurred while trying to clean `LocalScheduler` via `__del__` method")


class LogIterator:
    """Yields chunks of a growing log file until the app is terminal (``tail -f`` that knows when to stop)."""

    POLL_S = 0.1
    CHUNK = 64000

    def __init__(self, app_id: str, log_file: str, scheduler: Scheduler, should_tail: bool = True) -> None:
        self._app_id = app_id
        self._log_file = log_file
        self._scheduler = scheduler
        self._fp: Optional[TextIO] = None
        self._finished = not should_tail

    def _check_finished(self) -> None:
        desc = self._scheduler.describe(self._app_id)
        self._finished = desc is None or is_terminal(desc.state)

    def __iter__(self) -> "LogIterator":
        while True:
            self._check_finished()
            if os.path.isfile(self._log_file):
                time.sleep(self.POLL_S)  # let the writer get its first bytes out
                self._fp = open(self._log_file, mode="rt", newline="\n", errors="replace")
                return self
            if self._finished:
                raise RuntimeError(f"app: {self._app_id} finished without writing: {self._log_file}")
            time.sleep(self.POLL_S)

    def __next__(self) -> str:
        assert self._fp is not None
        while True:
            chunk = self._fp.read(self.CHUNK)
            if chunk:
                return chunk
            if self._finished:
                self._fp.close()
                raise StopIteration
            time.sleep(self.POLL_S)
            self._check_finished()


def create_scheduler(session_name: str, cache_size: int = 100, extra_paths: Optional[List[str]] = None,
                     image_provider_class: Callable[[LocalOpts], ImageProvider] = CWDImageProvider, **kwargs: Any) -> LocalScheduler:
    """Factory registered as ``local_cwd``.  ``**kwargs`` swallows the ``TORCHX_*``-derived parameters Runner passes to
    every factory (reference runner/api.py:128-134,626)."""
    return LocalScheduler(session_name=session_name, image_provider_class=image_provider_class, cache_size=cache_size,
                          extra_paths=extra_paths)
