"""``local_cuda``: the B200-native single-box scheduler.  One worker process per GPU, spawned directly.

It accepts the UNMODIFIED ``dist.ddp`` AppDef (``bash -c "torchrun ... --nproc_per_node N script.py"``, reference
torchx/components/dist.py:261-308) and replaces what ``local_cwd`` + ``torchrun`` do with it
(reference schedulers/local_scheduler.py:947-1022 + torch/distributed/elastic/agent/server/local_elastic_agent.py:309-323):

  * parses the torchrun command line back into (nnodes, nproc_per_node, script|-m module, args) and spawns the
    ``nnodes x nproc_per_node`` workers itself - no bash, no elastic agent, no c10d rendezvous round trips;
  * gives every worker the full torchrun environment contract (RANK, LOCAL_RANK, WORLD_SIZE, LOCAL_WORLD_SIZE,
    GROUP_RANK, GROUP_WORLD_SIZE, ROLE_RANK, ROLE_WORLD_SIZE, ROLE_NAME, MASTER_ADDR, MASTER_PORT,
    TORCHELASTIC_RESTART_COUNT / MAX_RESTARTS / RUN_ID / USE_AGENT_STORE / ERROR_FILE, OMP_NUM_THREADS) so existing
    scripts (``torchx.distributed.init_pg``, plain ``init_process_group("nccl")``) run unchanged;
  * pins each rank: ``B2_DEVICE`` = its CUDA ordinal, CPU affinity = the cores of that GPU's NUMA node split among the
    node's ranks;
  * hands out the peer-buffer rendezvous instead of a TCP store: ``B2_SHM_NAME=/b2_<app_id>`` + ``B2_EPOCH=<attempt>``
    name the POSIX-shm control block through which ``libb200ddp.so`` exchanges CUDA-IPC handles (include/b200ddp.h);
  * NEW behaviour the reference's local scheduler lacks (its num_restarts is hard-coded 0, local_scheduler.py:1057):
    ``Role.max_retries`` is honoured - when a worker dies the gang is torn down and re-launched with the next epoch
    (fresh shm name, TORCHELASTIC_RESTART_COUNT+1), so survivors never touch a dead peer's memory;
  * keeps the on-disk contract: ``<log_dir>/<session>/<app_id>/<role>/<replica>/{stdout,stderr,combined}.log`` now hold
    the ``[rank]:``-prefixed merge of the workers' streams (what ``--tee 3`` prints), per-worker files live under
    ``attempt_<n>/rank_<local_rank>/``; ``error.json`` / ``SUCCESS`` as before, so ``torchx log|status`` work unchanged.

  * app registry: every scheduled app leaves ``$TORCHX_HOME/apps/<session>/<app_id>.json`` (default ``~/.torchx_b200``) pointing
    at its log tree and, once finished, holding its final state, so ``torchx status | log | list`` also resolve from ANOTHER
    process (the reference's local scheduler keeps state in the submitting process only and ``list()`` raises,
    local_scheduler.py:615,1099-1102).  ``torchx cancel`` from another process drops ``<app_id>.cancel`` next to the entry; the
    launcher's supervisor thread takes the gang down within its 0.1 s poll and removes the file as the acknowledgement.

Roles whose command is not a torchrun line (e.g. ``utils.echo``) are launched exactly as ``local_cwd`` would.
"""
from __future__ import annotations

import io
import json
import logging
import os
import pprint
import shlex
import socket
import subprocess
import sys
import threading
import time
from dataclasses import asdict, dataclass, field
from typing import Any, Callable, Dict, List, Mapping, Optional, Tuple

from torchx_b200.schedulers.api import DescribeAppResponse, ListAppResponse, Stream, filter_regex, split_lines_iterator
from torchx_b200.schedulers.local_scheduler import (
    COMBINED_LOG,
    ENV_CUDA_VISIBLE_DEVICES,
    KILL_GRACE_S,
    STDERR_LOG,
    STDOUT_LOG,
    CWDImageProvider,
    ImageProvider,
    LocalOpts,
    LocalScheduler,
    LogIterator,
    Opts,
    PopenRequest,
    ReplicaParam,
    _LocalAppDef,
)
from torchx_b200.schedulers.ids import make_unique
from torchx_b200.specs.api import NONE, AppDef, AppDryRunInfo, AppState, CfgVal, Role, is_terminal, runopts

log = logging.getLogger(__name__)

MONITOR_POLL_S = 0.1


# ---------------------------------------------------------------------------------------------------------------
# run options
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class CudaOpts(Opts):
    devices: Optional[List[str]] = None
    """CUDA ordinals this job may use, in rank order (default: all GPUs reported by nvidia-smi)."""

    pin_cpus: bool = True
    """Bind each worker to its GPU's NUMA-local cores (split evenly among the ranks sharing the node)."""

    omp_num_threads: int = 1
    """OMP_NUM_THREADS for workers that do not set it (torchrun's default is 1)."""

    stage_mb: int = 0
    """Per-rank staging buffer size in MiB for the peer-buffer allreduce (0 = library default 512)."""

    master_port: int = 0
    """MASTER_PORT handed to the workers (0 = pick a free port at launch)."""

    job_timeout: float = 0.0
    """Seconds after which the whole gang is terminated (SIGTERM, then SIGKILL) and the app FAILED with exit code 34 -
    what wrapping every replica in the reference's ``torchx.apps.utils.process_monitor --timeout`` does, for the gang."""

    exit_on_file: Optional[str] = None
    """Terminate the gang as soon as this path exists (process_monitor's ``--exit_on_file``)."""

    start_on_file: Optional[str] = None
    """Hold the launch until this path exists (process_monitor's ``--start_on_file``); ``job_timeout`` covers the wait."""


# ---------------------------------------------------------------------------------------------------------------
# torchrun command line -> spec
# ---------------------------------------------------------------------------------------------------------------
_VALUE_OPTS = {
    "--rdzv_backend", "--rdzv_conf", "--rdzv_endpoint", "--rdzv_id", "--nnodes", "--nproc_per_node", "--tee", "-t", "--role",
    "--node_rank", "--max_restarts", "--master_addr", "--master_port", "--log_dir", "--redirects", "-r", "--monitor_interval",
    "--start_method", "--local_addr", "--local_ranks_filter", "--logs_specs", "--event_log_handler",
}
_FLAG_OPTS = {"--standalone", "--no_python", "--run_path"}


@dataclass
class TorchrunSpec:
    """What ``dist.ddp`` asked torchrun to do, recovered from the command line it emitted."""

    min_nnodes: int = 1
    max_nnodes: int = 1
    nproc_per_node: str = "1"  # an int literal, or "gpu"/"auto" (resolved against the device count)
    rdzv_id: str = ""
    role: str = ""
    tee: int = 0
    node_rank: Optional[int] = None
    max_restarts: Optional[int] = None
    module: Optional[str] = None
    script: Optional[str] = None
    no_python: bool = False
    script_args: List[str] = field(default_factory=list)

    def worker_cmd(self) -> List[str]:
        if self.module is not None:
            return [sys.executable, "-u", "-m", self.module, *self.script_args]
        assert self.script is not None
        if self.no_python:
            return [self.script, *self.script_args]
        return [sys.executable, "-u", self.script, *self.script_args]


def parse_torchrun(cmdline: str) -> Optional[TorchrunSpec]:
    """``None`` if ``cmdline`` is not a torchrun invocation; raises ``ValueError`` for a torchrun line this
    scheduler cannot honour."""
    try:
        toks = shlex.split(cmdline)
    except ValueError:
        return None
    if toks[:1] == ["torchrun"]:
        toks = toks[1:]
    elif len(toks) >= 3 and os.path.basename(toks[0]).startswith("python") and toks[1:3] == ["-m", "torch.distributed.run"]:
        toks = toks[3:]
    else:
        return None
    spec = TorchrunSpec()
    i = 0
    while i < len(toks):
        tok = toks[i]
        key, eq, inline = tok.partition("=")
        if key.startswith("--"):  # torchrun accepts both --nproc-per-node and --nproc_per_node
            key = "--" + key[2:].replace("-", "_")
        if key in ("-m", "--module"):
            if i + 1 >= len(toks):
                raise ValueError("torchrun -m needs a module name")
            spec.module, spec.script_args = toks[i + 1], toks[i + 2:]
            return spec
        if key in _FLAG_OPTS:
            if key == "--no_python":
                spec.no_python = True
            i += 1
            continue
        if key in _VALUE_OPTS:
            if eq:
                val, step = inline, 1
            else:
                if i + 1 >= len(toks):
                    raise ValueError(f"torchrun option {tok} needs a value")
                val, step = toks[i + 1], 2
            if key == "--nnodes":
                lo, _, hi = val.partition(":")
                spec.min_nnodes, spec.max_nnodes = int(lo), int(hi or lo)
            elif key == "--nproc_per_node":
                spec.nproc_per_node = val
            elif key == "--rdzv_id":
                spec.rdzv_id = val
            elif key == "--role":
                spec.role = val
            elif key in ("--tee", "-t"):
                spec.tee = int(val) if val.isdigit() else 0
            elif key == "--node_rank":
                spec.node_rank = int(val)
            elif key == "--max_restarts":
                spec.max_restarts = int(val)
            i += step
            continue
        if tok.startswith("-"):
            raise ValueError(f"local_cuda cannot interpret torchrun option {tok!r} in: {cmdline}")
        spec.script, spec.script_args = tok, toks[i + 1:]
        return spec
    raise ValueError(f"no training script or module in torchrun command: {cmdline}")


# ---------------------------------------------------------------------------------------------------------------
# request
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class WorkerGroup:
    """The workers one replica ("node") expands to."""

    role_name: str
    replica_id: int
    is_torchrun: bool
    nproc: int
    devices: List[int]  # CUDA ordinal per local rank (as seen by the worker, after CUDA_VISIBLE_DEVICES)
    cpu_sets: List[List[int]]  # affinity per local rank ([] = leave alone)
    cmd: List[str]
    env: Dict[str, str]  # node-level env (role env + local scheduler additions)
    rank_offset: int = 0
    world_size: int = 1
    group_world_size: int = 1
    cwd: Optional[str] = None
    role_label: str = ""


@dataclass
class CudaPopenRequest(PopenRequest):
    groups: Dict[str, List[WorkerGroup]] = field(default_factory=dict)
    max_retries: int = 0
    shm_name: str = ""
    master_port: int = 0
    stage_mb: int = 0
    omp_num_threads: int = 1
    job_timeout: float = 0.0
    exit_on_file: Optional[str] = None
    start_on_file: Optional[str] = None


TIMEOUT_EXIT_CODE = 34  # reference: torchx/apps/utils/process_monitor.py:16


# ---------------------------------------------------------------------------------------------------------------
# replica-level log multiplexer
# ---------------------------------------------------------------------------------------------------------------
class _ReplicaLogMux:
    """Follows the per-worker stdout/stderr files of one replica (across attempts) and writes the ``[rank]:``-prefixed
    merge into the replica-level stdout.log / stderr.log / combined.log that ``torchx log`` reads."""

    def __init__(self, stdout_path: str, stderr_path: str, combined_path: str) -> None:
        os.makedirs(os.path.dirname(stdout_path), exist_ok=True)
        self._sinks = {"out": io.open(stdout_path, "ab", buffering=0), "err": io.open(stderr_path, "ab", buffering=0)}
        self._combined = io.open(combined_path, "ab", buffering=0)
        self._lock = threading.Lock()
        self._sources: List[Dict[str, Any]] = []
        self._closed = False
        self._thread = threading.Thread(target=self._pump, name="replica-log-mux", daemon=True)
        self._thread.start()

    def add_source(self, path: str, kind: str, prefix: bytes) -> None:
        with self._lock:
            self._sources.append({"path": path, "kind": kind, "prefix": prefix, "fd": None, "tail": b""})

    def _drain(self) -> bool:
        moved = False
        with self._lock:
            sources = list(self._sources)
        for src in sources:
            if src["fd"] is None:
                if not os.path.exists(src["path"]):
                    continue
                src["fd"] = io.open(src["path"], "rb", buffering=0)
            data = src["fd"].read(1 << 16)
            if not data:
                continue
            moved = True
            data = src["tail"] + data
            lines = data.split(b"\n")
            src["tail"] = lines.pop()
            if lines:
                blob = b"".join(src["prefix"] + ln + b"\n" for ln in lines)
                self._sinks[src["kind"]].write(blob)
                self._combined.write(blob)
        return moved

    def _pump(self) -> None:
        while True:
            if not self._drain():
                if self._closed:
                    break
                time.sleep(0.05)
        for src in self._sources:
            if src["tail"]:
                blob = src["prefix"] + src["tail"] + b"\n"
                self._sinks[src["kind"]].write(blob)
                self._combined.write(blob)
                src["tail"] = b""

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        self._thread.join()
        for src in self._sources:
            if src["fd"] is not None:
                src["fd"].close()
        for f in (*self._sinks.values(), self._combined):
            f.close()


# ---------------------------------------------------------------------------------------------------------------
# host topology (best effort; everything degrades to "don't pin")
# ---------------------------------------------------------------------------------------------------------------
def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(device: int) -> List[int]:
    """Cores local to ``device``'s NUMA node, from nvidia-smi's PCI bus id and sysfs.  [] if unknown."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)],
                             capture_output=True, text=True, check=True, timeout=20).stdout.strip()
        bus = out.splitlines()[0].strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]  # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return []
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            allowed = set(os.sched_getaffinity(0))
            return [c for c in _parse_cpulist(f.read()) if c in allowed]
    except Exception:  # noqa: BLE001 - any failure just means "no pinning"
        return []


# ---------------------------------------------------------------------------------------------------------------
# cross-process app registry
# ---------------------------------------------------------------------------------------------------------------
def registry_dir(session_name: str) -> str:
    home = os.environ.get("TORCHX_HOME") or os.path.join(os.path.expanduser("~"), ".torchx_b200")
    return os.path.join(home, "apps", session_name or "default")


def _cancel_request_path(session_name: str, app_id: str) -> str:
    return os.path.join(registry_dir(session_name), f"{app_id}.cancel")


def _pid_alive(pid: int) -> bool:
    try:
        os.kill(pid, 0)
    except ProcessLookupError:
        return False
    except PermissionError:
        return True
    return True


class _AppRecord:
    """What another process can learn about an app: registry entry + the ``SUCCESS`` manifest in its log tree."""

    def __init__(self, path: str) -> None:
        with open(path) as f:
            self.entry: Dict[str, Any] = json.load(f)
        self.app_id: str = self.entry["app_id"]
        self.log_dir: str = self.entry["log_dir"]

    def manifest(self) -> Optional[Dict[str, Any]]:
        try:
            with open(os.path.join(self.log_dir, "SUCCESS")) as f:
                return json.load(f)
        except (OSError, ValueError):
            return None

    def state(self) -> Tuple[AppState, int]:
        m = self.manifest()
        if m is not None:
            return AppState[m.get("final_state", "UNKNOWN")], int(m.get("num_restarts", 0))
        if "final_state" in self.entry:  # the supervisor's own record; outlives a temporary log dir
            return AppState[self.entry["final_state"]], int(self.entry.get("num_restarts", 0))
        if _pid_alive(int(self.entry.get("launcher_pid", -1))):
            return AppState.RUNNING, 0
        return AppState.UNKNOWN, 0  # launcher gone without closing the app (killed -9): nothing supervises it any more


def _unlink_rendezvous_blocks(shm_name: str) -> None:
    """Rank 0 unlinks the control block once everyone has mapped it; a gang that died mid-rendezvous leaves it behind.
    POSIX shm objects live under /dev/shm on Linux: remove whatever this app's epochs left there."""
    import glob

    for path in glob.glob(f"/dev/shm/{shm_name.lstrip('/')}_*"):
        try:
            os.unlink(path)
        except OSError:
            pass


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return int(s.getsockname()[1])


# ---------------------------------------------------------------------------------------------------------------
# app bookkeeping
# ---------------------------------------------------------------------------------------------------------------
class _CudaApp(_LocalAppDef):
    def __init__(self, id: str, log_dir: str, request: CudaPopenRequest) -> None:
        super().__init__(id, log_dir)
        self.request = request
        self.muxes: List[_ReplicaLogMux] = []
        self.monitor: Optional[threading.Thread] = None
        self.stop_monitor = threading.Event()
        self.failure_msg = ""
        self.deadline: Optional[float] = None  # time.monotonic() at which job_timeout expires
        self.exit_code: Optional[int] = None   # TIMEOUT_EXIT_CODE when the supervisor ended the job


class LocalCudaScheduler(LocalScheduler):
    def __init__(self, session_name: str, image_provider_class: Callable[[LocalOpts], ImageProvider] = CWDImageProvider,
                 cache_size: int = 100, extra_paths: Optional[List[str]] = None) -> None:
        super().__init__(session_name, image_provider_class, cache_size, extra_paths, backend="local_cuda")

    # -- options ----------------------------------------------------------------------------------------------------
    def _run_opts(self) -> runopts:
        return CudaOpts.as_runopts()

    def _opts(self, cfg: LocalOpts) -> CudaOpts:  # type: ignore[override]
        return cfg if isinstance(cfg, CudaOpts) else CudaOpts.from_cfg(cfg)

    def _numa_cpus(self, device: int) -> List[int]:
        return gpu_numa_cpus(device)

    # -- request construction (pure: no dirs, no processes, no sockets) -------------------------------------------------
    def _device_pool(self, opts: CudaOpts) -> List[int]:
        if opts.devices:
            return [int(d) for d in opts.devices]
        return list(range(self._cuda_device_count()))

    def _to_popen_request(self, app: AppDef, cfg: LocalOpts) -> CudaPopenRequest:  # type: ignore[override]
        opts = self._opts(cfg)
        app_id = make_unique(app.name)
        provider = self._image_provider_class(cfg)
        app_log_dir = self._get_app_log_dir(app_id, opts)
        pool = self._device_pool(opts)
        cursor = 0  # next unclaimed entry of the device pool
        role_params: Dict[str, List[ReplicaParam]] = {}
        role_log_dirs: Dict[str, List[str]] = {}
        groups: Dict[str, List[WorkerGroup]] = {}
        max_retries = 0
        for role in app.roles:
            max_retries = max(max_retries, role.max_retries)
            params = role_params.setdefault(role.name, [])
            log_dirs = role_log_dirs.setdefault(role.name, [])
            role_groups = groups.setdefault(role.name, [])
            img_root = provider.fetch_role(role)
            cwd = provider.get_cwd(role.image)
            role.env["PATH"] = self._role_path_env(role, cwd, opts)
            for replica_id in range(role.num_replicas):
                rr, replica_log_dir = self._replica_role(role, img_root, app_id, replica_id, app_log_dir)
                node = provider.get_replica_param(img_root, rr, os.path.join(replica_log_dir, STDOUT_LOG),
                                                  os.path.join(replica_log_dir, STDERR_LOG), os.path.join(replica_log_dir, COMBINED_LOG))
                params.append(node)
                log_dirs.append(replica_log_dir)
                spec = None
                if rr.entrypoint == "bash" and len(rr.args) == 2 and rr.args[0] == "-c":
                    spec = parse_torchrun(rr.args[1])
                if spec is None:
                    role_groups.append(WorkerGroup(role.name, replica_id, False, 1, [-1], [[]], list(node.args), dict(node.env), cwd=node.cwd))
                    continue
                if spec.nproc_per_node in ("gpu", "auto"):
                    nproc = max(1, role.resource.gpu if role.resource.gpu > 0 else len(pool) // max(role.num_replicas, 1))
                else:
                    nproc = int(spec.nproc_per_node)
                wants_gpu = role.resource.gpu > 0 or len(pool) >= cursor + nproc
                devices: List[int] = []
                env = dict(node.env)
                if wants_gpu and len(pool) >= cursor + nproc:
                    mine = pool[cursor: cursor + nproc]
                    cursor += nproc
                    if role.num_replicas > 1:
                        # several "nodes" on one box: give each its own visible set so LOCAL_RANK -> cuda:LOCAL_RANK holds
                        env[ENV_CUDA_VISIBLE_DEVICES] = ",".join(str(d) for d in mine)
                        devices = list(range(nproc))
                    else:
                        devices = mine
                    physical = mine
                else:
                    if role.resource.gpu > 0 and pool:
                        raise ValueError(f"role {role.name!r} replica {replica_id} needs {nproc} GPUs but only {max(len(pool) - cursor, 0)} of {len(pool)} remain")
                    devices, physical = [-1] * nproc, []  # CPU job (gloo): no pinning
                cpu_sets: List[List[int]] = [[] for _ in range(nproc)]
                if opts.pin_cpus and physical:
                    by_node: Dict[Tuple[int, ...], List[int]] = {}
                    for lr, dev in enumerate(physical):
                        by_node.setdefault(tuple(self._numa_cpus(dev)), []).append(lr)
                    for cpus, ranks in by_node.items():
                        if not cpus:
                            continue
                        share = max(len(cpus) // len(ranks), 1)
                        for k, lr in enumerate(ranks):
                            cpu_sets[lr] = list(cpus[k * share: (k + 1) * share]) or list(cpus)
                role_groups.append(WorkerGroup(role.name, replica_id, True, nproc, devices, cpu_sets, spec.worker_cmd(), env, cwd=node.cwd,
                                               role_label=spec.role))
            # ranks: replicas of a role are consecutive "nodes"
            tr = [g for g in role_groups if g.is_torchrun]
            world = sum(g.nproc for g in tr)
            offset = 0
            for g in tr:
                g.rank_offset, g.world_size, g.group_world_size = offset, world, len(tr)
                offset += g.nproc
        return CudaPopenRequest(app_id, app_log_dir, role_params, role_log_dirs, groups=groups, max_retries=max_retries,
                                shm_name=f"/b2_{app_id}"[:200], master_port=opts.master_port, stage_mb=opts.stage_mb,
                                omp_num_threads=opts.omp_num_threads, job_timeout=float(opts.job_timeout or 0.0),
                                exit_on_file=opts.exit_on_file, start_on_file=opts.start_on_file)

    def _submit_dryrun(self, app: AppDef, cfg: LocalOpts) -> AppDryRunInfo[CudaPopenRequest]:  # type: ignore[override]
        return AppDryRunInfo(self._to_popen_request(app, cfg), lambda req: pprint.pformat(asdict(req), indent=2, width=100))

    # -- launching ----------------------------------------------------------------------------------------------------
    def _worker_env(self, req: CudaPopenRequest, g: WorkerGroup, local_rank: int, attempt: int, error_file: str) -> Dict[str, str]:
        """The torchrun contract (local_elastic_agent.py:309-323) plus the peer-buffer rendezvous variables."""
        env = dict(g.env)
        rank = g.rank_offset + local_rank
        env.update({
            "RANK": str(rank), "LOCAL_RANK": str(local_rank), "GROUP_RANK": str(g.replica_id), "ROLE_RANK": str(rank),
            "ROLE_NAME": g.role_label, "LOCAL_WORLD_SIZE": str(g.nproc), "WORLD_SIZE": str(g.world_size),
            "GROUP_WORLD_SIZE": str(g.group_world_size), "ROLE_WORLD_SIZE": str(g.world_size),
            "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(req.master_port),
            "TORCHELASTIC_RESTART_COUNT": str(attempt), "TORCHELASTIC_MAX_RESTARTS": str(req.max_retries),
            "TORCHELASTIC_RUN_ID": req.app_id, "TORCHELASTIC_USE_AGENT_STORE": "False",
            "TORCH_NCCL_ASYNC_ERROR_HANDLING": os.getenv("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1"),
            "TORCHELASTIC_ERROR_FILE": error_file,
            "B2_SHM_NAME": f"{req.shm_name}_{g.role_name}"[:240], "B2_EPOCH": str(attempt),
        })
        if g.devices[local_rank] >= 0:
            env["B2_DEVICE"] = str(g.devices[local_rank])
        if ENV_CUDA_VISIBLE_DEVICES in g.env and g.group_world_size > 1:
            # several "nodes" on one box, each with its own CUDA_VISIBLE_DEVICES: a peer's GPU is not visible to this process,
            # so the VMM import (cuMemSetAccess on an invisible device) and a multicast team cannot be set up; CUDA IPC can
            # open an invisible peer's memory - stay on that arena backend
            env.setdefault("B2_VMM", "0")
        if req.stage_mb:
            env["B2_STAGE_MB"] = str(req.stage_mb)
        if "OMP_NUM_THREADS" not in env and "OMP_NUM_THREADS" not in os.environ:
            env["OMP_NUM_THREADS"] = str(req.omp_num_threads)
        return env

    def _spawn_attempt(self, app: _CudaApp, attempt: int) -> None:
        req = app.request
        app.role_replicas = {}
        for role_name, role_groups in req.groups.items():
            for g, mux in zip(role_groups, [m for m in app.muxes if m.role == role_name]):  # type: ignore[attr-defined]
                replica_dir = req.role_log_dirs[role_name][g.replica_id]
                for lr in range(g.nproc):
                    if g.is_torchrun:
                        wdir = os.path.join(replica_dir, f"attempt_{attempt}", f"rank_{lr}")
                        prefix = f"[{g.rank_offset + lr}]:".encode()
                    else:
                        wdir = os.path.join(replica_dir, f"attempt_{attempt}")
                        prefix = b""
                    os.makedirs(wdir, exist_ok=True)
                    out, err = os.path.join(wdir, STDOUT_LOG), os.path.join(wdir, STDERR_LOG)
                    # the replica-level error.json stays the root-cause file of the LATEST attempt
                    error_file = os.path.join(wdir, "error.json")
                    env = self._worker_env(req, g, lr, attempt, error_file) if g.is_torchrun else dict(g.env)
                    p = ReplicaParam(g.cmd, env, out, err, None, g.cwd)
                    cpus = g.cpu_sets[lr] if lr < len(g.cpu_sets) else []
                    preexec = (lambda c=tuple(cpus): os.sched_setaffinity(0, c)) if cpus else None
                    rep = self._popen(role_name, g.replica_id, p, preexec_fn=preexec)
                    app.add_replica(role_name, rep)
                    mux.add_source(out, "out", prefix)
                    mux.add_source(err, "err", prefix)

    def schedule(self, dryrun_info: AppDryRunInfo[CudaPopenRequest]) -> str:  # type: ignore[override]
        req: CudaPopenRequest = dryrun_info.request
        if req.master_port == 0:
            req.master_port = _free_port()
        with self._apps_lock:
            self._reserve_slot(req.app_id)
            os.makedirs(req.log_dir)
            app = _CudaApp(req.app_id, req.log_dir, req)
            for role_name, params in req.role_params.items():
                for replica_id, node in enumerate(params):
                    os.makedirs(req.role_log_dirs[role_name][replica_id])
                    mux = _ReplicaLogMux(node.stdout, node.stderr, node.combined)  # type: ignore[arg-type]
                    mux.role = role_name  # type: ignore[attr-defined]
                    app.muxes.append(mux)
            app.extra_closers.append(lambda: [m.close() for m in app.muxes])
            app.extra_closers.append(lambda: _unlink_rendezvous_blocks(req.shm_name))
            app.deadline = (time.monotonic() + req.job_timeout) if req.job_timeout > 0 else None
            if req.start_on_file and not os.path.exists(req.start_on_file):
                app.set_state(AppState.PENDING)  # the monitor thread launches the gang when the file appears
            else:
                self._spawn_attempt(app, 0)
                app.set_state(AppState.RUNNING)
            self._apps[req.app_id] = app
            self._register(app)
        app.monitor = threading.Thread(target=self._monitor, args=(app,), name=f"monitor-{req.app_id}", daemon=True)
        app.monitor.start()
        return req.app_id

    # -- registry -----------------------------------------------------------------------------------------------------
    def _register(self, app: "_CudaApp") -> None:
        try:
            d = registry_dir(self.session_name)
            os.makedirs(d, exist_ok=True)
            entry = {"app_id": app.id, "log_dir": app.log_dir, "launcher_pid": os.getpid(), "created": getattr(app, "created", None) or time.time(),
                     "scheduler": self.backend, "roles": {r: len(g) for r, g in app.request.groups.items()}}
            app.created = entry["created"]
            if is_terminal(app.state):
                entry.update(final_state=app.state.name, num_restarts=app.num_restarts, finished=time.time())
            tmp = os.path.join(d, f".{app.id}.tmp")
            with open(tmp, "w") as f:
                json.dump(entry, f)
            os.replace(tmp, os.path.join(d, f"{app.id}.json"))
        except OSError as e:  # the registry is a convenience; never fail a launch over it
            log.debug("could not write the app registry entry: %s", e)

    def _record(self, app_id: str) -> Optional[_AppRecord]:
        path = os.path.join(registry_dir(self.session_name), f"{app_id}.json")
        try:
            return _AppRecord(path)
        except (OSError, ValueError, KeyError):
            return None

    def list(self, cfg: Optional[Mapping[str, CfgVal]] = None) -> List[ListAppResponse]:  # type: ignore[override]
        """Apps of this session known to the registry (newest first), whichever process launched them."""
        d = registry_dir(self.session_name)
        out: List[Tuple[float, ListAppResponse]] = []
        for name in (os.listdir(d) if os.path.isdir(d) else []):
            if not name.endswith(".json"):
                continue
            rec = self._record(name[:-5])
            if rec is None:
                continue
            live = self._apps.get(rec.app_id)
            state = live.state if live is not None else rec.state()[0]
            out.append((float(rec.entry.get("created", 0)), ListAppResponse(app_id=rec.app_id, state=state, name=rec.app_id.rsplit("-", 1)[0])))
        return [r for _, r in sorted(out, key=lambda t: -t[0])]

    # -- supervision --------------------------------------------------------------------------------------------------
    def _monitor(self, app: _CudaApp) -> None:
        try:
            self._supervise(app)
        finally:
            with app.lock:
                if is_terminal(app.state):  # keep the outcome where `torchx list/status` of other processes find it even
                    self._register(app)     # after a temporary log dir (and its SUCCESS manifest) is gone

    def _supervise(self, app: _CudaApp) -> None:
        """Gang supervision: any worker failure kills the attempt; with retries left the whole gang is re-launched
        under the next epoch (RetryPolicy.APPLICATION semantics), otherwise the app is FAILED."""
        req = app.request
        while not app.stop_monitor.wait(MONITOR_POLL_S):
            with app.lock:
                if is_terminal(app.state):
                    return
                # -- `torchx cancel` from another process: a request file next to the registry entry ------------------
                cancel_request = _cancel_request_path(self.session_name, app.id)
                if os.path.exists(cancel_request):
                    app.failure_msg = "cancelled on request of another process"
                    app.set_state(AppState.CANCELLED)
                    app.close()  # SIGTERM, grace period, SIGKILL; then the SUCCESS manifest other processes read the state from
                    try:
                        os.unlink(cancel_request)  # doubles as the acknowledgement the requester waits for
                    except OSError:
                        pass
                    return
                # -- process_monitor semantics for the gang (reference torchx/apps/utils/process_monitor.py:65-118) --
                timed_out = app.deadline is not None and time.monotonic() > app.deadline
                if app.state == AppState.PENDING:  # start_on_file: nothing launched yet
                    if timed_out:
                        app.failure_msg = "reached timeout before launching"
                        app.exit_code = TIMEOUT_EXIT_CODE
                        app.set_state(AppState.FAILED)
                        return
                    if os.path.exists(req.start_on_file or ""):
                        self._spawn_attempt(app, 0)
                        app.set_state(AppState.RUNNING)
                    continue
                if timed_out or (req.exit_on_file and os.path.exists(req.exit_on_file)):
                    app.failure_msg = (f"reached the job timeout of {req.job_timeout:g} s" if timed_out
                                       else f"{req.exit_on_file} exists") + ": gang terminated by the scheduler"
                    app.exit_code = TIMEOUT_EXIT_CODE
                    app.kill()  # SIGTERM, grace period, SIGKILL
                    app.set_state(AppState.FAILED)
                    return
                reps = app.replicas()
                failed = [r for r in reps if r.failed()]
                if failed:
                    r0 = failed[0]
                    app.failure_msg = f"{r0.role_name}[{r0.replica_id}] pid {r0.proc.pid} exited with code {r0.proc.returncode}"
                    app.kill()
                    if app.num_restarts < app.request.max_retries:
                        app.num_restarts += 1
                        log.warning("app %s: %s; re-launching (restart %d of %d)", app.id, app.failure_msg, app.num_restarts,
                                    app.request.max_retries)
                        self._spawn_attempt(app, app.num_restarts)
                        continue
                    app.set_state(AppState.FAILED)
                    return
                if reps and not any(r.is_alive() for r in reps):
                    app.set_state(AppState.SUCCEEDED)
                    return

    def _refresh_state(self, app: _LocalAppDef) -> AppState:
        return app.state  # the monitor thread owns state transitions

    def describe(self, app_id: str) -> Optional[DescribeAppResponse]:
        app = self._apps.get(app_id)
        if app is None:
            rec = self._record(app_id)  # launched by another process?
            if rec is None or not (os.path.isdir(rec.log_dir) or "final_state" in rec.entry):
                return None
            state, restarts = rec.state()
            return DescribeAppResponse(app_id=app_id, state=state, num_restarts=restarts, structured_error_msg=NONE,
                                       ui_url=f"file://{rec.log_dir}", msg="state recovered from the app registry")
        with app.lock:
            err = app.get_structured_error_msg()
            if is_terminal(app.state):
                app.close()
            msg = getattr(app, "failure_msg", "")
            return DescribeAppResponse(app_id=app_id, state=app.state, num_restarts=app.num_restarts, structured_error_msg=err,
                                       ui_url=f"file://{app.log_dir}", msg=msg or "<NONE>")

    def log_iter(self, app_id: str, role_name: str, k: int = 0, regex: Optional[str] = None, since: Optional[Any] = None,
                 until: Optional[Any] = None, should_tail: bool = False, streams: Optional[Stream] = None) -> Any:
        if app_id in self._apps:
            return super().log_iter(app_id, role_name, k, regex, since, until, should_tail, streams)
        rec = self._record(app_id)
        if rec is None:
            raise KeyError(app_id)
        name = {None: COMBINED_LOG, Stream.COMBINED: COMBINED_LOG, Stream.STDOUT: STDOUT_LOG, Stream.STDERR: STDERR_LOG}[streams]
        log_file = os.path.join(rec.log_dir, role_name, str(k), name)
        if not os.path.isfile(log_file):
            raise RuntimeError(f"app: {app_id} has no log file {log_file}")
        lines = split_lines_iterator(LogIterator(app_id, log_file, self, should_tail=should_tail))
        return filter_regex(regex, lines) if regex else lines

    def _cancel_existing(self, app_id: str) -> None:
        app = self._apps.get(app_id)
        if app is None:
            self._request_cancel(app_id)
            return
        if isinstance(app, _CudaApp):
            app.stop_monitor.set()
        with app.lock:
            app.close()
            app.state = AppState.CANCELLED

    def _request_cancel(self, app_id: str) -> None:
        """Cancel an app another process launched: leave ``<registry>/<app_id>.cancel`` for its supervisor thread (which
        polls every MONITOR_POLL_S) and wait until it has acknowledged by removing the file, i.e. the gang is down."""
        rec = self._record(app_id)
        if rec is None:
            raise RuntimeError(f"app {app_id} is not in this session's registry")
        state, _ = rec.state()
        if is_terminal(state):
            return
        launcher = int(rec.entry.get("launcher_pid", -1))
        if state == AppState.UNKNOWN or not _pid_alive(launcher):
            raise RuntimeError(f"app {app_id}: its launcher (pid {launcher}) is gone, so nothing supervises the workers any more; "
                               f"their pids are in the logs under {rec.log_dir}")
        path = _cancel_request_path(self.session_name, app_id)
        with open(path, "w") as f:
            f.write(f"requested by pid {os.getpid()} at {time.time():.3f}\n")
        deadline = time.monotonic() + KILL_GRACE_S + 5.0
        while os.path.exists(path):
            if time.monotonic() > deadline or not _pid_alive(launcher):
                try:
                    os.unlink(path)
                except OSError:
                    pass
                raise RuntimeError(f"app {app_id}: the launcher (pid {launcher}) did not act on the cancel request")
            time.sleep(MONITOR_POLL_S)

    def close(self) -> None:
        for app in list(self._apps.values()):
            if isinstance(app, _CudaApp):
                app.stop_monitor.set()
        super().close()


def create_scheduler(session_name: str, cache_size: int = 100, extra_paths: Optional[List[str]] = None,
                     image_provider_class: Callable[[LocalOpts], ImageProvider] = CWDImageProvider, **kwargs: Any) -> LocalCudaScheduler:
    return LocalCudaScheduler(session_name=session_name, image_provider_class=image_provider_class, cache_size=cache_size,
                              extra_paths=extra_paths)
