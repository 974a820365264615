"""``dist.ddp`` / ``dist.spmd``: the component every data-parallel job on this path is submitted through.

Signature, defaults and the emitted ``AppDef`` are byte-compatible with reference torchx/components/dist.py:162-308
(checked against the reference's own dry-run in tests/test_components.py): one role, ``entrypoint="bash"``,
``args=["-c", "torchrun --rdzv_backend c10d --rdzv_endpoint localhost:0 --rdzv_id '${app_id}' --nnodes N
--nproc_per_node M --tee 3 --role '' <script|-m module> <args>"]``.  ``local_cwd`` executes that string as is (bash ->
torchrun -> workers); ``local_cuda`` parses it back (schedulers/local_cuda_scheduler.py) and spawns the M workers per
node itself, one pinned process per GPU, with the same environment contract torchrun would have given them.
"""
from __future__ import annotations

import os
import re
import shlex
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Tuple

import torchx_b200
from torchx_b200 import specs
from torchx_b200.components.structured_arg import StructuredJArgument, StructuredNameArgument
from torchx_b200.specs import macros

_TORCH_DEBUG_FLAGS: Dict[str, str] = {
    "CUDA_LAUNCH_BLOCKING": "1",
    "NCCL_DESYNC_DEBUG": "1",
    "TORCH_DISTRIBUTED_DEBUG": "DETAIL",
    "TORCH_SHOW_CPP_STACKTRACES": "1",
}
"""Environment preset applied by ``--debug`` (reference dist.py:70-75)."""


class _noquote(str):
    """Marks a command token that must reach bash unquoted (it contains a shell expansion)."""


def _args_join(args: Iterable[str]) -> str:
    return " ".join(a if isinstance(a, _noquote) else shlex.quote(a) for a in args)


_J_FORMS = (
    re.compile(r"^(?P<min>\d+):(?P<max>\d+)x(?P<nproc>\d+)$"),  # 1:2x4  elastic
    re.compile(r"^(?P<max>\d+)x(?P<nproc>\d+)$"),  # 2x4
    re.compile(r"^(?P<nproc>\d+)$"),  # 4      one node
)


def parse_nnodes(j: str) -> Tuple[int, int, int, str]:
    """``[[min_nodes:]nodes x]nproc`` -> ``(min_nnodes, max_nnodes, nproc_per_node, nnodes literal for torchrun)``."""
    for form in _J_FORMS:
        m = form.match(j)
        if not m:
            continue
        g = m.groupdict()
        max_n = g.get("max") or "1"
        min_n = g.get("min") or max_n
        rep = f"{min_n}:{max_n}" if g.get("min") else max_n
        return int(min_n), int(max_n), int(g["nproc"]), rep
    raise ValueError(f"Invalid format for -j, usage example: 1:2x4 or 1x4 or 4. Given: {j}")


def get_role_name(script: Optional[str], m: Optional[str]) -> str:
    if script:
        return Path(script).stem
    if m:
        return m.rpartition(".")[2]
    raise ValueError("failed to compute role_name")


def ddp(
    *script_args: str,
    script: Optional[str] = None,
    m: Optional[str] = None,
    image: str = torchx_b200.IMAGE,
    name: str = "/",
    h: Optional[str] = None,
    cpu: int = 2,
    gpu: int = 0,
    memMB: int = 1024,
    j: str = "1x2",
    env: Optional[Dict[str, str]] = None,
    metadata: Optional[Dict[str, str]] = None,
    max_retries: int = 0,
    rdzv_port: int = 29500,
    rdzv_backend: str = "c10d",
    rdzv_conf: Optional[str] = None,
    mounts: Optional[List[str]] = None,
    debug: bool = False,
    tee: int = 3,
) -> specs.AppDef:
    """
    Distributed data parallel style application (one role, multi-replica).

    Launches ``nproc_per_node`` PyTorch workers on each of ``nnodes`` replicas with the torchrun environment
    contract (RANK, LOCAL_RANK, WORLD_SIZE, MASTER_ADDR, ...).  On a single node the rendezvous endpoint is
    ``localhost:0`` and ``rdzv_port`` is ignored.

    Note: (cpu, gpu, memMB) are mutually exclusive with ``h`` (named resource); ``h`` wins when given.

    Args:
        script_args: argv of the training program
        script: training script to launch on every worker
        m: training module to launch with ``python -m`` instead of a script
        image: recorded in the AppDef (container schedulers); the local schedulers run from the cwd
        name: ``{experiment}/{run}``, ``{experiment}/``, ``/{run}`` or ``{run}``; the run name defaults to the script or
            module name
        cpu: cores requested per replica
        gpu: GPUs requested per replica
        memMB: host memory requested per replica, MB
        h: named resource; wins over cpu / gpu / memMB when given
        j: [{min_nnodes}:]{nnodes}x{nproc_per_node}; on GPU hosts nproc_per_node must not exceed the GPU count
        env: extra environment for the workers, e.g. A=1,B=2
        metadata: scheduler metadata, e.g. K1=v1,K2=v2
        max_retries: scheduler-level retries (gang re-launch on local_cuda)
        rdzv_port: c10d store port on rank 0's host; multi-node only (a single node uses a random free port)
        rdzv_backend: torchrun rendezvous backend; multi-node only
        rdzv_conf: extra rendezvous settings, e.g. join_timeout=600,timeout=600
        mounts: (container schedulers only) e.g. type=bind,src=/host,dst=/job[,readonly]; recorded in the Role, not acted on
            by the local schedulers
        debug: apply the debug environment preset
        tee: which worker streams torchrun also copies to the console: 0 none, 1 stdout, 2 stderr, 3 both
    """
    if (script is None) == (m is None):
        raise ValueError("exactly one of --script and -m must be specified")
    min_nnodes, max_nnodes, nproc_per_node, nnodes_rep = parse_nnodes(j)
    if max_nnodes == 1:
        rdzv_endpoint: str = "localhost:0"  # single agent: let it pick any free port
    else:
        # bash resolves this to ${TORCHX_RANK0_HOST:=localhost}:29500; `$$` survives macro substitution as `$`
        rdzv_endpoint = _noquote(f"$${{{macros.rank0_env}:=localhost}}:{rdzv_port}")

    env = dict(env or {})
    argname = StructuredNameArgument.parse_from(name=name, m=m, script=script)
    env["TORCHX_TRACKING_EXPERIMENT_NAME"] = argname.experiment_name
    env["TORCHX_TRACKING_RUN_NAME"] = argname.run_name
    env.setdefault("LOGLEVEL", os.getenv("LOGLEVEL", "WARNING"))
    if debug:
        env.update(_TORCH_DEBUG_FLAGS)

    cmd: List[str] = ["torchrun", "--rdzv_backend", rdzv_backend]
    if rdzv_conf is not None:
        cmd += ["--rdzv_conf", rdzv_conf]
    cmd += ["--rdzv_endpoint", rdzv_endpoint, "--rdzv_id", f"{macros.app_id}", "--nnodes", nnodes_rep,
            "--nproc_per_node", str(nproc_per_node), "--tee", str(tee), "--role", ""]
    if rdzv_backend == "static":
        cmd += ["--node_rank", f"{macros.replica_id}"]
    cmd += [script] if script is not None else ["-m", m]  # type: ignore[list-item]
    cmd += list(script_args)

    role = specs.Role(
        name=get_role_name(script, m),
        image=image,
        min_replicas=min_nnodes,
        entrypoint="bash",
        num_replicas=int(max_nnodes),
        resource=specs.resource(cpu=cpu, gpu=gpu, memMB=memMB, h=h),
        args=["-c", _args_join(cmd)],
        env=env,
        port_map={"c10d": rdzv_port},
        max_retries=max_retries,
        mounts=specs.parse_mounts(mounts) if mounts else [],
    )
    return specs.AppDef(name=argname.run_name, roles=[role], metadata=dict(metadata or {}))


def spmd(
    *args: str,
    script: Optional[str] = None,
    m: Optional[str] = None,
    image: str = torchx_b200.IMAGE,
    name: str = "/",
    h: str = "gpu.small",
    j: str = "1x1",
    env: Optional[Dict[str, str]] = None,
    metadata: Optional[Dict[str, str]] = None,
    max_retries: int = 0,
    mounts: Optional[List[str]] = None,
    debug: bool = False,
) -> specs.AppDef:
    """
    Single-Process-Multiple-Data launch: ``n x m`` copies of the same program (``-j nxm``); ``-j n`` infers ``m``
    from the GPU count of the named resource ``h``.

    Args:
        args: argv of the program
        script: program file
        m: program module (``python -m``)
        image: recorded in the AppDef; the local schedulers run from the cwd
        name: ``{experiment}/{run}`` (either side optional)
        h: named resource describing one host
        j: {nnodes}x{nproc_per_node}; without the ``x`` part nproc_per_node is the host's GPU count
        env: extra environment, e.g. A=1,B=2
        metadata: scheduler metadata, e.g. K1=v1,K2=v2
        max_retries: scheduler-level retries
        mounts: as for ``ddp``
        debug: apply the debug environment preset
    """
    return ddp(*args, script=script, m=m, image=image, name=name, h=h, j=str(StructuredJArgument.parse_from(h, j)),
               env=env or {}, metadata=metadata, max_retries=max_retries, mounts=mounts, debug=debug)
