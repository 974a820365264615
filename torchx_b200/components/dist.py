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
        script_args: arguments to the main module
        script: script or binary to run within the image
        m: the python module path to run
        image: image (e.g. docker); ignored by the local schedulers, where the cwd is the image
        name: job name override in the following format: ``{experimentname}/{runname}`` or ``{experimentname}/`` or ``/{runname}`` or ``{runname}``.
            Uses the script or module name if ``{runname}`` not specified.
        cpu: number of cpus per replica
        gpu: number of gpus per replica
        memMB: cpu memory in MB per replica
        h: a registered named resource (if specified takes precedence over cpu, gpu, memMB)
        j: [{min_nnodes}:]{nnodes}x{nproc_per_node}, for gpu hosts, nproc_per_node must not exceed num gpus
        env: environment varibles to be passed to the run (e.g. ENV1=v1,ENV2=v2,ENV3=v3)
        metadata: metadata to be passed to the scheduler (e.g. KEY1=v1,KEY2=v2,KEY3=v3)
        max_retries: the number of scheduler retries allowed
        rdzv_port: the port on rank0's host to use for hosting the c10d store used for rendezvous.
                   Only takes effect when running multi-node. When running single node, this parameter
                   is ignored and a random free port is chosen.
        rdzv_backend: the rendezvous backend to use. Only takes effect when running multi-node.
        rdzv_conf: the additional rendezvous configuration to use (ex. join_timeout=600,close_timeout=600,timeout=600).
        mounts: mounts to mount into the worker environment/container (ex. type=<bind/volume>,src=/host,dst=/job[,readonly]).
                Not supported by the local schedulers.
        debug: whether to run with preset debug flags enabled
        tee: tees the specified std stream(s) to console + file. 0: none, 1: stdout, 2: stderr, 3: both
    """
    if (script is None) == (m is None):
        raise ValueError("exactly one of --script and -m must be specified")
    if mounts:
        raise ValueError("mounts are a container feature; the single-box launch path has none (SURVEY.md §2 row 16)")

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
        mounts=[],
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
        args: the arguments to the main module or script (e.g. my/trainer.py -foo bar)
        script: path of the main script
        m: the main module name (e.g. my.module.trainer), run as ``python -m``
        image: the base image of the job (ignored by the local schedulers)
        name: ``{experimentname}/{runname}`` or ``{experimentname}/`` or ``/{runname}`` or ``{runname}``
        h: the type of host to run on. Must be one of the registered named resources
        j: {nnodes}x{nproc_per_node}. For GPU hosts omitting nproc_per_node will infer it from the GPU count on the host
        env: environment variables to be passed to the run (e.g. ENV1=v1,ENV2=v2,ENV3=v3)
        metadata: metadata to be passed to the scheduler (e.g. KEY1=v1,KEY2=v2,KEY3=v3)
        max_retries: the number of scheduler retries allowed
        mounts: not supported on the single-box path
        debug: whether to run with preset debug flags enabled
    """
    return ddp(*args, script=script, m=m, image=image, name=name, h=h, j=str(StructuredJArgument.parse_from(h, j)),
               env=env or {}, metadata=metadata, max_retries=max_retries, mounts=mounts, debug=debug)
