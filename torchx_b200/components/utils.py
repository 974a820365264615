"""Small utility components used to smoke-test a scheduler's plumbing (reference torchx/components/utils.py:24-318;
only the ones that make sense on one box)."""
from __future__ import annotations

import shlex
from typing import Dict, Optional

import torchx_b200
from torchx_b200 import specs


def echo(msg: str = "hello world", image: str = torchx_b200.IMAGE, num_replicas: int = 1) -> specs.AppDef:
    """
    Echos a message to stdout (calls echo)

    Args:
        msg: message to echo
        image: image to use
        num_replicas: number of replicas to run
    """
    return specs.AppDef(name="echo", roles=[specs.Role(name="echo", image=image, entrypoint="echo", args=[msg],
                                                      num_replicas=num_replicas, resource=specs.resource(cpu=1, gpu=0, memMB=1024))])


def touch(file: str, image: str = torchx_b200.IMAGE) -> specs.AppDef:
    """
    Touches a file (calls touch)

    Args:
        file: file to create
        image: the image to use
    """
    return specs.AppDef(name="touch", roles=[specs.Role(name="touch", image=image, entrypoint="touch", args=[file],
                                                       num_replicas=1, resource=specs.resource(cpu=1, gpu=0, memMB=1024))])


def sh(*args: str, image: str = torchx_b200.IMAGE, num_replicas: int = 1, cpu: int = 1, gpu: int = 0, memMB: int = 1024,
       h: Optional[str] = None, env: Optional[Dict[str, str]] = None, max_retries: int = 0) -> specs.AppDef:
    """
    Runs the provided command via sh. Currently sh does not support
    environment variable substitution.

    Args:
        args: bash arguments
        image: image to use
        num_replicas: number of replicas to run
        cpu: number of cpus per replica
        gpu: number of gpus per replica
        memMB: cpu memory in MB per replica
        h: a registered named resource (if specified takes precedence over cpu, gpu, memMB)
        env: environment varibles to be passed to the run (e.g. ENV1=v1,ENV2=v2,ENV3=v3)
        max_retries: the number of scheduler retries allowed
    """
    escaped = " ".join(shlex.quote(a).replace("$", "\\$") for a in args)
    return specs.AppDef(name="sh", roles=[specs.Role(name="sh", image=image, entrypoint="sh", args=["-c", escaped],
                                                    num_replicas=num_replicas, resource=specs.resource(cpu=cpu, gpu=gpu, memMB=memMB, h=h),
                                                    env=dict(env or {}), max_retries=max_retries)])


def python(*args: str, m: Optional[str] = None, c: Optional[str] = None, script: Optional[str] = None, image: str = torchx_b200.IMAGE,
           name: str = "torchx_utils_python", cpu: int = 1, gpu: int = 0, memMB: int = 1024, h: Optional[str] = None,
           num_replicas: int = 1) -> specs.AppDef:
    """
    Runs ``python`` with the specified module, command or script on the specified
    image and host. Use ``--`` to separate component args and program args
    (e.g. ``torchx run utils.python --m foo.main -- --args to --main``)

    Args:
        args: arguments passed to the program in sys.argv[1:] (ignored with `--c`)
        m: run library module as a script
        c: program passed as string (may error if scheduler has a length limit on args)
        script: .py script to run
        image: image to run on
        name: name of the job
        cpu: number of cpus per replica
        gpu: number of gpus per replica
        memMB: cpu memory in MB per replica
        h: a registered named resource (if specified takes precedence over cpu, gpu, memMB)
        num_replicas: number of copies to run (each on its own container)
    """
    if sum(x is not None for x in (m, c, script)) != 1:
        raise ValueError("exactly one of `-m`, `-c` and `--script` needs to be specified")
    if script:
        cmd = [script]
    elif m:
        cmd = ["-m", m]
    else:
        cmd = ["-c", c]  # type: ignore[list-item]
    return specs.AppDef(name=name, roles=[specs.Role(name="python", image=image, entrypoint="python", num_replicas=num_replicas,
                                                    resource=specs.resource(cpu=cpu, gpu=gpu, memMB=memMB, h=h),
                                                    args=[*cmd, *args], env={"HYDRA_MAIN_MODULE": m} if m else {})])
