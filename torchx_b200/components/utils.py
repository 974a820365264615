"""Small utility components used to smoke-test a scheduler's plumbing (reference torchx/components/utils.py:24-318;
only the ones that make sense on one box)."""
from __future__ import annotations

import shlex
from typing import Dict, Optional

import torchx_b200
from torchx_b200 import specs


def echo(msg: str = "hello world", image: str = torchx_b200.IMAGE, num_replicas: int = 1) -> specs.AppDef:
    """
    Prints ``msg`` on every replica (the ``echo`` binary) - the smallest possible scheduler smoke test.

    Args:
        msg: text to print
        image: recorded in the AppDef; the local schedulers run from the cwd
        num_replicas: how many copies to start
    """
    return specs.AppDef(name="echo", roles=[specs.Role(name="echo", image=image, entrypoint="echo", args=[msg],
                                                      num_replicas=num_replicas, resource=specs.resource(cpu=1, gpu=0, memMB=1024))])


def touch(file: str, image: str = torchx_b200.IMAGE) -> specs.AppDef:
    """
    Creates an empty file (the ``touch`` binary); tests use it to observe macro substitution in arguments.

    Args:
        file: path of the file
        image: recorded in the AppDef; the local schedulers run from the cwd
    """
    return specs.AppDef(name="touch", roles=[specs.Role(name="touch", image=image, entrypoint="touch", args=[file],
                                                       num_replicas=1, resource=specs.resource(cpu=1, gpu=0, memMB=1024))])


def sh(*args: str, image: str = torchx_b200.IMAGE, num_replicas: int = 1, cpu: int = 1, gpu: int = 0, memMB: int = 1024,
       h: Optional[str] = None, env: Optional[Dict[str, str]] = None, max_retries: int = 0) -> specs.AppDef:
    """
    Joins ``args`` into one ``sh -c`` command line.  ``$`` is escaped, so no variable expansion happens in the shell.

    Args:
        args: words of the command
        image: recorded in the AppDef; the local schedulers run from the cwd
        num_replicas: how many copies to start
        cpu: cores requested per replica
        gpu: GPUs requested per replica
        memMB: host memory requested per replica, MB
        h: named resource; wins over cpu / gpu / memMB when given
        env: extra environment, e.g. A=1,B=2
        max_retries: scheduler-level retries (gang re-launch on local_cuda)
    """
    escaped = " ".join(shlex.quote(a).replace("$", "\\$") for a in args)
    return specs.AppDef(name="sh", roles=[specs.Role(name="sh", image=image, entrypoint="sh", args=["-c", escaped],
                                                    num_replicas=num_replicas, resource=specs.resource(cpu=cpu, gpu=gpu, memMB=memMB, h=h),
                                                    env=dict(env or {}), max_retries=max_retries)])


def python(*args: str, m: Optional[str] = None, c: Optional[str] = None, script: Optional[str] = None, image: str = torchx_b200.IMAGE,
           name: str = "torchx_utils_python", cpu: int = 1, gpu: int = 0, memMB: int = 1024, h: Optional[str] = None,
           num_replicas: int = 1) -> specs.AppDef:
    """
    Starts the interpreter on exactly one of: a module (``-m``), an inline program (``-c``) or a script file.  Program
    arguments follow a ``--`` separator: ``torchx run utils.python -m pkg.main -- --flag value``.

    Args:
        args: the program's own argv (unused with ``-c``)
        m: module to run as ``python -m``
        c: source text to run as ``python -c``
        script: path of a .py file
        image: recorded in the AppDef; the local schedulers run from the cwd
        name: job name
        cpu: cores requested per replica
        gpu: GPUs requested per replica
        memMB: host memory requested per replica, MB
        h: named resource; wins over cpu / gpu / memMB when given
        num_replicas: how many copies to start
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


def binary(*args: str, entrypoint: str, name: str = "torchx_utils_binary", num_replicas: int = 1, cpu: int = 1, gpu: int = 0,
           memMB: int = 1024, h: Optional[str] = None) -> specs.AppDef:
    """
    Runs an arbitrary executable with ``args`` - no shell, no interpreter (``torchx run utils.binary --entrypoint nvidia-smi -- -L``).

    Args:
        args: argv[1:] of the program
        entrypoint: the executable (on PATH, or a path)
        name: job name
        num_replicas: how many copies to start
        cpu: cores requested per replica
        gpu: GPUs requested per replica
        memMB: host memory requested per replica, MB
        h: named resource; wins over cpu / gpu / memMB when given
    """
    return specs.AppDef(name=name, roles=[specs.Role(name="binary", image=specs.NONE, entrypoint=entrypoint, args=list(args),
                                                    num_replicas=num_replicas, resource=specs.resource(cpu=cpu, gpu=gpu, memMB=memMB, h=h))])


def copy(src: str, dst: str, image: str = torchx_b200.IMAGE) -> specs.AppDef:
    """
    Copies one file between two fsspec locations (local paths, ``s3://``, ``memory://`` ...); no directories.

    Args:
        src: where to read the file
        dst: where to write it
        image: recorded in the AppDef; the local schedulers run from the cwd
    """
    return specs.AppDef(name="torchx-utils-copy", roles=[specs.Role(
        name="torchx-utils-copy", image=image, entrypoint="python", args=["-m", "torchx_b200.apps.utils.copy_main", "--src", src, "--dst", dst],
        resource=specs.Resource(cpu=1, gpu=0, memMB=1024))])
