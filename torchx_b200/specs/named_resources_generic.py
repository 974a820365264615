"""Generic t-shirt-size resources (reference torchx/specs/named_resources_generic.py:47-60; the AWS instance table is cloud-only and out of scope;
fractional variants come from ``plugins.register.named_resource(fractionals=...)``).  ``dist.ddp -h gpu.xlarge`` resolves here; on
``local_cuda`` the ``gpu`` count of the chosen resource caps how many devices a replica may claim."""
from typing import Callable, Mapping

from .api import Resource

GiB = 1024


def _make(cpu: int, gpu: int, mem_gib: int) -> Callable[[], Resource]:
    return lambda: Resource(cpu=cpu, gpu=gpu, memMB=mem_gib * GiB)


NAMED_RESOURCES: Mapping[str, Callable[[], Resource]] = {
    # cpu-only
    "cpu.nano": lambda: Resource(cpu=1, gpu=0, memMB=512),
    "cpu.micro": _make(1, 0, 1),
    "cpu.small": _make(1, 0, 2),
    "cpu.medium": _make(2, 0, 4),
    "cpu.large": _make(2, 0, 8),
    "cpu.xlarge": _make(8, 0, 32),
    # gpu
    "gpu.small": _make(8, 1, 32),
    "gpu.medium": _make(16, 2, 64),
    "gpu.large": _make(32, 4, 128),
    "gpu.xlarge": _make(64, 8, 256),
    # this box
    "b200.1": _make(24, 1, 180),
    "b200.2": _make(48, 2, 360),
    "b200.4": _make(96, 4, 720),
    "b200.8": _make(192, 8, 1440),
}
