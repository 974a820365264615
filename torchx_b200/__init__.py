"""torchx_b200 — a B200-native single-box DDP launch path behind TorchX's own surfaces.

Launcher side mirrors meta-pytorch/torchx (``specs``, ``schedulers``, ``components.dist.ddp``, ``runner``,
``cli``) and adds the ``local_cuda`` scheduler; worker side (``torchx_b200.ddp``) is a thin Python host over
``libb200ddp.so`` (hand-written sm_100a kernels, include/b200ddp.h).
"""
from .version import __version__  # noqa: F401

# Default "image" recorded in AppDefs.  The local schedulers ignore it (the cwd is the image) but components keep
# the field so AppDefs are interchangeable with TorchX's (reference torchx/version.py TORCHX_IMAGE).
IMAGE = f"ghcr.io/pytorch/torchx:{__version__}"
