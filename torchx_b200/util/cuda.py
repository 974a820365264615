"""Is there a GPU this process can actually use?  (reference torchx/util/cuda.py:12-31)"""
import torch


def has_cuda_devices() -> bool:
    """True only with BOTH a CUDA build of torch and at least one visible device: ``torch.cuda.is_available()`` alone can
    be true on a GPU-less host, and then ``tensor.cuda()`` still fails."""
    if not torch.cuda.is_available():
        return False
    return torch.cuda.device_count() >= 1
