"""``TorchxEvent`` / ``SourceType`` under the reference's module name (torchx/runner/events/api.py)."""
from torchx_b200.runner.events import SourceType, TorchxEvent  # noqa: F401
