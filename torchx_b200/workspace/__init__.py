"""Workspace building hook for image-based (plugin) schedulers; see :mod:`torchx_b200.workspace.api`."""
from torchx_b200.workspace.api import WorkspaceMixin  # noqa: F401
