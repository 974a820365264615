"""``WorkspaceMixin``: the hook through which an image-building scheduler turns ``Role.workspace`` (local project directories)
into a patched image before submission (reference torchx/workspace/api.py:83-178).

None of this package's own schedulers uses it - ``local_cuda`` and ``local_cwd`` run from the current directory - but
plugin schedulers are first-class here (``torchx_b200.plugins``), and a container scheduler written against TorchX mixes this
class in.  The Runner and ``Scheduler.submit`` call :meth:`build_workspaces` for such schedulers exactly as TorchX does.
"""
from __future__ import annotations

import abc
import logging
import tempfile
from typing import Dict, Generic, List, Mapping, TypeVar

from torchx_b200.specs import AppDef, CfgVal, Role, runopts

logger = logging.getLogger(__name__)
T = TypeVar("T")


class WorkspaceMixin(abc.ABC, Generic[T]):
    def __init__(self, *args: object, **kwargs: object) -> None:
        super().__init__(*args, **kwargs)

    def workspace_opts(self) -> runopts:
        """Extra ``-cfg`` options of the workspace builder (merged into the scheduler's runopts)."""
        return runopts()

    def build_workspaces(self, roles: List[Role], cfg: Mapping[str, CfgVal]) -> None:
        """Build the workspace of every role that has one; ``role.image`` (and possibly ``role.env``) are updated in place.
        One cache dict lives for the duration of the call so that roles sharing image + workspace are built once."""
        cache: Dict[object, object] = {}
        for i, role in enumerate(roles):
            if not role.workspace:
                continue
            before = role.image
            self.caching_build_workspace_and_update_role(role, cfg, cache)
            if role.image != before:
                logger.info("role[%d]=%s updated with new image to include workspace changes", i, role.name)

    def caching_build_workspace_and_update_role(self, role: Role, cfg: Mapping[str, CfgVal], build_cache: Dict[object, object]) -> None:
        """The method to implement.  The default bridges to the older single-directory hook: a workspace of one unmapped
        project is passed as is, anything else is first merged into a temporary directory."""
        workspace = role.workspace
        if not workspace:
            return
        if workspace.is_unmapped_single_project():
            self.build_workspace_and_update_role(role, str(workspace), cfg)
            return
        with tempfile.TemporaryDirectory(suffix="torchx_workspace_") as merged:
            workspace.merge_into(merged)
            self.build_workspace_and_update_role(role, merged, cfg)

    def build_workspace_and_update_role(self, role: Role, workspace: str, cfg: Mapping[str, CfgVal]) -> None:
        """Older hook (one directory); kept for schedulers that implement it."""
        raise NotImplementedError("implement `caching_build_workspace_and_update_role`")

    def dryrun_push_images(self, app: AppDef, cfg: Mapping[str, CfgVal]) -> T:
        """Remote schedulers: rewrite ``app`` to the final image names and return what :meth:`push_images` needs."""
        raise NotImplementedError("dryrun_push is not implemented")

    def push_images(self, images_to_push: T) -> None:
        raise NotImplementedError("push is not implemented")
