"""``unittest`` base class for component authors (reference torchx/components/component_test_base.py:27-124).

    class MyComponentTest(ComponentTestCase):
        def test_cli_can_parse_it(self):
            self.validate(my_components, "train")            # == `torchx run my_components.py:train --help`
        def test_it_runs(self):
            status = self.run_component(my_components.train, {"epochs": 1}, scheduler="local_cuda", timeout=120)
            self.assertEqual(status.state, AppState.SUCCEEDED)
"""
from __future__ import annotations

import os
import shutil
import tempfile
import time
import unittest
from types import ModuleType
from typing import Any, Callable, Dict, Optional

from torchx_b200.runner import get_runner
from torchx_b200.specs import AppDef, AppStatus
from torchx_b200.specs.builders import _create_args_parser
from torchx_b200.specs.finder import get_component


class ComponentTestCase(unittest.TestCase):
    """Gives each test a scratch ``self.test_dir`` and runs it from the package root, so components that name files
    relative to the repository resolve them."""

    def setUp(self) -> None:
        self.test_dir = tempfile.mkdtemp("torchx_component_test")
        self.old_cwd = os.getcwd()
        os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

    def tearDown(self) -> None:
        os.chdir(self.old_cwd)
        shutil.rmtree(self.test_dir, ignore_errors=True)

    def validate(self, module: ModuleType, function_name: str) -> None:
        """The component passes validation when addressed as ``/abs/file.py:function_name`` and its docstring / signature
        yield a working ``--help``."""
        path = getattr(module, "__file__", None)
        assert path, f"module must have __file__: {module}"
        comp = get_component(f"{os.path.abspath(path)}:{function_name}")
        with self.assertRaises(SystemExit):  # argparse prints the help text and exits 0
            _create_args_parser(comp.fn).parse_args(["--help"])

    def run_component(self, component: Callable[..., AppDef], args: Optional[Dict[str, Any]] = None,
                      scheduler_params: Optional[Dict[str, Any]] = None, scheduler: str = "local_cwd", interval: float = 0.1,
                      timeout: float = 1) -> Optional[AppStatus]:
        """Build the AppDef with ``args``, submit it to ``scheduler`` (its factory gets ``scheduler_params``) and poll every
        ``interval`` seconds; returns the status once it is terminal, but never before ``timeout`` seconds have passed (so a
        caller may also use it to observe a still-running app)."""
        app = component(**(args or {}))
        runner = get_runner(name=None, component_defaults=None, **(scheduler_params or {}))
        handle = runner.run(app, scheduler)
        waited = 0.0
        status = runner.status(handle)
        while waited < timeout or (status is not None and not status.is_terminal()):
            time.sleep(interval)
            waited += interval
            status = runner.status(handle)
        return status
