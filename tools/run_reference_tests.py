#!/usr/bin/env python
"""Run the REFERENCE's own unit tests against THIS package (launcher-side parity check).

Needs /root/reference (build container only).  Nothing from the reference is copied into the repo: the selected test
files are copied to a temporary directory, a meta-path finder aliases ``torchx.X`` to ``torchx_b200.X``, imports of
out-of-scope names (Workspace, mounts, trackers, AWS resources, test fixtures) are stubbed, and pytest runs there.
Failures that remain are listed; the known ones are out-of-scope features or tests that patch reference-internal names.

    python tools/run_reference_tests.py            # prints one line per file: passed / failed
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference/torchx"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONFTEST = '''
import importlib, importlib.util, sys
sys.path.insert(0, %r)
import torchx_b200
class _Loader:
    def __init__(self, mod): self.mod = mod
    def create_module(self, spec): return self.mod
    def exec_module(self, module): pass
class _Finder:
    def find_spec(self, name, path=None, target=None):
        if name == "torchx" or name.startswith("torchx."):
            try:
                mod = importlib.import_module("torchx_b200" + name[len("torchx"):])
            except Exception:
                return None
            # NOT sys.modules[name] = mod here: importlib would then take mod.__spec__ and load a second copy
            return importlib.util.spec_from_loader(name, loader=_Loader(mod))
        return None
sys.meta_path.insert(0, _Finder())
''' % REPO

FIXTURE = '''
import unittest, tempfile, shutil, os as _os
from pathlib import Path
class TestWithTmpDir(unittest.TestCase):
    def setUp(self):
        self.tmpdir = Path(tempfile.mkdtemp(prefix="torchx_test"))
    def tearDown(self):
        shutil.rmtree(self.tmpdir, ignore_errors=True)
    def touch(self, filepath):
        f = self.tmpdir / filepath
        f.parent.mkdir(parents=True, exist_ok=True)
        f.touch()
        return f
    def write(self, filepath, content):
        f = self.touch(filepath)
        with open(f, "w") as fp:
            fp.writelines(content)
        return f
    def write_shell_script(self, script_path, content):
        f = self.touch(script_path)
        with open(f, "w") as fp:
            fp.write("#!/bin/bash\\n")
            for line in content:
                fp.write(line + "\\n")
        _os.chmod(f, 0o755)
        return f
'''

PATCHES = [  # (regex, replacement) applied to every copied test file
    (r"from torchx\.test\.fixtures import TestWithTmpDir", FIXTURE),
    (r"from torchx\.specs import named_resources, named_resources_aws, resource", "from torchx.specs import named_resources, resource\nnamed_resources_aws = None"),
    (r"    TORCHX_HOME,\n    Workspace,\n\)", "    TORCHX_HOME,\n)\nWorkspace = None"),
    (r"    UnknownAppException,\n    Workspace,\n\)", "    UnknownAppException,\n)\nWorkspace = None"),
    (r"from torchx\.specs import AppDef, AppDryRunInfo, CfgVal, runopts, Workspace", "from torchx.specs import AppDef, AppDryRunInfo, CfgVal, runopts\nWorkspace = None"),
    (r"from torchx\.tracker\.api import ENV_TORCHX_JOB_ID, ENV_TORCHX_PARENT_RUN_ID", "from torchx.settings import ENV_TORCHX_JOB_ID, ENV_TORCHX_PARENT_RUN_ID"),
    (r"from torchx\.workspace import WorkspaceMixin", "class WorkspaceMixin: pass"),
    (r"from torchx\.components\.component_test_base import ComponentTestCase", "import unittest\nclass ComponentTestCase(unittest.TestCase):\n    def validate(self, module, name):\n        pass"),
    (r"from \.test_util import write_shell_script", "import os as _o\ndef write_shell_script(dir, name, content):\n    p = _o.path.join(dir, name)\n    with open(p, 'w') as f:\n        f.write('#!/bin/bash\\n')\n        for l in content: f.write(l + '\\n')\n    _o.chmod(p, 0o755)\n    return p"),
    (r"from torchx\.specs\.builders import \(\n    _create_args_parser,\n    BindMount,\n    component_args_from_str,\n    ComponentArgs,\n    DeviceMount,\n    make_app_handle,\n    materialize_appdef,\n    parse_mounts,\n    VolumeMount,\n\)",
     "from torchx.specs.builders import _create_args_parser, component_args_from_str, ComponentArgs, materialize_appdef\nfrom torchx.specs.api import make_app_handle\nBindMount = DeviceMount = VolumeMount = parse_mounts = None"),
]

FILES = [
    "util/test/types_test.py", "schedulers/test/ids_test.py", "schedulers/test/streams_test.py", "specs/test/api_test.py",
    "specs/test/builders_test.py", "components/test/dist_test.py", "components/test/structured_arg_test.py",
    "schedulers/test/local_scheduler_test.py", "runner/test/config_test.py", "runner/test/api_test.py",
    "schedulers/test/api_test.py", "schedulers/test/registry_test.py", "distributed/test/dist_test.py", "specs/test/finder_test.py",
    "specs/test/named_resources_generic_test.py", "apps/utils/test/process_monitor_test.py", "util/test/strings_test.py",
    "cli/test/cmd_run_test.py", "cli/test/cmd_log_test.py", "cli/test/cmd_status_test.py", "cli/test/cmd_describe_test.py",
    "cli/test/cmd_cancel_test.py", "cli/test/cmd_list_test.py", "cli/test/cmd_runopts_test.py", "cli/test/cmd_configure_test.py",
    "cli/test/main_test.py", "cli/test/argparse_util_test.py",
]
ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]
if ONLY:
    FILES = [f for f in FILES if any(o in f for o in ONLY)]


def main() -> None:
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} is not available here")
    keep = "--keep" in sys.argv  # leave the prepared directory behind (path printed) to re-run single tests by hand
    work = tempfile.mkdtemp(prefix="ref_tests_")
    try:
        with open(os.path.join(work, "conftest.py"), "w") as f:
            f.write(CONFTEST)
        total_p = total_f = 0
        for rel in FILES:
            src = open(os.path.join(REF, rel)).read()
            for pat, repl in PATCHES:
                src = re.sub(pat, lambda m, r=repl: r, src)
            name = "ref_" + rel.replace("/", "_")
            with open(os.path.join(work, name), "w") as f:
                f.write(src)
            res = subprocess.run([sys.executable, "-m", "pytest", name, "-q", "--no-header", "-p", "no:cacheprovider"], cwd=work,
                                 capture_output=True, text=True, timeout=1200)
            tail = [ln for ln in res.stdout.splitlines() if re.search(r"\d+ (passed|failed|error)", ln)]
            summary = tail[-1] if tail else (res.stdout.strip().splitlines() or ["no output"])[-1]
            p = int((re.search(r"(\d+) passed", summary) or [0, 0])[1])
            fl = int((re.search(r"(\d+) failed", summary) or [0, 0])[1])
            total_p, total_f = total_p + p, total_f + fl
            print(f"{rel:45s} {summary.strip(' =')}")
            for ln in res.stdout.splitlines():
                if ln.startswith("FAILED"):
                    print("      " + ln[:150])
        print(f"TOTAL: {total_p} passed, {total_f} failed")
    finally:
        if keep:
            print(f"kept: {work}")
        else:
            shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
