#!/usr/bin/env python
"""Run the REFERENCE's own unit tests against THIS package (launcher-side parity check).

Needs /root/reference (build container only).  Nothing from the reference is copied into the repo: the selected test
files (with the data files next to them) are copied to a temporary directory, a meta-path finder aliases ``torchx.X`` to
``torchx_b200.X``, the few imports of things this package does not have (the reference's test fixtures module, the AWS
resource table, the worker-side tracker backends) are stubbed, plugin fixtures are renamed to this package's namespace
(``torchx_b200_plugins``, ``torchx_b200.*`` entry-point groups), and pytest runs there, one process per file.
Failures that remain are listed; they need cloud pieces (AWS table, a kubernetes scheduler, the booth example app).

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

DIST_FIXTURE = FIXTURE + '''
IS_CI = IS_MACOS = False
class DistributedTestCase(TestWithTmpDir):
    def run_ddp(self, world_size, fn):
        from torch.distributed.launcher.api import LaunchConfig, elastic_launch
        cfg = LaunchConfig(min_nodes=1, max_nodes=1, nproc_per_node=world_size, rdzv_backend="c10d", rdzv_endpoint="127.0.0.1:0",
                           max_restarts=0, monitor_interval=0.01)
        return elastic_launch(cfg, entrypoint=fn)
'''

PATCHES = [  # (regex, replacement) applied to every copied test file
    (r"from torchx\.test\.fixtures import DistributedTestCase, IS_CI, IS_MACOS", DIST_FIXTURE),
    (r"from torchx\.runtime\.tracking import FsspecResultTracker", "FsspecResultTracker = None"),
    (r"from torchx\.util\.test\.entrypoints_test import EntryPoint_from_text",
     "def EntryPoint_from_text(text):\n    import configparser, importlib.metadata as _m\n    c = configparser.ConfigParser(delimiters='=')\n"
     "    c.read_string(text)\n    return [_m.EntryPoint(n, v, g) for g in c.sections() for n, v in c.items(g)]"),
    (r"from torchx\.test\.fixtures import TestWithTmpDir", FIXTURE),
    (r"from torchx\.specs import named_resources, named_resources_aws, resource", "from torchx.specs import named_resources, resource\nnamed_resources_aws = None"),
    (r"from torchx\.tracker\.api import ENV_TORCHX_JOB_ID, ENV_TORCHX_PARENT_RUN_ID", "from torchx.settings import ENV_TORCHX_JOB_ID, ENV_TORCHX_PARENT_RUN_ID"),
    (r"torchx\.util\.test\.entrypoints_test", "ref_entrypoints_test"),  # the test names ITSELF as an entry-point target
    (r"from \.test_util import write_shell_script", "import os as _o\ndef write_shell_script(dir, name, content):\n    p = _o.path.join(dir, name)\n    with open(p, 'w') as f:\n        f.write('#!/bin/bash\\n')\n        for l in content: f.write(l + '\\n')\n    _o.chmod(p, 0o755)\n    return p"),
]

FILES = [
    "util/test/types_test.py", "schedulers/test/ids_test.py", "schedulers/test/streams_test.py", "specs/test/api_test.py",
    "specs/test/builders_test.py", "components/test/dist_test.py", "components/test/structured_arg_test.py",
    "schedulers/test/local_scheduler_test.py", "runner/test/config_test.py", "runner/test/api_test.py",
    "schedulers/test/api_test.py", "distributed/test/dist_test.py", "specs/test/finder_test.py",
    "specs/test/named_resources_generic_test.py", "apps/utils/test/process_monitor_test.py",
    "cli/test/cmd_run_test.py", "cli/test/cmd_log_test.py", "cli/test/cmd_status_test.py", "cli/test/cmd_describe_test.py",
    "cli/test/cmd_cancel_test.py", "cli/test/cmd_list_test.py", "cli/test/cmd_runopts_test.py", "cli/test/cmd_configure_test.py",
    "cli/test/main_test.py", "cli/test/argparse_util_test.py", "plugins/test/register_test.py", "plugins/test/registry_test.py",
    "util/test/entrypoints_test.py", "util/test/cuda_test.py", "util/test/modules_test.py", "util/test/shlex_test.py", "util/test/strings_test.py",
    "cli/test/cmd_delete_test.py", "components/test/utils_test.py", "apps/utils/test/copy_test.py", "runner/events/test/lib_test.py",
]
ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]
if ONLY:
    FILES = [f for f in FILES if any(o in f for o in ONLY)]


def _plugin_names(text: str) -> str:
    text = text.replace("torchx_plugins", "torchx_b200_plugins")
    for grp in ("schedulers", "named_resources", "tracker", "cli.cmds"):
        text = text.replace(f'"torchx.{grp}"', f'"torchx_b200.{grp}"')
    return text


def _rename_plugin_namespace(root: str) -> None:
    for dirpath, dirnames, filenames in os.walk(root, topdown=False):
        for fn in filenames:
            if fn.endswith(".py"):
                path = os.path.join(dirpath, fn)
                with open(path) as f:
                    text = f.read()
                with open(path, "w") as f:
                    f.write(_plugin_names(text))
        for d in dirnames:
            if d == "torchx_plugins":
                os.rename(os.path.join(dirpath, d), os.path.join(dirpath, "torchx_b200_plugins"))


def main() -> None:
    if "-h" in sys.argv or "--help" in sys.argv:
        print(__doc__)
        print("    python tools/run_reference_tests.py cli/ plugins/     # only files whose path contains one of the words")
        print("    python tools/run_reference_tests.py --keep ...        # leave the prepared directory behind (path printed)")
        return
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} is not available here")
    keep = "--keep" in sys.argv  # leave the prepared directory behind (path printed) to re-run single tests by hand
    work = tempfile.mkdtemp(prefix="ref_tests_")
    try:
        with open(os.path.join(work, "_alias.py"), "w") as f:  # importable by spawned children too (sys.path travels)
            f.write(CONFTEST)
        with open(os.path.join(work, "conftest.py"), "w") as f:
            f.write("import sys, os\nsys.path.insert(0, os.path.dirname(__file__))\nimport _alias  # noqa: F401\n")
        total_p = total_f = 0
        for rel in FILES:
            src = open(os.path.join(REF, rel)).read()
            for pat, repl in PATCHES:
                src = re.sub(pat, lambda m, r=repl: r, src)
            # one sub-directory per reference test directory, with that directory's data files (component files, config
            # fixtures ...) next to the test, since tests address them relative to __file__
            sub = os.path.join(work, os.path.dirname(rel).replace("/", "_"))
            if not os.path.isdir(sub):
                os.makedirs(sub)
                src_dir = os.path.join(REF, os.path.dirname(rel))
                for entry in os.listdir(src_dir):
                    full = os.path.join(src_dir, entry)
                    if entry.endswith("_test.py") or entry in ("__init__.py", "__pycache__"):
                        continue
                    if os.path.isdir(full):
                        shutil.copytree(full, os.path.join(sub, entry), ignore=shutil.ignore_patterns("__pycache__"))
                    elif os.path.getsize(full) < 1 << 20:
                        shutil.copy(full, os.path.join(sub, entry))
                if rel.startswith("plugins/"):  # this package scans `torchx_b200_plugins.*` and its entry-point groups
                    _rename_plugin_namespace(sub)
            if rel.startswith("plugins/"):
                src = _plugin_names(src)
            if rel.endswith("finder_test.py"):  # this package's entry-point group, and the test's own module path
                src = src.replace("[torchx.components]", "[torchx_b200.components]").replace("torchx.specs.test.finder_test", "ref_finder_test")
                src = src.replace("torchx.specs.test.components", "components").replace("from importlib_metadata import EntryPoints", "from importlib.metadata import EntryPoints")
            if "_alias" not in src:  # processes spawned by a test import the test module without conftest.py
                hook = "import sys as _s, os as _o\n_s.path.insert(0, _o.path.dirname(_o.path.dirname(_o.path.abspath(__file__))))\nimport _alias  # noqa\n"
                m = re.search(r"^from __future__ import .*$", src, flags=re.M)
                src = src[:m.end()] + "\n" + hook + src[m.end():] if m else hook + src
            name = os.path.join(os.path.basename(sub), "ref_" + os.path.basename(rel))
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
