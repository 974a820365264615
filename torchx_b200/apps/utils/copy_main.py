"""``python -m torchx_b200.apps.utils.copy_main --src URL --dst URL``: one file from one fsspec location to another - what the
``utils.copy`` component runs (reference torchx/apps/utils/copy_main.py).  Same-filesystem copies use the filesystem's own
copy; otherwise the bytes are streamed in ``--bufsize`` pieces."""
from __future__ import annotations

import argparse
import os
import shutil
import sys
from typing import List, Optional


def parse_args(argv: List[str]) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="copies a file between fsspec locations")
    p.add_argument("--src", type=str, required=True, help="fsspec location of the file to read from")
    p.add_argument("--dst", type=str, required=True, help="fsspec location of where to copy the file to")
    p.add_argument("--bufsize", type=int, default=64 * 1024, help="bufsize to use for copying")
    return p.parse_args(argv)


def main(argv: Optional[List[str]] = None) -> None:
    import fsspec  # a missing fsspec is an error of this tool, not of the launcher

    a = parse_args(sys.argv[1:] if argv is None else argv)
    print(f"copying from {a.src} to {a.dst}")
    src_fs, src_path = fsspec.core.url_to_fs(a.src)
    dst_fs, dst_path = fsspec.core.url_to_fs(a.dst)
    try:
        dst_fs.mkdir(os.path.dirname(dst_path), create_parents=True)
    except FileExistsError:  # e.g. memory:// when the directory is already there
        pass
    if src_fs == dst_fs:
        print("filesystems are the same, using fs.copy() method")
        src_fs.copy(src_path, dst_path)
        return
    print("filesystems are different, using shutil.copyfileobj()")
    with src_fs.open(src_path, "rb") as fin, dst_fs.open(dst_path, "wb") as fout:
        shutil.copyfileobj(fin, fout, a.bufsize)


if __name__ == "__main__":
    main()
