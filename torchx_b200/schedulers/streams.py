"""Log fan-in: follow N growing source files and append what they produce to one sink, optionally prefixing every
line (``[3]: ...``) the way ``torchrun --tee`` labels worker output.  Runs a daemon thread; must be closed.
(Reference torchx/schedulers/streams.py:16-71 merges stdout+stderr into combined.log; the line-prefixing mode is what
``local_cuda`` uses to present one replica-level log for the workers it spawns itself.)"""
from __future__ import annotations

import io
import os
import threading
import time
from typing import BinaryIO, Dict, Optional, Sequence


class Tee:
    POLL_S = 0.05
    CHUNK = 1 << 16

    def __init__(self, out: BinaryIO, *sources: str, prefixes: Optional[Sequence[bytes]] = None) -> None:
        if not sources:
            raise ValueError("Tee needs at least one source file")
        if prefixes is not None and len(prefixes) != len(sources):
            raise ValueError("one prefix per source")
        self.out = out
        self._lock = threading.Lock()
        self._paths = list(sources)
        self._prefixes = list(prefixes) if prefixes is not None else None
        self._fds: Dict[int, io.FileIO] = {}
        self._partial: Dict[int, bytes] = {}
        self._closed = False
        self._thread = threading.Thread(target=self._pump, name="tee", daemon=True)
        self._thread.start()

    def _open_pending(self) -> None:
        for i, path in enumerate(self._paths):
            if i not in self._fds and os.path.exists(path):
                self._fds[i] = io.open(path, "rb", buffering=0)

    def _drain_once(self) -> bool:
        moved = False
        self._open_pending()
        for i, fd in self._fds.items():
            data = fd.read(self.CHUNK)
            if not data:
                continue
            moved = True
            if self._prefixes is None:
                self.write(data)
                continue
            data = self._partial.pop(i, b"") + data
            lines = data.split(b"\n")
            tail = lines.pop()
            if tail:
                self._partial[i] = tail
            if lines:
                self.write(b"".join(self._prefixes[i] + ln + b"\n" for ln in lines))
        return moved

    def _pump(self) -> None:
        while True:
            if not self._drain_once():
                if self._closed:
                    break
                time.sleep(self.POLL_S)
        for i, tail in self._partial.items():  # unterminated last lines
            self.write((self._prefixes[i] if self._prefixes else b"") + tail + b"\n")
        self._partial.clear()

    def write(self, data: bytes) -> int:
        with self._lock:
            return self.out.write(data)

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        self._thread.join()
        with self._lock:
            for fd in self._fds.values():
                fd.close()
            self.out.close()
