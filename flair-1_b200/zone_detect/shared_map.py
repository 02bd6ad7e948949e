"""The one output map of a zone that several ranks (one process per GPU) fill together.

The reference writes every tile's window straight into the single output raster (src/zone_detect/main.py:421-426).
With the zone sharded over ranks the equivalent is one [bands, H, W] uint8 array in POSIX shared memory: every
rank maps it, page-locks the rows it writes, and `fb_detect_zone_shard` copies its write rectangles from its GPU
directly to their place -- no rank's bytes pass through another rank's GPU or PCIe link, and the writer rank reads
the finished map after one barrier. A file under /dev/shm (tmpfs) backs it; when that is too small the system's
temporary directory is used instead (an ordinary page-cache-backed mapping, slower to fault in but identical
otherwise).
"""
from __future__ import annotations

import mmap
import os
import shutil
import tempfile
from pathlib import Path

import numpy as np

from .. import _native

PAGE = mmap.PAGESIZE


def backing_dir(nbytes: int) -> Path:
    shm = Path("/dev/shm")
    try:
        if shm.is_dir() and shutil.disk_usage(shm).free > nbytes + (64 << 20):
            return shm
    except OSError:
        pass
    return Path(tempfile.gettempdir())


class SharedHostMap:
    """uint8 [bands, rows, width] shared between the ranks of one node.

    create=True (one rank, before the others attach): creates and sizes the backing file. Every rank then calls
    pin_rows() for the rows it is going to write; close() unpins and unmaps, unlink() (creator) removes the file."""

    def __init__(self, path: str | os.PathLike, bands: int, rows: int, width: int, create: bool) -> None:
        self.path = str(path)
        self.shape = (int(bands), int(rows), int(width))
        self.nbytes = int(bands) * int(rows) * int(width)
        size = max(self.nbytes, 1)
        flags = os.O_RDWR | (os.O_CREAT | os.O_TRUNC if create else 0)
        fd = os.open(self.path, flags, 0o600)
        try:
            if create:
                os.ftruncate(fd, size)
            self._mm = mmap.mmap(fd, size, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        finally:
            os.close(fd)
        self.array = np.frombuffer(self._mm, dtype=np.uint8, count=self.nbytes).reshape(self.shape)
        self._base = self.array.ctypes.data if self.nbytes else 0
        self._pinned: list[int] = []
        self._creator = create

    @staticmethod
    def fresh_path(nbytes: int, tag: str = "map") -> str:
        return str(backing_dir(nbytes) / f"flairb200-{tag}-{os.getpid()}-{os.urandom(4).hex()}.u8")

    def band(self, b: int) -> np.ndarray:
        return self.array[b]

    def pin_rows(self, row0: int, row1: int) -> None:
        """Page-lock rows [row0, row1) of every band (rounded out to whole pages) for asynchronous D2H copies."""
        bands, rows, width = self.shape
        lib = _native.load_library()
        for b in range(bands):
            lo = (b * rows + row0) * width
            hi = (b * rows + row1) * width
            lo_p = (self._base + lo) // PAGE * PAGE
            hi_p = min(-(-(self._base + hi) // PAGE) * PAGE, self._base + -(-max(self.nbytes, 1) // PAGE) * PAGE)
            if hi_p <= lo_p:
                continue
            if self._pinned and lo_p < self._pinned[-1][1]:   # page shared with the previous band's span
                lo_p = self._pinned[-1][1]
                if hi_p <= lo_p:
                    continue
            rc = lib.fb_host_register(lo_p, hi_p - lo_p)
            if rc != 0:
                msg = lib.fb_last_error(None)
                raise _native.NativeError(rc, msg.decode() if msg else "fb_host_register failed")
            self._pinned.append((lo_p, hi_p))

    def close(self) -> None:
        lib = _native.load_library()
        for lo_p, _ in self._pinned:
            lib.fb_host_unregister(lo_p)
        self._pinned = []
        self.array = None
        try:
            self._mm.close()
        except (BufferError, ValueError):
            pass   # a caller still holds a view; the mapping goes away with the process

    def unlink(self) -> None:
        try:
            os.unlink(self.path)
        except FileNotFoundError:
            pass
