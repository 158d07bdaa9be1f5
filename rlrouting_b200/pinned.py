"""A small pool of page-locked host buffers whose lifetime follows the NumPy arrays handed out.

`Network.train` returns per-step vectors of R x T doubles (32 MB at 8,192 replicas x 500 steps).
Pinning that much memory costs ~18 ms per allocation and a pageable destination pays first-touch
page faults, so result buffers are page-locked once, returned zero-copy as NumPy arrays, and go
back to the pool when the last view of the array is garbage collected (weakref.finalize).
"""
import threading
import weakref

import numpy as np


class PinnedPool:
    def __init__(self, max_bytes=4 << 30):
        self.free = {}
        self.max_bytes = max_bytes
        self.held = 0
        self.lock = threading.Lock()

    def _release(self, key, buf):
        with self.lock:
            self.free.setdefault(key, []).append(buf)

    def take(self, shape, dtype=np.float64):
        """Returns (torch_tensor, numpy_array) sharing one pinned buffer of `shape`."""
        import torch
        tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.int64): torch.int64,
               np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
        key = (tuple(shape), np.dtype(dtype).str)
        with self.lock:
            lst = self.free.get(key)
            buf = lst.pop() if lst else None
        if buf is None:
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            with self.lock:
                if self.held + nbytes > self.max_bytes:       # drop cached buffers of other shapes first
                    self.free.clear()
                self.held += nbytes
            buf = torch.empty(tuple(shape), dtype=tdt, pin_memory=True)
        arr = buf.numpy()
        weakref.finalize(arr, self._release, key, buf)
        return buf, arr


POOL = PinnedPool()
