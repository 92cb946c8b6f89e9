"""Device context: owns one `inf_ctx` (stream, tables, scratch) per CUDA device."""
from __future__ import annotations

import ctypes as C
import threading

from . import _lib
from .errors import DeviceError, raise_for

_contexts = {}
_lock = threading.Lock()


class Context:
    def __init__(self, device: int = 0):
        lib = _lib.load()
        h = C.c_void_p()
        rc = lib.inf_init(int(device), C.byref(h))
        if rc != _lib.OK:
            raise DeviceError("inf_init(device=%d) failed: %s — infimum_b200 runs on CUDA only, "
                              "there is no CPU fallback" % (device, _lib.strerror(rc)))
        self.lib = lib
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.inf_destroy(self.handle)
            self.handle = None

    def check(self, rc: int):
        raise_for(rc, self.handle)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_context(device: int = 0) -> Context:
    """Process-wide context per device (calls on one context are serialised by
    the caller, as `&mut self` is in the reference)."""
    with _lock:
        ctx = _contexts.get(device)
        if ctx is None:
            ctx = _contexts[device] = Context(device)
        return ctx
