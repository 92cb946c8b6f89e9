"""Several GPUs from ONE process, through the C ABI's inf_multi_* entry points
(the form a Rust host would use; `sharded.py` is the one-process-per-GPU form
for torchrun).  Subtree roots are gathered with a single ncclAllGather inside
the library."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .errors import DeviceError, raise_for


class MultiGpu:
    def __init__(self, devices: Sequence[int], peer_copy: bool = False):
        self.lib = _lib.load()
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.inf_multi_init(arr, len(devices), _lib.MULTI_PEER_COPY if peer_copy else 0, C.byref(h))
        if rc != _lib.OK:
            raise DeviceError("inf_multi_init(%s) failed: %s" % (list(devices), _lib.strerror(rc)))
        self.handle = h
        self.devices = list(devices)

    def tree_merge(self, arity: int, full_depth: int, leaves, prepend_blank_leaf: bool, to_depth: bool
                   ) -> Tuple[Optional[bytes], int, int, int]:
        """Returns (root | None, depth field, root depth, rc) — rc 2 means the
        inserts alone completed the tree (TreeAlreadyMerged), root still set."""
        a = np.frombuffer(leaves, dtype=np.uint8) if not isinstance(leaves, np.ndarray) else leaves
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
        root = C.create_string_buffer(32)
        idp, rdp, has = C.c_uint32(), C.c_uint32(), C.c_int()
        rc = self.lib.inf_multi_tree_merge(self.handle, arity, full_depth, int(prepend_blank_leaf), int(to_depth),
                                           a.ctypes.data if a.size else None, a.shape[0], root, C.byref(idp),
                                           C.byref(rdp), C.byref(has))
        if rc not in (_lib.OK, _lib.ERR_TREE_ALREADY_MERGED):
            raise_for(rc)
        return (root.raw if has.value else None), idp.value, rdp.value, rc

    def hash_batch(self, n_inputs: int, inputs, little_endian: bool = False) -> np.ndarray:
        a = np.frombuffer(inputs, dtype=np.uint8) if not isinstance(inputs, np.ndarray) else inputs
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
        n = a.size // (32 * n_inputs)
        out = np.empty((n, 32), dtype=np.uint8)
        raise_for(self.lib.inf_multi_poseidon_hash_batch(self.handle, n_inputs, 1 if little_endian else 0, None,
                                                         a.ctypes.data, n, out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.lib.inf_multi_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
