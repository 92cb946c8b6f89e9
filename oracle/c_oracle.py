"""ctypes front end of the C oracle (oracle/poseidon_oracle.c).  TEST
INFRASTRUCTURE — see oracle/__init__.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "poseidon_oracle.c")):
        subprocess.run(["make", "-C", HERE, "-s", "PY=" + (os.environ.get("PYTHON") or "python3")], check=True)
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.oracle_hash.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_hash_batch.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int]
        _lib.oracle_tree_insert_merge.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                                  C.c_void_p, C.c_void_p, C.c_int]
        _lib.oracle_dense_tree_root.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int,
                                                C.c_int]
    return _lib


def hash_batch(n_inputs: int, data, threads: int = 0, faithful: bool = False) -> np.ndarray:
    a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data,
                             dtype=np.uint8).reshape(-1)
    n = a.size // (32 * n_inputs)
    out = np.empty((n, 32), dtype=np.uint8)
    rc = lib().oracle_hash_batch(n_inputs, a.ctypes.data, n, out.ctypes.data, threads or os.cpu_count() or 1,
                                 int(faithful))
    assert rc == 0
    return out


def hash_one(inputs, tag: bytes | None = None, faithful: bool = False) -> bytes:
    buf = b"".join(inputs)
    out = C.create_string_buffer(32)
    rc = lib().oracle_hash(len(inputs), buf, tag, out, int(faithful))
    assert rc == 0
    return out.raw


def tree_insert_merge(arity: int, full_depth: int, blank: bool, to_depth: bool, leaves, faithful: bool = False):
    """new + insert*N + merge, the reference's own sequence.  Returns
    (rc, root|None, depth_field, count)."""
    a = np.ascontiguousarray(np.frombuffer(leaves, dtype=np.uint8) if not isinstance(leaves, np.ndarray) else leaves,
                             dtype=np.uint8).reshape(-1)
    n = a.size // 32
    root = C.create_string_buffer(32)
    st = (C.c_uint32 * 3)()
    rc = lib().oracle_tree_insert_merge(arity, full_depth, int(blank), int(to_depth),
                                        a.ctypes.data if n else None, n, root, st, int(faithful))
    return rc, (root.raw if st[2] else None), int(st[0]), int(st[1])


def dense_tree_root(arity: int, depth: int, nodes, threads: int = 0, faithful: bool = False) -> bytes:
    a = np.ascontiguousarray(np.frombuffer(nodes, dtype=np.uint8) if not isinstance(nodes, np.ndarray) else nodes,
                             dtype=np.uint8).reshape(-1)
    root = C.create_string_buffer(32)
    rc = lib().oracle_dense_tree_root(arity, depth, a.ctypes.data if a.size else None, a.size // 32, root,
                                      threads or os.cpu_count() or 1, int(faithful))
    assert rc == 0
    return root.raw


def registration_leaves(public_keys, timestamps, threads: int = 0) -> np.ndarray:
    """hash4(pk.x, pk.y, 1, timestamp) per row (provider.rs:224-233)."""
    pk = np.ascontiguousarray(np.frombuffer(public_keys, dtype=np.uint8) if not isinstance(public_keys, np.ndarray)
                              else public_keys, dtype=np.uint8).reshape(-1, 64)
    n = pk.shape[0]
    rows = np.zeros((n, 4, 32), dtype=np.uint8)
    rows[:, 0] = pk[:, :32]
    rows[:, 1] = pk[:, 32:]
    rows[:, 2, 31] = 1
    ts = np.asarray(timestamps, dtype=np.uint64).reshape(-1)
    rows[:, 3, 24:] = ts.astype(">u8").view(np.uint8).reshape(n, 8)
    return hash_batch(4, rows, threads)


def interaction_leaves(public_keys, data, threads: int = 0) -> np.ndarray:
    """hash4(hash5(d[0..5]), hash5(d[5..10]), pk.x, pk.y) per row (provider.rs:249-278)."""
    pk = np.ascontiguousarray(np.frombuffer(public_keys, dtype=np.uint8) if not isinstance(public_keys, np.ndarray)
                              else public_keys, dtype=np.uint8).reshape(-1, 64)
    d = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data,
                             dtype=np.uint8).reshape(-1, 10, 32)
    n = pk.shape[0]
    left = hash_batch(5, np.ascontiguousarray(d[:, :5]), threads)
    right = hash_batch(5, np.ascontiguousarray(d[:, 5:]), threads)
    rows = np.empty((n, 4, 32), dtype=np.uint8)
    rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3] = left, right, pk[:, :32], pk[:, 32:]
    return hash_batch(4, rows, threads)
