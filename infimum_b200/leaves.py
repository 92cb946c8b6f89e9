"""Leaf hashing on the GPU: the step immediately before the trees.

Mirrors the leaf computation of pallet/src/poll/provider.rs:
    register_participant  :218-241   leaf = hash4(pk.x, pk.y, 1, timestamp)
    consume_interaction   :243-287   leaf = hash4(hash5(d[0..5]), hash5(d[5..10]), pk.x, pk.y)
in bulk, one participant / one message per GPU thread (csrc/leaves.cu), and the
replay calls that chain it into the tree on the device:
    replay_registrations  = register_participant x n + merge_registrations  :218-241, 289-311
    replay_interactions   = consume_interaction x n  + merge_interactions   :243-287, 313-327
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .context import Context, get_context


def _u8(a, width):
    a = np.frombuffer(a, dtype=np.uint8) if not isinstance(a, np.ndarray) else a
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
    if a.size % width:
        raise ValueError("expected a multiple of %d bytes" % width)
    return a


def registration_leaves(public_keys, timestamps, ctx: Optional[Context] = None, out=None) -> np.ndarray:
    """public_keys: n x 64 bytes (`PublicKey {x, y}`), timestamps: n u64 (block
    numbers).  Returns (n, 32) uint8 leaves."""
    ctx = ctx or get_context()
    pk = _u8(public_keys, 64)
    n = pk.size // 64
    ts = np.ascontiguousarray(np.asarray(timestamps, dtype=np.uint64).reshape(-1))
    if ts.size != n:
        raise ValueError("one timestamp per public key")
    if out is None:
        out = np.empty((n, 32), dtype=np.uint8)
    ctx.check(ctx.lib.inf_registration_leaves(ctx.handle, pk.ctypes.data, ts.ctypes.data, n, out.ctypes.data))
    return out.reshape(n, 32)


def interaction_leaves(public_keys, data, ctx: Optional[Context] = None, out=None) -> np.ndarray:
    """public_keys: n x 64 bytes, data: n x 320 bytes (`PollInteractionData`).
    Returns (n, 32) uint8 leaves."""
    ctx = ctx or get_context()
    pk = _u8(public_keys, 64)
    d = _u8(data, 320)
    n = pk.size // 64
    if d.size // 320 != n:
        raise ValueError("one 10-word message per public key")
    if out is None:
        out = np.empty((n, 32), dtype=np.uint8)
    ctx.check(ctx.lib.inf_interaction_leaves(ctx.handle, pk.ctypes.data, d.ctypes.data, n, out.ctypes.data))
    return out.reshape(n, 32)


def replay_registrations(registration_depth: int, public_keys, timestamps, ctx: Optional[Context] = None,
                         want_leaves: bool = False, retain: bool = False):
    """Raw registrations -> leaves -> merged registration tree, leaves staying on the
    device.  Returns (tree, process_commitment, leaves | None, RetainedTree | None);
    `tree` is the merged PollStateTree the reference would hold."""
    from .paths import RetainedTree
    from .tree import new_registration_tree
    ctx = ctx or get_context()
    pk = _u8(public_keys, 64)
    n = pk.size // 64
    ts = np.ascontiguousarray(np.asarray(timestamps, dtype=np.uint64).reshape(-1))
    if ts.size != n:
        raise ValueError("one timestamp per public key")
    root, commitment = C.create_string_buffer(32), C.create_string_buffer(32)
    idepth = C.c_uint32()
    leaves = np.empty((n, 32), dtype=np.uint8) if want_leaves else None
    handle = C.c_void_p()
    rc = ctx.lib.inf_replay_registrations(ctx.handle, registration_depth, pk.ctypes.data if n else None,
                                          ts.ctypes.data if n else None, n, root, commitment, C.byref(idepth),
                                          leaves.ctypes.data if want_leaves and n else None,
                                          C.byref(handle) if retain else None)
    if rc == _lib.ERR_TREE_ALREADY_MERGED:
        rc = _lib.OK                                   # the inserts completed the tree: root is set either way
    ctx.check(rc)
    t = new_registration_tree(registration_depth, ctx)
    t.depth, t.count, t.root, t._frontier, t._fresh = idepth.value, n, root.raw, [], False
    kept = None
    if retain:
        total, d = n + 1, 0
        while 2 ** d < total:
            d += 1
        kept = RetainedTree._adopt(handle, 2, registration_depth if total == 2 ** registration_depth else d, n, 1, ctx)
    return t, commitment.raw, leaves, kept


def replay_interactions(interaction_depth: int, public_keys, data, registrations_count: int, process_subtree_depth: int,
                        tally_subtree_depth: int, ctx: Optional[Context] = None, want_leaves: bool = False,
                        retain: bool = False):
    """Raw messages -> leaves -> merged interaction tree (merge(true)).  Returns
    (tree, expected_process, expected_tally, leaves | None, RetainedTree | None)."""
    from .paths import RetainedTree
    from .tree import new_interaction_tree
    ctx = ctx or get_context()
    pk = _u8(public_keys, 64)
    d = _u8(data, 320)
    n = pk.size // 64
    if d.size // 320 != n:
        raise ValueError("one 10-word message per public key")
    root = C.create_string_buffer(32)
    has, idepth, ep, et = C.c_int(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    leaves = np.empty((n, 32), dtype=np.uint8) if want_leaves else None
    handle = C.c_void_p()
    rc = ctx.lib.inf_replay_interactions(ctx.handle, interaction_depth, pk.ctypes.data if n else None,
                                         d.ctypes.data if n else None, n, registrations_count, process_subtree_depth,
                                         tally_subtree_depth, root, C.byref(has), C.byref(idepth), C.byref(ep), C.byref(et),
                                         leaves.ctypes.data if want_leaves and n else None,
                                         C.byref(handle) if retain else None)
    if rc == _lib.ERR_TREE_ALREADY_MERGED:
        rc = _lib.OK
    ctx.check(rc)
    t = new_interaction_tree(interaction_depth, ctx)
    t.depth, t.count, t._frontier, t._fresh = idepth.value, n, [], False
    t.root = root.raw if has.value else None
    kept = RetainedTree._adopt(handle, 5, interaction_depth, n, 0, ctx) if retain and handle.value else None
    return t, ep.value, et.value, leaves, kept
