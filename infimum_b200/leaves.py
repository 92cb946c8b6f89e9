"""Leaf hashing on the GPU: the step immediately before the trees.

Mirrors the leaf computation of pallet/src/poll/provider.rs:
    register_participant  :218-241   leaf = hash4(pk.x, pk.y, 1, timestamp)
    consume_interaction   :243-287   leaf = hash4(hash5(d[0..5]), hash5(d[5..10]), pk.x, pk.y)
in bulk, one participant / one message per GPU thread (csrc/leaves.cu).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

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
