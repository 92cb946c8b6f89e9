"""Retained trees, Merkle paths and outcome verification on the GPU.

Mirrors pallet/src/poll/provider.rs:
    compute_merkle_root_from_path(depth, index, leaf, path)   :396-436
    verify_outcome(outcome)                                   :76-139
and adds the producer side the off-chain coordinator needs (today maci-core in
cli/src/utils.ts:104-126): keep every level of a poll tree on the device and
serve sibling paths in bulk, in the layout the function above consumes.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from .context import Context, get_context
from .hasher import Poseidon

VOTE_TREE_ARITY = 5                      # provider.rs:403


class RetainedTree:
    """All levels of the dense zero-padded tree of `depth` levels, on the device."""

    def __init__(self, arity: int, depth: int, leaves, prepend_blank_leaf: bool = False,
                 ctx: Optional[Context] = None):
        self.ctx = ctx or get_context()
        a = np.frombuffer(leaves, dtype=np.uint8) if not isinstance(leaves, np.ndarray) else leaves
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
        self.arity, self.depth, self.n_leaves = int(arity), int(depth), a.shape[0]
        self.shift = 1 if prepend_blank_leaf else 0
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.inf_tree_build(self.ctx.handle, arity, depth, self.shift,
                                                   a.ctypes.data if a.size else None, a.shape[0], C.byref(h)))
        self.handle = h

    @classmethod
    def _adopt(cls, handle, arity: int, depth: int, n_leaves: int, shift: int, ctx: Context) -> "RetainedTree":
        """Wrap an inf_tree the library handed out (inf_replay_*)."""
        t = cls.__new__(cls)
        t.ctx, t.arity, t.depth, t.n_leaves, t.shift, t.handle = ctx, int(arity), int(depth), int(n_leaves), int(shift), handle
        return t

    def node_paths(self, level: int, node_indices) -> np.ndarray:
        """(n, depth - level, arity-1, 32) sibling paths from nodes of `level` to the
        root — with level = process_subtree_depth, the msgSubrootPathElements of
        every batch (circuits/process-messages.circom:57,85) in one call."""
        idx = np.ascontiguousarray(np.asarray(node_indices, dtype=np.uint64).reshape(-1))
        out = np.empty((idx.size, self.depth - level, self.arity - 1, 32), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.inf_tree_node_paths(self.handle, level, idx.ctypes.data, idx.size, out.ctypes.data))
        return out

    def level_nodes(self, level: int, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """(count, 32) nodes of `level` starting at `first` (zero subtrees included)."""
        if count is None:
            count = max(-(-(self.n_leaves + self.shift) // self.arity ** level) - first, 0)
        out = np.empty((count, 32), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.inf_tree_level_nodes(self.handle, level, first, count, out.ctypes.data))
        return out

    @property
    def root(self) -> bytes:
        buf = C.create_string_buffer(32)
        self.ctx.check(self.ctx.lib.inf_tree_root(self.handle, buf))
        return buf.raw

    def paths(self, leaf_indices) -> np.ndarray:
        """(n, depth, arity-1, 32) uint8 sibling paths.  Indices count the blank
        leaf as leaf 0 when the tree was built with prepend_blank_leaf."""
        idx = np.ascontiguousarray(np.asarray(leaf_indices, dtype=np.uint64).reshape(-1))
        out = np.empty((idx.size, self.depth, self.arity - 1, 32), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.inf_tree_paths(self.handle, idx.ctypes.data, idx.size, out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "handle", None):
            self.ctx.lib.inf_tree_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merkle_roots_from_paths(arity: int, depth: int, indices, leaves, paths, ctx: Optional[Context] = None) -> np.ndarray:
    """Batched compute_merkle_root_from_path: n paths -> (n, 32) roots."""
    ctx = ctx or get_context()
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64).reshape(-1))
    n = idx.size
    lv = np.ascontiguousarray(np.asarray(leaves, dtype=np.uint8).reshape(n, 32))
    pt = np.ascontiguousarray(np.asarray(paths, dtype=np.uint8).reshape(n, depth * (arity - 1) * 32))
    out = np.empty((n, 32), dtype=np.uint8)
    ctx.check(ctx.lib.inf_merkle_roots_from_paths(ctx.handle, arity, depth, idx.ctypes.data, lv.ctypes.data,
                                                  pt.ctypes.data if pt.size else None, n, out.ctypes.data))
    return out


def compute_merkle_root_from_path(depth: int, index: int, leaf: bytes, path: Sequence[Sequence[bytes]],
                                  ctx: Optional[Context] = None) -> Optional[bytes]:
    """provider.rs:396-436, one path (quinary, as in the reference)."""
    if len(path) < depth or any(len(p) < VOTE_TREE_ARITY - 1 for p in path[:depth]):
        return None                       # the reference would index out of bounds; it returns None upstream
    flat = b"".join(bytes(x) for lvl in path[:depth] for x in lvl[:VOTE_TREE_ARITY - 1])
    return merkle_roots_from_paths(VOTE_TREE_ARITY, depth, [index], np.frombuffer(bytes(leaf), dtype=np.uint8),
                                   np.frombuffer(flat, dtype=np.uint8), ctx)[0].tobytes()


def verify_outcome(vote_option_tree_depth: int, n_options: int, tally_commitment: bytes, outcome: dict,
                   ctx: Optional[Context] = None) -> Optional[int]:
    """verify_outcome (provider.rs:76-139) without the `is_proven` guard, with
    the per-option work batched: one path-root launch over all vote options,
    then two hash2 batches.  `outcome` has the fields of `PollOutcome`
    (coordinator.rs:53-77).  Returns the winning option index or None."""
    ctx = ctx or get_context()
    res: List[int] = list(outcome["tally_results"])
    proofs = outcome["tally_result_proofs"]
    if len(res) < n_options or len(proofs) < n_options:
        return None
    d = vote_option_tree_depth
    leaves = b"".join(int(r).to_bytes(32, "big") for r in res[:n_options])     # u32 in the last 4 bytes (:96-97)
    flat = b"".join(bytes(x) for o in range(n_options) for lvl in proofs[o][:d] for x in lvl[:4])
    roots = merkle_roots_from_paths(VOTE_TREE_ARITY, d, list(range(n_options)),
                                    np.frombuffer(leaves, dtype=np.uint8), np.frombuffer(flat, dtype=np.uint8), ctx)
    h2 = Poseidon.new_circom(2, ctx)
    salt, spent = bytes(outcome["tally_result_salt"]), bytes(outcome["spent_votes_hash"])
    a = h2.hash_batch(b"".join(roots[i].tobytes() + salt for i in range(n_options)))
    b = h2.hash_batch(b"".join(a[i].tobytes() + spent for i in range(n_options)))
    if any(b[i].tobytes() != bytes(tally_commitment) for i in range(n_options)):
        return None
    t = h2.hash_batch(bytes(outcome["total_spent"]) + bytes(outcome["total_spent_salt"]))
    t = h2.hash_batch(bytes(outcome["new_results_commitment"]) + t[0].tobytes())
    if t[0].tobytes() != bytes(tally_commitment):
        return None
    best, best_val = 0, 0
    for i in range(n_options):
        if res[i] > best_val:
            best, best_val = i, res[i]
    return best
