"""Host-side mirror of the reference's poll state tree, backed by the CUDA
kernels through the C ABI.

Mirrors pallet/src/poll/state.rs:
    PollStateTree {depth, full_depth, arity, count, hashes, root}   :70-91
    AmortizedIncrementalMerkleTree::new / insert / merge / hash     :120-302
and the two callers that define the contract, pallet/src/poll/provider.rs:
    merge_registrations  :289-311
    merge_interactions   :313-327

The reference inserts one leaf per extrinsic and keeps only the frontier; the
GPU path is the batch equivalent: leaves are buffered host-side by `insert`
(or handed over in bulk with `extend`) and folded into the stored frontier on
the device in one go (`inf_tree_append`) when the frontier is asked for; a tree
that still is what `new` made is reduced by `merge` in one shot
(`inf_tree_merge`), any other from its frontier (`inf_tree_merge_frontier`).
Either way the result is field-for-field what the reference would hold:
{depth, count, root, hashes}.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .context import Context, get_context
from .errors import MerkleTreeError
from .hasher import Poseidon


def get_merkle_zeroes(arity: int, ctx: Optional[Context] = None) -> List[bytes]:
    """zeroes.rs:81-85 — 33 levels; any arity other than 2 gets the quinary table."""
    ctx = ctx or get_context()
    buf = C.create_string_buffer(33 * 32)
    ctx.check(ctx.lib.inf_merkle_zeroes(ctx.handle, arity, buf))
    return [buf.raw[32 * i:32 * i + 32] for i in range(33)]


def empty_ballot_roots() -> List[bytes]:
    """zeroes.rs:73-79"""
    buf = C.create_string_buffer(5 * 32)
    _lib.load().inf_empty_ballot_roots(buf)
    return [buf.raw[32 * i:32 * i + 32] for i in range(5)]


class PollStateTree:
    """Field-for-field the reference struct (state.rs:70-91).  What has been hashed
    in so far is held the way the pallet stores it — the frontier `hashes` — plus
    a buffer of leaves not yet folded in, so a tree can be resumed from persisted
    state (`from_state`) and never needs its leaf history."""

    def __init__(self, arity: int, full_depth: int, zero_hash: Optional[Tuple[int, bytes]] = None,
                 ctx: Optional[Context] = None):
        self.arity = int(arity)
        self.full_depth = int(full_depth)
        self.depth = 0
        self.count = 0
        self.root: Optional[bytes] = None
        self.ctx = ctx or get_context()
        # state.rs:150-158: an optional pre-seeded (level, hash) entry.  The reference only ever
        # seeds level 0 with zeroes[0] (state.rs:48-52); that case also has a one-shot merge.
        self._frontier: List[Tuple[int, bytes]] = []
        self._blank = False
        self._fresh = True                   # frontier is still what new() made
        if zero_hash is not None:
            lvl, h = int(zero_hash[0]), bytes(zero_hash[1])
            if len(h) != 32:
                raise ValueError("hash must be 32 bytes")
            self._frontier = [(lvl, h)]
            self._blank = lvl == 0 and h == get_merkle_zeroes(self.arity, self.ctx)[0]
            self._fresh = self._blank
        self._chunks: List[np.ndarray] = []
        self._pending = 0
        self._depth_before_pending = 0

    # -- reference constructor name
    @classmethod
    def new(cls, arity: int, full_depth: int, zero_hash: Optional[Tuple[int, bytes]] = None,
            ctx: Optional[Context] = None) -> "PollStateTree":
        return cls(arity, full_depth, zero_hash, ctx)

    @classmethod
    def from_state(cls, arity: int, full_depth: int, depth: int, count: int,
                   hashes: Sequence[Tuple[int, bytes]], root: Optional[bytes] = None,
                   ctx: Optional[Context] = None) -> "PollStateTree":
        """Resume from the persisted struct (what the pallet reads back from storage,
        lib.rs:706-714): insert / merge continue from the stored frontier."""
        t = cls(arity, full_depth, None, ctx)
        t.depth, t.count = int(depth), int(count)
        t.root = bytes(root) if root is not None else None
        t._frontier = [(int(l), bytes(h)) for l, h in hashes]
        t._fresh = False
        t._depth_before_pending = t.depth
        return t

    def _logical(self) -> int:
        """Leaves the frontier stands for (the blank leaf included)."""
        return sum(self.arity ** l for l, _ in self._frontier)

    def _total(self) -> int:
        return self._logical() + self._pending

    def _flush(self):
        """Fold the buffered leaves into the frontier (inf_tree_append)."""
        if not self._pending:
            return
        lv = self._leaves()
        n_in = len(self._frontier)
        cap = 4 * 33
        in_levels = bytes(l for l, _ in self._frontier)
        in_hashes = b"".join(h for _, h in self._frontier)
        levels = C.create_string_buffer(cap)
        hashes = C.create_string_buffer(cap * 32)
        n, depth, has = C.c_uint32(), C.c_uint32(), C.c_int()
        root = C.create_string_buffer(32)
        rc = self.ctx.lib.inf_tree_append(self.ctx.handle, self.arity, self.full_depth, in_levels, in_hashes, n_in,
                                          self._depth_before_pending, lv.ctypes.data, lv.shape[0], levels, hashes, cap,
                                          C.byref(n), C.byref(depth), C.byref(has), root)
        self.ctx.check(rc)
        self._frontier = [(levels.raw[i], hashes.raw[32 * i:32 * i + 32]) for i in range(n.value)]
        self._chunks, self._pending, self._fresh = [], 0, False
        self.depth = depth.value
        self._depth_before_pending = self.depth
        if has.value:
            self.root = root.raw

    @property
    def hashes(self) -> List[Tuple[int, bytes]]:
        """The frontier `PollStateTree.hashes` (state.rs:85-86): empty once `root` is
        set; before that, the (level, hash) pairs the reference's insert cascade leaves."""
        if self.root is not None:
            return []
        self._flush()
        return list(self._frontier)

    def insert(self, leaf: bytes) -> "PollStateTree":
        """state.rs:176-225 (buffers the leaf; hashing happens in batches)."""
        if self.root is not None:
            raise MerkleTreeError("TreeAlreadyFull")
        if len(leaf) != 32:
            raise ValueError("leaf must be 32 bytes")
        return self.extend(np.frombuffer(bytes(leaf), dtype=np.uint8).reshape(1, 32))

    def extend(self, leaves) -> "PollStateTree":
        """Bulk insert: (n, 32) uint8 array or n*32 bytes, in insertion order."""
        if self.root is not None:
            raise MerkleTreeError("TreeAlreadyFull")
        a = np.frombuffer(leaves, dtype=np.uint8) if not isinstance(leaves, np.ndarray) else leaves
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
        cap = self.arity ** self.full_depth
        if self._total() + a.shape[0] > cap:
            raise MerkleTreeError("TreeAlreadyFull")
        if not self._pending:
            self._depth_before_pending = self.depth
        self._chunks.append(a)
        self._pending += a.shape[0]
        self.count += a.shape[0]
        self._update_depth()
        if self._total() == cap:
            # insert() completes the tree by itself (state.rs:218-222)
            if self._fresh:
                self._reduce(to_depth=True, completing=True)
            else:
                self._flush()
        return self

    def _update_depth(self):
        n, d = self._total(), 0
        while self.arity ** (d + 1) <= n and d < self.full_depth:
            d += 1
        self.depth = max(self.depth, d)                           # state.rs:212-213

    def _leaves(self) -> np.ndarray:
        if not self._chunks:
            return np.empty((0, 32), dtype=np.uint8)
        if len(self._chunks) > 1:
            self._chunks = [np.concatenate(self._chunks, axis=0)]
        return self._chunks[0]

    def _reduce(self, to_depth: bool, completing: bool = False):
        """One-shot new + insert x N + merge over the buffered leaves (fresh trees only)."""
        lv = self._leaves()
        root = C.create_string_buffer(32)
        idepth, rdepth, has = C.c_uint32(), C.c_uint32(), C.c_int()
        rc = self.ctx.lib.inf_tree_merge(self.ctx.handle, self.arity, self.full_depth,
                                         1 if self._blank else 0, 1 if to_depth else 0,
                                         lv.ctypes.data if lv.size else None, lv.shape[0], root,
                                         C.byref(idepth), C.byref(rdepth), C.byref(has))
        if rc == _lib.ERR_TREE_ALREADY_MERGED and completing:
            rc = _lib.OK
        self.ctx.check(rc)
        if has.value:
            self.root = root.raw
            self._chunks, self._pending, self._frontier = [], 0, []
        self.depth = idepth.value
        return rdepth.value

    def merge(self, to_depth: bool) -> "PollStateTree":
        """state.rs:230-281"""
        if self.root is not None:
            raise MerkleTreeError("TreeAlreadyMerged")
        if self._fresh:
            self._reduce(to_depth)
            return self
        self._flush()
        if self.root is not None:                               # cannot happen: extend() flushes on completion
            return self
        n = len(self._frontier)
        root = C.create_string_buffer(32)
        has, rdepth = C.c_int(), C.c_uint32()
        rc = self.ctx.lib.inf_tree_merge_frontier(self.ctx.handle, self.arity, self.full_depth,
                                                  bytes(l for l, _ in self._frontier),
                                                  b"".join(h for _, h in self._frontier), n, 1 if to_depth else 0,
                                                  root, C.byref(has), C.byref(rdepth))
        self.ctx.check(rc)
        if has.value:
            self.root = root.raw
            self._frontier = []
        return self

    @staticmethod
    def hash(inputs: Sequence[bytes], ctx: Optional[Context] = None) -> bytes:
        """state.rs:284-302: circom hasher of matching width over 32-byte BE
        inputs reduced mod p, canonical 32-byte BE out."""
        h = Poseidon.new_circom(len(inputs), ctx)
        return h.hash_batch(b"".join(bytes(b) for b in inputs), 1).tobytes()


def new_registration_tree(registration_depth: int, ctx: Optional[Context] = None) -> PollStateTree:
    """PollState::new, registrations (state.rs:48-52)."""
    ctx = ctx or get_context()
    return PollStateTree.new(2, registration_depth, (0, get_merkle_zeroes(2, ctx)[0]), ctx)


def new_interaction_tree(interaction_depth: int, ctx: Optional[Context] = None) -> PollStateTree:
    """PollState::new, interactions (state.rs:53-57)."""
    return PollStateTree.new(5, interaction_depth, None, ctx)


def merge_registrations(tree: PollStateTree) -> Tuple[PollStateTree, bytes]:
    """provider.rs:289-311: registrations.merge(false), then
    commitment.process = (0, H3(root, EMPTY_BALLOT_ROOTS[1], 0))."""
    tree.merge(False)
    if tree.root is None:
        raise MerkleTreeError("MergeFailed")
    commitment = PollStateTree.hash([tree.root, empty_ballot_roots()[1], bytes(32)], tree.ctx)
    return tree, commitment


def merge_interactions(tree: PollStateTree, registrations_count: int, process_subtree_depth: int,
                       tally_subtree_depth: int) -> Tuple[PollStateTree, int, int]:
    """provider.rs:313-327: interactions.merge(true) and the expected proof counts."""
    tree.merge(True)
    batch = tree.arity ** process_subtree_depth
    expected_process = tree.count // batch + (1 if tree.count % batch else 0)
    expected_tally = 1 + registrations_count // (2 ** tally_subtree_depth)
    return tree, expected_process, expected_tally
