"""Multi-GPU poll-tree merge: contiguous subtrees per rank, one all-gather of
subtree roots, top of the tree finished redundantly on every rank.

The reference has no parallel form; this is the sharding the
north star prescribes for `PollStateTree` (pallet/src/poll/state.rs:176-281)
on one 8xB200 box: the logical leaf array (blank leaf first for the
registration tree, state.rs:48-52) is cut at a shard level k into whole
level-k subtrees, each rank reduces its run of subtrees on its own GPU, the
<= a few dozen 32-byte subtree roots per rank are exchanged with ONE
`all_gather` (NCCL over NVLink under torchrun; gloo in the CPU tests), and
every rank finishes the top levels itself, so no broadcast is needed.

The data plane is pluggable so the host logic can be tested without a GPU:
a backend turns (nodes, level_in, n_levels, shift) into the reduced nodes.
`GpuBackend` is the product; tests inject the CPU oracle as the checker.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .context import Context, get_context
from .errors import MerkleTreeError


@dataclass
class ShardPlan:
    arity: int
    n_total: int                 # logical leaves, blank leaf included
    shift: int                   # 1 if the logical array starts with the blank leaf
    level: int                   # shard level k
    n_subtrees: int              # non-empty level-k subtrees
    subtree_ranges: List[Tuple[int, int]]   # per rank [first, last) subtree index
    root_depth: int
    insert_depth: int

    def leaf_range(self, rank: int) -> Tuple[int, int]:
        """[lo, hi) into the CALLER's leaf array (blank leaf not included)."""
        s0, s1 = self.subtree_ranges[rank]
        w = self.arity ** self.level
        lo = max(s0 * w - self.shift, 0)
        hi = max(min(s1 * w, self.n_total) - self.shift, 0)
        return lo, max(hi, lo)

    def rank_shift(self, rank: int) -> int:
        s0, s1 = self.subtree_ranges[rank]
        return self.shift if (s0 == 0 and s1 > 0) else 0


def make_plan(arity: int, full_depth: int, n_leaves: int, prepend_blank_leaf: bool, to_depth: bool,
              world: int, min_subtrees_per_rank: int = 64) -> ShardPlan:
    """Pick the shard level and the contiguous, count-balanced runs of
    NON-EMPTY subtrees (all-zero subtrees are never hashed).

    Ranks get whole subtrees, so their loads differ by up to one subtree: with
    at least 64 per rank that is <= 1.6 %.  (With 4 per rank a 2^26-leaf quinary
    tree on 8 ranks was cut into 35 subtrees, 4 or 5 per rank: 16 % imbalance,
    34.6 ms instead of 29.9.)  A finer cut costs nothing: the number of levels
    is the same, only more of the small top levels run after the gather, and
    the gather grows to a few tens of KB."""
    shift = 1 if prepend_blank_leaf else 0
    n_total = n_leaves + shift
    cap = arity ** full_depth
    if n_total > cap:
        raise MerkleTreeError("TreeAlreadyFull")                 # insert(): state.rs:182
    insert_depth = 0
    while n_total and arity ** (insert_depth + 1) <= n_total and insert_depth < full_depth:
        insert_depth += 1
    if to_depth or n_total == cap:
        root_depth = full_depth
    else:
        root_depth = 0
        while arity ** root_depth < n_total:
            root_depth += 1
    level = 0
    while level + 1 <= root_depth and -(-n_total // arity ** (level + 1)) >= min_subtrees_per_rank * world:
        level += 1
    n_sub = -(-n_total // arity ** level) if n_total else 0
    ranges = [(n_sub * r // world, n_sub * (r + 1) // world) for r in range(world)]
    return ShardPlan(arity, n_total, shift, level, n_sub, ranges, root_depth, insert_depth)


class GpuBackend:
    """Data plane on one GPU through the C ABI (inf_tree_reduce_dev)."""

    def __init__(self, ctx: Optional[Context] = None, device: Optional[int] = None):
        self.ctx = ctx or get_context(torch.cuda.current_device() if device is None else device)
        self.device = torch.device("cuda", self.ctx.device)

    def reduce(self, nodes: torch.Tensor, arity: int, level_in: int, n_levels: int, shift: int = 0) -> torch.Tensor:
        n_in = int(nodes.shape[0])
        n_out = -(-(n_in + shift) // arity ** n_levels) if (n_in + shift) else 0
        out = torch.empty((n_out, 32), dtype=torch.uint8, device=self.device)
        if n_out == 0:
            return out
        got = C.c_uint64()
        # torch's current stream; its legacy default stream has handle 0, which the
        # C ABI reads as "the context's own stream", so name it explicitly
        # (cudaStreamLegacy == 0x1) to stay ordered with the surrounding torch ops
        stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        rc = self.ctx.lib.inf_tree_reduce_dev(self.ctx.handle, arity, level_in, n_levels, shift,
                                              nodes.data_ptr() if n_in else None, n_in, out.data_ptr(),
                                              C.byref(got), stream)
        self.ctx.check(rc)
        assert got.value == n_out
        return out

    def empty(self, n: int) -> torch.Tensor:
        return torch.empty((n, 32), dtype=torch.uint8, device=self.device)


def sharded_tree_merge(local_leaves: torch.Tensor, plan: ShardPlan, backend, group=None):
    """Run the sharded merge.  `local_leaves` is this rank's slice
    (plan.leaf_range(rank)) of the leaf array, on the backend's device.
    Returns the 32-byte root as a (32,) uint8 tensor on the backend's device
    (identical on every rank), or None for an empty tree."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if plan.n_total == 0:
        return None
    a, k = plan.arity, plan.level
    s0, s1 = plan.subtree_ranges[rank]
    mine = backend.reduce(local_leaves, a, 0, k, plan.rank_shift(rank))
    assert mine.shape[0] == s1 - s0, (mine.shape, s0, s1)
    if world > 1:
        width = max(e - b for b, e in plan.subtree_ranges)
        send = backend.empty(width)
        send.zero_()
        send[: mine.shape[0]] = mine
        gathered = backend.empty(width * world)
        dist.all_gather_into_tensor(gathered, send, group=group) if hasattr(dist, "all_gather_into_tensor") \
            else dist.all_gather(list(gathered.view(world, width, 32).unbind(0)), send, group=group)
        parts = [gathered[r * width: r * width + (e - b)] for r, (b, e) in enumerate(plan.subtree_ranges)]
        level_nodes = torch.cat(parts, dim=0)
    else:
        level_nodes = mine
    assert level_nodes.shape[0] == plan.n_subtrees
    top = backend.reduce(level_nodes, a, k, plan.root_depth - k, 0)
    assert top.shape[0] == 1
    return top[0]


def emulated_sharded_merge(all_leaves: torch.Tensor, plan: ShardPlan, backend):
    """The same computation as `sharded_tree_merge` for every rank of the plan,
    run back to back on ONE device with the gather replaced by a concatenation.
    Lets a single GPU check a world-size-N plan (no concurrent ranks needed)."""
    if plan.n_total == 0:
        return None
    parts = []
    for r in range(len(plan.subtree_ranges)):
        lo, hi = plan.leaf_range(r)
        parts.append(backend.reduce(all_leaves[lo:hi], plan.arity, 0, plan.level, plan.rank_shift(r)))
    level_nodes = torch.cat(parts, dim=0)
    assert level_nodes.shape[0] == plan.n_subtrees
    top = backend.reduce(level_nodes, plan.arity, plan.level, plan.root_depth - plan.level, 0)
    return top[0]
