"""Host logic of the multi-GPU tree merge (infimum_b200/sharded.py) on CPU:
world_size-2 and -3 gloo process groups, with the oracle injected as the data
plane (the checker stands in for the GPU; on the GPU box the same code runs
with GpuBackend under NCCL, tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle
from oracle import poseidon_ref as O
from tests.util import random_fr_bytes


class OracleBackend:
    def reduce(self, nodes, arity, level_in, n_levels, shift=0):
        z = O.merkle_zeroes(arity)
        cur = nodes.numpy().reshape(-1, 32)
        if shift:
            cur = np.concatenate([np.frombuffer(z[level_in], dtype=np.uint8).reshape(1, 32)] * shift + [cur])
        for l in range(n_levels):
            pad = (-cur.shape[0]) % arity
            if pad:
                cur = np.concatenate([cur] + [np.frombuffer(z[level_in + l], dtype=np.uint8).reshape(1, 32)] * pad)
            cur = c_oracle.hash_batch(arity, np.ascontiguousarray(cur), threads=1)
        return torch.from_numpy(np.ascontiguousarray(cur))

    def empty(self, n):
        return torch.empty((n, 32), dtype=torch.uint8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cases, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from infimum_b200 import sharded
    out = []
    for (arity, full_depth, n, blank, to_depth) in cases:
        plan = sharded.make_plan(arity, full_depth, n, blank, to_depth, world, min_subtrees_per_rank=2)
        leaves = random_fr_bytes(max(n, 1), seed=arity * 1000 + n)[:n]
        lo, hi = plan.leaf_range(rank)
        root = sharded.sharded_tree_merge(torch.from_numpy(leaves[lo:hi].copy()), plan, OracleBackend())
        out.append((None if root is None else root.numpy().tobytes(), plan.insert_depth, plan.level))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


CASES = [(2, 12, 1000, True, False), (2, 12, 1024, False, True), (2, 12, 1023, True, False), (2, 10, 5, True, False),
         (5, 6, 700, False, True), (5, 5, 3125, False, True), (5, 6, 1, False, True), (2, 12, 0, True, False),
         (5, 6, 0, False, True), (2, 12, 2049, False, False)]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_matches_insert_merge(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, CASES, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i, (arity, full_depth, n, blank, to_depth) in enumerate(CASES):
        leaves = random_fr_bytes(max(n, 1), seed=arity * 1000 + n)[:n]
        rc, root, depth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, to_depth, leaves)
        for r in range(world):
            got_root, got_depth, level = results[r][i]
            assert got_root == root, (CASES[i], r)
            assert got_depth == depth, (CASES[i], r)


def test_plan_balances_non_empty_subtrees():
    from infimum_b200 import sharded
    # 2^20 registrations + blank leaf, 8 ranks: every rank gets work
    plan = sharded.make_plan(2, 21, 1 << 20, True, False, 8)
    sizes = [e - b for b, e in plan.subtree_ranges]
    assert min(sizes) >= 64 and max(sizes) - min(sizes) <= 1
    assert plan.root_depth == 21 and plan.insert_depth == 20
    covered = 0
    for r in range(8):
        lo, hi = plan.leaf_range(r)
        assert lo == covered
        covered = hi
    assert covered == 1 << 20
    # 2^26 quinary leaves, depth 12, 8 ranks
    plan = sharded.make_plan(5, 12, 1 << 26, False, True, 8)
    sizes = [e - b for b, e in plan.subtree_ranges]
    assert sum(sizes) == plan.n_subtrees == -(-(1 << 26) // 5 ** plan.level)
    assert min(sizes) >= 64 and max(sizes) - min(sizes) <= 1
    loads = [hi - lo for lo, hi in (plan.leaf_range(r) for r in range(8))]
    assert max(loads) <= 1.02 * (sum(loads) / 8)              # ranks get whole subtrees: within 2 % of each other
    with pytest.raises(Exception):
        sharded.make_plan(2, 3, 9, False, True, 2)


def test_plan_properties_random_shapes():
    """The plan, for random tree shapes and world sizes, with the default
    granularity: subtree runs are contiguous and cover every subtree, leaf
    ranges partition the leaves, the blank leaf lives on the first rank that
    has work, and the emulated sharded merge (every rank of the plan back to
    back, oracle data plane) reproduces insert x N + merge."""
    import random
    from infimum_b200 import sharded
    rng = random.Random(20261018)
    backend = OracleBackend()
    for case in range(60):
        arity = rng.choice((2, 5))
        full_depth = rng.randint(1, 11) if arity == 2 else rng.randint(1, 4)
        blank = rng.random() < 0.5
        cap = arity ** full_depth - (1 if blank else 0)
        n = rng.choice((0, 1, cap, rng.randint(0, cap), rng.randint(0, min(cap, 40))))
        to_depth = rng.random() < 0.5
        world = rng.choice((1, 2, 3, 8))
        plan = sharded.make_plan(arity, full_depth, n, blank, to_depth, world,
                                 min_subtrees_per_rank=rng.choice((1, 2, 4, 64)))
        assert plan.subtree_ranges[0][0] == 0 and plan.subtree_ranges[-1][1] == plan.n_subtrees
        assert all(plan.subtree_ranges[r][1] == plan.subtree_ranges[r + 1][0] for r in range(world - 1))
        covered = 0
        for r in range(world):
            lo, hi = plan.leaf_range(r)
            assert lo == covered and hi >= lo
            covered = hi
        assert covered == n
        assert sum(plan.rank_shift(r) for r in range(world)) == (1 if blank and plan.n_total else 0)
        leaves = random_fr_bytes(max(n, 1), seed=case)[:n]
        root = sharded.emulated_sharded_merge(torch.from_numpy(leaves.copy()), plan, backend)
        rc, exp, depth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, to_depth, leaves)
        assert rc in (0, 2)
        if plan.n_total == 0:
            assert root is None
        else:
            assert root.numpy().tobytes() == exp, (arity, full_depth, n, blank, to_depth, world)
            assert plan.insert_depth == depth
