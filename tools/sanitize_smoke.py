"""Small invocation of every device entry point, sized for compute-sanitizer
(memcheck): `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib  # noqa: E402
from infimum_b200.multi import MultiGpu  # noqa: E402
from oracle import c_oracle  # noqa: E402
from tests.util import random_fr_bytes  # noqa: E402

ctx = ib.get_context(0)
for k in (1, 2, 3, 4, 5, 7, 9, 12):
    raw = random_fr_bytes(257 * k, seed=k, canonical=False)
    h = ib.Poseidon.new_circom(k, ctx)
    assert (h.hash_batch(raw) == c_oracle.hash_batch(k, raw)).all()
    assert (h.hash_batch(raw, dense=True) == c_oracle.hash_batch(k, raw)).all()
for arity, depth, blank, to_depth, n in ((2, 12, True, False, 1001), (5, 5, False, True, 777), (2, 9, False, True, 512)):
    lv = random_fr_bytes(n, seed=n)
    t = ib.PollStateTree.new(arity, depth, (0, ib.get_merkle_zeroes(arity)[0]) if blank else None, ctx).extend(lv)
    fr = t.hashes
    if t.root is None:
        t.merge(to_depth)
    rc, root, d, c = c_oracle.tree_insert_merge(arity, depth, blank, to_depth, lv)
    assert t.root == root and t.depth == d, (arity, n)
    rt = ib.RetainedTree(arity, depth, lv, prepend_blank_leaf=blank, ctx=ctx)
    idx = np.arange(0, n, 37, dtype=np.uint64)
    paths = rt.paths(idx)
    logical = np.concatenate([np.frombuffer(ib.get_merkle_zeroes(arity)[0], dtype=np.uint8).reshape(1, 32), lv]) if blank else lv
    roots = ib.merkle_roots_from_paths(arity, depth, idx, logical[idx.astype(np.int64)], paths, ctx)
    assert all(roots[i].tobytes() == rt.root for i in range(len(idx)))
    rt.close()
pk = random_fr_bytes(2 * 300, seed=1).reshape(300, 64)
data = random_fr_bytes(10 * 300, seed=2).reshape(300, 320)
assert (ib.interaction_leaves(pk, data, ctx) == c_oracle.interaction_leaves(pk, data)).all()
assert (ib.registration_leaves(pk, np.arange(300, dtype=np.uint64), ctx) ==
        c_oracle.registration_leaves(pk, np.arange(300, dtype=np.uint64))).all()
mg = MultiGpu([0, 0, 0], peer_copy=True)
lv = random_fr_bytes(5000, seed=3)
root, d, rd, rc = mg.tree_merge(2, 14, lv, True, False)
assert root == c_oracle.tree_insert_merge(2, 14, True, False, lv)[1]
mg.close()
print("sanitize smoke ok")
