"""inf_multi_*: several GPUs from one process.  On a one-GPU box the sharding
logic is exercised with the same device listed several times and peer copies
standing in for the collective; with two or more GPUs the NCCL path runs."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests.util import random_fr_bytes

pytestmark = pytest.mark.gpu

CASES = [(2, 20, 100000, True, False), (2, 16, 1 << 16, False, True), (2, 21, (1 << 15) + 1, True, False),
         (5, 9, 100000, False, True), (5, 6, 5 ** 6 - 3, False, True), (2, 12, 3, True, False), (5, 4, 1, False, True),
         (2, 12, 0, True, False), (5, 3, 0, False, True)]


def _check(mg):
    for arity, full_depth, n, blank, to_depth in CASES:
        leaves = random_fr_bytes(max(n, 1), seed=n % 1000 + arity)[:n]
        root, depth, rdepth, rc = mg.tree_merge(arity, full_depth, leaves, blank, to_depth)
        orc, exp, odepth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, to_depth, leaves)
        assert rc == orc and root == exp and depth == odepth, (arity, n, mg.devices)
    raw = random_fr_bytes(2 * 70001, seed=9, canonical=False)
    assert (mg.hash_batch(2, raw) == c_oracle.hash_batch(2, raw)).all()


@pytest.mark.parametrize("n_dev", [1, 2, 3, 8])
def test_sharding_on_one_gpu_with_peer_copies(n_dev):
    from infimum_b200.multi import MultiGpu
    mg = MultiGpu([0] * n_dev, peer_copy=True)
    _check(mg)
    mg.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_all_gather_path():
    from infimum_b200.multi import MultiGpu
    mg = MultiGpu(list(range(min(torch.cuda.device_count(), 8))))
    _check(mg)
    mg.close()
