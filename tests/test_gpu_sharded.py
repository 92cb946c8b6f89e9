"""Sharded tree merge on the GPU data plane (GpuBackend over inf_tree_reduce_dev)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests.util import random_fr_bytes

pytestmark = pytest.mark.gpu

CASES = [(2, 20, 100000, True, False), (2, 17, 1 << 17, False, True), (2, 21, (1 << 16) + 1, True, False),
         (5, 9, 200000, False, True), (5, 7, 5 ** 7 - 3, False, True), (2, 12, 3, True, False), (5, 4, 1, False, True)]


@pytest.mark.parametrize("world", [1, 2, 8])
def test_emulated_world_on_one_gpu(world):
    from infimum_b200 import sharded
    backend = sharded.GpuBackend(device=0)
    for arity, full_depth, n, blank, to_depth in CASES:
        leaves = random_fr_bytes(n, seed=n % 1000 + arity)
        plan = sharded.make_plan(arity, full_depth, n, blank, to_depth, world)
        root = sharded.emulated_sharded_merge(torch.from_numpy(leaves).cuda(), plan, backend)
        rc, exp, depth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, to_depth, leaves)
        assert rc in (0, 2) and root.cpu().numpy().tobytes() == exp, (arity, n, world)   # rc 2: completed by insert
        assert plan.insert_depth == depth



def test_maximum_size_2_28_leaves():
    """SURVEY.md 8(d) config 5's largest size (8 GiB of leaves, 64-bit indexing on
    every level).  (1) Binary tree over [x]*2^28, merged to full depth, equals 28
    chained self-hashes from the oracle.  (2) Quinary tree (depth 13) over 2^28
    distinct leaves: the root of the whole equals the root over its 8-rank
    shards' roots."""
    import ctypes as C
    import infimum_b200 as ib
    from infimum_b200 import sharded
    ctx = ib.get_context(0)
    n = 1 << 28
    x = random_fr_bytes(1, seed=28)
    lv = torch.from_numpy(x).cuda().repeat(n, 1)
    root = C.create_string_buffer(32)
    a, b, h = C.c_uint32(), C.c_uint32(), C.c_int()

    def merge(arity, depth):
        torch.cuda.synchronize()        # the leaves were written on torch's stream, the merge runs on the context's
        rc = ctx.lib.inf_tree_merge_dev(ctx.handle, arity, depth, 0, 1, lv.data_ptr(), n, root, C.byref(a), C.byref(b),
                                        C.byref(h), None)
        assert rc in (0, 2), rc
        return root.raw

    node = x[0].tobytes()
    for _ in range(28):
        node = c_oracle.hash_one([node, node])
    assert merge(2, 28) == node and b.value == 28
    # distinct leaves: a counter in the low 8 bytes of every leaf
    idx = torch.arange(n, dtype=torch.int64, device="cuda")
    for k in range(8):
        lv[:, 31 - k] = ((idx >> (8 * k)) & 0xFF).to(torch.uint8)
    del idx
    whole = merge(5, 13)
    plan = sharded.make_plan(5, 13, n, False, True, 8)
    assert plan.n_subtrees >= 32
    got = sharded.emulated_sharded_merge(lv, plan, sharded.GpuBackend(ctx, 0))
    assert got.cpu().numpy().tobytes() == whole
    del lv
    torch.cuda.empty_cache()


WORKER = r"""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["INF_ROOT"])
from infimum_b200 import sharded
from tests.util import random_fr_bytes
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
backend = sharded.GpuBackend(device=rank)
out = []
for arity, full_depth, n, blank, to_depth in json.loads(os.environ["INF_CASES"]):
    plan = sharded.make_plan(arity, full_depth, n, blank, to_depth, world)
    leaves = random_fr_bytes(n, seed=n % 1000 + arity)
    lo, hi = plan.leaf_range(rank)
    root = sharded.sharded_tree_merge(torch.from_numpy(leaves[lo:hi].copy()).cuda(), plan, backend)
    out.append(root.cpu().numpy().tobytes().hex())
if rank == 0:
    print("ROOTS " + json.dumps(out))
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_two_ranks(tmp_path):
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, INF_ROOT=root, INF_CASES=json.dumps(CASES))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("ROOTS ")][0]
    roots = json.loads(line[6:])
    for (arity, full_depth, n, blank, to_depth), got in zip(CASES, roots):
        leaves = random_fr_bytes(n, seed=n % 1000 + arity)
        rc, exp, depth, count = c_oracle.tree_insert_merge(arity, full_depth, blank, to_depth, leaves)
        assert got == exp.hex()
