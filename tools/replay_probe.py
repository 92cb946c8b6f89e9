"""Replay raw messages -> root from pinned host memory vs (leaf kernel + tree merge) on device-resident data.

    [INF_REPLAY_CHUNK_LOG=k] [INF_NO_RAMP=1] python tools/replay_probe.py [tag]
"""
import ctypes as C, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib
from bench import device_random_fr
ctx = ib.get_context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
res = {}
for log_m in (20, 22):
    m = 1 << log_m
    blk = device_random_fr(12 * m, dev, seed=900)
    pk, dat = blk[: 2 * m], blk[2 * m:]
    ol = torch.empty((m, 32), dtype=torch.uint8, device=dev)
    depth = next(d for d in range(40) if 5 ** d >= m)
    root = C.create_string_buffer(32); a, b, h = C.c_uint32(), C.c_uint32(), C.c_int()

    def dev_run():
        assert ctx.lib.inf_interaction_leaves_dev(ctx.handle, pk.data_ptr(), dat.data_ptr(), m, ol.data_ptr(), stream.cuda_stream) == 0
        rc = ctx.lib.inf_tree_merge_dev(ctx.handle, 5, depth, 0, 1, ol.data_ptr(), m, root, C.byref(a), C.byref(b), C.byref(h), stream.cuda_stream)
        assert rc in (0, 2)
    best_dev = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dev_run(); torch.cuda.synchronize()
        best_dev = min(best_dev, time.perf_counter() - t0)
    dev_root = root.raw.hex()
    h_pk = torch.empty((m, 64), dtype=torch.uint8).pin_memory(); h_pk.copy_(pk.view(m, 64))
    h_dat = torch.empty((m, 320), dtype=torch.uint8).pin_memory(); h_dat.copy_(dat.view(m, 320))
    a_pk, a_dat = h_pk.numpy(), h_dat.numpy()
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        t, ep, et, _, _ = ib.replay_interactions(depth, a_pk, a_dat, 0, 2, 1, ctx)
        best = min(best, time.perf_counter() - t0)
    assert t.root.hex() == dev_root
    res["2^%d" % log_m] = {"replay_ms": round(best * 1e3, 3), "leaf_plus_tree_ms": round(best_dev * 1e3, 3), "over": round(best / best_dev, 4)}
    del blk, ol, h_pk, h_dat
print(json.dumps({"tag": sys.argv[1] if len(sys.argv) > 1 else "", "chunk_log": os.environ.get("INF_REPLAY_CHUNK_LOG", "15"),
                  "no_ramp": os.environ.get("INF_NO_RAMP", "0"), "res": res}))
