"""Block-shape sweep of the per-thread kernels (runtime overrides: INF_WIDE_BLOCK, INF_WIDE_PAD, INF_LEAF_BLOCK).

    INF_WIDE_BLOCK=384 INF_LEAF_BLOCK=384 python tools/geom_probe.py tag
Prints one JSON line: hash2..hash5 batch rates (2^22 / 2^21 hashes, inputs larger than L2), one bulk tree
level per arity, the fused interaction-leaf kernel, and the 2^24 / 2^20 trees.
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
buf = device_random_fr(12 << 20, dev, seed=3)          # 384 MB, larger than L2
out = torch.empty((1 << 22, 32), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()


def timed(fn, reps=3):
    fn(); fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for k, logn in ((2, 22), (3, 21), (4, 21), (5, 21)):
    n = 1 << logn
    h = ib.Poseidon.new_circom(k, ctx)
    ms = timed(lambda: h.hash_batch_device(buf.data_ptr(), n, out.data_ptr(), stream.cuda_stream))
    res["hash%d_M/s" % k] = round(n / ms / 1e3, 2)
got = C.c_uint64()
for arity, n_out in ((2, 1 << 22), (5, 1 << 21), (2, 1 << 19), (5, 1 << 18)):
    def lvl():
        rc = ctx.lib.inf_tree_reduce_dev(ctx.handle, arity, 0, 1, 0, buf.data_ptr(), n_out * arity, out.data_ptr(), C.byref(got), stream.cuda_stream)
        assert rc == 0
    res["level_a%d_%d_M/s" % (arity, n_out)] = round(n_out / timed(lvl) / 1e3, 2)
nm = 1 << 20
res["leaves_M/s"] = round(nm / timed(lambda: ctx.lib.inf_interaction_leaves_dev(ctx.handle, buf.data_ptr(), buf[2 * nm:].data_ptr(), nm, out.data_ptr(), stream.cuda_stream)) / 1e3, 2)
root = C.create_string_buffer(32); a, b, hh = C.c_uint32(), C.c_uint32(), C.c_int()
lv = device_random_fr(1 << 24, dev, seed=77)
for arity, logn, depth in ((2, 24, 24), (2, 20, 20), (5, 24, 11), (5, 20, 9)):
    n = 1 << logn
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = ctx.lib.inf_tree_merge_dev(ctx.handle, arity, depth, 0, 1, lv.data_ptr(), n, root, C.byref(a), C.byref(b), C.byref(hh), None)
        best = min(best, time.perf_counter() - t0); assert rc in (0, 2)
    res["tree_a%d_2^%d_ms" % (arity, logn)] = round(best * 1e3, 3)
    res["root_a%d_2^%d" % (arity, logn)] = root.raw.hex()[:12]
print(json.dumps({"tag": sys.argv[1] if len(sys.argv) > 1 else "", "wide": os.environ.get("INF_WIDE_BLOCK", "default"),
                  "pad": os.environ.get("INF_WIDE_PAD", "default"), "leaf": os.environ.get("INF_LEAF_BLOCK", "default"), "res": res}))
