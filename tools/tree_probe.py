"""Time device-resident tree merges of a few sizes (one library build / env)."""
import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib
ctx = ib.get_context(0)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
res = []
for arity, logn, depth, blank, to_depth in ((2, 10, 32, True, False), (2, 16, 32, True, False), (2, 20, 32, True, False), (2, 24, 24, False, True), (5, 20, 9, False, True), (5, 24, 11, False, True)):
    n = 1 << logn
    lv = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    lv[:, 0] %= 0x30
    root = C.create_string_buffer(32); a, b, h = C.c_uint32(), C.c_uint32(), C.c_int()
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = ctx.lib.inf_tree_merge_dev(ctx.handle, arity, depth, int(blank), int(to_depth), lv.data_ptr(), n, root, C.byref(a), C.byref(b), C.byref(h), None)
        best = min(best, time.perf_counter() - t0); assert rc in (0, 2)
    res.append("a%d 2^%d %.3f ms %s" % (arity, logn, best * 1e3, root.raw.hex()[:8]))
print(os.environ.get("INF_COOP_MAX", "default"), " | ".join(res))
