"""Time device-resident tree merges of a few sizes (one library build / env).

    [INF_COOP_MAX=n] [INF_NO_PDL=1] python tools/tree_probe.py [tag]
"""
import ctypes as C, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib
ctx = ib.get_context(0)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
res = {}
CASES = ((2, 10, 32, True, False), (2, 13, 32, True, False), (2, 16, 32, True, False), (2, 18, 32, True, False),
         (2, 20, 32, True, False), (2, 22, 22, False, True), (2, 24, 24, False, True),
         (5, 13, 6, False, True), (5, 16, 7, False, True), (5, 20, 9, False, True), (5, 24, 11, False, True))
for arity, logn, depth, blank, to_depth in CASES:
    n = 1 << logn
    lv = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=g)
    lv[:, 0] %= 0x30
    root = C.create_string_buffer(32); a, b, h = C.c_uint32(), C.c_uint32(), C.c_int()
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rc = ctx.lib.inf_tree_merge_dev(ctx.handle, arity, depth, int(blank), int(to_depth), lv.data_ptr(), n, root, C.byref(a), C.byref(b), C.byref(h), None)
        best = min(best, time.perf_counter() - t0); assert rc in (0, 2)
    res["a%d_2^%d" % (arity, logn)] = round(best * 1e3, 3)
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
print(json.dumps({"tag": tag, "coop_max": os.environ.get("INF_COOP_MAX", "default"), "no_pdl": os.environ.get("INF_NO_PDL", "0"), "ms": res, "root": root.raw.hex()[:16]}))
