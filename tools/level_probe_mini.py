"""tools/level_probe.py over a handful of sizes around the K2 / K3 crossover (argv: tag)."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib
ctx = ib.get_context(0)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(5)
stream = torch.cuda.Stream()
res = {}
for arity, sizes in ((2, (4096, 8192, 12288, 16384, 24576)), (5, (4096, 6144, 8192, 12288))):
    for n_out in sizes:
        n_in = n_out * arity
        src = torch.randint(0, 256, (n_in, 32), dtype=torch.uint8, device=dev, generator=g)
        src[:, 0] %= 0x30
        dst = torch.empty((n_out, 32), dtype=torch.uint8, device=dev)
        got = C.c_uint64()
        def run():
            rc = ctx.lib.inf_tree_reduce_dev(ctx.handle, arity, 0, 1, 0, src.data_ptr(), n_in, dst.data_ptr(), C.byref(got), stream.cuda_stream)
            assert rc == 0 and got.value == n_out
        run()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream); run(); e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res["a%d_%d" % (arity, n_out)] = round(best * 1e3, 1)
print(json.dumps({"tag": sys.argv[1] if len(sys.argv) > 1 else "", "coop_max": os.environ.get("INF_COOP_MAX", "default"), "us": res}))
