"""First-contact probe run on the B200 box: measured IMAD peaks and raw kernel
timings for a few sizes.  Writes gpurun_out/probe.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib  # noqa: E402
from tests.util import random_fr_bytes  # noqa: E402

out = {}
ctx = ib.get_context(0)
for kind, name in ((0, "imad_lo"), (1, "imad_wide_x2"), (2, "imad_wide_carry_chain_x2"), (3, "imad_hi")):
    v, clk = C.c_double(), C.c_double()
    ctx.check(ctx.lib.inf_measure_imad_peak(ctx.handle, kind, C.byref(v), C.byref(clk)))
    out[name] = {"imad_per_s": v.value, "clock_mhz": clk.value}
    print(name, "%.3f T IMAD/s" % (v.value / 1e12), "clock %.0f MHz" % clk.value)

dev = torch.device("cuda:0")
W = {2: 218592, 5: 731808, 4: 528000, 3: 1288 * 264}
for k, logn in ((2, 20), (2, 22), (2, 24), (5, 20), (5, 22), (4, 20), (3, 20)):
    n = 1 << logn
    src = torch.from_numpy(random_fr_bytes(min(n * k, 1 << 22), seed=k)).to(dev)
    reps = (n * k + src.shape[0] - 1) // src.shape[0]
    d_in = src.repeat(reps, 1)[: n * k].contiguous()
    d_out = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    h = ib.Poseidon.new_circom(k, ctx)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    for _ in range(2):
        h.hash_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        h.hash_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    rate = n / (ms * 1e-3)
    out["hash%d_2^%d" % (k, logn)] = {"ms": ms, "hashes_per_s": rate, "imad_eq_per_s": rate * W[k]}
    print("hash%d n=2^%d: %.2f ms, %.1f M hash/s, %.2f T IMAD-eq/s" % (k, logn, ms, rate / 1e6, rate * W[k] / 1e12))

# tree merges, device resident
for arity, logn, depth in ((2, 20, 20), (2, 24, 24), (5, 20, 9), (5, 24, 11)):
    n = 1 << logn
    src = torch.from_numpy(random_fr_bytes(min(n, 1 << 22), seed=arity)).to(dev)
    d_leaves = src.repeat((n + src.shape[0] - 1) // src.shape[0], 1)[:n].contiguous()
    root = C.create_string_buffer(32)
    idp, rdp, has = C.c_uint32(), C.c_uint32(), C.c_int()
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = ctx.lib.inf_tree_merge_dev(ctx.handle, arity, depth + 1, 0, 1, d_leaves.data_ptr(), n, root,
                                        C.byref(idp), C.byref(rdp), C.byref(has), None)
        ts.append(time.perf_counter() - t0)
        assert rc == 0, rc
    out["tree%d_2^%d" % (arity, logn)] = {"ms": min(ts) * 1e3, "root": root.raw.hex()}
    print("tree arity %d n=2^%d: %.2f ms" % (arity, logn, min(ts) * 1e3))

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
