"""Time the hash kernels of one library build (INFIMUM_B200_LIB) for the hot widths."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib
from tests.util import random_fr_bytes
ctx = ib.get_context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
res = []
for k, logn in ((2, 22), (3, 21), (4, 21), (5, 21)):
    n = 1 << logn
    src = torch.from_numpy(random_fr_bytes(1 << 20, seed=k)).to(dev)
    d_in = src.repeat((n * k + (1 << 20) - 1) >> 20, 1)[: n * k].contiguous()
    d_out = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    h = ib.Poseidon.new_circom(k, ctx)
    for _ in range(2):
        h.hash_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        h.hash_batch_device(d_in.data_ptr(), n, d_out.data_ptr(), stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res.append("hash%d %.1f M/s" % (k, n / ms / 1e3))
print(os.environ.get("INFIMUM_B200_LIB", "default"), " | ".join(res))
