import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib
ctx = ib.get_context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
nm = 1 << 20
g = torch.Generator(device=dev); g.manual_seed(1)
lv = torch.randint(0, 256, (12 * nm, 32), dtype=torch.uint8, device=dev, generator=g)
pk, dat = lv[: 2 * nm], lv[2 * nm:]
ol = torch.empty((nm, 32), dtype=torch.uint8, device=dev)
def run():
    assert ctx.lib.inf_interaction_leaves_dev(ctx.handle, pk.data_ptr(), dat.data_ptr(), nm, ol.data_ptr(), stream.cuda_stream) == 0
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(3): run()
e1.record(stream); torch.cuda.synchronize()
print(os.environ.get("INFIMUM_B200_LIB", "default"), "interaction leaves %.2f M msgs/s" % (nm * 3 / e0.elapsed_time(e1) / 1e3))
