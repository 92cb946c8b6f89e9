import time, sys, os
sys.path.insert(0, os.getcwd())
import torch, infimum_b200 as ib
ctx = ib.get_context(0)
ts = []
for i in range(6):
    t0 = time.perf_counter(); c = ib.Context(0); ts.append((time.perf_counter() - t0) * 1e3); c.close()
print("inf_init ms", [round(x, 2) for x in ts])
