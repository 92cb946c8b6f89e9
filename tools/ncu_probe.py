"""One launch of every kernel shape that matters, for `ncu --set full` (profiles/r02_ncu_*.md):

    python tools/ncu_probe.py && ncu --set full --clock-control none --import-source on \
        -k regex:'hash_batch|tree_level|_leaf_kernel' -o gpurun_out/r02_prof python tools/ncu_probe.py

Order of the profiled launches (inf_init's two hash_chain_coop launches are not matched by the filter):
  0 hash_batch_kernel<t=3>        2^22 hash2                      (the headline kernel)
  1 hash_batch_kernel<t=6>        2^20 hash5
  2 tree_level_kernel<t=3>        2^21 parents                    (a bulk level)
  3 tree_level_kernel<t=3>        32 768 parents                  (an under-filled level: 2 warps per sub-partition)
  4 tree_level_coop_kernel<t=3>   1 024 parents                   (K3, one block on 32 SMs)
  5 tree_level_kernel<t=6>        2^19 parents
  6 tree_level_coop_kernel<t=6>   625 parents
  7 interaction_leaf_kernel       2^18 messages
  8 hash_batch_coop_kernel<t=3>   1 hash                          (the O(1) callers)
"""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib
ctx = ib.get_context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
g = torch.Generator(device=dev); g.manual_seed(11)
buf = torch.randint(0, 256, (1 << 23, 32), dtype=torch.uint8, device=dev, generator=g)
buf[:, 0] %= 0x30
out = torch.empty((1 << 22, 32), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
h2, h5 = ib.Poseidon.new_circom(2, ctx), ib.Poseidon.new_circom(5, ctx)
got = C.c_uint64()


def level(arity, n_out):
    rc = ctx.lib.inf_tree_reduce_dev(ctx.handle, arity, 0, 1, 0, buf.data_ptr(), n_out * arity, out.data_ptr(), C.byref(got), stream.cuda_stream)
    assert rc == 0 and got.value == n_out


if len(sys.argv) > 1 and sys.argv[1] == "headline":
    # the bench's timed launch itself: 2^24 hash2 (roofline.traffic = its DRAM bytes, profiles/r02_hash2_traffic.json)
    big_in = torch.cat([buf, buf, buf, buf])
    big_out = torch.empty((1 << 24, 32), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    h2.hash_batch_device(big_in.data_ptr(), 1 << 24, big_out.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    print("ncu_probe headline ok")
    sys.exit(0)
h2.hash_batch_device(buf.data_ptr(), 1 << 22, out.data_ptr(), stream.cuda_stream)
h5.hash_batch_device(buf.data_ptr(), 1 << 20, out.data_ptr(), stream.cuda_stream)
level(2, 1 << 21)
level(2, 32768)
level(2, 1024)
level(5, 1 << 19)
level(5, 625)
nm = 1 << 18
assert ctx.lib.inf_interaction_leaves_dev(ctx.handle, buf.data_ptr(), buf[2 * nm:].data_ptr(), nm, out.data_ptr(), stream.cuda_stream) == 0
h2.hash_batch_device(buf.data_ptr(), 1, out.data_ptr(), stream.cuda_stream)
torch.cuda.synchronize()
print("ncu_probe ok")
