"""SURVEY.md section 8(d) config 5 on one GPU: leaves 2^10 .. 2^28 x {binary t=3, quinary t=6}.

Device-resident leaves, CUDA events on the launching stream, best of a few
runs after one warm-up.  Writes one JSON list in the section-8(d) report
schema to gpurun_out/sweep_n1.json and prints a table.  The two largest
sizes are also checked through a size-independent property: the root of the
whole tree equals the root over the roots of its shards (8-rank plan run back
to back on this device, sharded.emulated_sharded_merge).

    python tools/sweep.py [--max-log 28]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py

Under torchrun the same trees (same seeds, each rank generating only its own
leaves) are merged sharded over the N ranks (infimum_b200.sharded: contiguous
subtrees per rank, one NCCL all-gather of subtree roots); time = max over
ranks; the roots are compared with profiles/r01_sweep_n1.json when present.
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib  # noqa: E402
from infimum_b200 import sharded  # noqa: E402
from bench import device_random_fr, device_random_fr_range  # noqa: E402

W = {2: 218592, 5: 731808}


def n_hashes(arity, n, depth):
    h, c = 0, n
    for _ in range(depth):
        c = -(-c // arity)
        h += c
    return h


def main_sharded(args, world):
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ctx = ib.get_context(local)
    backend = sharded.GpuBackend(ctx)
    stream = torch.cuda.Stream(dev)
    ref = {}
    ref_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_sweep_n1.json")
    if os.path.exists(ref_path):
        ref = {(r["arity"], r["n_leaves"]): r["root"] for r in json.load(open(ref_path))}
    rows = []
    for logn in range(max(args.min_log, 16), args.max_log + 1, 2):
        n = 1 << logn
        for arity in (2, 5):
            depth = logn if arity == 2 else next(d for d in range(40) if 5 ** d >= n)
            plan = sharded.make_plan(arity, depth, n, False, True, world)
            lo, hi = plan.leaf_range(rank)
            lv = device_random_fr_range(lo, hi, dev, seed=500 + logn)
            best, root = None, None
            for it in range(4 if logn <= 24 else 3):
                torch.cuda.synchronize(dev)
                dist.barrier()
                with torch.cuda.stream(stream):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    root = sharded.sharded_tree_merge(lv, plan, backend)
                    e1.record(stream)
                torch.cuda.synchronize(dev)
                t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if it > 0:
                    best = float(t.item()) if best is None else min(best, float(t.item()))
            nh = n_hashes(arity, n, depth)
            root_hex = bytes(root.cpu().numpy().tobytes()).hex()
            row = {"config": "sweep", "t": arity + 1, "arity": arity, "n_leaves": n, "full_depth": depth, "n_hashes": nh,
                   "gpus": world, "ms_device": best, "hashes_per_s": nh / (best * 1e-3), "W_imad_per_hash": W[arity],
                   "shard_plan": {"level": plan.level, "n_subtrees": plan.n_subtrees}, "root": root_hex,
                   "bit_exact_vs_n1": (ref[(arity, n)] == root_hex) if (arity, n) in ref else None}
            assert row["bit_exact_vs_n1"] is not False, (arity, logn)
            rows.append(row)
            if rank == 0:
                print("N=%d arity %d  2^%-2d leaves  %9.3f ms  %7.1f M hashes/s  root %s  same as N=1: %s" % (
                    world, arity, logn, best, nh / best / 1e3, root_hex[:8], row["bit_exact_vs_n1"]), flush=True)
            del lv
            torch.cuda.empty_cache()
    if rank == 0:
        out = args.out.replace("_n1", "_n%d" % world)
        os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
        json.dump(rows, open(out, "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log", type=int, default=28)
    ap.add_argument("--min-log", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/sweep_n1.json")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return main_sharded(args, world)
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    ctx = ib.get_context(0)
    stream = torch.cuda.Stream(dev)
    prop = torch.cuda.get_device_properties(dev)
    peak_sm = prop.multi_processor_count * 64
    rows = []
    for logn in range(args.min_log, args.max_log + 1, 2):
        n = 1 << logn
        lv = device_random_fr(n, dev, seed=500 + logn)
        for arity in (2, 5):
            depth = logn if arity == 2 else next(d for d in range(40) if 5 ** d >= n)
            root = C.create_string_buffer(32)
            a, b, h = C.c_uint32(), C.c_uint32(), C.c_int()

            def run():
                rc = ctx.lib.inf_tree_merge_dev(ctx.handle, arity, depth, 0, 1, lv.data_ptr(), n, root, C.byref(a),
                                                C.byref(b), C.byref(h), stream.cuda_stream)
                assert rc in (0, 2), rc

            torch.cuda.synchronize(dev)      # leaves were generated on torch's default stream
            run()
            torch.cuda.synchronize(dev)
            best = None
            for _ in range(5 if logn <= 24 else 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                run()
                e1.record(stream)
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            nh = n_hashes(arity, n, depth)
            clock = 1.965e9
            row = {"config": "sweep", "t": arity + 1, "arity": arity, "n_leaves": n, "full_depth": depth,
                   "n_hashes": nh, "gpus": 1, "ms_device": best, "hashes_per_s": nh / (best * 1e-3),
                   "W_imad_per_hash": W[arity], "sm_count": prop.multi_processor_count,
                   "peak_imad_theoretical": peak_sm * clock,
                   "roofline": {"achieved": nh * W[arity] / (best * 1e-3) / (peak_sm * clock)},
                   "hbm_gbs_achieved": (n + nh) * 32 / (best * 1e-3) / 1e9, "root": root.raw.hex(),
                   "root_depth": b.value}
            if logn >= 26:
                plan = sharded.make_plan(arity, depth, n, False, True, 8)
                r2 = sharded.emulated_sharded_merge(lv, plan, sharded.GpuBackend(ctx, 0))
                row["shard_plan"] = {"level": plan.level, "n_subtrees": plan.n_subtrees}
                row["bit_exact_vs_sharded"] = bytes(r2.cpu().numpy().tobytes()).hex() == root.raw.hex()
                assert row["bit_exact_vs_sharded"], (arity, logn)
            rows.append(row)
            print("arity %d  2^%-2d leaves  depth %2d  %10d hashes  %9.3f ms  %6.1f M hashes/s  %5.1f %% roofline%s" % (
                arity, logn, depth, nh, best, nh / best / 1e3, 100 * row["roofline"]["achieved"],
                "  sharded-root ok" if row.get("bit_exact_vs_sharded") else ""), flush=True)
        del lv
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
