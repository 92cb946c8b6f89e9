"""SURVEY.md section 8(d) config 5 on one GPU: leaves 2^10 .. 2^28 x {binary t=3, quinary t=6}.

Device-resident leaves, CUDA events on the launching stream, best of a few
runs after one warm-up.  Writes one JSON list in the section-8(d) report
schema to gpurun_out/sweep_n1.json and prints a table.  The two largest
sizes are also checked through a size-independent property: the root of the
whole tree equals the root over the roots of its shards (8-rank plan run back
to back on this device, sharded.emulated_sharded_merge).

    python tools/sweep.py [--max-log 28]
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
from bench import device_random_fr  # noqa: E402

W = {2: 218592, 5: 731808}


def n_hashes(arity, n, depth):
    h, c = 0, n
    for _ in range(depth):
        c = -(-c // arity)
        h += c
    return h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log", type=int, default=28)
    ap.add_argument("--min-log", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/sweep_n1.json")
    args = ap.parse_args()
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
