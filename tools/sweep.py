"""SURVEY.md section 8(d) config 5 on one GPU: leaves 2^10 .. 2^28 x {binary t=3, quinary t=6}.

Device-resident leaves, CUDA events on the launching stream, best of a few
runs after warm-up (median of 10 up to 2^24 leaves).  Writes one JSON list in the
section-8(d) report schema to gpurun_out/sweep_n1.json and prints a table: every
row carries `bit_exact` — the root against the oracle (dense tree on all host
threads) up to 2^24 leaves, and for the two largest sizes a size-independent
property instead (the root of the whole tree equals the root over the roots of
its shards, 8-rank plan run back to back on this device) — the SM clock sampled
during the row, both roofline denominators and the CPU rate of the same box.

    python tools/sweep.py [--max-log 28]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py

Under torchrun the same trees (same seeds, each rank generating only its own
leaves) are merged sharded over the N ranks (infimum_b200.sharded: contiguous
subtrees per rank, one NCCL all-gather of subtree roots); time = max over
ranks; `bit_exact` there = the root equals the one-GPU root of
profiles/r02_sweep_n1.json (itself oracle-checked, see `bit_exact_via`).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import infimum_b200 as ib  # noqa: E402
from infimum_b200 import sharded  # noqa: E402
from bench import ClockSampler, device_random_fr, device_random_fr_range  # noqa: E402

W = {2: 218592, 5: 731808}
MODMUL = {2: 535, 5: 1189}          # field multiplications one hash executes (history recurrence); reference: 828 / 2772


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
    root_dir = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("r02_sweep_n1.json", "r01_sweep_n1.json"):
        ref_path = os.path.join(root_dir, "profiles", name)
        if os.path.exists(ref_path):
            ref = {(r["arity"], r["n_leaves"]): r for r in json.load(open(ref_path))}
            break
    prop = torch.cuda.get_device_properties(dev)
    peak_sm = prop.multi_processor_count * 64
    rows = []
    for logn in range(max(args.min_log, 16), args.max_log + 1, 2):
        n = 1 << logn
        for arity in (2, 5):
            depth = logn if arity == 2 else next(d for d in range(40) if 5 ** d >= n)
            plan = sharded.make_plan(arity, depth, n, False, True, world)
            lo, hi = plan.leaf_range(rank)
            lv = device_random_fr_range(lo, hi, dev, seed=500 + logn)
            best, root, times = None, None, []
            sampler = ClockSampler(local)
            if rank == 0:
                sampler.start()
            t_start = time.perf_counter()
            for it in range(11 if logn <= 24 else 4):
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
                    times.append(float(t.item()))
            clocks = sampler.stop(t_start, time.perf_counter()) if rank == 0 else {}
            best = statistics.median(times)
            nh = n_hashes(arity, n, depth)
            root_hex = bytes(root.cpu().numpy().tobytes()).hex()
            r1 = ref.get((arity, n))
            clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
            row = {"config": "sweep", "t": arity + 1, "arity": arity, "n_leaves": n, "full_depth": depth, "n_hashes": nh,
                   "gpus": world, "ms_device": best, "ms_min": min(times), "timing": "median of %d after 1 warm-up, max over ranks" % len(times),
                   "hashes_per_s": nh / (best * 1e-3), "W_imad_per_hash": W[arity],
                   "sm_count": prop.multi_processor_count, "sm_clock_mhz": clocks.get("sm_mhz"),
                   "peak_imad_theoretical": world * peak_sm * clock,
                   "roofline": {"achieved": nh * W[arity] / (best * 1e-3) / (world * peak_sm * clock),
                                "executed_modmul_per_hash": MODMUL[arity]},
                   "shard_plan": {"level": plan.level, "n_subtrees": plan.n_subtrees}, "root": root_hex,
                   "bit_exact": (r1["root"] == root_hex) if r1 else None,
                   "bit_exact_via": ("root == one-GPU root, which is " + str(r1.get("bit_exact_via", "the r01 one-GPU root"))) if r1 else None,
                   "speedup_vs_n1": (r1["ms_device"] / best) if r1 else None}
            assert row["bit_exact"] is not False, (arity, logn)
            rows.append(row)
            if rank == 0:
                print("N=%d arity %d  2^%-2d leaves  %9.3f ms  %7.1f M hashes/s  root %s  same as N=1: %s  x%.2f" % (
                    world, arity, logn, best, nh / best / 1e3, root_hex[:8], row["bit_exact"], row["speedup_vs_n1"] or 0), flush=True)
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
    ap.add_argument("--oracle-max-log", type=int, default=24, help="largest tree whose root the oracle recomputes")
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
    import numpy as np
    from oracle import c_oracle
    from bench import cpu_baseline_hash2
    v, clk = C.c_double(), C.c_double()
    ctx.check(ctx.lib.inf_measure_imad_peak(ctx.handle, 0, C.byref(v), C.byref(clk)))
    peak_measured = v.value
    cpu, _, _ = cpu_baseline_hash2(3.0)
    cores = os.cpu_count() or 1
    d5 = np.random.default_rng(5).integers(0, 256, size=(1 << 16) * 160, dtype=np.uint8)
    d5[::32] %= 0x30
    t0 = time.perf_counter()
    c_oracle.hash_batch(5, d5, threads=cores)
    cpu5 = (1 << 16) / (time.perf_counter() - t0)
    cpu_row = {2: {"variant": "B1 hoisted parameters (oracle/poseidon_oracle.c)", "cores": cores, "hashes_per_s": cpu["value"]},
               5: {"variant": "B1 hoisted parameters (oracle/poseidon_oracle.c)", "cores": cores, "hashes_per_s": cpu5}}
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
            run()
            torch.cuda.synchronize(dev)
            run()
            torch.cuda.synchronize(dev)
            times = []
            sampler = ClockSampler(0)
            sampler.start()
            t_start = time.perf_counter()
            reps = 10 if logn <= 24 else (5 if logn <= 26 else 3)
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                run()
                e1.record(stream)
                torch.cuda.synchronize(dev)
                times.append(e0.elapsed_time(e1))
            if time.perf_counter() - t_start < 0.35:        # give nvidia-smi a sample under load
                while time.perf_counter() - t_start < 0.35:
                    run()
                torch.cuda.synchronize(dev)
            clocks = sampler.stop(t_start, time.perf_counter())
            best = statistics.median(times)
            nh = n_hashes(arity, n, depth)
            clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
            row = {"config": "sweep", "t": arity + 1, "arity": arity, "n_leaves": n, "full_depth": depth,
                   "n_hashes": nh, "gpus": 1, "ms_device": best, "ms_min": min(times),
                   "timing": "median of %d after 3 warm-ups, CUDA events" % reps, "hashes_per_s": nh / (best * 1e-3),
                   "W_imad_per_hash": W[arity], "sm_count": prop.multi_processor_count,
                   "sm_clock_mhz": clocks.get("sm_mhz"), "clock_reasons": clocks.get("reasons"),
                   "peak_imad_theoretical": peak_sm * clock, "peak_imad_measured": peak_measured,
                   "roofline": {"achieved": nh * W[arity] / (best * 1e-3) / (peak_sm * clock),
                                "achieved_vs_measured_peak": nh * W[arity] / (best * 1e-3) / peak_measured,
                                "executed_modmul_per_hash": MODMUL[arity]},
                   "hbm_gbs_achieved": (n + nh) * 32 / (best * 1e-3) / 1e9, "root": root.raw.hex(),
                   "root_depth": b.value, "cpu": cpu_row[arity]}
            if logn <= args.oracle_max_log:
                exp = c_oracle.dense_tree_root(arity, depth, lv.cpu().numpy(), threads=cores)
                row["bit_exact"] = bytes(exp).hex() == root.raw.hex()
                row["bit_exact_via"] = "oracle dense tree over all leaves (oracle/poseidon_oracle.c)"
            else:
                plan = sharded.make_plan(arity, depth, n, False, True, 8)
                r2 = sharded.emulated_sharded_merge(lv, plan, sharded.GpuBackend(ctx, 0))
                row["shard_plan"] = {"level": plan.level, "n_subtrees": plan.n_subtrees}
                row["bit_exact"] = bytes(r2.cpu().numpy().tobytes()).hex() == root.raw.hex()
                row["bit_exact_via"] = "property: root(whole) == root(roots of an 8-rank shard plan); kernels oracle-checked at <= 2^%d leaves" % args.oracle_max_log
            assert row["bit_exact"], (arity, logn)
            rows.append(row)
            print("arity %d  2^%-2d leaves  depth %2d  %10d hashes  %9.3f ms  %6.1f M hashes/s  %5.1f %% roofline  %s MHz  bit-exact: %s" % (
                arity, logn, depth, nh, best, nh / best / 1e3, 100 * row["roofline"]["achieved"], clocks.get("sm_mhz"),
                row["bit_exact_via"][:24]), flush=True)
        del lv
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
