"""Markdown table of the metrics profiles/r02_ncu_kernels.md quotes, from an `ncu --page raw --csv` export:

    python tools/ncu_table.py profiles/r02_prof_raw.csv

One column per profiled launch, in launch order (tools/ncu_probe.py lists them)."""
import csv
import sys

ROWS = [
    ("duration ms", "gpu__time_duration.sum", "ms"),
    ("grid x block", None, None),
    ("registers / thread", "launch__registers_per_thread", None),
    ("local-memory instructions", None, None),
    ("multiply pipe busy, % of elapsed (`sm__pipe_fmaheavy_cycles_active`)", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", None),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active", None),
    ("resident warps per SM", "sm__warps_active.avg.per_cycle_active", None),
    ("instruction-cache hit rate % (`sm__icc_request_hit_rate`)", "sm__icc_request_hit_rate.pct", None),
    ("DRAM read MB", "dram__bytes_read.sum", "MB"),
    ("DRAM written MB", "dram__bytes_write.sum", "MB"),
    ("stall per issue: wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", None),
    ("math pipe throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", None),
    ("not selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", None),
    ("no instruction", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", None),
    ("dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", None),
    ("barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", None),
    ("short scoreboard (shared memory)", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", None),
]
SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def load(path):
    r = list(csv.reader(open(path)))
    hdr, units, rows = r[0], r[1], r[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(row, name, want=None):
        x = float(row[col[name]].replace(",", "") or 0)
        if want:
            x *= SCALE[units[col[name]]]
        return x
    return hdr, rows, col, val


def main(path):
    hdr, rows, col, val = load(path)
    print("| metric | " + " | ".join(str(i) for i in range(len(rows))) + " |")
    print("|---|" + "---|" * len(rows))
    print("| kernel | " + " | ".join("`%s`" % row[col["Kernel Name"]].split("(")[0].split("::")[-1][:40] for row in rows) + " |")
    for label, name, unit in ROWS:
        cells = []
        for row in rows:
            if label == "grid x block":
                cells.append("%d x %d" % (val(row, "launch__grid_size"), val(row, "launch__block_size")))
            elif label == "local-memory instructions":
                cells.append("%d" % (val(row, "sass__inst_executed_local_loads") + val(row, "sass__inst_executed_local_stores")))
            else:
                x = val(row, name, unit)
                cells.append(("%.3f" if label == "duration ms" else "%.0f" if name.startswith("launch") else "%.2f") % x)
        print("| " + label + " | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
