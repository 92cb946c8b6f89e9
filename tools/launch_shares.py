"""Device time per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list:

    python tools/launch_shares.py profiles/r02_launches_bench_steps2.csv"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        name = r[k].split("(")[0].split("::")[-1]
        tot[name] += float(r[v].replace(",", "")) * scale[r[u]]
        cnt[name] += 1
    allms = sum(tot.values())
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for name, ms in tot.most_common():
        print("| `%s` | %d | %.2f | %.1f %% |" % (name, cnt[name], ms, 100 * ms / allms))
    print("| all | %d | %.2f | |" % (sum(cnt.values()), allms))


if __name__ == "__main__":
    main(sys.argv[1])
