"""Dynamic multiply-pipe instruction count of one hash from the SASS of a kernel.

    python tools/sass_count.py infimum_b200/_build/poseidon_t3.o hash_batch_kernelILb0 4,28,3

Loops are found from backward branches; the trip counts are given in program
order (for the hash kernels: first-half full rounds, pairs of partial rounds,
second-half full rounds; see poseidon.cuh).  Everything outside a loop counts
once.  Prints totals per class: IMAD.WIDE(.X), IMAD.HI, plain IMAD (multiplying
forms only), everything else.
"""
import re
import subprocess
import sys


def count(obj, fn, trips, verbose=False):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    body, on = [], False
    for line in txt.splitlines():
        if "Function :" in line:
            on = fn in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                body.append((int(m.group(1), 16), m.group(2)))
    addr_idx = {a: i for i, (a, _) in enumerate(body)}
    loops = []
    for i, (a, ins) in enumerate(body):
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) in addr_idx and addr_idx[int(m.group(1), 16)] < i:       # (the trailing self-branch after EXIT is padding)
            loops.append((addr_idx[int(m.group(1), 16)], i))
    loops.sort()
    assert len(loops) == len(trips), (loops, trips)
    mult = [1] * len(body)
    for (lo, hi), t in zip(loops, trips):
        for i in range(lo, hi + 1):
            mult[i] = t

    def cls(ins):
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        if op.startswith("IMAD.WIDE"):
            return "wide"
        if op.startswith("IMAD.HI"):
            return "hi"
        if op in ("IMAD", "IMAD.U32") or op.startswith("IMAD.LO"):
            return "imad" if not re.search(r"\bRZ\b, \bRZ\b", ins) else "other"
        return "other"

    tot = {"wide": 0, "hi": 0, "imad": 0, "other": 0}
    for (a, ins), k in zip(body, mult):
        tot[cls(ins)] += k
    if verbose:
        print(fn, "loops", [(hi - lo + 1, t) for (lo, hi), t in zip(loops, trips)], tot, "all", sum(tot.values()))
    return tot


if __name__ == "__main__":
    count(sys.argv[1], sys.argv[2], [int(x) for x in sys.argv[3].split(",")], verbose=True)
