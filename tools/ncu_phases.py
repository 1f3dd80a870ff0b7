"""ncu report -> where the warp-stall samples of ONE kernel sit, by SASS segment.
Segments are cut at barriers / mbarrier waits / back-edges so that each one is a phase of the kernel; for each
segment: share of samples, instruction mix (DMMA / LDS / LDG / LDGSTS / SHFL / ...), top stall reasons.
usage: python tools/ncu_phases.py <report.ncu-rep> <kernel-regex> [min_share_pct]
"""
import csv
import re
import subprocess
import sys
from collections import Counter


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not starts:
        print("kernel not found")
        return
    hi = starts[0]
    end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
    h = rows[hi]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    print(rows[hi - 1][1] if hi else kern)
    segs, cur = [], None

    def new_seg(idx):
        return {"first": idx, "n": 0, "samples": 0, "ops": Counter(), "stalls": Counter(), "exec": 0, "last": ""}

    cur = new_seg(0)
    for k, r in enumerate(rows[hi + 1:end]):
        if len(r) <= isamp:
            continue
        src = r[isrc].strip()
        try:
            smp, ex = int(r[isamp]), int(r[iex])
        except ValueError:
            continue
        op = re.sub(r"^@!?U?P\d+\s+", "", src).split(" ")[0].split(".")[0]
        cur["n"] += 1
        cur["samples"] += smp
        cur["exec"] = max(cur["exec"], ex)
        cur["ops"][op] += 1
        cur["last"] = src
        for i, c in stall_cols:
            try:
                cur["stalls"][c[6:]] += int(r[i])
            except ValueError:
                pass
        if op in ("BAR", "BRA", "SYNCS", "EXIT", "WARPSYNC", "BSYNC") and (op != "BRA" or True):
            segs.append(cur)
            cur = new_seg(k + 1)
    segs.append(cur)
    tot = sum(s["samples"] for s in segs) or 1
    print("total samples %d, %d SASS instructions, %d segments" % (tot, sum(s["n"] for s in segs), len(segs)))
    acc = 0
    for s in segs:
        share = 100.0 * s["samples"] / tot
        acc += share
        if share < min_share:
            continue
        ops = " ".join("%s:%d" % kv for kv in s["ops"].most_common(6))
        st = " ".join("%s:%.0f%%" % (k, 100.0 * v / max(1, sum(s["stalls"].values()))) for k, v in s["stalls"].most_common(4))
        print("instr %5d..%5d  %5.1f%% (cum %5.1f%%)  exec/instr %-9d ops[%s]  stalls[%s]  ends: %s"
              % (s["first"], s["first"] + s["n"] - 1, share, acc, s["exec"], ops, st, s["last"][:50]))


if __name__ == "__main__":
    main()
