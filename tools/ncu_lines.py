"""(-gi: instructions of inlined functions are attributed to the outermost call-site line.)
Join an ncu SASS-level source page (samples per instruction) with nvdisasm line info so that
stall samples can be read per CUDA source line.  Usage:
  python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> <cubin> [top]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def line_map(cubin, kernel):
    out = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    m, cur, infn = {}, None, False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            infn = kernel in ln
        f = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if f:
            cur = (f.group(1).split("/")[-1], int(f.group(2)))
        a = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if a and infn:
            m[int(a.group(1), 16)] = (cur, a.group(2).strip())
    return m


def main():
    rep, kernel, cubin = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    ia, isamp, isrc = h.index("Address"), h.index("# Samples"), h.index("Source")
    iex = h.index("Instructions Executed")
    lm = line_map(cubin, kernel)
    base = None
    per_line, per_line_inst, tot = defaultdict(int), defaultdict(int), 0
    for r in rows[hi + 1:]:
        try:
            addr, s, ex = int(r[ia], 16), int(r[isamp]), int(r[iex])
        except Exception:
            continue
        if base is None:
            base = addr
        key = lm.get(addr - base, (("?", 0), ""))[0]
        per_line[key] += s
        per_line_inst[key] += ex
        tot += s
    print("total samples", tot)
    for key, s in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
        print("%7d %5.1f%%  inst %10d  %s:%s" % (s, 100.0 * s / max(tot, 1), per_line_inst[key], key[0], key[1]))


if __name__ == "__main__":
    main()
