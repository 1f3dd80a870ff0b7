"""Per-instruction view of an ncu report for a range of CUDA source lines (inlined code is attributed to the
outermost call-site line, nvdisasm -gi).

  python tools/ncu_sass.py dump  <report.ncu-rep> <cubin> <kernel-substring> <line_lo> <line_hi> [top]
      every SASS instruction of those lines with its samples and its two dominant stall reasons
      (top = only the N most-sampled instructions)
  python tools/ncu_sass.py stalls <report.ncu-rep> <cubin> <kernel-substring> <line_lo> <line_hi>
      stall-reason totals and the opcodes carrying the samples for those lines

cubin: `cuobjdump -xelf all scenelib2_b200/libsl2b200.so` of the SAME build that was profiled.
"""
import collections
import csv
import re
import subprocess
import sys


def load(rep, cubin, kernel):
    out = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    m, cur, infn = {}, None, False
    for ln in out.splitlines():
        if ln.startswith(".text."):
            infn = kernel in ln
        f = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if f:
            cur = int(f.group(2))
        a = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if a and infn:
            m[int(a.group(1), 16)] = (cur, a.group(2).strip())
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    return m, rows[hi], rows[hi + 1:]


def main():
    mode, rep, cubin, kernel, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
    top = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    m, h, rows = load(rep, cubin, kernel)
    ia, isamp, iex = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    base, out = None, []
    tot, opc, n = collections.Counter(), collections.Counter(), 0
    for r in rows:
        try:
            addr, s, ex = int(r[ia], 16), int(r[isamp]), int(r[iex])
        except Exception:
            continue
        if base is None:
            base = addr
        line, ins = m.get(addr - base, (0, "?"))
        if not (lo <= line <= hi):
            continue
        n += s
        for i in stall_cols:
            tot[h[i][6:]] += int(r[i] or 0)
        op = ins.split()
        opc[op[1] if ins.startswith("@") and len(op) > 1 else (op[0] if op else "?")] += s
        best = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
        out.append((s, "%6x L%-5d s=%-6d ex=%-9d %-58s %s" % (addr - base, line, s, ex, ins[:58],
                                                             " ".join("%s=%d" % (nm, v) for v, nm in best if v))))
    if mode == "stalls":
        print("samples", n)
        print(" ".join("%s=%d" % kv for kv in tot.most_common(10)))
        print(" ".join("%s=%d" % kv for kv in opc.most_common(12)))
    else:
        if top:
            out = sorted(out, reverse=True)[:top]
        for _, ln in out:
            print(ln)


if __name__ == "__main__":
    main()
