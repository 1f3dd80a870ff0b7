"""ncu report(s) -> profiles/r02_dram_bytes.json: dram__bytes_read + write per launch of each kernel, per camera stream,
keyed by BASELINE config (bench.py multiplies by the streams of the run; a config without a capture has no entry).
usage: python tools/ncu_traffic.py <config>=<report.ncu-rep>[:streams] ... > profiles/r02_dram_bytes.json"""
import csv
import json
import subprocess
import sys

NAMES = {"upd_hp_kernel": "upd_hp", "upd_hp2_kernel": "upd_hp", "upd_chol_kernel": "upd_chol", "upd_solve_kernel": "upd_solve",
         "upd_syrk_kernel": "upd_syrk", "upd_finish_kernel": "upd_finish", "search_kernel": "search",
         "predict_kernel": "predict", "cull_kernel": "cull"}
out = {}
for arg in sys.argv[1:]:
    cfg, rest = arg.split("=")
    rep, _, streams = rest.partition(":")
    streams = int(streams or 296)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    ik, ir, iw, it = (h.index(c) for c in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                            "gpu__time_duration.sum"))
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        key = next((v for k, v in NAMES.items() if k in r[ik]), None)
        if key is None:
            continue
        b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        acc.setdefault(key, []).append(b)
    out[cfg] = {k: {"bytes_per_stream": sum(v) / len(v) / streams, "launches_captured": len(v),
                    "streams_in_capture": streams, "report": rep.split("/")[-1]} for k, v in acc.items()}
print(json.dumps(out, indent=1))
