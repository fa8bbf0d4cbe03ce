"""Per-source-line stall summary of an ncu report (run where ncu is installed, no GPU needed):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]
Uses `ncu --page source --print-source cuda,sass --csv`; lines of CUDA source carry the summed samples of
their SASS instructions."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
cur, hdr, data = None, None, []
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
    elif r[0] == "Line No":
        hdr = r
    elif hdr and cur and r[0].isdigit() and len(r) == len(hdr):
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)           # "Source" appears twice (CUDA text, SASS text): keep the CUDA text
        d["file"] = cur
        data.append(d)
tot = sum(int(d["# Samples"] or 0) for d in data)
print("total samples", tot)
agg = {}
for d in data:
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"):
            agg[k] = agg.get(k, 0) + int(v)
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for d in sorted(data, key=lambda d: -int(d["# Samples"] or 0))[:top_n]:
    s = int(d["# Samples"])
    st = sorted(((k[6:], int(v)) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")),
                key=lambda kv: -kv[1])[:3]
    print(f"{d['file']}:{d['Line No']:>4} {100 * s / tot:5.1f}%  {d['Source'].strip()[:80]:80s} {st}")
