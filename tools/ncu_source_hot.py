"""Hottest CUDA source lines of an ncu report (warp stall samples aggregated per line, top stall reasons):
   python tools/ncu_source_hot.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(raw.splitlines()))
hdr, fpath = None, ""
lines = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].rsplit("/", 1)[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        ie = hdr.index("Instructions Executed")
        stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    key = (fpath, r[0])
    st = {hdr[i]: int(r[i] or 0) for i in stalls}
    d = lines.setdefault(key, [0, r[1], {}, 0])
    d[0] += n
    d[3] += int(r[ie] or 0)
    for k, v in st.items():
        d[2][k] = d[2].get(k, 0) + v
tot = sum(d[0] for d in lines.values()) or 1
print("total samples", tot)
for (f, ln), d in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(((v, k) for k, v in d[2].items() if v), reverse=True)[:3]
    print(f"{d[0]:7d} {100 * d[0] / tot:5.1f}%  inst {d[3]:9d}  {f}:{ln:>4}  {d[1].strip()[:90]}   {[(k[6:], v) for v, k in st]}")
