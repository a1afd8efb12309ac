"""Correlates an ncu SASS-level source page (csv) with CUDA source lines using nvdisasm line info.

usage: python profiles/ncu_by_line.py <report.ncu-rep> <cubin> <kernel-name-substring> [top]
Prints, per source line, the share of warp-stall samples and the dominant stall reasons.
(ncu's own `--print-source cuda` view carries no metrics in csv mode.)
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
# nvdisasm: map instruction offsets to source lines for the kernel
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}
cur_line = None
in_k = False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        in_k = kname in m.group(1)
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
base = int(body[0][col["Address"]], 16)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
per_line = defaultdict(lambda: defaultdict(float))
tot = 0.0
for r in body:
    off = int(r[col["Address"]], 16) - base
    ln = line_of.get(off, (None, ""))[0]
    n = float(r[col["# Samples"]] or 0)
    tot += n
    per_line[ln]["samples"] += n
    per_line[ln]["inst"] += float(r[col["Instructions Executed"]] or 0)
    for sc in stall_cols:
        per_line[ln][sc] += float(r[col[sc]] or 0)
print(f"total samples {tot:.0f}")
for ln, d in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    reasons = sorted(((d[sc], sc) for sc in stall_cols), reverse=True)[:3]
    rs = ", ".join(f"{n[6:]} {v / max(d['samples'], 1) * 100:.0f}%" for v, n in reasons if v > 0)
    print(f"{str(ln):38s} {d['samples'] / tot * 100:6.2f}%  inst {d['inst']:12.0f}  {rs}")
