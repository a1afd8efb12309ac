"""Annotated SASS of one kernel from an ncu report: executed warp instructions per unit (e.g. per hop), stall samples and the source line
of every instruction.  usage: python profiles/ncu_sass_annotate.py <report.ncu-rep> <cubin> <kernel-substring> <units> > out.txt"""
import csv, io, re, subprocess, sys
rep, cubin, kname, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address"); hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of = {}; cur = None; ink = False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m: ink = kname in m.group(1); continue
    if not ink: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
base = int(body[0][col["Address"]], 16)
for r in body:
    off = int(r[col["Address"]], 16) - base
    n = float(r[col["Instructions Executed"]] or 0); s = float(r[col["# Samples"]] or 0)
    k, ins = line_of.get(off, (None, r[col["Source"]] if "Source" in col else ""))
    print(f"{off:05x} {n / units:6.2f} {int(s):5d}  {(k[0][:14] + ':' + str(k[1])) if k else '':22s} {ins}")
