"""Instruction counts and stall samples per CUDA source line / per function from an ncu report.
usage: python profiles/ncu_inst_by_line.py <report.ncu-rep> <cubin> <kernel-substring> <hops*queries> [top]"""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, cubin, kname, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="Address"); hdr=rows[hi]; col={h:i for i,h in enumerate(hdr)}
body=[r for r in rows[hi+1:] if len(r)==len(hdr)]
dis=subprocess.run(["nvdisasm","-g","-c",cubin],capture_output=True,text=True).stdout
line_of={}; cur=None; ink=False
for ln in dis.splitlines():
    m=re.match(r"\s*\.section\s+\.text\.(\S+)",ln)
    if m: ink=kname in m.group(1); continue
    if not ink: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: cur=(m.group(1).split("/")[-1],int(m.group(2))); continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);",ln)
    if m: line_of[int(m.group(1),16)]=(cur,m.group(2).strip())
base=int(body[0][col["Address"]],16)
per=defaultdict(float); samp=defaultdict(float); tot=0; ts=0
for r in body:
    off=int(r[col["Address"]],16)-base
    n=float(r[col["Instructions Executed"]] or 0); tot+=n
    k=line_of.get(off,(None,""))[0]
    per[k]+=n; samp[k]+=float(r[col["# Samples"]] or 0); ts+=float(r[col["# Samples"]] or 0)
print(f"total warp instructions {tot:.0f} = {tot/units:.1f} per unit; samples {ts:.0f}")
import os
srcl=open(os.path.join(os.path.dirname(os.path.abspath(__file__)),"..","bang-billion-scale-ann_b200","csrc","search_kernel.cuh")).read().splitlines()
for k,n in sorted(per.items(), key=lambda kv:-kv[1])[:top]:
    t = srcl[k[1]-1].strip()[:78] if k and k[0]=="search_kernel.cuh" and k[1]<=len(srcl) else ""
    print(f"{str(k):34s} inst {n/tot*100:5.1f}% {n/units:6.1f}/unit  stall {samp[k]/ts*100:5.1f}%  {t}")
