#!/usr/bin/env python3
"""Instruction count of k_fast_cells by phase from an ncu report with source info (developer tool).
usage: tools/ncu_phases.py report.ncu-rep [images]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--kernel-name","k_fast_cells"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
src=open("iv_slam_b200/csrc/k_fast.cuh").read().split("\n")
# phase boundaries from the "// ---- X:" markers in the source
marks=[(i+1,l.strip()) for i,l in enumerate(src) if l.strip().startswith("// ----")]
cur=None;hdr=None;per={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if r[0].isdigit() and r[2]=="-": per[(cur,int(r[0]))]=int(r[hdr.index("Instructions Executed")])
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv','--metrics','smsp__inst_executed.sum,gpu__time_duration.sum'],capture_output=True,text=True).stdout
rr=list(csv.reader(io.StringIO(raw)))
T=float(rr[2][rr[0].index('smsp__inst_executed.sum')]); dur=rr[2][rr[0].index('gpu__time_duration.sum')]
px=n_img*1444097
print("kernel: %.1fM warp instr, %.1f thread-instr/px, %s us"%(T/1e6,T*32/px,dur))
bounds=[(1,"defs/score fn")]+[(ln,txt) for ln,txt in marks]+[(10**6,"")]
first_kernel_line=next(i+1 for i,l in enumerate(src) if "__global__" in l)
for (a,name),(b,_) in zip(bounds,bounds[1:]):
    if name=="defs/score fn":
        n1=sum(v for (f,l),v in per.items() if f=="k_fast.cuh" and l<first_kernel_line)
        n2=sum(v for (f,l),v in per.items() if f=="k_fast.cuh" and first_kernel_line<=l<b)
        print("%-60s %5.1f thread-instr/px"%("inlined helpers (vmin3/vmax3/score network)",n1*32/px))
        print("%-60s %5.1f"%("kernel prologue (incl. hoisted address math)",n2*32/px)); continue
    n=sum(v for (f,l),v in per.items() if f=="k_fast.cuh" and a<=l<b)
    print("%-60s %5.1f"%(name[:60],n*32/px))
other=sum(v for (f,l),v in per.items() if f!="k_fast.cuh")
print("%-60s %5.1f"%("other files (intrinsics headers)",other*32/px))
