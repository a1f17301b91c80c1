#!/usr/bin/env python3
"""Per basic-block executed count (per warp) of a kernel from an ncu report (first instance). usage: ncu_trips.py rep kernel warps_total"""
import csv, io, subprocess, sys
rep, kern, W = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass","--kernel-name",kern],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
h=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
hdr=rows[h]; ie=hdr.index("Instructions Executed"); isrc=hdr.index("Source"); ismp=hdr.index("# Samples")
prev=None; start=None; n=0; last=-1; smp=0; tot=0; first_op=""
def flush():
    global tot
    if prev is not None:
        print("%s  %4d instr x %7.2f /warp = %8.1f   samples %6d   %s"%(start,n,prev,n*prev,smp,first_op)); tot+=n*prev
for r in rows[h+1:]:
    if len(r)<=ie or not r[0].startswith("0x"): continue
    a=int(r[0],16)
    if a<last: break
    last=a
    key=round(int(r[ie])/W,2)
    if key!=prev:
        flush(); prev=key; start=r[0][-4:]; n=0; smp=0; first_op=r[isrc].strip()[:40]
    n+=1; smp+=int(r[ismp] or 0)
flush(); print("total per warp %.0f"%tot)
