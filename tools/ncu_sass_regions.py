#!/usr/bin/env python3
"""Executed warp instructions of one kernel split at its barriers / by opcode (exact: per SASS instruction, no line table).
usage: tools/ncu_sass_regions.py report.ncu-rep kernel [pixels_total]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
px = float(sys.argv[3]) if len(sys.argv) > 3 else 64 * 1444097
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass","--kernel-name",kern],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; ie = hdr.index("Instructions Executed"); isrc = hdr.index("Source")
region, regs, ops = 0, collections.OrderedDict(), collections.Counter()
tot = 0
first = {}
for r in rows[h + 1:]:
    if len(r) <= ie or not r[0].startswith("0x"): continue
    n = int(r[ie]); s = r[isrc].strip(); op = s.split()[0] if not s.startswith("@") else s.split()[1]
    regs[region] = regs.get(region, 0) + n; first.setdefault(region, r[0]); ops[op.split(".")[0]] += n; tot += n
    if op.startswith("BAR"): region += 1
print("total %.2fM warp instr = %.1f thread-instr/px" % (tot / 1e6, tot * 32 / px))
for k, v in regs.items(): print("region %d (from %s): %6.2f%%  %5.1f thread-instr/px" % (k, first[k][-5:], 100.0 * v / tot, v * 32 / px))
print("by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(18)))
