#!/usr/bin/env python3
"""Per-source-line instruction counts of one kernel from an ncu report (needs -lineinfo and --import-source on).
usage: tools/ncu_lines.py report.ncu-rep kernel_name [file_substring] [top_n]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
fsub = sys.argv[3] if len(sys.argv) > 3 else ""
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None
per = collections.OrderedDict(); total = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0].isdigit() and r[2] == "-":          # a source line row (aggregated over its SASS)
        ie = hdr.index("Instructions Executed")
        try: n = int(r[ie])
        except ValueError: continue
        per[(cur, int(r[0]))] = (n, r[1].strip()[:110], r[hdr.index("# Samples")])
        total += n
print("total warp instructions:", total)
items = [(k, v) for k, v in per.items() if fsub in k[0]]
byline = sorted(items, key=lambda kv: -kv[1][0])[:top]
for (f, ln), (n, src, smp) in sorted(byline, key=lambda kv: (kv[0][0], kv[0][1])):
    print("%6.2f%%  %-14s:%4d  smp %-6s %s" % (100.0 * n / total, f.split("/")[-1], ln, smp, src))
