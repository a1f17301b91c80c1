"""Post-process a full ncu capture (developer tool, runs where ncu is installed, no GPU needed):
   ncu_summary.py <report.ncu-rep> <images per launch> <summary.csv> <dram_bytes_per_image.json> [warp_inst_per_image.json]"""
import csv, io, json, re, subprocess, sys
rep, n_img, out_csv, out_json = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread']
idx = [hdr.index(c) for c in cols]
seen, out, dram, inst = {}, [], {}, {}
def to_bytes(v, u):
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
for r in data:
    if len(r) <= max(idx):
        continue
    name = re.sub(r'<.*>', '', r[idx[0]].split('(')[0].replace('void ', ''))
    key = (name, r[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
    if key in seen:            # left and right eye launch the same kernels: keep the first of each shape
        continue
    seen[key] = 1
    out.append([r[i][:16] if i == idx[0] else r[i] for i in idx])
    b = to_bytes(r[idx[5]], units[idx[5]]) + to_bytes(r[idx[6]], units[idx[6]])
    dram[name] = dram.get(name, 0.0) + b / n_img
    inst[name] = inst.get(name, 0.0) + float(r[idx[4]]) / n_img
with open(out_csv, 'w', newline='') as f:
    w = csv.writer(f); w.writerow(cols); w.writerow([units[i] for i in idx]); w.writerows(out)
json.dump({k: round(v) for k, v in dram.items()}, open(out_json, 'w'), indent=1)
if len(sys.argv) > 5:      # warp instructions executed per image per kernel (feeds roofline.issue_frac in bench.py)
    json.dump({k: round(v) for k, v in inst.items()}, open(sys.argv[5], 'w'), indent=1)
print(json.dumps({k: round(v) for k, v in dram.items()}))
