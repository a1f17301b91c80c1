"""Copies a tools/record_run.sh result set (gpurun_out/<tag>_*) into profiles/ and prints the numbers profiles/README.md quotes
(developer tool; runs where ncu is installed, no GPU needed):  python tools/publish_record.py <tag> [prefix]"""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r2"      # file prefix under profiles/
g = lambda n: os.path.join(ROOT, "gpurun_out", "%s_%s" % (tag, n))
p = lambda n: os.path.join(ROOT, "profiles", n)
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), g("full.ncu-rep"), "64", p(rnd + "_ncu_summary.csv"), "/tmp/dram.json", "/tmp/inst.json"], stdout=subprocess.DEVNULL)
wi = json.load(open("/tmp/inst.json"))
wi["k_stereo_match"] = wi.get("k_stereo_match", 0) + wi.pop("k_stereo_index", 0)
json.dump(dict({"_note": "smsp__inst_executed.sum (warp instructions) per image (per pair for k_stereo_match incl. k_stereo_index; all 7 levels for k_resize_level) from the same capture"}, **wi), open(p("ncu_warp_inst_per_image.json"), "w"), indent=1)
d = json.load(open("/tmp/dram.json"))
d["k_stereo_match"] += d.pop("k_stereo_index", 0)
out = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per image (per pair for k_stereo_match incl. k_stereo_index; all 7 levels for k_resize_level) from one ncu --set full capture at 64 images per launch (tools/record_run.sh, tools/ncu_summary.py)"}
out.update(d)
json.dump(out, open(p("ncu_dram_bytes_per_image.json"), "w"), indent=1)
shutil.copy(g("bench.json"), p(rnd + "_bench_n1.json"))
shutil.copy(g("bench_ref.json"), p(rnd + "_bench_reference_arm.json"))
shutil.copy(g("launches.csv"), p(rnd + "_ncu_launches.csv"))
open(p(rnd + "_kernel_times.txt"), "w").write(open(g("kernel_times.txt")).read() + open(g("latency.txt")).read())
open(p(rnd + "_n2_n4_measurements.txt"), "w").write(open(g("n2.txt")).read() + open(g("n4.txt")).read())
b = json.load(open(g("bench.json")))
r = json.load(open(g("bench_ref.json")))
print("value %.0f  ms/step %.2f  e2e %.0f  launches %d" % (b["value"], b["ms_per_step"], b["e2e"]["value"], b["gpu_launches"]))
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in b["roofline"].items()})
print("shares", b["kernel_shares"])
print("cpu", round(b["cpu_baseline"]["value"], 1), round(b["cpu_baseline"]["reference_threading_2plus1"]["value"], 1), "reference arm", round(r["value"], 1))
rows = list(csv.reader(io.StringIO("\n".join(l for l in open(g("launches.csv")).read().splitlines() if l.startswith('"')))))
ik, iv = rows[0].index("Kernel Name"), rows[0].index("Metric Value")
tot, cnt = {}, {}
for row in rows[1:]:
    if len(row) > iv:
        n = row[ik].split("(")[0].replace("void ", "").split("<")[0]
        try:
            v = float(row[iv].replace(",", ""))
        except ValueError:
            continue
        tot[n] = tot.get(n, 0) + v; cnt[n] = cnt.get(n, 0) + 1
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-20s %4d launches %8.0f us  share %.3f" % (k, cnt[k], v / 1e3, v / T))
print(open(g("kernel_times.txt")).read())
print(open(g("latency.txt")).read().splitlines()[-1])
