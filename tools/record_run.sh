#!/bin/bash
# Round record run (on the GPU box, under gpurun): bench line, reference arm, ncu launch list, one full ncu capture,
# isolated kernel times, shim latency, N2/N4 measurements.  Everything lands in gpurun_out/ with the given tag.
tag=${1:-rec}
o=gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 2>$o/${tag}_bench.err > $o/${tag}_bench_ref.json
python bench.py --steps 10 --warmup 3 2>>$o/${tag}_bench.err > $o/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 160 --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --batch 256 --no-cpu-baseline --no-configs > $o/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -s 0 -c 40 -o $o/${tag}_full -f python tools/profile_run.py 64 1 > $o/${tag}_full.log 2>&1
python tools/kernel_times.py 512 3 > $o/${tag}_kernel_times.txt 2>&1
python tools/kernel_times.py 512 3 1241 376 2000 intro >> $o/${tag}_kernel_times.txt 2>&1
python tools/kernel_times.py 16 3 3840 2160 8000 >> $o/${tag}_kernel_times.txt 2>&1
python tools/latency.py > $o/${tag}_latency.txt 2>&1
python tools/run_time_b1.py >> $o/${tag}_latency.txt 2>&1
python tools/projection_bench.py > $o/${tag}_n2.txt 2>&1
python tools/prologue_bench.py 256 > $o/${tag}_n4.txt 2>&1
grep -E "fast|sum" $o/${tag}_kernel_times.txt; cat $o/${tag}_latency.txt | tail -4; head -c 400 $o/${tag}_bench.json
