#!/bin/bash
# Developer GPU iteration (under gpurun): quick parity subset, isolated kernel times, optional ncu capture of one kernel.
#   tools/iter.sh <tag> [kernel-regex-for-ncu] [pytest -k expression]
tag=${1:-it}; kern=${2:-}; kexpr=${3:-"c1 or c2 or c3 or golden or random_geometries or cost_map or octree or other_reference"}
o=gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$kexpr" 2>&1 | tail -4
python tools/kernel_times.py 512 3 > $o/${tag}_kernel_times.txt 2>&1; cat $o/${tag}_kernel_times.txt
if [ -n "$kern" ]; then
  ncu --set full --clock-control none --import-source on -k "regex:$kern" -c 2 -o $o/${tag}_ncu -f python tools/profile_run.py 64 1 > $o/${tag}_ncu.log 2>&1
  tail -2 $o/${tag}_ncu.log
fi
