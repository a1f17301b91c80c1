#!/bin/bash
# developer tool: per-kernel times for every library variant under iv_slam_b200/lib/var (built with make OUT=... EXTRA=...)
n=${1:-512}
echo "== default"; python tools/kernel_times.py $n 3 2>&1 | grep -E "${2:-k_fast_cells}|sum per pair"
for f in iv_slam_b200/lib/var/*.so; do
  echo "== $f"; IVSLAM_GPU_LIB=$PWD/$f python tools/kernel_times.py $n 3 2>&1 | grep -E "${2:-k_fast_cells}|sum per pair|rror"
done
