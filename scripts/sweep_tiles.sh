#!/bin/bash
# Tuning sweep of the K1/K2 tile geometry on the GPU box (rebuilds the library per config).
set -u
mkdir -p gpurun_out
for cfg in "256 8 4" "256 8 5" "256 4 8" "512 4 4" "256 6 6" "128 8 8" "256 5 6" "512 8 2" "384 4 5" "128 16 6"; do
  set -- $cfg
  export DFD_NVCC_DEFS="-DDFD_TILE_THREADS=$1 -DDFD_TILE_K=$2 -DDFD_TILE_MIN_CTAS=$3"
  python datafusion_distributed_b200/build.py --force >/dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  echo -n "T=$1 K=$2 CTAS=$3 : "
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step %.3f  scatter %.3f ms (%.1f%%)  hist %.3f  scan %.3f' % (d['ms_per_step'], r['kernel_ms'], 100*r['frac'], r['hist_ms'], r['scan_ms']))"
done 2>&1 | tee gpurun_out/sweep_tiles.txt
unset DFD_NVCC_DEFS
python datafusion_distributed_b200/build.py --force >/dev/null
