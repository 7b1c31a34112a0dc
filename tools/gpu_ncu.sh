#!/bin/bash
# ncu captures of the solve kernel at 4 warps/SM and at 1 warp/SM (same command run plain first, as the recipe requires);
# the reports are reduced to CSV on the box (the .ncu-rep files are too large to bring back together)
set -u
mkdir -p gpurun_out
for occ in 0 1; do
  CMD="python tools/gpu_cycles.py slab_mode=1 occ=$occ"
  STEPS=5 $CMD > gpurun_out/ncu_plain_occ$occ.log 2>&1 &&
  STEPS=5 ncu --set full --clock-control none --import-source on -k regex:bmpc_solve_kernel -s 4 -c 1 -f -o /tmp/prof_occ$occ $CMD > gpurun_out/ncu_occ$occ.log 2>&1
  echo "occ=$occ rc=$?"
  ncu -i /tmp/prof_occ$occ.ncu-rep --page raw --csv > gpurun_out/prof_occ${occ}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_occ$occ.ncu-rep --page source --csv > gpurun_out/prof_occ${occ}_source.csv 2>/dev/null
done
cp /tmp/prof_occ0.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out | tail -10
