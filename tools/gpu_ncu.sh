#!/bin/bash
# ncu --set full capture of one warm solve-kernel launch (same command run plain first, as the recipe requires); the report
# is reduced to CSV on the box (raw metrics + per-source-line page)
set -u
mkdir -p gpurun_out
CMD="python tools/gpu_cycles.py"
STEPS=5 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
STEPS=5 ncu --set full --clock-control none --import-source on -k regex:bmpc_solve_kernel -s 4 -c 1 -f -o /tmp/prof_solve $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"
ncu -i /tmp/prof_solve.ncu-rep --page raw --csv > gpurun_out/prof_solve_raw.csv 2>/dev/null
ncu -i /tmp/prof_solve.ncu-rep --page source --csv > gpurun_out/prof_solve_source.csv 2>/dev/null
tail -3 gpurun_out/ncu_plain.log
ls -la gpurun_out | tail -5
