#!/bin/bash
# A/B of library builds (build/lib_*.so) on the same box
set -u
mkdir -p gpurun_out; : > gpurun_out/ab.log
for rep in 1 2; do
for v in ${VARIANTS:-v0 v1}; do
  cp build/lib_$v.so belief-planning_b200/libbranchmpc.so
  echo "== $v rep $rep" >> gpurun_out/ab.log
  python tools/bench_configs.py ${CFGS:-cfg3 cfg4} 2>&1 | cut -c1-400 >> gpurun_out/ab.log
done; done
cat gpurun_out/ab.log
