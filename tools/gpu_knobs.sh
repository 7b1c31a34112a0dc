#!/bin/bash
set -u
mkdir -p gpurun_out; : > gpurun_out/knobs.log
run() { echo "== $*" >> gpurun_out/knobs.log; env "$@" BMPC_SWEEP=3x3 python tools/bench_configs.py ${CFGS:-cfg3 cfg4 cfg5} 2>&1 | cut -c1-400 >> gpurun_out/knobs.log; }
while read -r line; do [ -n "$line" ] && run $line; done < tools/knobs.txt
cat gpurun_out/knobs.log
