#!/bin/bash
# One GPU session: parity tests, bench, ncu launch list + one full capture of the solve kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$BCMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$BCMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bmpc_solve_kernel -s 3 -c 1 -f -o gpurun_out/prof_solve $BCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
