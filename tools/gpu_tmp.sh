set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v20.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_v20.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_v20.json 2> gpurun_out/bench_ref_v20.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench_1gpu_v20.json 2> gpurun_out/bench_v20.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/bench_1gpu_v20.json
BCMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-others"
$BCMD > gpurun_out/plain_v20.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_bench_v20.csv $BCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
bash tools/gpu_ncu.sh
