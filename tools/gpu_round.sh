#!/bin/bash
# one GPU call of the development loop: A/B of library builds, knob probes, parity tests, phase accounting
set -u
mkdir -p gpurun_out
TAG=${TAG:-run}
VARIANTS="${VARIANTS:-}" CFGS="${CFGS:-cfg3 cfg4}" bash tools/gpu_ab.sh > /dev/null 2>&1
cp gpurun_out/ab.log gpurun_out/ab_$TAG.log
cp build/lib_${FINAL}.so belief-planning_b200/libbranchmpc.so
: > gpurun_out/knobs_$TAG.log
while IFS= read -r line; do
  [ -z "$line" ] && continue
  echo "== $line" >> gpurun_out/knobs_$TAG.log
  env $line python tools/bench_configs.py ${KNOB_CFGS:-cfg3} 2>&1 | cut -c1-420 >> gpurun_out/knobs_$TAG.log
done <<< "${KNOBS:-}"
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log
python tools/gpu_phases.py hw > gpurun_out/phases_$TAG.log 2>&1
python - <<'PY'
import re,os,glob
tag=os.environ.get("TAG","run")
for f in ("gpurun_out/ab_%s.log"%tag,"gpurun_out/knobs_%s.log"%tag):
    for l in open(f):
        l=l.strip()
        if l.startswith("=="): print(l); continue
        m=re.search(r'"config": "(\S+).*?cold_solves_per_s": (\d+).*?warm_solves_per_s": (\d+).*?status_warm_last": (\[[^\]]*\]).*?mean_nfact_cold_warm": (\[[^\]]*\]).*?mean_nsolve_cold_warm": (\[[^\]]*\])',l)
        if m: print("   ",*m.groups())
        elif "rror" in l: print("   ",l[:200])
PY
tail -12 gpurun_out/phases_$TAG.log
