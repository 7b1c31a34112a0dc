#!/bin/bash
# Knob scan on the warm closed loop (16 384 highway episodes, mean kernel time over four rho-refresh periods): every line of
# $KNOBS_FILE (default tools/knobs_warm.txt) is a set of VAR=value pairs for tools/gpu_warm_stats.py
set -u
mkdir -p gpurun_out; OUT=gpurun_out/knob_scan.log; : > $OUT
for rep in 1 2; do
while IFS= read -r line; do
  [ -z "$line" ] && continue
  echo "== $line (rep $rep)" >> $OUT
  env $line BRIEF=1 STEPS=${STEPS:-46} python tools/gpu_warm_stats.py > /tmp/ws.log 2>&1
  python - >> $OUT <<'PY'
import re
ms, adm, nf, bad = [], [], [], 0
for l in open("/tmp/ws.log"):
    m = re.match(r"step (\d+)\s+([\d.]+) ms\s+status \[(\d+), (\d+), (\d+), (\d+)\]\s+ADMM path ([\d.]+) %\s+iters ([\d.]+) nfact ([\d.]+)", l)
    if not m or int(m.group(1)) < 10:
        continue
    ms.append(float(m.group(2))); adm.append(float(m.group(7))); nf.append(float(m.group(9))); bad += int(m.group(5)) + int(m.group(6))
n = max(len(ms), 1)
print("steps %d  mean %.3f ms  (%.3f M solves/s)  ADMM path %.1f %%  nfact %.2f  not-solved %d" % (n, sum(ms) / n, 16384 / (sum(ms) / n + 1e-9) / 1e3, sum(adm) / n, sum(nf) / n, bad))
PY
done < ${KNOBS_FILE:-tools/knobs_warm.txt}
done
cat $OUT
