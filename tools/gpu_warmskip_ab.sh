#!/bin/bash
# A/B of the warm-polish skip knobs (bmpc_config.reserved[0]: bit 0 = no rebalance, bit 1 = warm polish on rho-refresh solves,
# bits 4..7 = solves without a warm attempt after one that ended on the ADMM path; 15 = never skip): mean kernel time over
# four rho-refresh periods of the closed loop, 16 384 highway episodes
set -u
mkdir -p gpurun_out; : > gpurun_out/warmskip.log
for rep in 1 2; do
for r0 in ${R0S:-240 0 16 32 96 242 2 18 34}; do
  echo "== R0=$r0 rep $rep" >> gpurun_out/warmskip.log
  BMPC_R0=$r0 BRIEF=1 STEPS=46 python tools/gpu_warm_stats.py > /tmp/ws.log 2>&1
  python - >> gpurun_out/warmskip.log <<'PY'
import re
ms, adm, nf, bad = [], [], [], 0
for l in open("/tmp/ws.log"):
    m = re.match(r"step (\d+)\s+([\d.]+) ms\s+status \[(\d+), (\d+), (\d+), (\d+)\]\s+ADMM path ([\d.]+) %\s+iters ([\d.]+) nfact ([\d.]+)", l)
    if not m:
        continue
    if int(m.group(1)) < 10:
        continue
    ms.append(float(m.group(2))); adm.append(float(m.group(7))); nf.append(float(m.group(9))); bad += int(m.group(5)) + int(m.group(6))
n = len(ms)
print("steps %d  mean %.3f ms  (%.3f M solves/s)  ADMM path %.1f %%  nfact %.2f  not-solved %d" % (n, sum(ms) / n, 16384 / (sum(ms) / n) / 1e3, sum(adm) / n, sum(nf) / n, bad))
PY
done; done
cat gpurun_out/warmskip.log
