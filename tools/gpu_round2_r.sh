#!/bin/bash
# GPU session R of round 2: sanitizer over the four-warp tile kernel and the new bounds kernel, ncu capture of k_qps4, config 3 G = 64
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r_memcheck.log 2>&1; tail -3 gpurun_out/r_memcheck.log
SANITIZE_MAX_ITER=60 timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/r_racecheck.log 2>&1; tail -3 gpurun_out/r_racecheck.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qps4 -c 1 -o gpurun_out/r_qps4_full python tools/dbg_shared.py 4096 8 > gpurun_out/r_ncu.log 2>&1; tail -5 gpurun_out/r_ncu.log
timeout 600 python bench.py --config 3 --groups 64 --steps 3 --warmup 3 > gpurun_out/r_bench3_g64.json 2> gpurun_out/r_bench3_g64.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r_bench3_g64.json").read().strip().splitlines()[-1])
print("g64 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", d["roofline"]["frac"])
PY
