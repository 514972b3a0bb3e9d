#!/bin/bash
# GPU session O of round 2: bounds kernel test + bench line with side-kernel rooflines, compute-sanitizer over the round-2 kernels
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "upstream or downstream" 2>&1 | tail -3
timeout 500 python bench.py > gpurun_out/o_bench2.json 2> gpurun_out/o_bench2.err; tail -c 300 gpurun_out/o_bench2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/o_bench2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
for k in d["roofline_side_kernels"]: print({x:k.get(x) for x in ("kernel","ms_per_launch","achieved","frac","error")})
PY
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/o_memcheck.log 2>&1; tail -4 gpurun_out/o_memcheck.log
SANITIZE_MAX_ITER=60 timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/o_racecheck.log 2>&1; tail -4 gpurun_out/o_racecheck.log
