#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "shared_kkt" > gpurun_out/c_pytest.log 2>&1; tail -25 gpurun_out/c_pytest.log
timeout 600 python bench.py --config 3 --groups 8 --steps 6 --warmup 3 > gpurun_out/c_bench3.json 2> gpurun_out/c_bench3.err; tail -c 1500 gpurun_out/c_bench3.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c_bench3.json").read().strip().splitlines()[-1])
    print("config3 value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["kernel_ms_per_step"], "iters", d["config"]["mean_axis_iters"], "solved", d["config"]["solved_fraction"])
    print(d["roofline"]["per_class"])
except Exception as e: print("ERR", e)
PY
