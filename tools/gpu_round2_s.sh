#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/s_bench2.json 2> gpurun_out/s_bench2.err; tail -c 300 gpurun_out/s_bench2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s_bench2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["kernel_ms_per_step"]["qp"])
PY
timeout 300 python tools/dbg_shared.py 2048 8 > gpurun_out/s_dbg_new.log 2>&1; tail -4 gpurun_out/s_dbg_new.log
SPECTRAL_LEGACY_QPS=1 timeout 300 python tools/dbg_shared.py 2048 8 > gpurun_out/s_dbg_legacy.log 2>&1; tail -4 gpurun_out/s_dbg_legacy.log
