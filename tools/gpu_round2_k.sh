#!/bin/bash
# GPU session K of round 2: bench after the check change, wrapper tests, launch list + full ncu capture of k_qpa
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/k_bench2.json 2> gpurun_out/k_bench2.err; tail -c 300 gpurun_out/k_bench2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/k_bench2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["kernel_ms_per_step"])
print([ (c["class"], round(c["ms_per_step"],2)) for c in d["roofline"]["per_class"]])
PY
timeout 300 python -m pytest tests/test_gpu_parity.py -q -rs -k "wrapper" > gpurun_out/k_pytest.log 2>&1; tail -3 gpurun_out/k_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/k_launches.csv python bench.py --steps 2 --warmup 1 --streams 1 > gpurun_out/k_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qpa -c 2 -o gpurun_out/k_qpa_full python bench.py --steps 1 --warmup 1 --streams 1 > gpurun_out/k_ncu_full.log 2>&1
ls -la gpurun_out/k_*
