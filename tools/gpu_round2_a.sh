#!/bin/bash
# GPU session A of round 2: parity tests, A/B bench (anchor vs legacy full-row), launch list, ncu of the new kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/a_pytest.log 2>&1
tail -30 gpurun_out/a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_anchor.json 2> gpurun_out/a_bench_anchor.err
SPECTRAL_LEGACY_QPD=1 timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_legacy.json 2> gpurun_out/a_bench_legacy.err
tail -c 600 gpurun_out/a_bench_anchor.err
python - <<'PY'
import json
for n in ("anchor","legacy"):
    try:
        d=json.loads(open("gpurun_out/a_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "qp ms", d["kernel_ms_per_step"])
    except Exception as e: print(n, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 1 --streams 1 > gpurun_out/a_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qpa -c 2 -o gpurun_out/a_qpa_full python bench.py --steps 1 --warmup 1 --streams 1 > gpurun_out/a_ncu_full.log 2>&1
ls -la gpurun_out | head -30
