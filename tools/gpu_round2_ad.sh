#!/bin/bash
set -x
mkdir -p gpurun_out
SPECTRAL_DEFER_FINISH=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fixture or config2 or mixed or weights or weight_sweep or edge or async or device_resident or obstacle" 2>&1 | tail -3 | cut -c1-300
SPECTRAL_DEFER_FINISH=1 timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/ad_bench2_defer.json 2> gpurun_out/ad.err; tail -c 300 gpurun_out/ad.err
timeout 300 python bench.py --steps 40 --warmup 5 > gpurun_out/ad_bench2_inline.json 2> gpurun_out/ad2.err
python - <<'PY'
import json
for n in ("defer","inline"):
    d=json.loads(open("gpurun_out/ad_bench2_%s.json"%n).read().strip().splitlines()[-1])
    print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "qp ms", round(d["kernel_ms_per_step"]["qp"],2), "solved", d["config"]["solved_fraction"], "nopolish", round(d["config"]["with_reference_polish_setting"]["value"]))
PY
