#!/bin/bash
# GPU session L of round 2: full GPU suite (parity summary kept), config 3 (G = 1 / 8 / 64), config 4, reference arm
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/l_pytest.log 2>&1; tail -4 gpurun_out/l_pytest.log | cut -c1-600
for g in 8 1 64; do
  timeout 600 python bench.py --config 3 --groups $g --steps 4 --warmup 3 > gpurun_out/l_bench3_g$g.json 2> gpurun_out/l_bench3_g$g.err; tail -c 300 gpurun_out/l_bench3_g$g.err
done
timeout 600 python bench.py --config 4 --steps 1 --warmup 1 > gpurun_out/l_bench4.json 2> gpurun_out/l_bench4.err; tail -c 300 gpurun_out/l_bench4.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/l_bench2_ref.json 2> gpurun_out/l_bench2_ref.err
python - <<'PY'
import json
for n in ("bench3_g8","bench3_g1","bench3_g64","bench4","bench2_ref"):
    try:
        d=json.loads(open("gpurun_out/l_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", d.get("e2e",{}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "iters", d["config"].get("mean_axis_iters"))
    except Exception as e: print(n, "ERR", e)
PY
