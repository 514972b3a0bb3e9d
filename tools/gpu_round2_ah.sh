#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "shared_kkt" 2>&1 | tail -2 | cut -c1-300
for g in 1 64 8; do
  timeout 600 python bench.py --config 3 --groups $g --steps 3 --warmup 3 > gpurun_out/ah_g$g.json 2>/dev/null
  SPECTRAL_QPS_NOHINT=1 timeout 600 python bench.py --config 3 --groups $g --steps 3 --warmup 3 > gpurun_out/ah_g${g}_nohint.json 2>/dev/null
done
python - <<'PY'
import json
for n in ("g1","g1_nohint","g64","g64_nohint","g8","g8_nohint"):
    try:
        d=json.loads(open("gpurun_out/ah_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],1), "iters", round(d["config"]["mean_axis_iters"]), "solved", round(d["config"]["solved_fraction"],4))
    except Exception as e: print(n,"ERR",e)
PY
