#!/bin/bash
# experiment: lane-per-segment kernels (block-tridiagonal solve) for every class vs the dense-operator kernels
set -x
mkdir -p gpurun_out
for st in 4 8 16; do
SPECTRAL_FORCE_LANES=1 timeout 300 python bench.py --steps 40 --warmup 5 --streams $st > gpurun_out/t_lanes_s$st.json 2> gpurun_out/t_lanes_s$st.err; tail -c 200 gpurun_out/t_lanes_s$st.err
done
python - <<'PY'
import json
for st in (4,8,16):
    try:
        d=json.loads(open("gpurun_out/t_lanes_s%d.json"%st).read().strip().splitlines()[-1])
        print("lanes streams",st,"value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "qp ms", d["kernel_ms_per_step"]["qp"], "iters", d["config"]["mean_axis_iters"], "solved", d["config"]["solved_fraction"])
    except Exception as e: print(st,"ERR",e)
PY
SPECTRAL_FORCE_LANES=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "config2_cub or mixed_variable" 2>&1 | tail -4 | cut -c1-300
