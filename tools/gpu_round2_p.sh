#!/bin/bash
# GPU session P of round 2: four-warp shared-KKT tile kernel (k_qps4) -- parity tests, A/B against the two-warp kernel, bounds kernel
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "shared_kkt or upstream" > gpurun_out/p_pytest.log 2>&1; tail -12 gpurun_out/p_pytest.log | cut -c1-400
timeout 600 python bench.py --config 3 --groups 8 --steps 3 --warmup 3 > gpurun_out/p_bench3_g8.json 2> gpurun_out/p_bench3_g8.err; tail -c 300 gpurun_out/p_bench3_g8.err
SPECTRAL_LEGACY_QPS=1 timeout 600 python bench.py --config 3 --groups 8 --steps 3 --warmup 3 > gpurun_out/p_bench3_g8_legacy.json 2> gpurun_out/p_bench3_g8_legacy.err
timeout 600 python bench.py --config 3 --groups 1 --steps 3 --warmup 3 > gpurun_out/p_bench3_g1.json 2> gpurun_out/p_bench3_g1.err
python - <<'PY'
import json
for n in ("bench3_g8","bench3_g8_legacy","bench3_g1"):
    try:
        d=json.loads(open("gpurun_out/p_%s.json"%n).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", r["frac"], "iters", d["config"]["mean_axis_iters"], "solved", d["config"]["solved_fraction"])
        print("   ", [(c["class"],round(c["ms_per_step"],1),round(c["frac"],4)) for c in r["per_class"]])
        for k in d["roofline_side_kernels"]: print("   ", {x:k.get(x) for x in ("kernel","ms_per_launch","frac","error")})
    except Exception as e: print(n, "ERR", e)
PY
