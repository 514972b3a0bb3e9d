#!/bin/bash
# Final GPU session of round 2, one GPU: smoke, full GPU suite, default bench, reference arm, configs 3 / 4, launch list
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; tail -3 gpurun_out/z_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/z_pytest.log 2>&1; tail -3 gpurun_out/z_pytest.log | cut -c1-900
timeout 400 python bench.py > gpurun_out/z_bench2.json 2> gpurun_out/z_bench2.err; tail -c 300 gpurun_out/z_bench2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_bench2_ref.json 2> gpurun_out/z_bench2_ref.err
for g in 8 1 64; do
  timeout 600 python bench.py --config 3 --groups $g --steps 4 --warmup 3 > gpurun_out/z_bench3_g$g.json 2> gpurun_out/z_bench3_g$g.err
done
timeout 600 python bench.py --config 4 --steps 1 --warmup 1 > gpurun_out/z_bench4.json 2> gpurun_out/z_bench4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 1 --streams 1 > gpurun_out/z_ncu_list.log 2>&1
python - <<'PY'
import json
for n in ("bench2","bench2_ref","bench3_g8","bench3_g1","bench3_g64","bench4"):
    try:
        d=json.loads(open("gpurun_out/z_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", round((d.get("e2e") or {}).get("value") or 0), "frac", (d.get("roofline") or {}).get("frac"), "ms", round(d["ms_per_step"],2))
    except Exception as e: print(n, "ERR", e)
PY
