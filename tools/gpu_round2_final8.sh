#!/bin/bash
# Final GPU session of round 2, 8 GPUs: default bench at N = 2, 4, 8 (weak scaling) and the 1 M-scenario sweep (config 5) at N = 8, 4, 2, 1
set -x
mkdir -p gpurun_out
P=29600
for n in 8 4 2; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 40 --warmup 5 > gpurun_out/z_bench2_n$n.json 2> gpurun_out/z_bench2_n$n.err
  tail -c 200 gpurun_out/z_bench2_n$n.err
done
for n in 8 4 2; do
  P=$((P+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --config 5 --steps 1 --warmup 1 > gpurun_out/z_bench5_n$n.json 2> gpurun_out/z_bench5_n$n.err
  tail -c 200 gpurun_out/z_bench5_n$n.err
done
timeout 400 python bench.py --config 5 --steps 1 --warmup 1 > gpurun_out/z_bench5_n1.json 2> gpurun_out/z_bench5_n1.err
python - <<'PY'
import json
for n in ("bench2_n8","bench2_n4","bench2_n2","bench5_n8","bench5_n4","bench5_n2","bench5_n1"):
    try:
        d=json.loads(open("gpurun_out/z_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", round((d.get("e2e") or {}).get("value") or 0), "ms", round(d["ms_per_step"],2), d["config"].get("winner"))
    except Exception as e: print(n, "ERR", e)
PY
