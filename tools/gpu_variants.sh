#!/bin/bash
# usage: tools/gpu_variants.sh name1 name2 ...   (libraries built by tools/build_variant.sh)
mkdir -p gpurun_out
for v in "$@"; do
  SPECTRAL_LIB_DIR=$PWD/spectral_b200/lib_$v timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python - "$v" <<'PY'
import json, sys
v=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/var_%s.json"%v).read().strip().splitlines()[-1])
    print("VARIANT %-14s value %8.0f e2e %8.0f qp_ms %.3f frac %.4f" % (v, d["value"], d["e2e"]["value"], d["kernel_ms_per_step"]["qp"], d["roofline"]["frac"]))
except Exception as e:
    print("VARIANT", v, "ERR", e, open("gpurun_out/var_%s.err"%v).read()[-300:])
PY
done
