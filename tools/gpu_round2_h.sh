#!/bin/bash
# GPU session H of round 2: full GPU suite, default bench, config 3, config 5 (N=1 small)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; tail -8 gpurun_out/h_pytest.log
timeout 300 python bench.py > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err; tail -c 600 gpurun_out/h_bench2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench2_ref.json 2> gpurun_out/h_bench2_ref.err
timeout 600 python bench.py --config 5 --total 131072 --steps 1 --warmup 1 > gpurun_out/h_bench5_small.json 2> gpurun_out/h_bench5.err; tail -c 600 gpurun_out/h_bench5.err
cut -c1-600 gpurun_out/h_bench2.json
