#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "weight_sweep or sweep_argmin or shared_kkt" > gpurun_out/f_pytest.log 2>&1; tail -15 gpurun_out/f_pytest.log
timeout 600 python bench.py --config 5 --total 131072 --steps 1 --warmup 1 > gpurun_out/f_bench5_small.json 2> gpurun_out/f_bench5.err; tail -c 800 gpurun_out/f_bench5.err; cut -c1-900 gpurun_out/f_bench5_small.json
