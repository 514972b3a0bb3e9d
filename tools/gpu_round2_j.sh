#!/bin/bash
# GPU session J of round 2 (2 GPUs): default bench and config 5 (small) under torchrun, wrapper drop-in tests
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -rs -k "wrapper or many_segments or edge" > gpurun_out/j_pytest.log 2>&1; tail -5 gpurun_out/j_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/j_bench2_n2.json 2> gpurun_out/j_bench2_n2.err; tail -c 400 gpurun_out/j_bench2_n2.err; cut -c1-300 gpurun_out/j_bench2_n2.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 5 --total 262144 --steps 1 --warmup 1 > gpurun_out/j_bench5_n2.json 2> gpurun_out/j_bench5_n2.err; tail -c 600 gpurun_out/j_bench5_n2.err; cut -c1-1200 gpurun_out/j_bench5_n2.json
