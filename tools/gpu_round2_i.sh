#!/bin/bash
# GPU session I of round 2: the new upstream/downstream/wrapper tests + full suite
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/i_pytest.log 2>&1; tail -12 gpurun_out/i_pytest.log
