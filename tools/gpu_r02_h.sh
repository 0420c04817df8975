#!/bin/bash
# Round 2, GPU call H (1 GPU): the GPU suite the way the driver runs it (serial), with the slowest tests listed
set -o pipefail
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -x --durations=25 ) 2>&1 | tail -45 | tee gpurun_out/r02h_pytest_gpu.log
