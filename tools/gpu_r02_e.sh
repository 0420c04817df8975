#!/bin/bash
# Round 2, GPU call E (1 GPU): the whole GPU suite on the current library (new: device SpMV exchange, Graph500 stream, SpGEMM)
set -o pipefail
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x -n 3 2>&1 | tail -15 | tee gpurun_out/r02e_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02e_smoke.log
