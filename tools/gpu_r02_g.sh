#!/bin/bash
# Round 2, 2-GPU call G: distributed ingestion, device SpMV exchange and the C++ drivers on a process grid
set -o pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest "tests/test_summa_gpu.py::test_summa_2gpu" tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02g_pytest_2.log
