#!/bin/bash
# Round 2, GPU call F (1 GPU): narrow panels (16- / 32-byte rows) on the 4-lane layout vs the 1- / 2-lane layouts, final bench line.
set -o pipefail
mkdir -p gpurun_out
for NARROW in 0 1; do
  for K in 1 4 8; do CB_K2_NARROW=$NARROW timeout 300 python tools/kbench.py c2 c5 --k $K --steps 5 2>&1 | sed "s/^{/{\"narrow\": $NARROW, /" ; done
  CB_K2_NARROW=$NARROW timeout 300 python tools/kbench.py c5b --steps 5 2>&1 | sed "s/^{/{\"narrow\": $NARROW, /"
done | tee gpurun_out/r02f_narrow.jsonl
timeout 2400 python -m pytest tests -m gpu -q -x -n 3 2>&1 | tail -8 | tee gpurun_out/r02f_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -c 300 gpurun_out/r02f_bench.err
