#!/bin/bash
# Round 2, GPU call C (1 GPU): K2 with the entry prefetch against K2, then the whole GPU suite and a first bench run.
set -o pipefail
mkdir -p gpurun_out
V="k2 pf pf:point=0 pf:point=1 pf:slab=256"
timeout 600 python tools/kbench.py c2 c5 c5b c3 s24f32 c4 --steps 5 --check --variants $V 2>&1 | tee gpurun_out/r02c_sweep.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02c_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -c 600 gpurun_out/r02c_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err; tail -c 600 gpurun_out/r02c_bench_ref.err
du -sh gpurun_out
