#!/bin/bash
# Round 2, 8-GPU call: bench.py at N=8 on the default grid (2x4) with targets and e2e, then C3 alone on 4x2 for the grid choice.
set -o pipefail
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 \
    > gpurun_out/r02m_bench_8.json 2> gpurun_out/r02m_bench_8.err; tail -c 600 gpurun_out/r02m_bench_8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 3 --grid 4x2 --no-target --e2e-steps 0 \
    > gpurun_out/r02m_bench_8_4x2.json 2> gpurun_out/r02m_bench_8_4x2.err; tail -c 400 gpurun_out/r02m_bench_8_4x2.err
du -sh gpurun_out
