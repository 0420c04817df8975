#!/bin/bash
# Round 2, multi-GPU call: SUMMA parity tests that need N GPUs (dense, host-panel path, sparse right-hand side), then bench.py at N.
#   gpurun --gpus N -- 'bash tools/gpu_r02_multi.sh N [extra bench grids]'
N=${1:-2}
set -o pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r02m_gpus_$N.txt
timeout 1500 python -m pytest tests/test_summa_gpu.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r02m_pytest_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/r02m_bench_$N.json 2> gpurun_out/r02m_bench_$N.err; tail -c 800 gpurun_out/r02m_bench_$N.err
shift
for G in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 --grid $G \
      > gpurun_out/r02m_bench_${N}_$G.json 2> gpurun_out/r02m_bench_${N}_$G.err; tail -c 400 gpurun_out/r02m_bench_${N}_$G.err
done
du -sh gpurun_out
