#!/bin/bash
# scaling run: bash tools/gpu_scale.sh <workload> "<N list>" [extra bench args]
W=$1; NS=$2; shift 2
mkdir -p gpurun_out
for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --workload $W --no-cpu-baseline "$@" 2> gpurun_out/scale_${W}_$N.err | tee gpurun_out/scale_${W}_$N.json | cut -c1-900
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --workload $W --no-cpu-baseline "$@" 2> gpurun_out/scale_${W}_$N.err | tee gpurun_out/scale_${W}_$N.json | cut -c1-900
  fi
  tail -2 gpurun_out/scale_${W}_$N.err
done
