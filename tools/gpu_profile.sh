#!/bin/bash
# ncu evidence for one workload: launch list + one full capture of the multiply kernel.
# Usage (under gpurun): bash tools/gpu_profile.sh <workload> <tag>
W=${1:-c2}; TAG=${2:-r01}
mkdir -p gpurun_out
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > gpurun_out/plain_$W.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${W}_$TAG.csv $CMD > gpurun_out/ncu_list_$W.log 2>&1
$CMD > gpurun_out/plain2_$W.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:cb_spmm_kernel -s 3 -c 2 -o gpurun_out/prof_${W}_$TAG $CMD > gpurun_out/ncu_full_$W.log 2>&1
tail -3 gpurun_out/ncu_full_$W.log
ls -la gpurun_out/
