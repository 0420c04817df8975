#!/bin/bash
# First GPU call of round 2 (1 GPU, ~12 min): validate the hub variant (K2H) on hardware, then measure it.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
# 1. the regular GPU suite (must stay green), 2. K2H / K2R parity tests (bit-identity with K2), 3. K2 vs K2H vs K2R sweep,
# 4. launch list + one full ncu capture of the best-looking K2H point on C2.  Everything lands in gpurun_out/.
set -o pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
CB_TEST_NEW=1 timeout 900 python -m pytest tests/test_new_variants_gpu.py -x -q 2>&1 | tail -25 | tee gpurun_out/r02_pytest_hub.log
if grep -q "passed" gpurun_out/r02_pytest_hub.log && ! grep -q "failed" gpurun_out/r02_pytest_hub.log; then
  timeout 1200 bash tools/tune_hub.sh "c2 c5 s24f32" 2>&1 | tee gpurun_out/r02_tune_hub.log
  CS=${HUB_CS:-4}; SLAB=${HUB_SLAB:-0}
  CB_SPMM_HUB=1 CB_SPMM_HUB_CLUSTER=$CS CB_SPMM_HUB_SLAB=$SLAB timeout 300 python tools/kbench.py c2 --steps 3 > gpurun_out/r02_hub_plain.log 2>&1 &&
  CB_SPMM_HUB=1 CB_SPMM_HUB_CLUSTER=$CS CB_SPMM_HUB_SLAB=$SLAB timeout 600 ncu --set full --clock-control none --import-source on -k "regex:cb_spmm_(hub|ring)_kernel" -c 2 \
      -o gpurun_out/r02_prof_c2_hub python tools/kbench.py c2 --steps 1 > gpurun_out/r02_ncu_hub.log 2>&1
else
  echo "K2H parity tests did not pass: no measurements taken" | tee gpurun_out/r02_tune_hub.log
fi
