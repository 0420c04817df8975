#!/bin/bash
# Round 2, GPU call A (1 GPU): the regular suite, the variants that were written without a GPU, ceilings, sweep, ncu on the target.
set -o pipefail
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02a_pytest_gpu.log
CB_TEST_NEW=1 timeout 900 python -m pytest tests/test_new_variants_gpu.py -q -n 4 2>&1 | tail -30 | tee gpurun_out/r02a_pytest_new.log
timeout 300 tools/gather_probe 2>&1 | tee gpurun_out/r02a_gather_probe.jsonl
export CB_SPMM_HUB_MIN_COVER_PCT=1
K2V="k2 k2:point=0 k2:point=1 k2:slab=256 k2:slab=128 k2:slab=128,point=1 k2:slab=64"
HUBV="hub:c=1,slab=128 hub:c=1,slab=256 hub:c=1 hub:c=2,slab=128 hub:c=2,slab=256 hub:c=4,slab=128 hub:c=4,slab=256 hub:c=4 hub:c=8,slab=128 hub:c=8,slab=256 ring hub:c=2,slab=128,ring=8 hub:c=4,slab=256,ring=8"
timeout 900 python tools/kbench.py c2 c5 s24f32 c4 --steps 5 --check --variants $K2V $HUBV 2>&1 | tee gpurun_out/r02a_sweep.jsonl
timeout 300 python tools/kbench.py c3 c5b --steps 5 --check --variants $K2V ring 2>&1 | tee -a gpurun_out/r02a_sweep.jsonl
CB_LIB=$PWD/combblas-spmm-test_b200/lib_var/libcombblas_b200_fma.so timeout 300 python tools/kbench.py c2 s24f32 c4 --steps 5 --variants k2 k2:slab=128 2>&1 | tee gpurun_out/r02a_fma.jsonl
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cb_spmm_kernel -s 3 -c 1 -o gpurun_out/r02a_prof_s24f32_k2 python tools/kbench.py s24f32 --steps 1 > gpurun_out/r02a_ncu_s24f32.log 2>&1
for V in k2:slab=128 k2:slab=256 hub:c=1,slab=128 hub:c=4,slab=128; do
  timeout 400 ncu --metrics $M --clock-control none -k "regex:cb_spmm_(hub_|ring_)?kernel" -s 3 -c 1 --csv --log-file gpurun_out/r02a_ncu_s24f32_${V//[:=,]/_}.csv python tools/kbench.py s24f32 --steps 1 --variants $V > /dev/null 2>&1
done
timeout 400 ncu --metrics $M --clock-control none -k "regex:cb_spmm_(hub_|ring_)?kernel" -s 3 -c 1 --csv --log-file gpurun_out/r02a_ncu_c4_k2.csv python tools/kbench.py c4 --steps 1 > /dev/null 2>&1
for V in hub:c=1,slab=128 hub:c=4,slab=128; do
  timeout 400 ncu --metrics $M --clock-control none -k "regex:cb_spmm_(hub_|ring_)?kernel" -s 3 -c 1 --csv --log-file gpurun_out/r02a_ncu_c2_${V//[:=,]/_}.csv python tools/kbench.py c2 --steps 1 --variants $V > /dev/null 2>&1
done
ls -la gpurun_out | tail -30
