#!/bin/bash
# Round 2, GPU call B (1 GPU): the pipelined walk K2P (+ L2 hints) against K2 on every workload (bit-compared), the device
# sparse x sparse product, ncu on C2 / s24.
set -o pipefail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_spgemm_gpu.py tests/test_spmm_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r02b_pytest.log
V="k2 pipe:d=4 pipe:d=8 pipe:d=4,slab=256 pipe:d=8,slab=256 pipe:d=8,slab=128 pipe:d=8,l2=32 pipe:d=8,l2=64 pipe:d=8,l2=96 pipe:d=8,slab=256,l2=64 pipe:d=8,slab=128,l2=64"
timeout 900 python tools/kbench.py c2 c5 c5b c3 s24f32 c4 --steps 5 --check --variants $V 2>&1 | tee gpurun_out/r02b_sweep.jsonl
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,smsp__inst_executed.sum"
for W in c2 s24f32 c4; do for V1 in pipe:d=4 pipe:d=8 pipe:d=8,l2=64; do
  timeout 400 ncu --metrics $M --clock-control none -k "regex:cb_spmm_(pipe_)?kernel" -s 3 -c 1 --csv --log-file gpurun_out/r02b_ncu_${W}_${V1//[:=,]/_}.csv python tools/kbench.py $W --steps 1 --variants $V1 > /dev/null 2>&1
done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cb_spmm_pipe_kernel -s 3 -c 1 -o /tmp/r02b_prof_c2_pipe8 python tools/kbench.py c2 --steps 1 --variants pipe:d=8 > gpurun_out/r02b_ncu_c2_full.log 2>&1
# the report itself can exceed what gpurun copies back: keep the two pages that are read afterwards
ncu -i /tmp/r02b_prof_c2_pipe8.ncu-rep --page raw --csv > gpurun_out/r02b_prof_c2_pipe8_raw.csv 2>/dev/null
ncu -i /tmp/r02b_prof_c2_pipe8.ncu-rep --page source --csv --print-source sass > gpurun_out/r02b_prof_c2_pipe8_source.csv 2>/dev/null
du -sh gpurun_out
ls -la gpurun_out | tail -12
