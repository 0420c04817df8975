#!/bin/bash
# round 2: one ncu --set full capture each of K2T (bulk-copy ring) on C2 and K2W (hub panel under an L2 window) on R-MAT s24 fp32
# (--clock-control none; times under ncu are not bench values).  The reports stay on the box (2 x 48 MB); the summaries come back.
ncu --set full --clock-control none --import-source on --kernel-name regex:cb_spmm_tma_kernel --launch-skip 3 --launch-count 1 -f -o /tmp/r02_k2t_c2 \
    python tools/kbench.py c2 --steps 1 --variants tma:s=2,w=4 > gpurun_out/ncu_k2t.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:cb_spmm_kernel --launch-skip 3 --launch-count 1 -f -o /tmp/r02_k2w_s24f32 \
    python tools/kbench.py s24f32 --steps 1 --variants win:mb=64 > gpurun_out/ncu_k2w.log 2>&1
python tools/ncu_summary.py full /tmp/r02_k2t_c2.ncu-rep gpurun_out/r02_k2t_c2_full.md
python tools/ncu_summary.py full /tmp/r02_k2w_s24f32.ncu-rep gpurun_out/r02_k2w_s24f32_full.md
ls -la gpurun_out/*.md
