#!/bin/bash
# DRAM bytes and L2 hit rate of K2 against K2W on R-MAT s24 fp32 (ncu metric pass, --clock-control none; times under ncu are not bench values)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum \
    --clock-control none --kernel-name regex:cb_spmm_kernel --launch-count 8 --csv --log-file gpurun_out/r02_ncu_l2window_s24f32.csv \
    python tools/kbench.py s24f32 --steps 1 --variants k2 win:mb=64 > gpurun_out/ncu_win.log 2>&1
tail -n 12 gpurun_out/r02_ncu_l2window_s24f32.csv | cut -c1-300
