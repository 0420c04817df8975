#!/bin/bash
# round 2, K2T (bulk-copy ring) against K2: bit-identity and time on C2 / C5 / C3 / R-MAT s24 fp32
out=gpurun_out/r02_sweep_d_tma.jsonl
: > $out
timeout 240 python tools/kbench.py c2 --check --steps 10 --variants k2 tma tma:s=2 tma:s=2,w=12 tma:s=4,w=3 tma:s=2,w=4 >> $out 2>gpurun_out/tma_c2.err || echo '{"w":"c2","error":"timeout or crash"}' >> $out
timeout 240 python tools/kbench.py c5 --check --steps 10 --variants k2 tma tma:s=2 tma:s=4,w=6 >> $out 2>gpurun_out/tma_c5.err || echo '{"w":"c5","error":"timeout or crash"}' >> $out
timeout 300 python tools/kbench.py c3 --check --steps 5 --variants k2 tma tma:w=3 tma:w=4 >> $out 2>gpurun_out/tma_c3.err || echo '{"w":"c3","error":"timeout or crash"}' >> $out
timeout 300 python tools/kbench.py s24f32 --check --steps 5 --variants k2 tma tma:w=4 >> $out 2>gpurun_out/tma_s24.err || echo '{"w":"s24f32","error":"timeout or crash"}' >> $out
cat $out | cut -c1-330
tail -3 gpurun_out/tma_*.err
