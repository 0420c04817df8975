#!/bin/bash
# K2 tuning sweep over variant builds (lib_var/), 1 GPU
for L in "" $(ls combblas-spmm-test_b200/lib_var/*.so); do
  echo "== ${L:-default}"
  CB_LIB=$L python tools/kbench.py c2 c5 s24f32 --steps 5 | cut -c1-200
  CB_LIB=$L python tools/kbench.py c3 --steps 5 --k 32 | cut -c1-200
done
