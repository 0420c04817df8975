#!/bin/bash
# K2 tuning sweep over variant builds x forced operating point, 1 GPU.  usage: tools/tune.sh "<workloads>" [lib ...]
W=${1:-"c2 c5 c3 s24f32"}; shift
for L in "" "$@"; do
  for P in 0 1; do
    echo "== ${L:-default} point=$P (0 deep, 1 wide)"
    CB_K2_POINT=$P CB_LIB=$L python tools/kbench.py $W --steps 5 | cut -c1-200
  done
done
