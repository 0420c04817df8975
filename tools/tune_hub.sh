#!/bin/bash
# Hub variant (K2H) against plain K2 on one GPU: cluster size x slab width sweep.  usage: tools/tune_hub.sh "<workloads>"
# First hardware run of K2H: run `CB_TEST_NEW=1 python -m pytest tests/test_new_variants_gpu.py -x -q` before trusting any number.
W=${1:-"c2 c5 s24f32"}
echo "== plain K2"; python tools/kbench.py $W --steps 5 | cut -c1-220
for CS in 1 2 4 8; do
  for SLAB in 0 128 256; do
    echo "== hub cluster=$CS slab=$SLAB"
    CB_SPMM_HUB=1 CB_SPMM_HUB_CLUSTER=$CS CB_SPMM_HUB_SLAB=$SLAB timeout 600 python tools/kbench.py $W --steps 5 | cut -c1-220
  done
done
echo "== ring only (K2R, no hub rows)"; CB_SPMM_RING=8 timeout 600 python tools/kbench.py $W c3 --steps 5 | cut -c1-220
for CS in 2 4; do
  echo "== ring + hub cluster=$CS"
  CB_SPMM_HUB=1 CB_SPMM_HUB_CLUSTER=$CS CB_SPMM_RING=8 timeout 600 python tools/kbench.py $W --steps 5 | cut -c1-220
done
echo "== narrow panels: 4-lane layout vs CB_K2_NARROW=1 (boolean k=32, SpMV-like k=1,4,8)"
python tools/kbench.py c5b --steps 5 | cut -c1-220; CB_K2_NARROW=1 python tools/kbench.py c5b --steps 5 | cut -c1-220
for K in 1 4 8; do python tools/kbench.py c2 --k $K --steps 5 | cut -c1-220; CB_K2_NARROW=1 python tools/kbench.py c2 --k $K --steps 5 | cut -c1-220; done
