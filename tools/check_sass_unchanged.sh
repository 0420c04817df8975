#!/bin/bash
# Proves that a change did not alter any kernel that was already validated on hardware: builds the library from a
# baseline commit into /tmp and compares the SASS of every kernel object by object with the working tree's build.
# usage: tools/check_sass_unchanged.sh <baseline-commit>      (round 1: 3eac267 = the library the GPU suite passed with)
set -e
BASE=${1:?baseline commit}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OLD=/tmp/sass_baseline_$BASE
rm -rf "$OLD" && mkdir -p "$OLD"
git -C "$ROOT" archive "$BASE" combblas-spmm-test_b200/csrc include | tar -x -C "$OLD"
make -s -j8 -C "$OLD/combblas-spmm-test_b200/csrc"
make -s -j8 -C "$ROOT/combblas-spmm-test_b200/csrc"
rc=0
for o in "$OLD"/combblas-spmm-test_b200/csrc/build/*.o; do
  f=$(basename "$o")
  printf "%-20s " "$f"
  python "$ROOT/tools/sass_compare.py" "$o" "$ROOT/combblas-spmm-test_b200/csrc/build/$f" | tail -1 || rc=1
done
exit $rc
