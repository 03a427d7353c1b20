#!/bin/bash
# build libsnb.so; non-zero exit (and the compiler errors) when the build fails
set -o pipefail
out=$(python -m switch_nerf_b200.build 2>&1); rc=$?
if [ $rc -ne 0 ]; then echo "$out" | grep -v "^ptxas info" | grep -B2 -A6 -i "error" | head -40; echo "BUILD FAILED"; exit 1; fi
grep -n "k_select\|k_front_ts\|k_back_ts" -A 3 switch_nerf_b200/csrc/build/ptxas.log | grep "Used" | sed 's/ptxas info    ://'
echo "BUILD OK"
