#!/bin/bash
tag=${1:-r02af}
out=gpurun_out
mkdir -p $out
for v in prev cur; do
  if [ $v = prev ]; then export VITS_MAS_LIB=vits_b200/build_prev/libvits_mas_prev.so; else unset VITS_MAS_LIB; fi
  echo "=== $v"; timeout 200 python tools/period_dp.py --dp2 2>&1 | cut -c1-200
done > $out/${tag}_period_prev_vs_cur.txt 2>&1; cat $out/${tag}_period_prev_vs_cur.txt
