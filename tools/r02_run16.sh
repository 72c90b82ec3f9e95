#!/bin/bash
tag=${1:-r02ac}
out=gpurun_out
mkdir -p $out
for rep in 1 2; do
for v in prev cur; do
  if [ $v = prev ]; then export VITS_MAS_LIB=vits_b200/build_prev/libvits_mas_prev.so; else unset VITS_MAS_LIB; fi
  echo "=== $v (rep $rep)"; timeout 300 python tools/ab_dp2.py c2 c3 --modes 33:0,1:0 --no-fuzz 2>&1 | grep -E "wf=" | awk 'NR%2==1' | cut -c1-120
done; done > $out/${tag}_prev_vs_cur.txt 2>&1; cat $out/${tag}_prev_vs_cur.txt
