#!/bin/bash
tag=${1:-r02am}
out=gpurun_out
mkdir -p $out
for v in prev cur; do
  if [ $v = prev ]; then export VITS_MAS_LIB=vits_b200/build_prev/libvits_mas_prev.so; m=33:0; else unset VITS_MAS_LIB; m=33:0,37:0; fi
  echo "=== $v"; timeout 300 python tools/ab_dp2.py c2 c3 --modes $m --no-fuzz 2>&1 | grep -E "wf=" | cut -c1-120
done > $out/${tag}_prev_vs_cur.txt 2>&1; cat $out/${tag}_prev_vs_cur.txt
