#!/bin/bash
tag=${1:-r02ae}
out=gpurun_out
mkdir -p $out
for v in prev cur; do
  if [ $v = prev ]; then export VITS_MAS_LIB=vits_b200/build_prev/libvits_mas_prev.so; else unset VITS_MAS_LIB; fi
  for wl in c2 c3; do echo "=== $v $wl"; timeout 200 python tools/timeline_gap.py $wl 2>&1 | tail -11 | grep -E "call [0-9]:|DP|period" | head -12 | cut -c1-260; done
done > $out/${tag}_tl_prev_vs_cur.txt 2>&1; cat $out/${tag}_tl_prev_vs_cur.txt
