#!/bin/bash
tag=${1:-r02ar}
out=gpurun_out
mkdir -p $out
for v in prev cur; do
  if [ $v = prev ]; then export VITS_MAS_LIB=vits_b200/build_prev/libvits_mas_prev.so; else unset VITS_MAS_LIB; fi
  echo "=== $v"; timeout 300 python tools/ab_dp2.py c2 c3 --modes 33:0 --no-fuzz 2>&1 | grep -E "wf=" | awk 'NR%2==1' | cut -c1-120
  timeout 200 python tools/timeline_gap.py c2 2>&1 | tail -11 | grep -E "DP|period" | head -3 | cut -c1-200
done > $out/${tag}_prev_vs_cur.txt 2>&1; cat $out/${tag}_prev_vs_cur.txt
unset VITS_MAS_LIB
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $out/${tag}_pytest.log
