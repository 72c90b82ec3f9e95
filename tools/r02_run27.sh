#!/bin/bash
tag=${1:-r02at}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $out/${tag}_pytest.log
for v in prev cur; do
  if [ $v = prev ]; then export VITS_MAS_LIB=vits_b200/build_prev/libvits_mas_prev.so; else unset VITS_MAS_LIB; fi
  echo "=== $v"; timeout 300 python tools/ab_dp2.py c2 c3 --modes 33:0 --no-fuzz 2>&1 | grep -E "wf=" | awk 'NR%2==1' | cut -c1-120
done > $out/${tag}_prev_vs_cur.txt 2>&1; cat $out/${tag}_prev_vs_cur.txt
unset VITS_MAS_LIB
timeout 200 python tools/timeline_gap.py c2 2>&1 | tail -12 | grep -E "after|DP|period" | cut -c1-260 | tee $out/${tag}_tail.txt
timeout 600 python tools/ab_dp2.py --modes 33:0,1:0 > $out/${tag}_fuzz.txt 2>&1; echo "fuzz lines with failures:"; grep "bad reps" $out/${tag}_fuzz.txt | grep -E ":[1-9]" | cut -c1-300; grep -c "bad reps" $out/${tag}_fuzz.txt
timeout 300 python tools/fuzz.py 3 2>&1 | tail -16 | cut -c1-200
