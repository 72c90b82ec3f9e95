#!/bin/bash
tag=${1:-r02ah}
out=gpurun_out
mkdir -p $out
for v in prev cur; do
  if [ $v = prev ]; then export VITS_MAS_LIB=vits_b200/build_prev/libvits_mas_prev.so; else unset VITS_MAS_LIB; fi
  echo "=== $v"; timeout 300 python tools/ab_dp2.py c2 c3 --modes 33:0 --no-fuzz 2>&1 | grep -E "wf=" | awk 'NR%2==1' | cut -c1-120
done > $out/${tag}_prev_vs_cur.txt 2>&1; cat $out/${tag}_prev_vs_cur.txt
unset VITS_MAS_LIB
timeout 600 python tools/ab_dp2.py --modes 33:0,33:2 > $out/${tag}_fuzz.txt 2>&1; echo "fuzz lines with failures:"; grep "bad reps" $out/${tag}_fuzz.txt | grep -E ":[1-9]" | cut -c1-300; grep -c "bad reps" $out/${tag}_fuzz.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $out/${tag}_pytest.log
TL=vits_b200/build_trace/libvits_mas_trace.so
for wl in c2; do VITS_MAS_LIB=$TL timeout 120 python tools/trace_dp.py $wl 0 0 33 0; done > $out/${tag}_trace.txt 2>&1; grep -E "^(c2|c3|warp)" $out/${tag}_trace.txt | cut -c1-330
