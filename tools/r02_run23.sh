#!/bin/bash
tag=${1:-r02an}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/ab_dp2.py c2 c3 --modes 33:0,35:0 --no-fuzz 2>&1 | grep -E "wf=" | cut -c1-120 > $out/${tag}_prefetch_ab.txt; cat $out/${tag}_prefetch_ab.txt
timeout 600 python tools/ab_dp2.py --modes 35:0,35:2 > $out/${tag}_fuzz.txt 2>&1; echo "fuzz lines with failures:"; grep "bad reps" $out/${tag}_fuzz.txt | grep -E ":[1-9]" | cut -c1-300; grep -c "bad reps" $out/${tag}_fuzz.txt
