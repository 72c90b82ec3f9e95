#!/bin/bash
tag=${1:-r02bf}
out=gpurun_out
mkdir -p $out
timeout 600 python tools/ab_dp2.py --modes 33:0,41:0,49:0,49:2 > $out/${tag}_fuzz.txt 2>&1; echo "fuzz rc=$?"; echo "fuzz lines with failures:"; grep "bad reps" $out/${tag}_fuzz.txt | grep -E ":[1-9]" | cut -c1-400; grep -c "bad reps" $out/${tag}_fuzz.txt; tail -3 $out/${tag}_fuzz.txt | cut -c1-300
timeout 600 python tools/ab_dp2.py c2 c3 c4 --modes 41:0,33:0,49:0 --no-fuzz 2>&1 | grep -E "wf=|Error|error" | awk 'NR%2==1' | cut -c1-120 | tee $out/${tag}_ab.txt
