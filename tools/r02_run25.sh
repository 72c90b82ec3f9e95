#!/bin/bash
tag=${1:-r02aq}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/${tag}_pytest.log
timeout 300 python tools/mask_traffic.py 2>&1 | tee $out/${tag}_mask_traffic.txt
timeout 300 python tools/ab_dp2.py c2 c3 --modes 33:0,1:0 --no-fuzz 2>&1 | grep -E "wf=" | awk 'NR%2==1' | cut -c1-120 | tee $out/${tag}_ab.txt
timeout 600 python tools/ab_dp2.py --modes 33:0,1:0 > $out/${tag}_fuzz.txt 2>&1; echo "fuzz lines with failures:"; grep "bad reps" $out/${tag}_fuzz.txt | grep -E ":[1-9]" | cut -c1-300; grep -c "bad reps" $out/${tag}_fuzz.txt
