#!/bin/bash
tag=${1:-r02bi}
out=gpurun_out
for ps in 64 200 500 1000; do echo "=== MAS_DP2_PSLEEP=$ps"; MAS_DP2_PSLEEP=$ps timeout 300 python tools/ab_dp2.py c3 c4 --modes 33:0 --no-fuzz 2>&1 | grep -E "wf=" | awk 'NR%2==1' | cut -c1-120; done | tee $out/${tag}_psleep.txt
