#!/bin/bash
tag=${1:-r02bg}
out=gpurun_out
for dbg in 0 1 2 3; do echo "=== MAS_DP2_DBG=$dbg"; MAS_DP2_DBG=$dbg timeout 300 python tools/ab_dp2.py c3 --modes 33:0 --no-fuzz 2>&1 | grep -E "wf=" | awk 'NR%2==1' | cut -c1-120; done | tee $out/${tag}_cluster_dbg.txt
