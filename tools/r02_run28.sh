#!/bin/bash
tag=${1:-r02bd}
out=gpurun_out
for fd in 1 2 3 4 6 8; do echo "=== MAS_FILL_DIV=$fd"; MAS_FILL_DIV=$fd timeout 200 python tools/ab_dp2.py c2 --modes 33:0 --no-fuzz 2>&1 | grep -E "wf=" | awk 'NR%2==1' | cut -c1-100; done > $out/${tag}_fill_div.txt 2>&1; cat $out/${tag}_fill_div.txt
