#!/bin/bash
tag=${1:-r02ao}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/ab_dp2.py c2 c3 --modes 33:0,37:0,41:0,45:0 --no-fuzz 2>&1 | grep -E "wf=" | cut -c1-120 > $out/${tag}_maskdelay_ab.txt; cat $out/${tag}_maskdelay_ab.txt
