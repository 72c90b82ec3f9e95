#!/bin/bash
tag=${1:-r02i}
out=gpurun_out
for n1 in 0 1 2 3; do
  echo "== MAS_FUSED_N1=$n1"; MAS_FUSED_N1=$n1 timeout 100 python tools/timeline_fused.py c2 2>&1 | tee -a $out/${tag}_tl.txt
done
echo "== auto"; timeout 300 python tools/check_fused.py 2>&1 | tee $out/${tag}_fused.txt
