#!/bin/bash
tag=${1:-r02w}
out=gpurun_out
mkdir -p $out
timeout 200 python tools/timeline_gap.py c2 --mask > $out/${tag}_tlgap_mask.txt 2>&1; echo "tl rc=$?"; tail -14 $out/${tag}_tlgap_mask.txt
timeout 200 python tools/timeline_gap.py c2 --mask --ragged > $out/${tag}_tlgap_mask_ragged.txt 2>&1; tail -14 $out/${tag}_tlgap_mask_ragged.txt
timeout 600 ncu --set full --clock-control none --import-source on --sampling-interval 0 -k regex:mas_dp2_kernel -s 2 -c 1 -f -o $out/${tag}_dp2_full python tools/prof_one.py c2 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 $out/${tag}_ncu_full.log
