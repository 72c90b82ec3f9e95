#!/bin/bash
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
TL=vits_b200/build_trace/libvits_mas_trace.so
for m in 0 1; do VITS_MAS_LIB=$TL timeout 120 python tools/trace_dp.py c2 0 0 33 $m; done > $out/${tag}_trace_mask.txt 2>&1; echo "trace rc=$?"; cut -c1-600 $out/${tag}_trace_mask.txt
timeout 600 python tools/ab_dp2.py c2 --modes 33:0,35:0,1:0 --no-fuzz > $out/${tag}_ab_dp2.txt 2>&1; echo "ab rc=$?"; cat $out/${tag}_ab_dp2.txt | cut -c1-200
