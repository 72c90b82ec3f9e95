#!/bin/bash
tag=${1:-r02t}
out=gpurun_out
mkdir -p $out
timeout 600 python tools/ab_dp2.py c2 c3 > $out/${tag}_ab_dp2.txt 2>&1; echo "ab rc=$?"; cat $out/${tag}_ab_dp2.txt | cut -c1-250
TL=vits_b200/build_trace/libvits_mas_trace.so
for wf in 1 33; do VITS_MAS_LIB=$TL timeout 120 python tools/trace_dp.py c2 0 0 $wf; done > $out/${tag}_trace_dp2.txt 2>&1; echo "trace rc=$?"; grep -E "^(c2|warp)" $out/${tag}_trace_dp2.txt; grep -A5 "wf=33" $out/${tag}_trace_dp2.txt | cut -c1-400
