#!/bin/bash
tag=${1:-r02ag}
out=gpurun_out
TL=vits_b200/build_trace/libvits_mas_trace.so
for wl in c2 c3; do VITS_MAS_LIB=$TL timeout 120 python tools/trace_dp.py $wl 0 0 33 0; done > $out/${tag}_trace.txt 2>&1; grep -E "^(c2|c3|warp)|fix-up|period" $out/${tag}_trace.txt | cut -c1-330
