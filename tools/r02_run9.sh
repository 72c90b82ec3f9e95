#!/bin/bash
# sanitizer logs, superstep trace of the wavefront DP, streamed-path mismatch diagnosis
tag=${1:-r02s}
out=gpurun_out
mkdir -p $out
TL=vits_b200/build_trace/libvits_mas_trace.so
for a in "c2 0 0" "c2 0 1" "c2 1 0"; do VITS_MAS_LIB=$TL timeout 120 python tools/trace_dp.py $a; done > $out/${tag}_trace_dp.txt 2>&1; echo "trace rc=$?"; grep -E "^(c2|warp)" $out/${tag}_trace_dp.txt
timeout 200 python tools/dbg_fused.py > $out/${tag}_dbg_fused.txt 2>&1; echo "dbg rc=$?"; cat $out/${tag}_dbg_fused.txt | cut -c1-400
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > $out/${tag}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $out/${tag}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > $out/${tag}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $out/${tag}_sanitizer_racecheck.log
