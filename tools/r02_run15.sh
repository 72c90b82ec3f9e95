#!/bin/bash
tag=${1:-r02aa}
out=gpurun_out
mkdir -p $out
timeout 600 python tools/ab_dp2.py c2 c3 --modes 33:0,1:0,33:2 > $out/${tag}_ab_dp2.txt 2>&1; echo "ab rc=$?"; grep -v "bad reps" $out/${tag}_ab_dp2.txt | cut -c1-200; echo "fuzz lines with failures:"; grep "bad reps" $out/${tag}_ab_dp2.txt | grep -E ":[1-9]" | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/${tag}_pytest.log
TL=vits_b200/build_trace/libvits_mas_trace.so
VITS_MAS_LIB=$TL timeout 120 python tools/trace_dp.py c2 0 0 33 0 > $out/${tag}_trace.txt 2>&1; cut -c1-500 $out/${tag}_trace.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['other_variant']['ms_per_step'], d['e2e']['value'], d['path_breakdown']['stats_to_path_us'])"
