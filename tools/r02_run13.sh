#!/bin/bash
tag=${1:-r02x}
out=gpurun_out
mkdir -p $out
timeout 600 python tools/ab_dp2.py c2 c3 --modes 1:0,33:0,33:2 --no-fuzz > $out/${tag}_ab_dp2.txt 2>&1; echo "ab rc=$?"; grep -v "bad reps" $out/${tag}_ab_dp2.txt | cut -c1-200; grep "bad reps" $out/${tag}_ab_dp2.txt | grep -c -E ":[1-9]"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/${tag}_pytest.log
timeout 200 python tools/timeline_gap.py c2 --mask > $out/${tag}_tlgap_mask.txt 2>&1; echo "tl rc=$?"; tail -11 $out/${tag}_tlgap_mask.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['other_variant']['ms_per_step'], d['e2e']['value'], d['path_breakdown']['stats_to_path_us'])"
