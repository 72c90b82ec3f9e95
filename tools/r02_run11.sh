#!/bin/bash
tag=${1:-r02v}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/${tag}_pytest.log
timeout 300 python tools/period_dp.py > $out/${tag}_period.txt 2>&1; echo "period rc=$?"; cat $out/${tag}_period.txt
timeout 200 python tools/timeline_gap.py c2 > $out/${tag}_tlgap.txt 2>&1; echo "tl rc=$?"; tail -45 $out/${tag}_tlgap.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['other_variant']['ms_per_step'], d['e2e']['value'], d['path_breakdown']['stats_to_path_us'])"
