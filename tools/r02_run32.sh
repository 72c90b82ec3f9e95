#!/bin/bash
tag=${1:-r02bj}
out=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/${tag}_pytest.log
timeout 300 python tools/fuzz.py 5 2>&1 | tail -16 | cut -c1-160
for wl in c3 c4; do timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-e2e > $out/${tag}_bench_$wl.json 2> $out/${tag}_bench_$wl.err; python -c "
import json; d=json.load(open('$out/${tag}_bench_$wl.json')); print('$wl', d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['other_variant']['ms_per_step'], d['config']['other_variant']['roofline_frac'])"; done
