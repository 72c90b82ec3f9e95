#!/bin/bash
# Final verification pass: GPU tests, smoke, the bench's default run (both arms).
tag=${1:-r02_final}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); e=d['e2e']; print(d['steps'], d['warmup'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['other_variant']['ms_per_step'], d['config']['other_variant']['roofline_frac'], 'e2e', e['value'], e['c_entry_pageable']['value'], e['python_api']['value'], d['path_breakdown']['neg_cent_us'], d['path_breakdown']['stats_to_path_us'], d['clocks'], d['gpu_launches'])"
timeout 300 python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; cut -c1-220 $out/${tag}_bench_ref.json
