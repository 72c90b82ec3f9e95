#!/bin/bash
tag=${1:-r02bl}
out=gpurun_out
timeout 600 python -m pytest tests/test_mas_gpu.py tests/test_binding.py tests/test_abi.py -m gpu -x -q 2>&1 | tail -2
for ns in 0 1; do MAS_HOST_NOSTAGE=$ns timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-breakdown > $out/${tag}_bench_nostage$ns.json 2> $out/${tag}_bench_nostage$ns.err; python -c "
import json; d=json.load(open('$out/${tag}_bench_nostage$ns.json')); e=d['e2e']; print('NOSTAGE=$ns', 'step', d['ms_per_step'], 'e2e pinned', e['value'], 'pageable C entry', e['c_entry_pageable']['value'], 'python api', e['python_api']['value'])"; done
