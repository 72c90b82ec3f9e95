#!/bin/bash
tag=${1:-r02q}
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $out/${tag}_pytest.log
timeout 300 python tools/check_fused.py 2>&1 | tee $out/${tag}_fused.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); print(d['ms_per_step'], d['e2e']['value'], d['e2e']['python_api']['value'], d['path_breakdown']['stats_to_path_us'])"
