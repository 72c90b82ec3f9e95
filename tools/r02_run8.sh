#!/bin/bash
tag=${1:-r02r}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $out/${tag}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); e=d['e2e']; print(d['ms_per_step'], e['value'], e['python_api']['value'], e['vs_cpu'], e['pcie_link']['h2d_bound'], d['path_breakdown']['stats_to_path_us'])"
MAS_HOST_DENSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-breakdown > $out/${tag}_bench_dense_d2h.json 2> /dev/null; python -c "
import json; d=json.load(open('$out/${tag}_bench_dense_d2h.json')); print('dense d2h e2e', d['e2e']['value'])"
for t in 2 4 16; do MAS_HOST_THREADS=$t timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-breakdown 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('threads $t e2e', d['e2e']['value'])"; done
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; cut -c1-160 $out/${tag}_bench_ref.json
