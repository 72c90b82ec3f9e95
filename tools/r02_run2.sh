#!/bin/bash
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_c5.py -x -q 2>&1 | grep -v Warning | tail -60 > $out/${tag}_pytest_c5.log; tail -40 $out/${tag}_pytest_c5.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_c5.py 2>&1 | tail -25 | tee $out/${tag}_pytest.log
timeout 200 python tools/sweep_wf.py c2 > $out/${tag}_sweep_c2.txt 2>&1; cat $out/${tag}_sweep_c2.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); print(d['ms_per_step'], d['e2e']['value'], d['e2e']['python_api'])"
