#!/bin/bash
# c4 / cluster pass: tests on the cluster path, A/B of the cluster forms, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mas_gpu.py -x -q -m gpu -k "16_bit or configs or full_size or wide_texts or generations" > gpurun_out/c4_pytest.log 2>&1
tail -3 gpurun_out/c4_pytest.log
for wl in c3 c4; do timeout 300 python tools/ab_dp2.py $wl --no-fuzz --modes 49:0,51:0,33:0,35:0 ; done > gpurun_out/c4_ab2.txt 2>&1
timeout 300 python bench.py --workload c4 --steps 50 --warmup 5 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
tail -c 300 gpurun_out/c4_bench.json
true
