#!/bin/bash
# Round-2 profiling pass at the final code: bench (both arms), ncu launch list, full capture of the dominant kernel
# (mas_dp2_kernel, dense warp sampling), full capture of one call's three kernels with variable lengths (DRAM traffic).
tag=${1:-r02p}
out=gpurun_out
mkdir -p $out
timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('$out/${tag}_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['other_variant']['ms_per_step'], d['e2e']['value'], d['path_breakdown'])"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; cut -c1-200 $out/${tag}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu --no-breakdown > $out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:mas_dp2_kernel -s 2 -c 1 -f -o $out/${tag}_mas_dp2_full \
  python tools/prof_one.py c2 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 $out/${tag}_ncu_full.log
timeout 600 ncu --set full --clock-control none -k regex:mas_ -s 6 -c 3 -f -o $out/${tag}_chain_c2_variable_lengths \
  python tools/prof_one.py c2 --ragged > $out/${tag}_ncu_chain.log 2>&1; echo "ncu chain rc=$?"; tail -2 $out/${tag}_ncu_chain.log
ls -la $out/${tag}_*
