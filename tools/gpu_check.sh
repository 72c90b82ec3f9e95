#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list, one full ncu capture of the forward kernel.
# usage (from the repo root, on the GPU box):  bash tools/gpu_check.sh [tag]
tag=${1:-run}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cat $out/${tag}_bench.json | cut -c1-600
python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $out/${tag}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
  python bench.py --steps 8 --warmup 3 --no-graph --no-e2e --no-cpu --no-breakdown > $out/${tag}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mas_dp_kernel -s 2 -c 1 -f -o $out/${tag}_fwd_full \
  python tools/prof_one.py c2 > $out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:neg_cent_tc_kernel -s 1 -c 1 -f -o $out/${tag}_nc_full \
  python tools/trace_nc.py > $out/${tag}_ncu_nc_full.log 2>&1; echo "ncu neg_cent full rc=$?"
