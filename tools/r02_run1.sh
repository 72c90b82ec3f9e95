#!/bin/bash
# Round-2 GPU pass 1: parity tests, bench (driver's flags + long), reference arm, c5, compute-sanitizer logs.
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader | tee $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/${tag}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_driver.json 2> $out/${tag}_bench_driver.err; echo "bench(driver flags) rc=$?"; cut -c1-400 $out/${tag}_bench_driver.json
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-300 $out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; cut -c1-200 $out/${tag}_bench_ref.json
timeout 600 python bench.py --workload c5 --steps 5 --warmup 2 > $out/${tag}_c5.json 2> $out/${tag}_c5.err; echo "c5 rc=$?"; cut -c1-1500 $out/${tag}_c5.json; tail -5 $out/${tag}_c5.err
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file $out/${tag}_sanitizer_$tool.log python tools/sanitize_small.py > $out/${tag}_sanitizer_$tool.out 2>&1
  echo "$tool rc=$?"; tail -3 $out/${tag}_sanitizer_$tool.log; tail -2 $out/${tag}_sanitizer_$tool.out
done
