#!/bin/bash
# Round-2 multi-GPU pass (run with gpurun --gpus 8): c5 at 2/4/8 GPUs, the weak-scaling c2 line with its c3_strong key,
# c3 strong scaling as a headline, and the reference arm at 8 ranks.
tag=${1:-r02_multi}
out=gpurun_out
mkdir -p $out
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; }
for n in 2 4 8; do
  run $n --workload c5 --steps 5 --warmup 2 > $out/${tag}_c5_n$n.json 2> $out/${tag}_c5_n$n.err; echo "c5 n=$n rc=$?"; cut -c1-200 $out/${tag}_c5_n$n.json
done
run 8 --steps 20 --warmup 5 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err; echo "bench n=8 rc=$?"; cut -c1-300 $out/${tag}_bench_n8.json
run 8 --steps 20 --warmup 5 --scaling strong --workload c3 --no-e2e > $out/${tag}_c3_strong_n8.json 2> $out/${tag}_c3_strong_n8.err; echo "c3 strong n=8 rc=$?"; cut -c1-300 $out/${tag}_c3_strong_n8.json
run 8 --impl reference --steps 20 --warmup 5 > $out/${tag}_ref_n8.json 2> $out/${tag}_ref_n8.err; echo "ref n=8 rc=$?"; cut -c1-300 $out/${tag}_ref_n8.json
timeout 300 python -m pytest tests/test_mas_gpu.py -m gpu -q -k two_devices 2>&1 | tail -2
