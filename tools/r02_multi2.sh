#!/bin/bash
# Round-2 final multi-GPU pass (gpurun --gpus 8) with mas_dp2: weak-scaling c2 line (+ its c3_strong key), c5 at 8 GPUs,
# c3 strong scaling as a headline, reference arm at 8 ranks.
tag=${1:-r02_multi2}
out=gpurun_out
mkdir -p $out
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; }
run 8 --steps 20 --warmup 5 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err; echo "bench n=8 rc=$?"; cut -c1-260 $out/${tag}_bench_n8.json
run 8 --workload c5 --steps 5 --warmup 2 > $out/${tag}_c5_n8.json 2> $out/${tag}_c5_n8.err; echo "c5 n=8 rc=$?"; cut -c1-200 $out/${tag}_c5_n8.json
run 8 --steps 20 --warmup 5 --scaling strong --workload c3 --no-e2e > $out/${tag}_c3_strong_n8.json 2> $out/${tag}_c3_strong_n8.err; echo "c3 strong n=8 rc=$?"; cut -c1-260 $out/${tag}_c3_strong_n8.json
run 2 --steps 20 --warmup 5 --no-e2e > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err; echo "bench n=2 rc=$?"; cut -c1-260 $out/${tag}_bench_n2.json
