#!/bin/bash
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests/test_binding.py tests/test_c5.py -m gpu -x -q 2>&1 | tail -3 | tee $out/${tag}_order.log
timeout 120 python tools/check_fused.py small 2>&1 | tee $out/${tag}_fused_small.txt
timeout 300 python tools/check_fused.py 2>&1 | tee $out/${tag}_fused.txt
