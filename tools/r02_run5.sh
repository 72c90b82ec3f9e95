#!/bin/bash
tag=${1:-r02k}
out=gpurun_out
run() { echo "== $*"; env "$@" timeout 100 python tools/timeline_fused.py c2 2>&1 | tee -a $out/${tag}_tl.txt; }
run MAS_FUSED_N1=3 MAS_FUSED_NOFILL=1
run MAS_FUSED_N1=3 MAS_FUSED_NOFILL=400
run MAS_FUSED_N1=3 MAS_FUSED_NOFILL=1 MAS_FUSED_GEMM_CTAS=54 MAS_FUSED_BTW=16
run MAS_FUSED_N1=3 MAS_FUSED_NOFILL=1 MAS_FUSED_GEMM_CTAS=64 MAS_FUSED_BTW=8
run MAS_FUSED_N1=2 MAS_FUSED_NOFILL=1 MAS_FUSED_GEMM_CTAS=64 MAS_FUSED_BTW=8
