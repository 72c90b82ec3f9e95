#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mas_gpu.py -x -q -m gpu -k "16_bit or configs or full_size or wide_texts or generations or fuzz or edge" 2>&1 | tail -2
timeout 300 python tools/ab_dp2.py c2 --modes 33:0,33:2,49:0 > gpurun_out/bnd4_fuzz.txt 2>&1
{ for wl in c2 c3 c4; do echo "== 16-byte hand-off stores (MAS_BND4=1)"; timeout 300 python tools/ab_dp2.py $wl --no-fuzz --modes 33:0
  echo "== one store per step (MAS_BND4=0)"; VITS_MAS_LIB=vits_b200/build_bnd0/libvits_mas_bnd0.so timeout 300 python tools/ab_dp2.py $wl --no-fuzz --modes 33:0; done; } > gpurun_out/bnd4_ab.txt 2>&1
VITS_MAS_LIB=vits_b200/build_trace/libvits_mas_trace.so timeout 120 python tools/trace_dp.py c2 1 0 -1 0 > gpurun_out/c2_trace_bnd4.txt 2>&1
true
