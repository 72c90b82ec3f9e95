#!/bin/bash
# Final pass of the round: tests, smoke, default bench (both arms), c3/c4 bench, launch list and ncu captures.
bash tools/r02_final.sh r02_final
bash tools/r02_prof.sh r02_finalp
timeout 300 python bench.py --workload c3 --steps 50 --warmup 5 --no-e2e > gpurun_out/r02_final_bench_c3.json 2> gpurun_out/r02_final_bench_c3.err
timeout 300 python bench.py --workload c4 --steps 50 --warmup 5 --no-e2e > gpurun_out/r02_final_bench_c4.json 2> gpurun_out/r02_final_bench_c4.err
python - <<'P'
import json
for w in ("c3", "c4"):
    d = json.loads(open(f"gpurun_out/r02_final_bench_{w}.json").read().strip().splitlines()[-1])
    print(w, d["ms_per_step"], d["value"], d["roofline"]["frac"], d["config"]["other_variant"]["ms_per_step"], d["config"]["other_variant"]["roofline_frac"])
P
