#!/bin/bash
mkdir -p gpurun_out
for w in kamani sine; do
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${w}_graph.json 2> gpurun_out/r2_bench_${w}_graph.err
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --graph off > gpurun_out/r2_bench_${w}_eager.json 2> gpurun_out/r2_bench_${w}_eager.err
tail -2 gpurun_out/r2_bench_${w}_graph.err
done
python - <<'PY'
import json
for w in ("kamani","sine"):
    for m in ("graph","eager"):
        d=json.load(open(f"gpurun_out/r2_bench_{w}_{m}.json"))
        print(w,m,round(d["value"],1),round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"graph",d["cuda_graph"],"launches",d["gpu_launches"])
PY
