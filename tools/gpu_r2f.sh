#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "coarse_solve or benchmarked or layer_vs_reference or chain or kamani or interleaved" > gpurun_out/r2f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_gl32.json 2> gpurun_out/r2f_bench_gl32.err
PDEOP_FACTOR_LOOKAHEAD=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_gl32_nolook.json 2> gpurun_out/r2f_bench_gl32_nolook.err
python - <<'PY'
import json
for n in ("r2f_bench_gl32","r2f_bench_gl32_nolook"):
    d=json.load(open(f"gpurun_out/{n}.json"))
    print(n, round(d["value"],2), round(d["ms_per_step"],1), {k:round(v["ms_per_step"],1) for k,v in d["breakdown"].items()})
PY
